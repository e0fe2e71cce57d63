"""Host-side mirror of the reference's `Spectral` / `MechBase` call surface for the device path.

Names follow /root/reference/src/grid/spectral.h:27-41 and mech_base.h:43-73; every method forwards to the C ABI
(include/dmk_grid.h).  Fields are torch CUDA tensors of float64 holding the reference's Fortran memory order:
shape (nz_local, ny, nx, 9) with component c = i + 3*j, i.e. natural[..., i, j] == tensor[..., i + 3*j]."""
import ctypes as C
import sys

import numpy as np
import torch

from . import lib as L


def natural_to_records(a):
    """numpy [..., 3, 3] (natural i,j) -> [..., 9] Fortran component order c = i + 3 j"""
    a = np.asarray(a, float)
    return np.ascontiguousarray(np.swapaxes(a, -1, -2)).reshape(a.shape[:-2] + (9,))


def records_to_natural(r):
    r = np.asarray(r)
    return np.swapaxes(r.reshape(r.shape[:-1] + (3, 3)), -1, -2)


def c3333_to_fortran(Cn):
    """natural C[i,j,k,l] -> flat Fortran order (i fastest)"""
    return np.ascontiguousarray(np.transpose(np.asarray(Cn, float), (3, 2, 1, 0))).reshape(-1)


def fortran_to_c3333(f):
    return np.transpose(np.asarray(f, float).reshape(3, 3, 3, 3), (3, 2, 1, 0))


def _p(t):
    return C.c_void_p(t.data_ptr())


def _hp(a):
    return a.ctypes.data_as(C.c_void_p)


class GridContext:
    """Owns a dmk_ctx for one z-slab on the current CUDA device."""

    def __init__(self, cells, size, derivative=0, divergence_correction=2, world_size=1, rank=0, kt_y=0, kt_z=0,
                 device=None):
        if not torch.cuda.is_available():
            raise L.DmkError('no CUDA device: solver_cpp_b200 has no CPU fallback')
        self.lib = L.load()
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        g = L.GridDesc()
        g.cells[:] = [int(c) for c in cells]
        g.size[:] = [float(s) for s in size]
        g.world_size, g.rank = int(world_size), int(rank)
        g.derivative, g.divergence_correction, g.kt_y, g.kt_z = derivative, divergence_correction, kt_y, kt_z
        self.ctx = C.c_void_p()
        L.check(self.lib.dmk_create(C.byref(self.ctx), C.byref(g)))
        L.check(self.lib.dmk_set_stream(self.ctx, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        self.cells = tuple(int(c) for c in cells)
        self.size = tuple(float(s) for s in size)
        nzl, zo, nyl, yo = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        L.check(self.lib.dmk_local_cells(self.ctx, C.byref(nzl), C.byref(zo), C.byref(nyl), C.byref(yo)))
        self.nz_local, self.z_offset, self.ny_local, self.y_offset = nzl.value, zo.value, nyl.value, yo.value
        nx, ny, nz = self.cells
        self.nxr = nx // 2 + 1
        self.n_local = nx * ny * self.nz_local
        self.field_shape = (self.nz_local, ny, nx, 9)
        self.fourier_shape = (nz, self.ny_local, self.nxr, 9, 2)
        self.world_size, self.rank = world_size, rank
        self._comm_keepalive = None

    def close(self):
        if self.ctx:
            self.lib.dmk_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers
    def new_field(self):
        return torch.empty(self.field_shape, dtype=torch.float64, device=self.device)

    def to_device(self, natural):
        return torch.from_numpy(natural_to_records(natural)).to(self.device).reshape(self.field_shape).contiguous()

    def to_natural(self, t):
        return records_to_natural(t.detach().cpu().numpy())

    def synchronize(self):
        L.check(self.lib.dmk_synchronize(self.ctx))

    def kernel_launches(self):
        n = C.c_longlong()
        L.check(self.lib.dmk_kernel_launches(self.ctx, C.byref(n)))
        return n.value

    # ---- FFT<5>::forward / backward (fft.h:37-60)
    def fft_forward(self, field):
        out = torch.empty(self.fourier_shape, dtype=torch.float64, device=self.device)
        L.check(self.lib.dmk_fft_forward(self.ctx, _p(field), _p(out)))
        return out

    def fft_backward(self, hat, wgt=1.0):
        out = self.new_field()
        L.check(self.lib.dmk_fft_backward(self.ctx, _p(hat), _p(out), C.c_double(wgt)))
        return out

    # ---- MechBase::update_gamma (mech_base.cpp:156)
    def update_gamma(self, C_natural):
        f = c3333_to_fortran(C_natural)
        L.check(self.lib.dmk_update_gamma(self.ctx, _hp(f)))

    # ---- calculate_divergence_rms + gamma_convolution (mech_base.cpp:302-477), in place on `field`
    def divergence_and_gamma(self, field, field_aim_natural):
        aim = natural_to_records(np.asarray(field_aim_natural, float))
        err = C.c_double()
        L.check(self.lib.dmk_divergence_and_gamma(self.ctx, _p(field), _hp(aim), C.byref(err)))
        return err.value

    # ---- MechBase::base_update_coords (mech_base.cpp:45-154): deformed node / cell-centre coordinates of the local slab
    def update_coords(self, F):
        nx, ny, _ = self.cells
        x_n = np.zeros((self.nz_local + 1, ny + 1, nx + 1, 3))
        x_p = np.zeros((self.nz_local, ny, nx, 3))
        L.check(self.lib.dmk_update_coords(self.ctx, _p(F), _hp(x_n), _hp(x_p)))
        return x_n, x_p

    # ---- MechBase::forward_field / calculate_rate (mech_base.cpp:213-247, 479-497)
    def forward_field(self, dt, field_last, rate, aim_natural=None):
        out = self.new_field()
        aim = None if aim_natural is None else natural_to_records(np.asarray(aim_natural, float))
        L.check(self.lib.dmk_forward_field(self.ctx, C.c_double(dt), _p(field_last), _p(rate),
                                           None if aim is None else _hp(aim), _p(out)))
        return out

    def calculate_rate(self, heterogeneous, field0, field, dt, av_rate_natural):
        out = self.new_field()
        av = natural_to_records(np.asarray(av_rate_natural, float))
        L.check(self.lib.dmk_calculate_rate(self.ctx, C.c_int(int(heterogeneous)), _p(field0), _p(field),
                                            C.c_double(dt), _hp(av), _p(out)))
        return out

    # ---- f_materialpoint_initDamask + f_discretization_init for the hot path
    def material_init(self, material, numerics=None):
        """material: solver_cpp_b200.material.Material for THIS slab (cell order x fastest)"""
        from . import material as M
        if self.lib.dmk_sizeof_phase() != C.sizeof(L.Phase) or self.lib.dmk_sizeof_numerics() != C.sizeof(L.Numerics):
            raise L.DmkError('ctypes layout of dmk_phase / dmk_numerics does not match the library')
        if material.n_cells != self.n_local:
            raise L.DmkError('material has %d cells, slab has %d' % (material.n_cells, self.n_local))
        self.numerics = numerics if numerics is not None else M.make_numerics()
        self._material = material
        self.n_constituents = material.n_constituents
        self.n_points = self.n_local * self.n_constituents       # rows of the per-point arrays (get_state / set_state)
        if self.n_constituents == 1:
            L.check(self.lib.dmk_material_init(self.ctx, material.phases, C.c_int(len(material.phases)),
                                               _hp(material.phase_of_cell), _hp(material.O_of_cell), C.byref(self.numerics)))
        else:
            # isostrain homogenisation: point k * n_local + cell is constituent k of the cell
            L.check(self.lib.dmk_material_init_constituents(self.ctx, material.phases, C.c_int(len(material.phases)),
                                                            C.c_int(self.n_constituents), _hp(material.phase_of_point),
                                                            _hp(material.O_of_point), _hp(material.v_of_point),
                                                            C.byref(self.numerics)))

    # ---- f_homogenization_mechanical_response (+ P of _response2)
    def constitutive_response(self, F, dt, P=None):
        P = self.new_field() if P is None else P
        L.check(self.lib.dmk_constitutive_response(self.ctx, _p(F), _p(P), C.c_double(dt)))
        return P

    # ---- dPdF reductions of Spectral::constitutive_response (spectral.cpp:189-218)
    def tangent_response(self, dt, materialise=False):
        Cv, Cm = np.zeros(81), np.zeros(81)
        cells = (C.c_longlong * 2)()
        d = torch.empty((self.n_local, 81), dtype=torch.float64, device=self.device) if materialise else None
        L.check(self.lib.dmk_tangent_response(self.ctx, C.c_double(dt), _hp(Cv), _hp(Cm), None if d is None else _p(d), cells))
        out = (fortran_to_c3333(Cv), fortran_to_c3333(Cm), (cells[0], cells[1]))
        return out + (d,) if materialise else out

    def homogenise(self, P, dPdF):
        """P_av, C_volAvg, C_minMaxAvg and the two selected cells from GIVEN device fields (Spectral::constitutive_response,
        spectral.cpp:160-218): P records (nz_local, ny, nx, 9), dPdF (n_local, 81) in the reference's Fortran order"""
        Pav, Cv, Cm = np.zeros(9), np.zeros(81), np.zeros(81)
        cells = (C.c_longlong * 2)()
        L.check(self.lib.dmk_homogenise(self.ctx, _p(P), _p(dPdF), C.c_longlong(self.n_local), _hp(Pav), _hp(Cv), _hp(Cm), cells))
        return Pav.reshape(3, 3).T, fortran_to_c3333(Cv), fortran_to_c3333(Cm), (cells[0], cells[1])

    def materialpoint_forward(self):
        L.check(self.lib.dmk_materialpoint_forward(self.ctx))

    def terminally_ill(self, reset=False):
        f = C.c_int()
        L.check(self.lib.dmk_terminally_ill(self.ctx, C.byref(f), C.c_int(int(reset))))
        return bool(f.value)

    def counters(self):
        a = (C.c_longlong * 7)()
        L.check(self.lib.dmk_get_counters(self.ctx, a))
        return dict(zip(('Lp_iter', 'stress_calls', 'state_iter', 'substeps', 'cutbacks', 'ill_cells', 'deferred'), list(a)))

    def schedule_stats(self):
        a = (C.c_longlong * 8)()
        L.check(self.lib.dmk_get_schedule_stats(self.ctx, a))
        return dict(zip(('lp_trips', 'lp_lanes', 'sb_trips', 'sb_lanes', 'refill_trips', 'refill_lanes'), list(a)[:6]))

    def get_state(self, name, comps):
        out = np.zeros((getattr(self, 'n_points', self.n_local), comps))
        L.check(self.lib.dmk_get_state(self.ctx, name.encode(), _hp(out), C.c_size_t(out.size)))
        return out

    RESTART_ARRAYS = (('F', 9), ('F0', 9), ('Fp', 9), ('Fp0', 9), ('Fe', 9), ('Lp', 9), ('Lp0', 9), ('S', 9), ('S0', 9))

    def set_state(self, name, array):
        a = np.ascontiguousarray(array, dtype=float)
        L.check(self.lib.dmk_set_state(self.ctx, name.encode(), _hp(a), C.c_size_t(a.size)))

    def save_point_state(self, path):
        """restart data of the per-point integrator (phase_mechanical.f90:1264-1298 writes the same set to the restart file)"""
        d = {k: self.get_state(k, n) for k, n in self.RESTART_ARRAYS}
        ms = int(self.lib.dmk_max_state(self.ctx))
        if ms > 0:
            d['state'], d['state0'] = self.get_state('state', ms), self.get_state('state0', ms)
        md = int(self.lib.dmk_max_dep(self.ctx))
        if md > 0:
            d['dep'] = self.get_state('dep', md)
        np.savez(path, **d)

    def load_point_state(self, path):
        d = np.load(path)
        for k in d.files:
            self.set_state(k, d[k])

    def set_option(self, name, value):
        L.check(self.lib.dmk_set_option(self.ctx, name.encode(), C.c_int(int(value))))

    # ---- measurement aids
    def set_profiling(self, enable=True):
        L.check(self.lib.dmk_set_profiling(self.ctx, C.c_int(int(enable))))

    def get_profile(self, reset=True):
        out = (C.c_double * 6)()
        L.check(self.lib.dmk_get_profile(self.ctx, out, C.c_int(int(reset))))
        keys = ('constitutive_ms', 'reduce_aim_ms', 'fft_forward_ms', 'gamma_zpass_ms', 'fft_inverse_ms', 'evaluations')
        return dict(zip(keys, list(out)))

    def fp64_peak_tflops(self):
        v = C.c_double()
        L.check(self.lib.dmk_fp64_peak(C.byref(v)))
        return v.value

    # ---- multi-GPU: collective callbacks (solver_cpp_b200.parallel.TorchComm)
    def enable_fused_exchange(self, group=None):
        """all-gathers the CUDA IPC handles of every rank's spectral buffers and switches the distributed FFT's transposes to
        stores over peer memory inside the y / z passes (dmk_p2p_import).  Returns True if the fused path is on."""
        import torch.distributed as dist
        ok, on = 1, C.c_int(0)
        mine = (C.c_ubyte * 128)()
        try:
            L.check(self.lib.dmk_p2p_export(self.ctx, mine))
        except L.DmkError:
            ok = 0
        t = torch.tensor(list(bytes(mine)), dtype=torch.uint8, device=self.device)
        out = torch.empty(128 * self.world_size, dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(out, t, group=group)
        if ok:
            allh = (C.c_ubyte * (128 * self.world_size))(*out.cpu().tolist())
            try:
                L.check(self.lib.dmk_p2p_import(self.ctx, allh, C.byref(on)))
            except L.DmkError as e:                    # e.g. CUDA IPC not permitted in this environment
                sys.stderr.write('solver_cpp_b200: fused exchange unavailable (%s); using the all-to-all path\n' % e)
                ok = 0
        # the path is a collective property: every rank on the fused exchange, or none
        flag = torch.tensor([1.0 if (ok and on.value) else 0.0], device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if flag.item() == 0.0:
            L.check(self.lib.dmk_p2p_disable(self.ctx))
            on = C.c_int(0)
        self.fused_exchange = bool(on.value)
        return self.fused_exchange

    def set_comm(self, comm):
        """parallel.TorchComm (callback table) or parallel.NcclComm (the library's native NCCL binding)"""
        self._comm_keepalive = comm
        if hasattr(comm, 'attach'):
            comm.attach(self)
        else:
            L.check(self.lib.dmk_set_comm(self.ctx, C.byref(comm.struct)))
