"""Host-side mirror of the reference's `MechSolverBasic` and `DamaskGrid` call surface over the device-resident solver.

Names follow /root/reference/src/grid/mech_solver_basic.h:10-29 (init, forward, calculate_solution, form_residual,
converged) and /root/reference/src/grid/damask_grid.cpp:36-196 (loop_over_loadcase: increments, cut-back, forwarding);
load-case files are parsed like /root/reference/src/grid/config.cpp:247-358.  Every numerical step runs in
libdamask_grid_b200.so; this file only sequences calls."""
import ctypes as C

import numpy as np

from . import lib as L
from . import material as M
from .grid import GridContext, _hp, _p, fortran_to_c3333


class BoundaryCondition:
    def __init__(self, values=None, mask=None, kind=''):
        self.values = np.zeros((3, 3)) if values is None else np.asarray(values, float)
        self.mask = np.zeros((3, 3), bool) if mask is None else np.asarray(mask, bool)      # True = 'x'
        self.type = kind


def _parse_bc(entry):
    vals, mask = np.zeros((3, 3)), np.zeros((3, 3), bool)
    for i in range(3):
        for j in range(3):
            v = entry[i][j]
            if isinstance(v, str) and v.strip().lower() == 'x':
                mask[i, j] = True
            else:
                vals[i, j] = float(v)
    return vals, mask


def parse_load(load):
    """load-case dict (YAML) -> list of load steps (Config::LoadStep, config.h:74-84)"""
    steps = []
    for st in load['loadstep']:
        mech = st['boundary_conditions']['mechanical']
        step = dict(R=np.eye(3))
        for k in ('F', 'dot_F', 'L'):
            if k in mech:
                step['deformation'] = BoundaryCondition(*_parse_bc(mech[k]), k)
        for k in ('P', 'dot_P'):
            if k in mech:
                step['stress'] = BoundaryCondition(*_parse_bc(mech[k]), k)
        if 'deformation' not in step:
            raise L.DmkError('load step needs one of F / dot_F / L')
        if 'stress' not in step:
            # upstream DAMASK_grid.f90: no stress entry = no component is stress controlled (mask all 'x'); the deformation must then
            # prescribe all nine (next check).  Config::parse_load (config.cpp:309-311) throws instead -- on the reference's own
            # example decks examples/grid/shearXY.yaml and shearZX.yaml
            step['stress'] = BoundaryCondition(np.zeros((3, 3)), np.ones((3, 3), bool), 'P')
        if np.any(step['deformation'].mask == step['stress'].mask):
            raise L.DmkError('each component must be controlled by exactly one of deformation and stress')
        if 'R' in mech:
            step['R'] = M.qu2om(mech['R'])
        d = st['discretization']
        step.update(t=float(d['t']), N=int(d['N']), r=float(d.get('r', 1)), estimate_rate=bool(st.get('estimate_rate', True)),
                    f_out=int(st.get('f_out', 1)), f_restart=int(st.get('f_restart', 0)))             # config.h:93-95 defaults
        steps.append(step)
    return steps


def _i9(mask):
    a = np.ascontiguousarray(np.asarray(mask).reshape(9), dtype=np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int))


class MechSolverBasic:
    """grid_mechanical_spectral_basic with fields, NGMRES history and per-point state resident on the GPU"""

    def __init__(self, grid: GridContext):
        self.g = grid
        self.lib = grid.lib
        self.ctx = grid.ctx

    # MechSolverBasic::init (mech_solver_basic.cpp:19-108)
    def init(self, gamma_pair_cells=None):
        if gamma_pair_cells is None:
            L.check(self.lib.dmk_mech_init(self.ctx, None))
        else:
            a = (C.c_longlong * 2)(int(gamma_pair_cells[0]), int(gamma_pair_cells[1]))
            L.check(self.lib.dmk_mech_init(self.ctx, a))

    # MechSolverBasic::forward (mech_solver_basic.cpp:163-220)
    def forward(self, cut_back, guess, dt, dt_old, t_remaining, deformation_bc, stress_bc, R):
        dv = np.ascontiguousarray(deformation_bc.values.reshape(9))
        sv = np.ascontiguousarray(stress_bc.values.reshape(9))
        dm, dmp = _i9(deformation_bc.mask)
        sm, smp = _i9(stress_bc.mask)
        Rm = np.ascontiguousarray(np.asarray(R, float).reshape(9))
        L.check(self.lib.dmk_mech_forward(self.ctx, C.c_int(int(cut_back)), C.c_int(int(guess)), C.c_double(dt), C.c_double(dt_old),
                                          C.c_double(t_remaining), deformation_bc.type.encode(), _hp(dv), dmp,
                                          stress_bc.type.encode(), _hp(sv), smp, _hp(Rm)))

    # MechSolverBasic::calculate_solution (mech_solver_basic.cpp:110-161)
    def calculate_solution(self):
        s = L.SolutionState()
        L.check(self.lib.dmk_mech_calculate_solution(self.ctx, C.byref(s)))
        return dict(converged=bool(s.converged), required_iterations=s.required_iterations,
                    terminally_ill=bool(s.terminally_ill), stag_converged=bool(s.stag_converged))

    # MechSolverBasic::form_residual (mech_solver_basic.cpp:286-339): one evaluation on device fields
    def form_residual(self, F, r=None):
        r = self.g.new_field() if r is None else r
        L.check(self.lib.dmk_mech_form_residual(self.ctx, _p(F), _p(r)))
        return r

    def form_residual_host(self, F_host, r_host):
        """same through HOST buffers in the reference's memory order (H2D + evaluation + D2H)"""
        L.check(self.lib.dmk_mech_form_residual_host(self.ctx, _hp(F_host), _hp(r_host)))
        return r_host

    def update_coords(self):
        """MechSolverBasic::update_coords (mech_solver_basic.cpp:222-240): node and cell-centre coordinates of the current F"""
        return self.g.update_coords(self.field('F'))

    def prepare_evaluation(self):
        """the pre-SNES part of calculate_solution (tangent reductions, masked compliance)"""
        L.check(self.lib.dmk_mech_prepare(self.ctx))

    def set_history(self, F=None, F_lastInc=None, F_aim=None, F_aim_lastInc=None, P_aim=None):
        """restart-style injection of solver state (device tensors for fields, natural 3x3 for aims)"""
        for name, t in (('F', F), ('F_lastInc', F_lastInc)):
            if t is not None:
                L.check(self.lib.dmk_mech_set_field_device(self.ctx, name.encode(), _p(t)))
        for name, a in (('F_aim', F_aim), ('F_aim_lastInc', F_aim_lastInc), ('P_aim', P_aim)):
            if a is not None:
                v = np.ascontiguousarray(np.asarray(a, float).T.reshape(9))
                L.check(self.lib.dmk_mech_set(self.ctx, name.encode(), _hp(v)))

    # ---- state access
    def get(self, what, n=1):
        out = np.zeros(n)
        L.check(self.lib.dmk_mech_get(self.ctx, what.encode(), _hp(out), C.c_size_t(n)))
        return out

    def get33(self, what):
        return self.get(what, 9).reshape(3, 3).T            # Fortran order -> natural [i, j]

    def get3333(self, what):
        return fortran_to_c3333(self.get(what, 81))

    def field(self, what):
        """device copy (torch tensor) of a solver field: "F", "F_lastInc", "F_dot", "r" """
        out = self.g.new_field()
        L.check(self.lib.dmk_mech_copy_field(self.ctx, what.encode(), _p(out)))
        return out

    def field_host(self, what):
        out = np.zeros(self.g.field_shape)
        L.check(self.lib.dmk_mech_get_field(self.ctx, what.encode(), _hp(out)))
        return out

    def eval_log(self):
        n = C.c_size_t()
        L.check(self.lib.dmk_mech_eval_log(self.ctx, None, C.c_size_t(0), C.byref(n)))
        out = np.zeros((n.value, 12))
        if n.value:
            L.check(self.lib.dmk_mech_eval_log(self.ctx, _hp(out), C.c_size_t(n.value), C.byref(n)))
        return out

    def eval_stats(self):
        """per residual evaluation: cells of this slab whose full first sub-step failed (the reference cuts those back)"""
        n = C.c_size_t()
        L.check(self.lib.dmk_mech_eval_stats(self.ctx, None, C.c_size_t(0), C.byref(n)))
        out = np.zeros((n.value, 2))
        if n.value:
            L.check(self.lib.dmk_mech_eval_stats(self.ctx, _hp(out), C.c_size_t(n.value), C.byref(n)))
        return out

    def capture(self, enable=True):
        L.check(self.lib.dmk_mech_capture(self.ctx, C.c_int(int(enable))))

    def captured(self):
        """(F, r) of every residual evaluation since capture(True): arrays [n_eval, 9 * n_local], reference memory order"""
        n = C.c_size_t()
        L.check(self.lib.dmk_mech_get_capture(self.ctx, None, None, C.c_size_t(0), C.byref(n)))
        F = np.zeros((n.value, 9 * self.g.n_local)); r = np.zeros_like(F)
        if n.value:
            L.check(self.lib.dmk_mech_get_capture(self.ctx, _hp(F), _hp(r), C.c_size_t(n.value), C.byref(n)))
        return F, r

    def reports(self):
        n = C.c_size_t()
        L.check(self.lib.dmk_mech_reports(self.ctx, None, C.c_size_t(0), C.byref(n)))
        out = np.zeros((n.value, 6))
        if n.value:
            L.check(self.lib.dmk_mech_reports(self.ctx, _hp(out), C.c_size_t(n.value), C.byref(n)))
        return out


class MechSolverPolarisation(MechSolverBasic):
    """grid_mechanical_spectral_polarisation of upstream DAMASK on the device (the reference tree holds only its numerics keys,
    config.cpp:450-457): solution vector (F | F_tau), same call surface as MechSolverBasic"""

    def __init__(self, grid, alpha=1.0, beta=1.0, eps_abs_curl=1.0e-10, eps_rel_curl=5.0e-4):
        super().__init__(grid)
        self.pol = (alpha, beta, eps_abs_curl, eps_rel_curl)

    def init(self, gamma_pair_cells=None):
        if gamma_pair_cells is not None:
            raise L.DmkError('the polarisation solver has no injected reference-stiffness pair')
        L.check(self.lib.dmk_mech_init_polarisation(self.ctx, *[C.c_double(v) for v in self.pol]))

    def new_vector(self):
        import torch
        return torch.empty((2,) + tuple(self.g.field_shape), dtype=torch.float64, device=self.g.device)

    def form_residual(self, X, r=None):
        """X = (F | F_tau) as one device tensor [2, nz_local, ny, nx, 9] -> (r_F | r_F_tau)"""
        r = self.new_vector() if r is None else r
        L.check(self.lib.dmk_mech_form_residual(self.ctx, _p(X), _p(r)))
        return r

    def form_residual_host(self, F_host, r_host):
        raise L.DmkError('form_residual_host: basic scheme only')


def loop_over_loadcase(solver: MechSolverBasic, steps, max_cut_back=3, callback=None, max_increments=None):
    """DamaskGrid::loop_over_loadcase (damask_grid.cpp:82-196), upstream cut-back bookkeeping.
    Returns one record per converged increment."""
    g = solver.g
    records = []
    cut_back = False
    sub_step_factor, cut_back_level, total_inc = 2, 0, 0
    t, dt, dt_prev = 0.0, 1.0, 0.0
    for l, st in enumerate(steps):
        t0 = t
        guess = st['estimate_rate'] and l > 0
        for inc in range(st['N']):
            if max_increments is not None and total_inc >= max_increments:
                return records
            total_inc += 1
            dt_prev = dt
            if abs(st['r'] - 1.0) < 1e-9:
                dt = st['t'] / st['N']
            else:
                dt = st['t'] * (st['r'] ** inc - st['r'] ** (inc + 1)) / (1.0 - st['r'] ** st['N'])
            dt = dt * float(sub_step_factor) ** (-cut_back_level)
            step_fraction = 0
            while step_fraction < sub_step_factor ** cut_back_level:
                t_remaining = st['t'] + t0 - t
                t += dt
                step_fraction += 1
                solver.forward(cut_back, guess, dt, dt_prev, t_remaining, st['deformation'], st['stress'], st['R'])
                if not cut_back:
                    g.materialpoint_forward()                                    # f_materialpoint_forward
                sol = solver.calculate_solution()
                found = sol['converged'] and sol['stag_converged'] and not sol['terminally_ill']
                if found:
                    dt_prev, cut_back, guess = dt, False, True
                elif cut_back_level < max_cut_back:
                    cut_back = True
                    step_fraction = (step_fraction - 1) * sub_step_factor
                    cut_back_level += 1
                    t -= dt
                    dt = dt / sub_step_factor
                else:
                    raise L.DmkError('no solution found for increment %d' % total_inc)
            cut_back_level = max(0, cut_back_level - 1)
            rec = dict(inc=total_inc, step=l, inc_in_step=inc + 1, f_out=st.get('f_out', 1), t=t, converged=sol['converged'],
                       iterations=sol['required_iterations'],
                       P_av=solver.get33('P_av'), F_av=solver.get('F_av', 9).reshape(3, 3).T, F_aim=solver.get33('F_aim'))
            records.append(rec)
            if callback:
                callback(rec, solver)
    return records


def make_solver(cells, size, material, material_at, numerics=None, derivative=0, divergence_correction=2, world_size=1, rank=0,
                comm=None, scheme='spectral_basic', **scheme_numerics):
    """GridContext + material + MechSolverBasic (or MechSolverPolarisation: scheme = 'spectral_polarisation', with alpha, beta,
    eps_abs_curl, eps_rel_curl) for one z-slab (material_at: the slab's cells, x fastest)"""
    g = GridContext(cells, size, derivative=derivative, divergence_correction=divergence_correction, world_size=world_size, rank=rank)
    if comm is not None:
        g.set_comm(comm)
        from . import parallel
        if world_size > 1 and parallel.fused_exchange_wanted():
            g.enable_fused_exchange(getattr(comm, 'group', None))
    g.material_init(M.Material(material, material_at), numerics)
    if scheme == 'spectral_polarisation':
        return MechSolverPolarisation(g, **scheme_numerics)
    if scheme != 'spectral_basic':
        raise L.DmkError('unknown mechanical solver %r (spectral_basic, spectral_polarisation)' % scheme)
    return MechSolverBasic(g)
