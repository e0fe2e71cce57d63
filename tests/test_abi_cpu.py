"""CPU tests of the C-ABI boundary: the shared library loads without a GPU, exports every symbol include/dmk_grid.h
declares, refuses compute without a device (no CPU fallback), and its pure host helpers (the f_math_* / f_rotate_* of
/root/reference/src/grid/utilities_fortran.h:8-15, crystal tables) agree with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from solver_cpp_b200 import lib as L
    return L.load()


def test_every_declared_symbol_is_exported():
    hdr = open(os.path.join(ROOT, 'include', 'dmk_grid.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    names = set(re.findall(r'\b((?:dmk|f)_[a-zA-Z0-9_]+)\s*\(', hdr))
    names -= {'dmk_ctx'}
    assert len(names) > 60
    lib = _lib()
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from solver_cpp_b200 import lib as L
    from solver_cpp_b200.grid import GridContext
    with pytest.raises(L.DmkError):
        GridContext((4, 4, 4), (1e-5,) * 3)
    g = L.GridDesc()
    g.cells[:] = [4, 4, 4]; g.size[:] = [1e-5] * 3; g.world_size = 1
    ctx = C.c_void_p()
    rc = _lib().dmk_create(C.byref(ctx), C.byref(g))
    assert rc != 0 and b'no CUDA device' in _lib().dmk_last_error()


def test_math_helpers_known_answers():
    lib = _lib()
    rng = np.random.default_rng(0)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    # f_math_invert (math.f90:578): Fortran order n x n
    A = rng.standard_normal((5, 5)) + 5 * np.eye(5)
    Af = np.asfortranarray(A); inv = np.zeros((5, 5), order='F'); err = C.c_int(7); n = C.c_int(5)
    lib.f_math_invert(P(inv), C.byref(err), P(Af), C.byref(n))
    assert err.value == 0 and np.abs(inv - np.linalg.inv(A)).max() < 1e-13
    Z = np.zeros((3, 3), order='F'); inv3 = np.ones((3, 3), order='F'); n = C.c_int(3)
    lib.f_math_invert(P(inv3), C.byref(err), P(Z), C.byref(n))
    assert err.value != 0
    # 3333 <-> 99 with row-major pairs (MAPPLAIN), mul3333xx33
    T = rng.standard_normal((3, 3, 3, 3)); Tf = np.asfortranarray(T)
    m99 = np.zeros((9, 9), order='F')
    lib.f_math_3333to99(P(Tf), P(m99))
    assert np.array_equal(m99, T.reshape(9, 9))
    back = np.zeros((3, 3, 3, 3), order='F')
    lib.f_math_99to3333(P(m99), P(back))
    assert np.array_equal(back, T)
    B = rng.standard_normal((3, 3)); Bf = np.asfortranarray(B); res = np.zeros((3, 3), order='F')
    lib.f_math_mul3333xx33(P(Tf), P(Bf), P(res))
    assert np.abs(res - np.einsum('ijkl,kl->ij', T, B)).max() < 1e-14
    # rotations against the oracle (rotations.f90:288-359)
    from oracle import crystal as oc, spectral as osp
    q = rng.standard_normal(4); q /= np.linalg.norm(q)
    R = oc.qu2om(q)
    for act in (0, 1):
        out = np.zeros((3, 3), order='F'); a = C.c_int(act)
        lib.f_rotate_tensor2(P(np.ascontiguousarray(q)), P(Bf), P(out), C.byref(a))
        assert np.abs(out - osp.rotate2(R, B, active=bool(act))).max() < 1e-14
        out4 = np.zeros((3, 3, 3, 3), order='F')
        lib.f_rotate_tensor4(P(np.ascontiguousarray(q)), P(Tf), P(out4), C.byref(a))
        assert np.abs(out4 - osp.rotate4(R, T, active=bool(act))).max() < 1e-13


def test_crystal_tables_match_oracle():
    from solver_cpp_b200 import material as M
    from oracle import crystal as oc
    for N in ([12], [12, 6], [5, 3]):
        assert np.abs(M.schmid_slip('cF', N).reshape(-1, 3, 3) - oc.schmid_slip(N, 'cF')).max() < 1e-15
        v = np.arange(1, 14.0)
        assert np.array_equal(M.interaction('sl-sl', 'cF', N, N, v), oc.interaction_slip_slip(N, v, 'cF'))
        assert np.array_equal(M.interaction('sl-tw', 'cF', N, [12], v[:4]), oc.interaction_slip_twin(N, [12], v[:4], 'cF'))
        f = np.repeat([0.3, 0.8][:len(N)], N)
        assert np.abs(M.forest_projection('cF', N, f) - oc.forest_projection(N, 'cF', f)).max() < 1e-15
    assert np.abs(M.schmid_twin('cF', [12]).reshape(-1, 3, 3) - oc.schmid_twin([12], 'cF')).max() < 1e-15
    assert np.array_equal(M.twin_nucleation_pairs([12]), oc.CF_TWIN_NUCLEATION_SLIP_PAIR - 1)
    C66 = oc.symmetrize_C66(175e9, 115e9, 135e9)
    assert np.abs(M.C66_twin('cF', [12], C66).reshape(-1, 6, 6) - oc.C66_twin([12], C66, 'cF')).max() < 1e-4
    for N in ([12], [12, 12, 24]):
        for sense in (1, -1):
            c = [0.1, 0.2, 0.3, 0.05, 0.02, 0.01]
            assert np.abs(M.schmid_slip('cI', N, c, sense).reshape(-1, 3, 3) - oc.schmid_slip(N, 'cI', c, sense)).max() < 1e-15


def test_material_tables_match_oracle_phase_setup():
    """material.yaml -> dmk_phase (solver_cpp_b200.material) against the oracle's independent set-up"""
    from solver_cpp_b200 import material as M, workloads as W
    from oracle import constitutive as oc
    for name in ('Al_pheno', 'Cu_pheno', 'Al_iso', 'TWIP', 'Al_elastic'):
        a = M.make_phase(W.PHASES[name]); b = oc.make_phase(name, W.PHASES[name])
        assert (a.law, a.n_sl, a.n_tw, a.n_state) == (b.law, b.n_sl, b.n_tw, b.n_state)
        for f in ('C66', 'P_sl', 'P_tw', 'h_sl_sl', 'h_sl_tw', 'h_tw_tw', 'dot_gamma_0_sl', 'n_sl_exp', 'a_sl', 'xi_inf_sl', 'dt_forest',
                  'dt_C66_tw', 'dt_b_sl', 'dt_tau_0', 'dt_pair', 'atol', 'state_init'):
            x, y = np.ctypeslib.as_array(getattr(a, f)), np.ctypeslib.as_array(getattr(b, f))
            sl = tuple(slice(0, min(s, t)) for s, t in zip(x.shape, y.shape))
            assert np.allclose(x[sl], y[sl], rtol=1e-14, atol=1e-15 * max(np.abs(y[sl]).max(), 1e-300)), (name, f)
        for f in ('iso_xi_inf', 'iso_n', 'iso_a', 'dt_mu', 'dt_nu', 'dt_kB_T', 'dt_Gamma_sf', 'dt_omega', 'dt_D', 'dt_L_tw', 'dt_x_c'):
            assert np.isclose(getattr(a, f), getattr(b, f), rtol=1e-14, atol=0), (name, f)


def test_native_driver_reaches_the_device_and_refuses_without_one(golden_dir):
    """tests/abi/native_driver.cpp (the reference's main.cpp reduced to calls into the library, all C++): on a box without a GPU
    it must read the geometry through the f_VTI_* symbols, hand the cells to f_discretization_init, parse material.yaml -- and
    then fail loudly at the context: there is no CPU path to fall back to."""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    import __graft_entry__ as ge
    exe = ge.build_native_driver()
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and 'usage' in r.stderr
    r = subprocess.run([exe, os.path.join(golden_dir, 'g6_unit_test_geometry.vti'), os.path.join(golden_dir, 'material_20grains_Al.yaml'),
                        os.path.join(golden_dir, 'load_tensionX.yaml')], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and 'no CUDA device' in r.stderr and 'NATIVE_DRIVER_OK' not in r.stdout
    r = subprocess.run([exe, os.path.join(golden_dir, 'material_20grains_Al.yaml'), 'x', 'y'], capture_output=True, text=True)
    assert r.returncode == 1 and 'f_VTI_readCellsSizeOrigin' in r.stderr         # not a geometry file


def test_public_header_is_plain_c(tmp_path):
    """the boundary is a C ABI: include/dmk_grid.h must compile as C99 (what a cgo / Fortran-bind(C) / ctypes binding generator
    would read), with the return codes and the load-case types it declares"""
    import subprocess
    src = tmp_path / 'c_abi.c'
    src.write_text('#include "dmk_grid.h"\n'
                   'int main(void) { dmk_load_step s; dmk_loop_ops o; dmk_grid_numerics g; dmk_numerics n; dmk_phase p;\n'
                   '  (void)s; (void)o; (void)g; (void)n; (void)p; return DMK_OK + 0 * (DMK_ERR_CUDA + DMK_ERR_ARG + DMK_ERR_UNSUPPORTED + DMK_ERR_STATE); }\n')
    r = subprocess.run(['gcc', '-std=c99', '-Wall', '-Wextra', '-pedantic', '-Werror', '-I', os.path.join(ROOT, 'include'), '-fsyntax-only', str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def _phase_tables_agree(phase, tol=5e-16):
    """product tables (solver_cpp_b200.material.make_phase: crystallography from the library's C++) against the oracle's
    (oracle.constitutive.make_phase: crystallography in Python), field by field over the leading block both layouts hold"""
    from solver_cpp_b200 import lib as L, material as M
    from oracle import constitutive as OC
    a, b = M.make_phase(phase, M.T_ROOM), OC.make_phase('X', phase, OC.T_ROOM)
    oracle_fields = {f for f, _ in OC.Phase._fields_}

    def arr(v):
        return np.ctypeslib.as_array(v).astype(float) if hasattr(v, '_length_') else np.asarray(v, float)
    compared = 0
    for f, _ in L.Phase._fields_:
        assert f in oracle_fields, f
        x, y = arr(getattr(a, f)), arr(getattr(b, f))
        sl = tuple(slice(0, min(p, q)) for p, q in zip(x.shape, y.shape))
        x, y = x[sl], y[sl]
        scale = np.abs(y).max() if y.size else 0.0
        assert np.abs(x - y).max() <= tol * max(scale, 1e-300) if y.size else True, (f, np.abs(x - y).max(), scale)
        compared += 1
    return compared


def test_phase_tables_product_vs_oracle():
    """two independent set-ups of the same phase parameters (Schmid and non-Schmid projections, interaction matrices, forest
    projections, twin systems, stiffness, law parameters): every workload phase of this repo, and -- in the build container --
    every parameter set of the reference's examples that the device path offers"""
    import glob
    import yaml
    from solver_cpp_b200 import workloads as W
    for name, phase in W.PHASES.items():
        assert _phase_tables_agree(phase) > 80, name
    base = '/root/reference/examples/config/phase/mechanical/'
    offered = {'dislotungsten_W': ('W', 'cI'), 'isotropic_free-surface': ('Al', 'cF'), 'kinehardening_317L': ('AISI304', 'cF'),
               'phenopowerlaw_AA6022-T4': ('Al', 'cF'), 'phenopowerlaw_Al': ('Al', 'cF'), 'phenopowerlaw_Au': ('Au', 'cF'),
               'phenopowerlaw_Cu': ('Cu', 'cF'), 'phenopowerlaw_Pt-5%Cu': ('Pt', 'cF')}
    if not glob.glob(base + 'plastic/*.yaml'):
        return
    for name, (el, lat) in offered.items():
        phase = {'lattice': lat, 'mechanical': {'elastic': yaml.safe_load(open(base + 'elastic/Hooke_%s.yaml' % el)),
                                                'plastic': yaml.safe_load(open(base + 'plastic/%s.yaml' % name))}}
        assert _phase_tables_agree(phase) > 80, name
