"""GPU parity tests of the isostrain homogenisation (N_constituents > 1: homogenization_mechanical_isostrain.f90:36-45,
mechanical_homogenize homogenization_mechanical.f90:125-145 of the reference) through the C ABI against the oracle.
PARITY UNPINNED for the homogenisation step itself (the reference holds no fixture with several constituents per cell); the
oracle's isostrain path is checked against single-constituent oracles -- which G1/G2/G4 pin -- in tests/test_isostrain_cpu.py."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from oracle import spectral as osp                      # noqa: E402  (checker only)
from oracle.constitutive import ConstitutiveOracle      # noqa: E402

RTOL = 1.0e-9


def _rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


def _pointwise_rel(a, b):
    a = a.reshape(a.shape[0], -1); b = b.reshape(b.shape[0], -1)
    return np.abs(a - b).max(1) / np.maximum(np.abs(b).max(1), 1e-300)


def _F_path(cells, steps, rng, rate=2.0e-3, dt=0.5, noise=1.0e-4):
    nx, ny, nz = cells
    fluct = rng.standard_normal((nz, ny, nx, 3, 3))
    out = []
    for i in range(1, steps + 1):
        Fbar = np.eye(3)
        Fbar[0, 0] += rate * dt * i; Fbar[1, 1] -= 0.3 * rate * dt * i; Fbar[2, 2] -= 0.3 * rate * dt * i
        out.append(Fbar + noise * i * fluct)
    return out


@pytest.mark.parametrize('phases,nc,cells', [(['Al_pheno'], 2, (8, 6, 5)), (['Al_pheno', 'Al_elastic'], 3, (6, 5, 4)),
                                               (['Al_pheno', 'TWIP'], 2, (6, 5, 4))])
def test_isostrain_response_and_tangent_match_oracle(phases, nc, cells):
    """cell stress = sum_k v_k P_k within 1e-9 of the oracle's, the state of every constituent (point k * n + cell) within
    1e-8, the ill flag, the homogenised tangent per cell within 1e-8 and its reductions against numpy on the device's own
    field.  Cells with a constituent on a rounding-chaotic cut-back path are exempt as in test_gpu_constitutive.py."""
    from solver_cpp_b200 import workloads as W, material as M
    from solver_cpp_b200.grid import GridContext, fortran_to_c3333
    material, material_at, size = W.taylor_polycrystal(cells, 9, phases, nc, seed=3)
    g = GridContext(cells, size)
    g.set_option('pheno_mass_deferral', 0)
    g.material_init(M.Material(material, material_at))
    assert g.n_constituents == nc and int(g.lib.dmk_n_constituents(g.ctx)) == nc
    o = ConstitutiveOracle(material, material_at)
    n = int(np.prod(cells))
    rng = np.random.default_rng(7)
    dt = 0.5
    clean = np.ones(n, bool)
    compared = 0
    for inc, F in enumerate(_F_path(cells, 5, rng)):
        ill_o = o.mechanical_response(F.reshape(n, 9), dt)
        P_o, dPdF_o = o.mechanical_response2(dt, tangent=True)
        P = g.to_natural(g.constitutive_response(g.to_device(F), dt)).reshape(n, 3, 3)
        assert g.terminally_ill() == ill_o
        cut_pt = (o.point_cutbacks != 0) | (g.get_state('cutbacks', 1)[:, 0] != 0)            # per point
        flip_pt = g.get_state('lp_iter', 1)[:, 0] != o.point_lp_iter
        quiet = ~(cut_pt | flip_pt).reshape(nc, n).any(0)                                       # per cell
        err = _pointwise_rel(P, P_o)
        aux_pt = _pointwise_rel(g.get_state('Fp', 9), o.a['Fp'])
        aux = aux_pt.reshape(nc, n).max(0)
        assert g.get_state('F', 9).shape == (nc * n, 9)
        assert np.array_equal(g.get_state('F', 9), o.a['F'])                                    # every constituent saw the cell's F
        bad = ((err >= RTOL) | (aux >= 10 * RTOL)) & clean
        assert not (bad & quiet).any(), (inc, err[bad & quiet].max(), aux[bad & quiet].max())
        assert err[clean].max() < 0.2
        clean &= ~bad
        if inc in (1, 4):
            Cv, Cm, sel, d = g.tangent_response(dt, materialise=True)
            assert d.shape == (n, 81)
            dP = np.stack([fortran_to_c3333(r) for r in d.cpu().numpy()])
            et = _pointwise_rel(dP, dPdF_o)
            assert (et[clean & quiet] < 1e-8).all(), (inc, et[clean & quiet].max())
            _, C_vol, C_mm, sel_np = osp.homogenise(P_o, dP, osp.Grid(cells, g.size))          # numpy reduction of the DEVICE tangents
            assert _rel(Cv, C_vol) < 1e-12 and tuple(sel) == tuple(sel_np) and _rel(Cm, C_mm) < 1e-12
        compared += int((clean & quiet).sum())
        g.materialpoint_forward()
        o.forward()
    assert clean.mean() > 0.9 and compared > 0.7 * 5 * n, (clean.mean(), compared)
    assert np.abs(g.get_state('Lp', 9)).max() > 1e-6                                            # the path went plastic
    g.close()


def test_isostrain_increment_loop_matches_free_running_oracle():
    """a Taylor aggregate of 2 crystallites per cell through the whole solver: identical NGMRES iteration counts and flags,
    homogenised P and F per increment within 1e-9"""
    from solver_cpp_b200 import mech, workloads as W
    from oracle.solver import MechSolverBasicOracle, run_loadcase, parse_load as oparse
    cells = (8, 8, 8)
    material, material_at, size = W.taylor_polycrystal(cells, 12, ['Al_pheno'], 2, seed=1)
    o = MechSolverBasicOracle(cells, size, material, material_at, tangent_every_eval=False)
    o.log_fields = False
    o.init()
    steps_o = oparse(W.TENSION_X)
    steps_o[0]['N'] = 10
    s = mech.make_solver(cells, size, material, material_at)
    s.init()
    assert tuple(int(v) for v in s.get('minmax_cells', 2)) == tuple(o.minmax_cells)
    assert _rel(s.get3333('C_volAvg'), o.C_volAvg) < 1e-12
    n_inc = 3
    recs = mech.loop_over_loadcase(s, mech.parse_load(W.TENSION_X), max_increments=n_inc)
    orecs = []

    class Stop(Exception):
        pass

    def cb(rec, solver):
        orecs.append(rec)
        if len(orecs) == n_inc:
            raise Stop()
    try:
        run_loadcase(o, steps_o, callback=cb)
    except Stop:
        pass
    assert [r['iterations'] for r in recs] == [r['iterations'] for r in orecs]
    assert [r['converged'] for r in recs] == [r['converged'] for r in orecs]
    for a, b in zip(recs, orecs):
        assert _rel(a['P_av'], b['P_av']) < 1e-9 and _rel(a['F_av'], b['F_av']) < 1e-9
    assert _rel(s.g.to_natural(s.field('F')), o.F) < 1e-9


def test_isostrain_through_the_fortran_named_symbols():
    """f_homogenization_mechanical_response / _response2 on the host mirrors with several constituents per cell: the caller's
    F0 reaches every constituent, P and dPdF come back homogenised per cell -- identical to the device-pointer API"""
    from solver_cpp_b200 import workloads as W, material as M, lib as L
    from solver_cpp_b200.grid import GridContext, natural_to_records
    cells = (6, 5, 4)
    n = int(np.prod(cells))
    material, material_at, size = W.taylor_polycrystal(cells, 5, ['Al_pheno'], 2, seed=0)
    ga, gb = GridContext(cells, size), GridContext(cells, size)
    ga.material_init(M.Material(material, material_at)); gb.material_init(M.Material(material, material_at))
    lib = L.load()
    L.check(lib.dmk_bind_global(ga.ctx))
    try:
        pF0, pF, pP, pT, pill = (C.POINTER(C.c_double)(), C.POINTER(C.c_double)(), C.POINTER(C.c_double)(), C.POINTER(C.c_double)(), C.c_void_p())
        lib.f_homogenization_fetch_tensor_pointers(C.byref(pF0), C.byref(pF), C.byref(pP), C.byref(pT), C.byref(pill))
        F = np.ctypeslib.as_array(pF, shape=(n * 9,)); Pm = np.ctypeslib.as_array(pP, shape=(n * 9,)); T = np.ctypeslib.as_array(pT, shape=(n * 81,))
        rng = np.random.default_rng(0)
        Fnat = np.eye(3) + np.diag([1.5e-3, -5e-4, -5e-4]) + 2e-4 * rng.standard_normal((cells[2], cells[1], cells[0], 3, 3))
        F[:] = natural_to_records(Fnat).reshape(-1)
        dt, one, nn = C.c_double(0.5), C.c_int(1), C.c_int(n)
        lib.f_homogenization_mechanical_response(C.byref(dt), C.byref(one), C.byref(nn))
        ip = (C.c_int * 2)(1, 1); el = (C.c_int * 2)(1, n)
        lib.f_homogenization_mechanical_response2(C.byref(dt), ip, el)
        P_b = gb.constitutive_response(gb.to_device(Fnat), 0.5).cpu().numpy().reshape(-1)
        _, _, _, d = gb.tangent_response(0.5, materialise=True)
        assert np.array_equal(Pm, P_b)
        assert np.array_equal(T, d.cpu().numpy().reshape(-1))
        assert C.cast(pill, C.POINTER(C.c_int))[0] == 0
        assert ga.get_state('F0', 9).shape == (2 * n, 9) and np.array_equal(ga.get_state('F0', 9), gb.get_state('F0', 9))
        lib.f_materialpoint_forward()
        gb.materialpoint_forward()
        assert np.array_equal(ga.get_state('Fp0', 9), gb.get_state('Fp0', 9))
    finally:
        L.check(lib.dmk_bind_global(None))


def test_isostrain_result_file_from_the_solver(tmp_path):
    """DADF5 output of a Taylor-aggregate job: every constituent's datasets through /cell_to/phase [N, C]; each constituent
    carries the cell's F, and the volume-fraction weighted mean of the written P is the solver's homogenised stress"""
    from solver_cpp_b200 import mech, workloads as W, material as M
    from solver_cpp_b200.dadf5 import ResultWriter, ResultReader
    cells = (6, 5, 4)
    n, nc = int(np.prod(cells)), 2
    material, material_at, size = W.taylor_polycrystal(cells, 7, ['Al_pheno', 'Cu_pheno'], nc, seed=5)
    s = mech.make_solver(cells, size, material, material_at)
    s.init()
    w = ResultWriter(cells, size, material, material_at)
    recs = mech.loop_over_loadcase(s, mech.parse_load(W.TENSION_X), max_increments=2, callback=w.callback())
    path = str(tmp_path / 'taylor.hdf5')
    w.save(path)
    r = ResultReader(path)
    assert r.n_const == nc and r.increments == [1, 2]
    F = r.place('F', 2)
    assert F.shape == (n, nc, 3, 3)
    F_cell = s.g.to_natural(s.field('F')).reshape(n, 3, 3)
    assert np.array_equal(F[:, 0], F[:, 1])                                       # isostrain
    assert np.abs(F[:, 0] - F_cell).max() < 1e-12
    Fp = r.place('F_p', 2)
    assert np.array_equal(np.swapaxes(Fp, 0, 1).reshape(nc * n, 3, 3), s.g.get_state('Fp', 9).reshape(nc * n, 3, 3))
    assert np.abs(Fp[:, 0] - Fp[:, 1]).max() > 1e-3                                # two different crystallites per cell
    v = M.Material(material, material_at).v_of_point.reshape(nc, n).T            # [cell, constituent]
    P_av = (r.place('P', 2) * v[:, :, None, None]).sum(1).mean(0)
    assert _rel(P_av, recs[-1]['P_av']) < 1e-6
    assert r.place('xi_sl', 2).shape == (n, nc, 12)
