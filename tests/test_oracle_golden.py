"""CPU tests: the oracle (oracle/) against the reference's golden vectors G1/G2.
These pin the checker; they do not exercise the product path."""
import os

import numpy as np
import pytest
import yaml

from oracle import spectral as sp
from oracle.constitutive import ConstitutiveOracle
from oracle.solver import MechSolverBasicOracle, parse_load, run_loadcase


@pytest.fixture(scope='module')
def deck(golden_dir):
    g1 = np.load(os.path.join(golden_dir, 'g1_fortran_run.npz'))
    g2 = np.load(os.path.join(golden_dir, 'g2_constitutive_known_answer.npz'))
    mat = yaml.safe_load(open(os.path.join(golden_dir, 'material_20grains_Al.yaml')))
    load = yaml.safe_load(open(os.path.join(golden_dir, 'load_tensionX.yaml')))
    return g1, g2, mat, load


def test_g2_constitutive_known_answer(deck):
    """output_damask_cpp.sh:290-303: P of all 27 cells for F = diag(1.0005,1,1), dt = 0.5 s, virgin state."""
    g1, g2, mat, load = deck
    o = ConstitutiveOracle(mat, g2['grain_map'])
    F = sp.fortran_to_natural(g2['F'], (3, 3, 3))
    assert not o.mechanical_response(F, 0.5)
    P, dPdF = o.mechanical_response2(0.5)
    Pref = sp.fortran_to_natural(g2['P'], (3, 3, 3)).reshape(27, 3, 3)
    assert np.abs(P - Pref).max() / np.abs(Pref).max() < 1e-11
    P_av = P.mean(0)
    assert np.abs(P_av.T.reshape(-1) - g2['P_av']).max() / np.abs(g2['P_av']).max() < 1e-11
    # divergence RMS of that stress field: raw sum and scaled value (mech_base.cpp:391,394 prints)
    grid = sp.Grid((3, 3, 3), (1e-5,) * 3)
    xi1, _ = sp.xi_tables(grid)
    rms, raw = sp.divergence_rms(sp.fft_forward(Pref.reshape(3, 3, 3, 3, 3)), xi1, grid, sp.scaled_geom_size(grid))
    assert abs(raw - float(g2['rms_raw'])) / float(g2['rms_raw']) < 1e-13
    assert abs(rms - float(g2['rms_scaled'])) / float(g2['rms_scaled']) < 1e-13


def test_g1_free_run_reproduces_fortran_log(deck):
    """Free-running oracle (own S, NGMRES, constitutive) vs all 108 logged evaluations of output_damask_f.sh.
    Only the noise-selected reference-stiffness pair is injected (grains 11 and 15)."""
    g1, g2, mat, load = deck
    s = MechSolverBasicOracle((3, 3, 3), (1e-5,) * 3, mat, g2['grain_map'])
    s.init(gamma_pair_cells=(9, 21))
    recs = run_loadcase(s, parse_load(load))
    assert [r['iterations'] for r in recs] == list(g1['total_iter_at_convergence'])
    assert len(s.log) == 108
    for i, ev in enumerate(s.log):
        assert np.abs(sp.natural_to_fortran(ev['F']) - g1['F'][i]).max() < 1e-12, i
        # residual: FP64 noise level in units of F
        assert np.abs(sp.natural_to_fortran(ev['r']) - g1['r'][i]).max() < 1e-13, i
        assert np.abs(ev['P_av'] * 1e-6 - g1['P_av_MPa'][i]).max() < 5.1e-5, i
    assert len(s.reports) == len(g1['err_div'])
    for rep, ed, eb, dt_, bt in zip(s.reports, g1['err_div'], g1['err_BC'], g1['err_div_tol'], g1['err_BC_tol']):
        assert abs(rep['err_div'] - ed) <= 0.006 * ed + 1e-300
        assert abs(rep['err_BC'] - eb) <= 0.006 * eb + 1e-300
        assert abs(rep['divTol'] - dt_) <= 0.006 * dt_
        assert abs(rep['BCTol'] - bt) <= 0.006 * bt
    P11 = [round(r['P_av'][0, 0] * 1e-6, 4) for r in recs]
    assert P11 == [35.2362, 64.6648, 72.9485, 75.6207, 76.7790, 77.3081, 77.6004, 77.8005, 77.9353, 78.0432]


def test_tie_rule_is_order_independent():
    rng = np.random.default_rng(0)
    base = 1.0e22
    nrm = base * (1.0 + 1e-15 * rng.standard_normal(50))
    assert sp.select_minmax(nrm) == (0, 0)
    nrm[17] *= 1.001; nrm[30] *= 0.999
    assert sp.select_minmax(nrm) == (17, 30)


def test_cutback_path_is_rounding_sensitive(tmp_path):
    """Documents a property of the reference ALGORITHM, not of any implementation: for a soft virgin material the first
    trial sub-step overshoots (Lp ~ 1e27), and whether the following cut-back sub-step converges is decided by
    rounding.  The oracle and the same source compiled with FMA contraction disagree on O(1%) of the cells by O(1e-3);
    cells that never cut back agree to 1e-9.  GPU parity tests therefore exempt exactly those cells."""
    import ctypes as C
    import subprocess
    from solver_cpp_b200 import workloads as W
    import oracle.constitutive as oc
    src = os.path.join(os.path.dirname(oc.__file__), 'constitutive.c')
    lib = str(tmp_path / 'liboracle_fma.so')
    subprocess.check_call(['gcc', '-O2', '-fopenmp', '-ffp-contract=fast', '-mfma', '-std=c11', '-shared', '-fPIC', '-o', lib, src, '-lm'],
                          cwd=os.path.dirname(src))
    cells = (16, 16, 8)
    material, material_at, size = W.polycrystal(cells, 20, 'Cu_pheno', 0)
    n = int(np.prod(cells))
    rng = np.random.default_rng(11)
    F = np.eye(3) + np.diag([5e-4, -1.5e-4, -1.5e-4]) + 1e-4 * rng.standard_normal((n, 3, 3))
    a = oc.ConstitutiveOracle(material, material_at)
    b = oc.ConstitutiveOracle(material, material_at)
    b.lib = C.CDLL(lib)
    b.lib.orc_mechanical_response.restype = C.c_int
    a.mechanical_response(F.reshape(n, 9), 0.5)
    b.mechanical_response(F.reshape(n, 9), 0.5)
    err = np.abs(a.a['P'] - b.a['P']).max(1) / np.abs(a.a['P']).max(1)
    quiet = (a.point_cutbacks == 0) & (b.point_cutbacks == 0)
    assert err[quiet].max(initial=0.0) < 1e-9
    assert (a.point_cutbacks > 0).any()
    assert (err >= 1e-9).mean() < 0.1


@pytest.fixture(scope='module')
def g3(golden_dir):
    return np.load(os.path.join(golden_dir, 'g3_unit_test_literals.npz'))


def test_g3_divergence_rms_unit_test_literal(g3):
    """test_mech_base.cpp:324-385 (MechBaseTestDivergenceRMS): the reference's own literal field on a 4x1x1 grid and its
    expected value 1481.2323577332318; the xi table the oracle builds must be the literal one of the test."""
    cells = tuple(int(c) for c in g3['div_cells'])
    grid = sp.Grid(cells, tuple(1e-5 * c for c in cells))             # conftest.h:39-40: geom_size = cells * 1e-5
    xi1, _ = sp.xi_tables(grid, 0)
    lit = np.transpose(g3['div_xi1st'], (2, 3, 1, 0))                  # (3, kx, kz, ky) -> [kz, ky, kx, 3]
    assert np.abs(xi1 - lit).max() <= 1e-9 * np.abs(lit).max()
    field = np.transpose(g3['div_field'], (4, 3, 2, 0, 1))             # (i, j, x, y, z) -> [z, y, x, i, j]
    rms, _ = sp.divergence_rms(sp.fft_forward(field), xi1, grid, sp.scaled_geom_size(grid))
    assert abs(rms - float(g3['div_rms'])) <= 1e-12 * float(g3['div_rms'])


def test_g3_masked_compliance_unit_test_literal(g3):
    """test_mech_base.cpp:239-321 (MechBaseTestMaskedCompliance): literal stiffness, tensionX-like mask, literal result."""
    S = sp.masked_compliance(g3['mc_C'], np.eye(3), g3['mc_mask_stress'])
    exp = g3['mc_expected']
    assert np.abs(S - exp).max() <= 1e-12 * np.abs(exp).max()
    assert (S[exp == 0] == 0).all()


def test_g3_homogenisation_unit_test_literal(g3):
    """test_spectral.cpp:58-216 (SpectralTestConstitutiveResponse): literal P and dPdF of two cells ('10th iteration of
    tensionX') -> literal P_av, C_volAvg, C_minMaxAvg."""
    grid = sp.Grid((2, 1, 1), (2e-5, 1e-5, 1e-5))
    P = np.transpose(g3['hom_P'], (2, 0, 1))                           # (i, j, cell) -> [cell, i, j]
    dPdF = np.transpose(g3['hom_dPdF'], (4, 0, 1, 2, 3))
    P_av, C_vol, C_mm, _ = sp.homogenise(P, dPdF, grid)
    for got, key in ((P_av, 'hom_P_av'), (C_vol, 'hom_C_vol_avg'), (C_mm, 'hom_C_minmax_avg')):
        assert np.abs(got - g3[key]).max() <= 1e-12 * np.abs(g3[key]).max(), key


def test_g3_xi_tables_unit_test_literal(g3):
    """test_spectral.cpp:15-48 (SpectralTestInit): 2x1x1 grid of size 2e-5 x 1e-5 x 1e-5: xi2nd(0, kx=1) = i 314159.2653589793,
    xi1st identically zero (the only non-zero frequency is the Nyquist one)."""
    cells = tuple(int(c) for c in g3['xi_cells'])
    grid = sp.Grid(cells, tuple(1e-5 * c for c in cells))
    xi1, xi2 = sp.xi_tables(grid, 0)
    lit1 = np.transpose(g3['xi_expected_1st'], (2, 3, 1, 0))           # (3, kx, kz, ky) -> [kz, ky, kx, 3]
    lit2 = np.transpose(g3['xi_expected_2nd'], (2, 3, 1, 0))
    assert (xi1 == lit1).all()
    assert np.abs(xi2 - lit2).max() <= 1e-8                            # the reference compares with tensor_eq(eps = 1e-8)
    assert abs(xi2[0, 0, 1, 0].imag - 314159.2653589793) <= 1e-15 * 314159.2653589793


def _fld(a):
    return np.transpose(a, (4, 3, 2, 0, 1))                            # (i, j, x, y, z) -> [z, y, x, i, j]


def test_g3_forward_field_and_calculate_rate_unit_test_literals(g3):
    """test_mech_base.cpp:199-237 (MechBaseTestForwardField) and :492-545 (TestCalculateRate)"""
    grid = sp.Grid((2, 1, 1), (2e-5, 1e-5, 1e-5))
    got = sp.forward_field(float(g3['ff_delta_t']), _fld(g3['ff_field_last_inc']), _fld(g3['ff_rate']), grid, aim=g3['ff_aim'])
    assert np.abs(got - _fld(g3['ff_expected'])).max() <= 1e-15
    rate = sp.calculate_rate(bool(g3['cr_heterogeneous']), _fld(g3['cr_field_0']), _fld(g3['cr_field']), float(g3['cr_delta_t']),
                             g3['cr_av_rate'])
    exp = _fld(g3['cr_expected'])
    assert np.abs(rate - exp).max() <= 1e-17                           # literals printed to 17 digits
    assert (rate[exp == 0] == 0).all()
    # the homogeneous branch hands out the average rate
    assert np.array_equal(sp.calculate_rate(False, _fld(g3['cr_field_0']), _fld(g3['cr_field']), 1.0, g3['cr_av_rate'])[0, 0, 1], g3['cr_av_rate'])


def test_g3_form_residual_glue_unit_test_literal(g3):
    """test_mech_solver_basic.cpp:232-376 (TestFormResidual): with the constitutive response mocked to a literal P field on
    the 2x1x1 grid, the residual glue must give err_div = 0, err_BC = 11717501649.529392 (EXPECT_EQ, exact), the literal
    F_aim, and hand the literal rotated Delta F_aim to the Gamma convolution -- which on this grid (xi1st = 0) returns it in
    every cell.  The reference test also mocks divergence and Gamma; here the oracle's own run on the literal field."""
    from oracle.solver import MechSolverBasicOracle
    from solver_cpp_b200 import workloads as W
    cells = (2, 1, 1)
    material, material_at, _ = W.polycrystal(cells, 2, 'Al_pheno', 0)
    o = MechSolverBasicOracle(cells, (2e-5, 1e-5, 1e-5), material, material_at, tangent_every_eval=False)
    P_lit = _fld(g3['fr_P']).reshape(2, 3, 3)

    class MockedResponse:                                              # EXPECT_CALL(spectral, constitutive_response).WillOnce(Return(P))
        counters = np.zeros(6, dtype=np.int64)
        terminally_ill = False

        def mechanical_response(self, F, dt):
            return False

        def mechanical_response2(self, dt, tangent=True):
            return P_lit.copy(), None
    o.con = MockedResponse()
    o.S = g3['fr_S'].copy()
    o.P_av = g3['fr_P_av_before'].copy()
    o.F_aim = g3['fr_F_aim'].copy()
    o.P_aim = np.zeros((3, 3))
    o.params = dict(stress_mask=g3['fr_stress_mask'].copy(), R=np.eye(3), dt=1.0)
    o.C_ref = np.einsum('ik,jl->ijkl', np.eye(3), np.eye(3)) * 1.0e11
    r = o.form_residual(_fld(g3['fr_F']).reshape(-1), 0, 0).reshape(2, 3, 3)
    assert o.err_div == float(g3['fr_err_div']) == 0.0
    assert o.err_BC == float(g3['fr_err_BC'])                          # exact, like the reference's EXPECT_EQ
    assert np.abs(o.P_av - g3['fr_P_av_before']).max() <= 2e-6         # the literal P_av is the mean of the literal P (1.2e10 Pa)
    assert np.abs(o.F_aim - g3['fr_expected_F_aim']).max() <= 1e-8     # tensor_eq's eps
    assert np.abs(o.F_aim - g3['fr_expected_F_aim']).max() <= 5e-16
    for cell in range(2):
        assert np.abs(r[cell] - g3['fr_expected_gamma_aim']).max() <= 5e-17


def test_converged_reports_a_positive_reason_on_zero_errors():
    """test_mech_solver_basic.cpp:91-102 (TestConverged): total_iter = 1, zero errors -> converged.  (The reference test expects
    the enum value 2; mech_solver_basic.cpp:264 assigns SNES_BREAKOUT_INNER_ITER, 6 in PETSc 3.18's enum -- any positive
    reason ends the SNES loop as converged, which is what is pinned here.)"""
    from oracle.solver import MechSolverBasicOracle, SNES_BREAKOUT_INNER_ITER
    from solver_cpp_b200 import workloads as W
    cells = (2, 1, 1)
    material, material_at, _ = W.polycrystal(cells, 2, 'Al_pheno', 0)
    o = MechSolverBasicOracle(cells, (2e-5, 1e-5, 1e-5), material, material_at)
    o.total_iter = 1
    o.err_div = o.err_BC = 0.0
    assert o.converged(4) == SNES_BREAKOUT_INNER_ITER > 0
    o.total_iter = 0                                                   # below itmin = 1: keep iterating
    assert o.converged(0) == 0
    o.total_iter = o.num['itmax']; o.err_BC = 1.0e30
    assert o.converged(0) < 0


def test_g3_gamma_convolution_and_update_gamma_unit_test_literals(g3):
    """test_mech_base.cpp:389-490 (MechBaseTestGammaConvolution): literal stress field, literal C_ref, xi1st = 0 on the 2x1x1
    grid -> every cell holds field_aim (DC path, unnormalised inverse).  :154-197 (TestUpdateGamma): with xi = 0 the Gamma
    operator of any stiffness is zero -- here: applying it returns zero at every frequency."""
    grid = sp.Grid((2, 1, 1), (2e-5, 1e-5, 1e-5))
    xi1, _ = sp.xi_tables(grid, 0)
    assert (xi1 == 0).all()
    got = sp.gamma_convolution(_fld(g3['gc_field']), g3['gc_field_aim'], xi1, g3['gc_C_ref'], grid)
    exp = _fld(g3['gc_expected'])
    assert np.abs(got - exp).max() <= 1e-20 and (got[exp == 0] == 0).all()
    hat = sp.fft_forward(_fld(g3['gc_field']))
    for C in (g3['ug_C'], g3['gc_C_ref']):
        assert (sp.gamma_hat_apply(hat, xi1, C) == 0).all()


def test_g3_divergence_correction_unit_test_literals(g3):
    """test_mech_base.cpp:54-76 (MechBaseTestInitDivergenceCorrection): geometry 3 x 4 x 5 on 2 x 1 x 1 cells -> scaled
    geometry for divergence_correction none / size / size+grid (ASSERT_EQ: exact)"""
    grid = sp.Grid(tuple(int(c) for c in g3['dc_cells']), tuple(float(v) for v in g3['dc_geom_size']))
    assert sp.scaled_geom_size(grid, sp.DIVCORR_NONE).tolist() == g3['dc_geom_size'].tolist()
    assert sp.scaled_geom_size(grid, sp.DIVCORR_SIZE).tolist() == g3['dc_expected_size'].tolist()
    # size+grid: the test's literal (0.6, 0.8, 1) contradicts the code it tests -- mech_base.cpp:20-31 (and upstream DAMASK,
    # grid_mechanical_spectral_*: the axis that is neither minloc nor maxloc of size / cells) normalises by the MIDDLE one of
    # size / cells = (1.5, 4, 5), i.e. by axis 1: size / 4 * cells[1] = (0.75, 1, 1.25); the literal is size / 5, the largest.
    # The oracle follows the code; the literal is kept in the fixture to document the discrepancy.
    assert g3['dc_expected_size_grid'].tolist() == [0.6, 0.8, 1.0]
    assert sp.scaled_geom_size(grid, sp.DIVCORR_SIZE_GRID).tolist() == [0.75, 1.0, 1.25]


def test_g3_update_coords_unit_test_literal(g3):
    """test_mech_base.cpp:78-152 (MechBaseTestUpdateCoords): F = I on the 2x1x1 grid -> undeformed nodes (3, nx+1, ny+1, nz+1)
    and cell centres (3, nx, ny, nz)"""
    grid = sp.Grid((2, 1, 1), (2e-5, 1e-5, 1e-5))
    x_n, x_p = sp.update_coords(_fld(g3['uc_F']), grid)
    assert np.abs(np.transpose(x_n, (3, 2, 1, 0)) - g3['uc_expected_x_n']).max() <= 1e-20
    assert np.abs(np.transpose(x_p, (3, 2, 1, 0)) - g3['uc_expected_x_p']).max() <= 1e-20


def test_solver_init_and_forward_unit_test_literals():
    """test_mech_solver_basic.cpp:32-89 (TestMechSolverBasicInit): after init F_last_inc (and homogenization_F0) are the identity in
    every cell.  :159-231 (TestBasicForward): forward(cut_back = false, guess = false, dt = 100.2, dt_old = 1, t_remaining = 1002)
    with dot_F_11 = 1e-3 from F = I gives F_aim_11 = 1.1002000000000001, the forwarded field equal to the aim in both cells, and
    params.delta_t = 100.2, the rotation unchanged."""
    from oracle.solver import MechSolverBasicOracle, parse_load as oparse
    from solver_cpp_b200 import workloads as W
    cells = (2, 1, 1)
    material, material_at, _ = W.polycrystal(cells, 2, 'Al_pheno', 0)
    o = MechSolverBasicOracle(cells, (2e-5, 1e-5, 1e-5), material, material_at)
    o.init()
    I = np.broadcast_to(np.eye(3), (1, 1, 2, 3, 3))
    assert np.array_equal(o.F_last_inc, I) and np.array_equal(o.F, I)
    assert np.array_equal(o.con.a['F0'].reshape(2, 3, 3), I.reshape(2, 3, 3))
    st = oparse({'loadstep': [{'boundary_conditions': {'mechanical': {'dot_F': [[1e-3, 0, 0], [0, 'x', 0], [0, 0, 'x']],
                                                                       'P': [['x', 'x', 'x'], ['x', 0, 'x'], ['x', 'x', 0]],
                                                                       'R': [1.0, 0.0, 0.0, 0.0]}},
                               'discretization': {'t': 1002, 'N': 10}}]})[0]
    o.forward(False, False, 100.2, 1, 1002, st['deformation'], st['stress'], st['R'])
    expected_aim = np.diag([1.1002000000000001, 1.0, 1.0])
    assert np.array_equal(o.F_aim, expected_aim)
    assert np.array_equal(o.F_last_inc, I)
    assert np.abs(o.F - expected_aim).max() <= 2.3e-16                # tensor_eq(F, forwarded_field)
    assert o.params['dt'] == 100.2 and np.array_equal(o.params['R'], np.eye(3))
    assert np.array_equal(o.params['stress_mask'], [[True, True, True], [True, False, True], [True, True, False]])


def test_fft_round_trip_of_the_reference_unit_test():
    """test_fft.cpp:10-58 (TestInitForwardBackwardCopy / ...DirectAssignment): 2x1x1 cells, 3x3 components, random data in
    (1, 1000): forward (unnormalised) then backward with wgt = 1/N = 0.5 returns the field (tensor_eq: 1e-8 absolute)."""
    rng = np.random.default_rng(2)
    field = rng.uniform(1.0, 1000.0, size=(1, 1, 2, 3, 3))
    hat = sp.fft_forward(field)
    assert hat.shape == (1, 1, 2, 3, 3)                                # nx/2 + 1 = 2 complex coefficients along x
    assert np.abs(hat[0, 0, 0].real - field.sum(axis=(0, 1, 2))).max() <= 1e-12 * 2000.0      # unnormalised: DC = sum
    back = sp.fft_backward(hat, (2, 1, 1), 0.5)
    assert np.abs(back - field).max() <= 1e-12


def test_forward_boundary_condition_types():
    """MechSolverBasic::forward (mech_solver_basic.cpp:163-220) for the deformation types F / dot_F / L and the stress types
    P / dot_P.  Two lines of the C++ port are not followed (upstream DAMASK's grid_mechanical_spectral_basic_forward is): its "L"
    branch repeats the "F" formula (:188-189; upstream: L . F_aim_lastInc), and its P_aim update selects the components the mask
    marks 'x' (:205-209; upstream: the stress-controlled ones).  Properties pinned here: an F target is reached exactly at the end
    of its load step however the step is discretised, L gives the rate L . F_aim, P ramps linearly to its target, dot_P integrates."""
    from oracle.solver import MechSolverBasicOracle, parse_load as oparse
    from solver_cpp_b200 import workloads as W
    cells = (2, 1, 1)
    material, material_at, _ = W.polycrystal(cells, 2, 'Al_pheno', 0)

    def fresh():
        o = MechSolverBasicOracle(cells, (2e-5, 1e-5, 1e-5), material, material_at)
        o.init()
        return o

    def step_of(mech):
        return oparse({'loadstep': [{'boundary_conditions': {'mechanical': mech}, 'discretization': {'t': 4.0, 'N': 4}}]})[0]
    X = 'x'
    # F: target reached at the end of the step, for equal and for unequal time steps
    target = [[1.02, 0.01, 0], [0, X, 0], [0, 0, X]]
    st = step_of({'F': target, 'P': [[X, X, X], [X, 0, X], [X, X, 0]]})
    for dts in ([1.0, 1.0, 1.0, 1.0], [0.5, 1.5, 0.25, 1.75]):
        o, t = fresh(), 0.0
        for dt in dts:
            o.forward(False, False, dt, 1.0, 4.0 - t, st['deformation'], st['stress'], st['R'])
            t += dt
        assert abs(o.F_aim[0, 0] - 1.02) < 1e-15 and abs(o.F_aim[0, 1] - 0.01) < 1e-16 and o.F_aim[1, 1] == 1.0
        assert np.abs(o.F.mean(axis=(0, 1, 2)) - o.F_aim).max() < 1e-15        # the field follows its aim
    # L: rate = L . F_aim_lastInc on the controlled components
    L = np.array([[1e-3, 2e-3, 0], [0, 0, 0], [0, 0, 0]])
    st = step_of({'L': [[1e-3, 2e-3, 0], [0, X, 0], [0, 0, X]], 'P': [[X, X, X], [X, 0, X], [X, X, 0]]})
    o = fresh()
    o.F_aim = np.array([[1.1, 0.0, 0.0], [0.0, 0.9, 0.0], [0.0, 0.0, 1.0]])
    o.forward(False, False, 0.5, 1.0, 4.0, st['deformation'], st['stress'], st['R'])
    expect = np.where(st['deformation'].mask, 0.0, L @ np.diag([1.1, 0.9, 1.0]))
    assert np.array_equal(o.F_aim_dot, expect) and np.abs(o.F_aim - (np.diag([1.1, 0.9, 1.0]) + 0.5 * expect)).max() < 1e-16
    assert o.F_aim_dot[0, 1] == 2e-3 * 0.9                                      # not the "F" formula of the C++ port
    # P ramps to its target over the step; dot_P integrates; components marked x stay untouched
    st = step_of({'dot_F': [[1e-3, 0, 0], [0, X, 0], [0, 0, X]], 'P': [[X, X, X], [X, 4.0e6, X], [X, X, -2.0e6]]})
    o, t = fresh(), 0.0
    o.P_aim[0, 0] = 7.0
    for dt in (0.5, 1.5, 0.25, 1.75):
        o.forward(False, False, dt, 1.0, 4.0 - t, st['deformation'], st['stress'], st['R'])
        t += dt
        assert abs(o.P_aim[1, 1] - 4.0e6 * t / 4.0) < 1e-8 and abs(o.P_aim[2, 2] + 2.0e6 * t / 4.0) < 1e-8 and o.P_aim[0, 0] == 7.0
    st = step_of({'dot_F': [[1e-3, 0, 0], [0, X, 0], [0, 0, X]], 'dot_P': [[X, X, X], [X, 1.0e6, X], [X, X, 0]]})
    o = fresh()
    for k in range(3):
        o.forward(False, False, 0.5, 1.0, 4.0 - 0.5 * k, st['deformation'], st['stress'], st['R'])
    assert o.P_aim[1, 1] == 1.5e6 and o.P_aim[2, 2] == 0.0
    # a cut-back repeats the step from the same F_aim_lastInc with the smaller dt
    st = step_of({'dot_F': [[1e-3, 0, 0], [0, X, 0], [0, 0, X]], 'P': [[X, X, X], [X, 0, X], [X, X, 0]]})
    o = fresh()
    o.forward(False, False, 1.0, 1.0, 4.0, st['deformation'], st['stress'], st['R'])
    assert o.F_aim[0, 0] == 1.001
    o.forward(True, False, 0.5, 1.0, 4.0, st['deformation'], st['stress'], st['R'])
    assert o.F_aim[0, 0] == 1.0005 and np.array_equal(o.F_aim_last_inc, np.eye(3))


def test_g4_real_damask_run_reproduced_within_solver_tolerance(golden_dir):
    """Golden set G4: a real DAMASK_grid (3.0.0-alpha7) result file shipped with the reference
    (python/tests/resources/Result/6grains6x7x8_single_phase_tensionY.hdf5; non-cubic 6x7x8 grid, mixed boundary
    conditions, 40 increments), read with the repo's HDF5 reader (tests/golden/make_golden_g4.py).  Two converged
    equilibrium solves agree to the solver's own tolerances (eps_rel_div = 5e-4, eps_rel_P = 1e-3), not to rounding:
    per-point F to 1e-5 absolute, per-point P to 5e-4 of max|P|, homogenised P to 3e-4, final slip resistances to 1e-4."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('dmk_vti', os.path.join(golden_dir, '..', '..', 'solver_cpp_b200', 'vti.py'))
    vti = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(vti)
    g4 = np.load(os.path.join(golden_dir, 'g4_damask_run.npz'))
    geo = vti.read_vti(os.path.join(golden_dir, '6grains6x7x8_single_phase.vti'))
    mat = yaml.safe_load(open(os.path.join(golden_dir, '6grains6x7x8_single_phase.material.yaml')))
    load = yaml.safe_load(open(os.path.join(golden_dir, '6grains6x7x8_tensionY.load.yaml')))
    o = MechSolverBasicOracle(tuple(int(c) for c in geo['cells']), tuple(geo['size']), mat, geo['material'].reshape(-1))
    o.init()
    snap = {}

    def cb(rec, s):
        if rec['inc'] in g4['increments']:
            snap[rec['inc']] = (s.F.reshape(-1, 3, 3).copy(), s.con.a['P'].reshape(-1, 3, 3).copy(), s.con.a['state'].copy())
    recs = run_loadcase(o, parse_load(load), callback=cb)
    assert len(recs) == 40 and all(r['converged'] for r in recs)
    for k, inc in enumerate(g4['increments']):
        F, P, _ = snap[int(inc)]
        Pd = g4['P'][k]
        assert np.abs(F - g4['F'][k]).max() < 1e-5
        assert np.abs(P - Pd).max() < 5e-4 * np.abs(Pd).max()
        assert np.abs(P.mean(0) - Pd.mean(0)).max() < 3e-4 * np.abs(Pd.mean(0)).max()
        dFav = np.abs(F.mean(0) - g4['F'][k].mean(0))
        free = np.zeros((3, 3), bool)
        free[0, 0] = free[2, 2] = True                                  # stress-controlled components of the average
        assert dFav[~free].max() < 1e-12 and dFav[free].max() < 1e-6
    assert np.abs(snap[40][2][:, :12] - g4['xi_sl_final']).max() < 1e-4 * g4['xi_sl_final'].max()


def test_g4_update_coords_reproduces_damask_displacements(golden_dir):
    """u_p and u_n (cell-centre and nodal displacements) written by DAMASK itself, recomputed from the F it stored:
    agreement to rounding pins base_update_coords (and shows the C++ port's sign / double-weight defects, oracle docstring)."""
    g4 = np.load(os.path.join(golden_dir, 'g4_damask_run.npz'))
    grid = sp.Grid((6, 7, 8), (0.75, 0.875, 1.0))
    x_n0, x_p0 = sp.update_coords(np.broadcast_to(np.eye(3), (8, 7, 6, 3, 3)), grid)
    for k in range(len(g4['increments'])):
        x_n, x_p = sp.update_coords(g4['F'][k].reshape(8, 7, 6, 3, 3), grid)
        assert np.abs((x_p - x_p0).reshape(-1, 3) - g4['u_p'][k]).max() < 1e-15
        assert np.abs((x_n - x_n0).reshape(-1, 3) - g4['u_n'][k]).max() < 1e-15


def test_g4b_two_phase_real_damask_run_within_solver_tolerance(golden_dir):
    """12grains6x7x8_tensionY.hdf5 of the reference: the same load case on a 12-grain microstructure with an fcc and a bcc
    (cI, <111>{110}) phenopowerlaw phase -- pins the multi-phase path and the cI tables against a real DAMASK run."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('dmk_vti', os.path.join(golden_dir, '..', '..', 'solver_cpp_b200', 'vti.py'))
    vti = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(vti)
    g4 = np.load(os.path.join(golden_dir, 'g4b_damask_run_two_phase.npz'))
    geo = vti.read_vti(os.path.join(golden_dir, '12grains6x7x8.vti'))
    mat = yaml.safe_load(open(os.path.join(golden_dir, '12grains6x7x8.material.yaml')))
    load = yaml.safe_load(open(os.path.join(golden_dir, '6grains6x7x8_tensionY.load.yaml')))
    o = MechSolverBasicOracle(tuple(int(c) for c in geo['cells']), tuple(geo['size']), mat, geo['material'].reshape(-1))
    o.init()
    snap = {}

    def cb(rec, s):
        if rec['inc'] in g4['increments']:
            snap[rec['inc']] = (s.F.reshape(-1, 3, 3).copy(), s.con.a['P'].reshape(-1, 3, 3).copy())
    recs = run_loadcase(o, parse_load(load), callback=cb)
    assert len(recs) == 40 and all(r['converged'] for r in recs)
    for k, inc in enumerate(g4['increments']):
        F, P = snap[int(inc)]
        Pd = g4['P'][k]
        assert np.abs(F - g4['F'][k]).max() < 1e-5
        assert np.abs(P - Pd).max() < 6e-4 * np.abs(Pd).max()
        assert np.abs(P.mean(0) - Pd.mean(0)).max() < 3e-4 * np.abs(Pd.mean(0)).max()


def test_g6_cell_centre_and_node_coordinates_unit_test_literals(golden_dir):
    """test_discretization_grid.cpp:70-101 (TestCalculateIpCoords0 / TestCalculateNodes0): undeformed cell centres and nodes
    of the 2x1x1 grid of size 2e-5 x 1e-5 x 1e-5, x fastest.  The oracle has no separate routine for them: they are what
    update_coords (mech_base.cpp:45-154) returns at F = I, which is how the result writer obtains them."""
    g6 = np.load(os.path.join(golden_dir, 'g6_discretization_literals.npz'))
    cells = tuple(int(c) for c in g6['coords_cells'])
    grid = sp.Grid(cells, tuple(float(s) for s in g6['coords_size']))
    nx, ny, nz = cells
    x_n, x_p = sp.update_coords(np.broadcast_to(np.eye(3), (nz, ny, nx, 3, 3)).copy(), grid)
    assert np.abs(x_p.reshape(-1, 3).T - g6['ip_coords0']).max() <= 1e-20
    assert np.abs(x_n.reshape(-1, 3).T - g6['node_coords0']).max() <= 1e-20


def test_g6_voronoi_generator_reproduces_the_reference_geometry(golden_dir):
    """The 16^3 VTI text embedded in the reference's reader test (test_discretization_grid.cpp:103-124) is the periodic Voronoi
    tessellation of examples/grid/20grains.seeds (damask.Grid.from_Voronoi_tessellation, python/damask/_grid.py:563-571).
    solver_cpp_b200.workloads.voronoi_material_at -- the generator of every synthetic polycrystal in bench.py and the tests --
    must give the same material ID in all 4096 cells, and workloads.reference_20grains() (BASELINE.json configs[0]) must be
    that deck: seeds of 20grains.seeds, orientations and phase of examples/grid/material.yaml."""
    import importlib.util
    import yaml
    from solver_cpp_b200 import workloads as W
    spec = importlib.util.spec_from_file_location('dmk_vti', os.path.join(os.path.dirname(golden_dir), '..', 'solver_cpp_b200', 'vti.py'))
    vti = importlib.util.module_from_spec(spec); spec.loader.exec_module(vti)
    geo = vti.read_vti(os.path.join(golden_dir, 'g6_unit_test_geometry.vti'))
    rows = [l.split() for l in open(os.path.join(golden_dir, '20grains.seeds'))]
    seeds = np.array([[float(v) for v in r] for r in rows if len(r) == 6 and r[0].replace('.', '').isdigit()])
    assert seeds.shape == (20, 6)
    assert np.array_equal(np.asarray(W.REFERENCE_20GRAINS_SEEDS), seeds[:, :3])
    got = W.voronoi_material_at((16, 16, 16), tuple(geo['size']), seeds[:, :3] * geo['size'])
    assert np.array_equal(got, geo['material'].ravel())
    # slabs of the same tessellation (what each rank of a multi-GPU job computes for itself)
    parts = [W.voronoi_material_at((16, 16, 16), tuple(geo['size']), seeds[:, :3] * geo['size'], z_range=(a, b)) for a, b in ((0, 5), (5, 16))]
    assert np.array_equal(np.concatenate(parts), got)
    material, material_at, size = W.reference_20grains()
    assert np.array_equal(material_at, got) and size == (1.0, 1.0, 1.0)
    deck = yaml.safe_load(open(os.path.join(golden_dir, 'material_20grains_Al.yaml')))
    assert [m['constituents'][0]['O'] for m in material['material']] == [m['constituents'][0]['O'] for m in deck['material']]
    ours, theirs = material['phase']['Aluminum'], deck['phase']['Aluminum']
    assert ours['lattice'] == theirs['lattice'] and ours['mechanical']['elastic'] == theirs['mechanical']['elastic']
    assert ours['mechanical']['plastic'] == {k: v for k, v in theirs['mechanical']['plastic'].items() if k != 'output'}


def test_reference_16cubed_deck_solves_like_the_recorded_27_cell_run():
    """BASELINE.json configs[0] (the reference's own example deck, workloads.reference_20grains: 16^3 cells, 20 grains, Al
    phenopowerlaw, tensionX with the example numerics N_iter_min 4 / N_iter_max 40) through the oracle's solver.  No recorded
    output of this deck exists in the reference tree; the recorded Fortran run G1 is the same material and load on 27 cells with
    the same 20 orientations, so the homogenised stress-strain curve must lie close to G1's (35.24, 64.66, 72.95 MPa)."""
    from oracle.solver import MechSolverBasicOracle, run_loadcase, parse_load as oparse
    from solver_cpp_b200 import workloads as W
    material, material_at, size = W.reference_20grains()
    o = MechSolverBasicOracle((16, 16, 16), size, material, material_at, tangent_every_eval=False, grid_mech=dict(itmin=4, itmax=40))
    o.log_fields = False
    o.init()
    recs = []

    class Stop(Exception):
        pass

    def cb(rec, s):
        recs.append(rec)
        if len(recs) == 3:
            raise Stop
    try:
        run_loadcase(o, oparse(W.TENSION_X), callback=cb)
    except Stop:
        pass
    assert [r['converged'] for r in recs] == [True] * 3 and all(4 <= r['iterations'] <= 12 for r in recs)
    P11 = np.array([r['P_av'][0, 0] for r in recs]) * 1e-6
    assert np.abs(P11 / np.array([35.2362, 64.6648, 72.9485]) - 1.0).max() < 0.03
    for r in recs:                                                     # uniaxial stress state of tensionX
        assert abs(r['P_av'][1, 1]) < 1e-3 * r['P_av'][0, 0] and abs(r['P_av'][2, 2]) < 1e-3 * r['P_av'][0, 0]


def test_fully_deformation_controlled_load_case_like_the_reference_shear_decks():
    """examples/grid/shearXY.yaml of the reference: dot_F prescribes all nine components, no stress entry.  The masked compliance
    of an all-'x' stress mask is zero (oracle and the library's host helper), F_aim follows the prescribed rate exactly, and the
    solver converges on the equilibrium alone."""
    import ctypes as C
    from oracle.solver import MechSolverBasicOracle, run_loadcase, parse_load as oparse
    from solver_cpp_b200 import workloads as W, lib as L
    load = {'solver': {'mechanical': 'spectral_basic'},
            'loadstep': [{'boundary_conditions': {'mechanical': {'dot_F': [[0, 0, 0], [1.0e-3, 0, 0], [0, 0, 0]]}},
                          'discretization': {'t': 60, 'N': 120}, 'f_out': 20}]}
    Cs = np.einsum('ik,jl->ijkl', np.eye(3), np.eye(3)) * 1.0e11
    assert np.abs(sp.masked_compliance(Cs, np.eye(3), np.ones((3, 3), bool))).max() == 0.0
    lib = L.load()
    S, mask, R = np.ones(81), np.ones(9, dtype=np.int32), np.eye(3).reshape(9).copy()
    Cf = np.ascontiguousarray(np.transpose(Cs, (3, 2, 1, 0))).reshape(-1)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    assert lib.dmk_masked_compliance(P(Cf), P(R), P(mask), P(S)) == 0 and np.abs(S).max() == 0.0
    cells = (6, 6, 6)
    material, material_at, size = W.polycrystal(cells, 6, 'Al_pheno', 0)
    o = MechSolverBasicOracle(cells, size, material, material_at, tangent_every_eval=False)
    o.log_fields = False
    o.init()
    recs = []

    class Stop(Exception):
        pass

    def cb(rec, s):
        recs.append(rec)
        if len(recs) == 3:
            raise Stop
    try:
        run_loadcase(o, oparse(load), callback=cb)
    except Stop:
        pass
    assert [r['converged'] for r in recs] == [True] * 3
    for k, r in enumerate(recs):
        expect = np.eye(3); expect[1, 0] = 1.0e-3 * 0.5 * (k + 1)
        assert np.abs(r['F_aim'] - expect).max() < 1e-18 and np.abs(r['F_av'] - expect).max() < 1e-15
        assert o.err_BC == 0.0                                         # nothing is stress controlled
    assert recs[2]['P_av'][1, 0] > recs[1]['P_av'][1, 0] > recs[0]['P_av'][1, 0] > 1.0e7


def _g8_job(golden_dir):
    import yaml
    g8 = np.load(os.path.join(golden_dir, 'g8_reference_run_2x2x2.npz'))
    material = yaml.safe_load(str(g8['material_yaml']))
    numerics = yaml.safe_load(str(g8['numerics_yaml']))['solver']['grid']['mechanical']
    load = {'solver': {'mechanical': 'spectral_basic'},
            'loadstep': [{'boundary_conditions': {'mechanical': {'dot_F': [[1.0e-3, 0, 0], [0, 'x', 0], [0, 0, 'x']],
                                                                  'P': [['x', 'x', 'x'], ['x', 0, 'x'], ['x', 'x', 0]]}},
                          'discretization': {'t': 10, 'N': 10}, 'f_out': 1}]}                # dt = t of the stored increment = 1 s
    return g8, material, dict(itmin=int(numerics['N_iter_min']), itmax=int(numerics['N_iter_max'])), load


def test_g8_result_file_of_the_reference_fork_itself(golden_dir):
    """Golden set G8: increment 1 of a result file the reference's author wrote with the C++ fork (2 x 2 x 2 cells of one
    cube-oriented Al crystal, tensionX, N_iter_min 4): a complete converged increment -- forward, four NGMRES iterations, stress BC --
    in the reference's own output.  The oracle's free-running solve must land on it: F, F_e, F_p, xi_sl to rounding (3e-16),
    P and L_p to 1e-11 (the zero-stress components are converged to 6e-5 Pa of 6e7 in both)."""
    from oracle.solver import MechSolverBasicOracle, run_loadcase, parse_load as oparse
    g8, material, num, load = _g8_job(golden_dir)
    assert float(g8['t']) == 1.0 and 'alpha8' in str(g8['creator'])
    for k in ('F', 'F_e', 'F_p', 'L_p', 'P', 'xi_sl'):
        assert np.abs(g8[k] - g8[k][0]).max() == 0.0                    # homogeneous: all 8 cells alike
    o = MechSolverBasicOracle((2, 2, 2), (1.0, 1.0, 1.0), material, np.zeros(8, dtype=int), grid_mech=num, tangent_every_eval=False)
    o.log_fields = False
    o.init()
    recs = []

    class Stop(Exception):
        pass

    def cb(rec, s):
        recs.append(rec)
        raise Stop
    try:
        run_loadcase(o, oparse(load), callback=cb)
    except Stop:
        pass
    assert recs[0]['converged'] and recs[0]['iterations'] == 4 and recs[0]['t'] == float(g8['t'])
    got = {'F': o.con.a['F'], 'F_e': o.con.a['Fe'], 'F_p': o.con.a['Fp'], 'L_p': o.con.a['Lp'], 'P': o.con.a['P'], 'xi_sl': o.con.a['state'][:, :12]}
    for k, tol in (('F', 3e-16), ('F_e', 3e-16), ('F_p', 3e-16), ('xi_sl', 3e-16), ('P', 1e-11), ('L_p', 1e-11)):
        ref = g8[k].reshape(8, -1)
        assert np.abs(got[k].reshape(8, -1) - ref).max() <= tol * np.abs(ref).max(), (k, np.abs(got[k].reshape(8, -1) - ref).max() / np.abs(ref).max())
    assert abs(o.con.a['F'][0, 4] - 0.9996331955299964) < 2e-16            # the lateral contraction the stress BC found


def test_g8_first_evaluation_known_answer_of_the_fortran_log(golden_dir):
    """The same 2 x 2 x 2 job, first residual evaluation, as the Fortran side printed it (test.patch at the root of the reference
    tree, the author's diff of the Fortran and the C++ run logs; lines 706-746): F = diag(1.001, 1, 1), dt = 1 s, virgin cube-oriented
    Al -> per point F_e11 = 1.0009998313594122, S11 = 106795532.01732938, S22 = 60444088.135403559 and the homogenised
    P11 = 106902291.52922715, P22 = 60444098.318547405.  17 printed digits: the oracle must reproduce every one."""
    import yaml
    g8 = np.load(os.path.join(golden_dir, 'g8_reference_run_2x2x2.npz'))
    o = ConstitutiveOracle(yaml.safe_load(str(g8['material_yaml'])), np.zeros(8, dtype=int))
    o.mechanical_response(np.tile(np.diag([1.001, 1.0, 1.0]).reshape(9), (8, 1)), 1.0)
    P, _ = o.mechanical_response2(1.0, tangent=False)
    assert np.abs(P - P[0]).max() == 0.0                                                   # homogeneous: every cell alike
    assert P[0, 0, 0] == 106902291.52922715 and P[0, 1, 1] == 60444098.318547405          # the same doubles as the log's P_av2
    assert o.a['Fe'][0, 0] == 1.0009998313594122 and abs(o.a['Fe'][0, 4] - 1.0000000842360579) <= 2.3e-16
    S = o.a['S'][0].reshape(3, 3)
    assert S[0, 0] == 106795532.01732938 and abs(S[1, 1] - 60444088.135403559) <= 1.5e-8


def test_g9_third_real_damask_run_on_an_odd_two_phase_grid(golden_dir):
    """Golden set G9: python/tests/resources/Result/2phase_irregularGrid_tensionX_*.hdf5 of the reference (DAMASK_grid
    3.0.0-alpha7-880; 13 x 8 x 10 cells: an odd, non-power-of-two grid; 18 grains of an fcc phenopowerlaw matrix and an ELASTIC bcc
    precipitate; tensionX, 40 increments of 0.25 s; inputs from the result file's own /setup group, tests/golden/make_golden_g9.py).
    Two converged solves agree to the solver's tolerances (eps_rel_div 5e-4, eps_rel_P 1e-3): per-cell F to 1e-4, P to 3e-4 of
    max|P|, homogenised P to 1e-4, the matrix's final slip resistances to 2e-4 (observed 4.6e-5, 1e-4, 1.8e-5, 6.5e-5)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('dmk_vti', os.path.join(golden_dir, '..', '..', 'solver_cpp_b200', 'vti.py'))
    vti = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(vti)
    g9 = np.load(os.path.join(golden_dir, 'g9_damask_run_irregular.npz'))
    geo = vti.read_vti(os.path.join(golden_dir, 'g9_irregular.vti'))
    assert tuple(int(c) for c in geo['cells']) == (13, 8, 10) and 'DAMASK_grid' in str(g9['creator'])
    mat = yaml.safe_load(open(os.path.join(golden_dir, 'g9_irregular.material.yaml')))
    load = yaml.safe_load(open(os.path.join(golden_dir, 'g9_irregular.load.yaml')))
    assert mat['phase']['Precipitate']['lattice'] == 'cI' and 'plastic' not in mat['phase']['Precipitate']['mechanical']
    o = MechSolverBasicOracle(tuple(int(c) for c in geo['cells']), tuple(geo['size']), mat, geo['material'].reshape(-1), tangent_every_eval=False)
    o.log_fields = False
    o.init()
    recs = run_loadcase(o, parse_load(load))
    assert len(recs) == 40 and all(r['converged'] for r in recs) and abs(recs[-1]['t'] - float(g9['t'])) < 1e-12
    F, P = o.F.reshape(-1, 3, 3), o.con.a['P'].reshape(-1, 3, 3)
    assert np.abs(F - g9['F']).max() < 1e-4
    assert np.abs(P - g9['P']).max() < 3e-4 * np.abs(g9['P']).max()
    assert np.abs(P.mean(0) - g9['P'].mean(0)).max() < 1e-4 * np.abs(g9['P'].mean(0)).max()
    matrix = g9['phase_of_cell'] == 'Matrix'
    assert 0 < matrix.sum() < matrix.size and np.isnan(g9['xi_sl'][~matrix]).all()          # the precipitate carries no slip resistance
    assert np.abs(o.con.a['state'][matrix, :12] - g9['xi_sl'][matrix]).max() < 2e-4 * g9['xi_sl'][matrix].max()
    dFav = np.abs(F.mean(0) - g9['F'].mean(0))
    free = np.zeros((3, 3), bool)
    free[1, 1] = free[2, 2] = True                                         # stress-controlled components of the average
    assert dFav[~free].max() < 1e-12 and dFav[free].max() < 1e-6
