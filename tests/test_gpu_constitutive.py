"""GPU parity tests of the per-point constitutive kernel, the tangent kernel and the homogenisation reductions against
the oracle (oracle/constitutive.c, pinned by G1/G2), through the C ABI.

Tolerance (BASELINE.json north_star): per-point stress within 1e-9 relative in FP64, identical iteration counts and
convergence flags.  The kernel assembles the Newton Jacobian in closed form instead of multiplying 9x9 matrices, so
rounding differs from the oracle at the 1e-16 level; converged iterates agree to ~1e-12."""
import os

import numpy as np
import pytest
import yaml

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from oracle import spectral as osp                      # noqa: E402  (checker only)
from oracle.constitutive import ConstitutiveOracle      # noqa: E402

RTOL = 1.0e-9


def _ctx(cells, size):
    from solver_cpp_b200.grid import GridContext
    return GridContext(cells, size)


def _setup(cells, grains, phase, seed=0, num_phase_mech=None, mass_deferral=0):
    """mass_deferral = 0: points whose first sub-step fails ALWAYS go to the reference-order kernel, the path whose per-cell
    parity these tests pin; the default (in-lane cut-backs when more than 1/32 of the cells gave up, i.e. on a virgin-state
    evaluation) has its own test: test_mass_cutback_evaluation_default_path"""
    from solver_cpp_b200 import workloads as W, material as M
    material, material_at, size = W.polycrystal(cells, grains, phase, seed)
    g = _ctx(cells, size)
    g.set_option('pheno_mass_deferral', mass_deferral)
    g.material_init(M.Material(material, material_at), M.make_numerics(num_phase_mech) if num_phase_mech else None)
    o = ConstitutiveOracle(material, material_at, num_phase_mech=num_phase_mech)
    return g, o


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _F_sequence(cells, n_eval, rng, rate=1.0e-3, dt=0.5, noise=2.0e-4):
    """tensionX-like deformation path with spatial fluctuations, several evaluations per increment"""
    nx, ny, nz = cells
    out = []
    F = np.broadcast_to(np.eye(3), (nz, ny, nx, 3, 3)).copy()
    fluct = rng.standard_normal((nz, ny, nx, 3, 3))
    for inc in range(n_eval):
        Fbar = np.eye(3)
        Fbar[0, 0] += rate * dt * (inc + 1)
        Fbar[1, 1] -= 0.3 * rate * dt * (inc + 1)
        Fbar[2, 2] -= 0.3 * rate * dt * (inc + 1)
        out.append(Fbar + noise * (inc + 1) * fluct)
    return out


def _pointwise_rel(a, b):
    a = a.reshape(a.shape[0], -1); b = b.reshape(b.shape[0], -1)
    return np.abs(a - b).max(1) / np.maximum(np.abs(b).max(1), 1e-300)


@pytest.mark.parametrize('phase,cells,grains,n_inc', [('Al_pheno', (8, 6, 5), 11, 6), ('Cu_pheno', (16, 16, 8), 20, 5),
                                                       ('Al_iso', (8, 8, 8), 9, 6), ('Al_elastic', (6, 5, 4), 7, 2),
                                                       ('TWIP', (8, 8, 4), 9, 5)])
def test_constitutive_matches_oracle_over_increments(phase, cells, grains, n_inc):
    """Every cell must match the oracle to 1e-9 -- except cells whose sub-step cut-back path is rounding-chaotic:
    when the first trial step overshoots by tens of orders of magnitude (soft virgin Cu), whether the next, smaller
    sub-step converges is decided by the last bit, and ANY two IEEE-valid builds of the reference algorithm differ
    there (tests/test_oracle_golden.py::test_cutback_path_is_rounding_sensitive shows it for the oracle and its own
    FMA-contracted build).  Such cells must have cut back (in the oracle or on the device), stay rare, and stay close.
    Second, milder case: a stopping / line-search decision of the Lp Newton loop whose two sides agree to rounding may
    flip; the cell then ends an iteration apart from the oracle (per-cell Lp iteration counts differ), both answers
    inside the integrator's own tolerance (1e-6).  Such cells must agree to 1e-6 (observed: 1e-11).  On this deliberately
    rough path (white-noise F, most Cu cells cut back, the others need 50-90 Newton iterations with long line searches)
    they are a sizeable part of the few quiet cells; test_newton_decisions_rarely_flip_on_a_solver_iterate bounds their
    share on the kind of iterate a running solve produces."""
    g, o = _setup(cells, grains, phase)
    rng = np.random.default_rng(11)
    dt = 0.5
    n = int(np.prod(cells))
    nst = o.phases[0].n_state
    clean = np.ones(n, bool)
    seq = _F_sequence(cells, n_inc, rng, rate=(2.0e-3 if phase != 'TWIP' else 6.0e-3))
    for inc, Ftarget in enumerate(seq):
        # two evaluations per increment (an intermediate iterate, then the target), like the NGMRES driver does
        for Fe_ in (0.5 * (Ftarget + (seq[inc - 1] if inc else np.broadcast_to(np.eye(3), Ftarget.shape))), Ftarget):
            ill_o = o.mechanical_response(Fe_.reshape(n, 9), dt)
            P_o, _ = o.mechanical_response2(dt, tangent=False)
            P = g.to_natural(g.constitutive_response(g.to_device(Fe_), dt)).reshape(n, 3, 3)
            assert g.terminally_ill() == ill_o
            err = _pointwise_rel(P, P_o)                                   # the north-star quantity: 1e-9
            aux = _pointwise_rel(g.get_state('Fp', 9), o.a['Fp'])          # internal variables: 1e-8
            if nst:
                aux = np.maximum(aux, _pointwise_rel(g.get_state('state', o.max_state)[:, :nst], o.a['state'][:, :nst]))
            bad = ((err >= RTOL) | (aux >= 10 * RTOL)) & clean
            quiet = (o.point_cutbacks == 0) & (g.get_state('cutbacks', 1)[:, 0] == 0)      # no cut-back on either side
            flipped = quiet & clean & (g.get_state('lp_iter', 1)[:, 0] != o.point_lp_iter)   # a Newton stopping decision flipped
            assert (err[flipped] < 1e-6).all(), (inc, flipped.mean(), err[flipped].max())
            quiet &= ~flipped
            assert not (bad & quiet).any(), (inc, err[bad & quiet].max(), aux[bad & quiet].max())
            assert err[clean].max() < 0.2
            assert _rel(g.get_state('Lp', 9)[clean & ~bad], o.a['Lp'][clean & ~bad]) < 1e-7
            clean &= ~bad
        g.materialpoint_forward()
        o.forward()
    assert clean.mean() > 0.7, clean.mean()
    if phase in ('Al_pheno', 'Al_elastic'):
        assert clean.all()
    # plasticity was active (otherwise the test pins only Hooke)
    if phase not in ('Al_elastic',):
        assert np.abs(o.a['Lp']).max() > 1e-6


@pytest.mark.parametrize('integrator', ['FPI', 'RK4'])
def test_kinehardening_matches_oracle_through_a_load_reversal(integrator):
    """kinehardening (phase_mechanical_plastic_kinehardening.f90, parameters of kinehardening_317L.yaml; unpinned: the reference
    tree holds no fixture for it) in the reference-order kernel against the oracle on a tension -> unloading -> compression
    path with spatial fluctuations.  The shear direction of the active systems flips on the way back, so the deltaState jump
    (plastic_deltaState, phase_mechanical_plastic.f90:372-413: sgn_gamma, chi_flip, gamma_flip) and the part of the state
    vector the integrators leave alone are exercised: P 1e-9, state 1e-9 (or a thousandth of the integrator's absolute
    tolerance where chi runs through zero), equal Lp iteration counts, the terminally-ill flag; tangent 1e-8."""
    cells, dt = (8, 6, 5), 0.5
    npm = None if integrator == 'FPI' else {'plastic': {'integrator_state': integrator}}
    g, o = _setup(cells, 9, 'Steel_kine', num_phase_mech=npm)
    n = int(np.prod(cells))
    ns, nsl = o.phases[0].n_state, o.phases[0].n_sl
    assert ns == 6 * nsl
    atol = np.ctypeslib.as_array(o.phases[0].atol)[:ns].copy()
    atol[atol == 0.0] = 1e-300
    rng = np.random.default_rng(5)
    nx, ny, nz = cells
    fluct = rng.standard_normal((nz, ny, nx, 3, 3))
    compared = 0
    clean = np.ones(n, bool)                    # no cut-back anywhere in the cell's history, on either side: same start state
    for k, e in enumerate([1.5e-3, 3.0e-3, 4.5e-3, 3.0e-3, 1.5e-3, 0.0, -1.5e-3, -3.0e-3]):
        F = np.eye(3) + np.diag([e, -0.3 * e, -0.3 * e]) + 1.0e-4 * (k + 1) * fluct
        ill_o = o.mechanical_response(F.reshape(n, 9), dt)
        P_o, dPdF_o = o.mechanical_response2(dt, tangent=(k in (2, 6)))
        P = g.to_natural(g.constitutive_response(g.to_device(F), dt)).reshape(n, 3, 3)
        assert g.terminally_ill() == ill_o and not ill_o
        # first yield cuts back in part of the cells (rounding-chaotic, see above): they leave the comparison for good
        clean &= (o.point_cutbacks == 0) & (g.get_state('cutbacks', 1)[:, 0] == 0)
        same = clean & (g.get_state('lp_iter', 1)[:, 0] == o.point_lp_iter)
        assert clean.mean() > 0.5 and same.sum() >= 0.98 * clean.sum(), (k, clean.mean(), same.sum() / max(clean.sum(), 1))
        err = _pointwise_rel(P, P_o)
        assert (err[same] < RTOL).all() and (err[clean] < 1e-6).all(), (k, err[same].max(), err[clean].max())
        st, st_o = g.get_state('state', o.max_state)[:, :ns], o.a['state'][:, :ns]
        ds = np.abs(st - st_o)
        # 1e-9 of the entry (or of its absolute tolerance); the back stress chi moves by h_0_chi = 55 GPa per unit of shear and runs
        # through zero on the way back, so its rounding differences are those of gamma times h_0_chi: 1e-9 of h_0_chi max|gamma|
        scale = np.maximum(np.abs(st_o), atol)
        scale[:, nsl:2 * nsl] = np.maximum(scale[:, nsl:2 * nsl], 55.0e9 * np.abs(st_o[:, 2 * nsl:3 * nsl]).max())
        assert (ds[same] <= 1e-9 * scale[same]).all(), (k, (ds[same] / scale[same]).max())
        if k in (2, 6):
            from solver_cpp_b200.grid import fortran_to_c3333
            _, _, _, d = g.tangent_response(dt, materialise=True)
            dP = np.stack([fortran_to_c3333(r) for r in d.cpu().numpy()])
            et = _pointwise_rel(dP, dPdF_o)
            assert (et[same] < 1e-8).all(), (k, et[same].max())
        compared += int(same.sum())
        clean &= same
        g.materialpoint_forward()
        o.forward()
    st = o.a['state'][:, :ns]
    sgn, chi_flip = st[:, 3 * nsl:4 * nsl], st[:, 4 * nsl:5 * nsl]
    assert compared > 0.5 * 8 * n
    assert (sgn > 0.5).any() and (sgn < -0.5).any() and (chi_flip > 0.0).any()      # directions were set and flipped
    assert np.abs(st[:, nsl:2 * nsl]).max() > 1e5                                   # a back stress developed


def test_dislotungsten_matches_oracle_through_first_yield():
    """dislotungsten (phase_mechanical_plastic_dislotungsten.f90; cI tungsten, parameters and non-Schmid coefficients of
    dislotungsten_W.yaml; unpinned: no fixture in the reference tree) in the reference-order kernel against the oracle from the
    elastic range through first yield (P11 about 2 GPa, where this parameter set stays integrable): P 1e-9, state and the
    dependent state (Lambda_sl, tau_pass; carried from call to call in HBM) 1e-9, equal Lp iteration counts, tangent 1e-8."""
    cells, dt = (8, 6, 5), 1.0
    g, o = _setup(cells, 9, 'W_tungsten')
    n = int(np.prod(cells))
    ns, nsl = o.phases[0].n_state, o.phases[0].n_sl
    assert ns == 3 * nsl
    nd = 2 * nsl
    atol = np.ctypeslib.as_array(o.phases[0].atol)[:ns].copy()
    dep0 = g.get_state('dep', int(g.lib.dmk_max_dep(g.ctx)))[:, :nd]
    assert _rel(dep0, o.a['dep'][:, :nd]) < 1e-13                        # crystallite_init: dependent state of the initial state
    rng = np.random.default_rng(9)
    nx, ny, nz = cells
    fluct = rng.standard_normal((nz, ny, nx, 3, 3))
    clean = np.ones(n, bool)
    compared = 0
    for k, e in enumerate([1.5e-3, 3.0e-3, 4.0e-3, 4.4e-3, 4.8e-3, 5.1e-3]):
        F = np.eye(3) + np.diag([e, -0.3 * e, -0.3 * e]) + 5.0e-5 * (k + 1) * fluct
        ill_o = o.mechanical_response(F.reshape(n, 9), dt)
        P_o, dPdF_o = o.mechanical_response2(dt, tangent=(k == 4))
        P = g.to_natural(g.constitutive_response(g.to_device(F), dt)).reshape(n, 3, 3)
        assert g.terminally_ill() == ill_o and not ill_o
        clean &= (o.point_cutbacks == 0) & (g.get_state('cutbacks', 1)[:, 0] == 0)
        same = clean & (g.get_state('lp_iter', 1)[:, 0] == o.point_lp_iter)
        assert clean.mean() > 0.5 and same.sum() >= 0.98 * clean.sum(), (k, clean.mean(), same.sum() / max(clean.sum(), 1))
        err = _pointwise_rel(P, P_o)
        assert (err[same] < RTOL).all() and (err[clean] < 1e-6).all(), (k, err[same].max(), err[clean].max())
        st, st_o = g.get_state('state', o.max_state)[:, :ns], o.a['state'][:, :ns]
        es = np.abs(st - st_o) / np.maximum(np.abs(st_o), atol)
        assert (es[same] < 10 * RTOL).all(), (k, es[same].max())
        dep = g.get_state('dep', int(g.lib.dmk_max_dep(g.ctx)))[:, :nd]
        assert (_pointwise_rel(dep, o.a['dep'][:, :nd])[same] < 10 * RTOL).all()
        if k == 4:
            from solver_cpp_b200.grid import fortran_to_c3333
            _, _, _, d = g.tangent_response(dt, materialise=True)
            dP = np.stack([fortran_to_c3333(r) for r in d.cpu().numpy()])
            et = _pointwise_rel(dP, dPdF_o)
            assert (et[same] < 1e-8).all(), (k, et[same].max())
        compared += int(same.sum())
        clean &= same
        g.materialpoint_forward()
        o.forward()
    assert compared > 0.5 * 6 * n
    assert np.abs(o.a['Lp']).max() > 1e-6 and o.a['state'][:, :nsl].max() > 10 * 1.0e9       # plastic flow, dislocations multiplied


def test_newton_decisions_rarely_flip_on_a_solver_iterate():
    """Cu polycrystal, homogeneous pre-strain, then a smooth in-grain fluctuation of the size a running solve produces
    (no cut-backs; every third cell needs the line search): per-cell Lp iteration counts equal the oracle's on all but
    <= 0.5 % of the cells, P to 1e-9 on the rest and 1e-6 on those."""
    cells = (16, 16, 8)
    g, o = _setup(cells, 12, 'Cu_pheno')
    n = int(np.prod(cells))
    nx, ny, nz = cells
    dt = 0.5
    rng = np.random.default_rng(21)
    z, y, x = np.meshgrid(np.arange(nz) / nz, np.arange(ny) / ny, np.arange(nx) / nx, indexing='ij')
    smooth = sum(rng.standard_normal((3, 3)) * np.cos(2 * np.pi * (k[0] * x + k[1] * y + k[2] * z + rng.random()))[..., None, None]
                 for k in ((1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (2, 1, 1), (1, 2, 3)))
    clean = np.ones(n, bool)                    # no cut-back anywhere in the cell's history, on either side
    e = 0.0
    for de in (1e-5, 1e-5, 2e-5, 4e-5, 8e-5) + (1.6e-4,) * 8:       # gentle ramp: virgin Cu cuts back on anything faster
        e += de
        F = np.broadcast_to(np.diag([1 + e, 1 - 0.5 * e, 1 - 0.5 * e]), (nz, ny, nx, 3, 3)).copy()
        o.mechanical_response(F.reshape(n, 9), dt)
        g.constitutive_response(g.to_device(F), dt)
        clean &= (o.point_cutbacks == 0) & (g.get_state('cutbacks', 1)[:, 0] == 0)
        g.materialpoint_forward(); o.forward()
    n_quiet = n_flipped = 0
    for amp in (1.5e-5, 3.0e-5, 4.5e-5):
        F = np.diag([1 + e + 1.6e-4, 1 - 0.5 * (e + 1.6e-4), 1 - 0.5 * (e + 1.6e-4)]) + amp * smooth
        o.mechanical_response(F.reshape(n, 9), dt)
        P_o, _ = o.mechanical_response2(dt, tangent=False)
        P = g.to_natural(g.constitutive_response(g.to_device(F), dt)).reshape(n, 3, 3)
        quiet = clean & (o.point_cutbacks == 0) & (g.get_state('cutbacks', 1)[:, 0] == 0)
        flipped = quiet & (g.get_state('lp_iter', 1)[:, 0] != o.point_lp_iter)
        err = _pointwise_rel(P, P_o)
        assert (err[quiet & ~flipped] < RTOL).all() and (err[flipped] < 1e-6).all()
        n_quiet += quiet.sum(); n_flipped += flipped.sum()
    assert n_quiet > 2.5 * n and n_flipped <= 0.005 * n_quiet, (n_quiet, n_flipped)
    assert o.point_lp_iter.mean() > 8 and g.counters()['deferred'] == 0


@pytest.mark.parametrize('integrator', ['Euler', 'AdaptiveEuler', 'RK4', 'RKCK45'])
def test_other_state_integrators_match_oracle(integrator):
    """numerics `integrator_state`: Euler / AdaptiveEuler / RK4 / RKCK45 (phase_mechanical.f90:699-932) through the C ABI:
    terminally-ill flag, per-cell Lp iteration counts and P (1e-9) on cells that do not cut back"""
    cells = (8, 6, 5)
    g, o = _setup(cells, 11, 'Al_pheno', num_phase_mech={'plastic': {'integrator_state': integrator}})
    rng = np.random.default_rng(5)
    n = int(np.prod(cells))
    n_clean = 0
    for F in _F_sequence(cells, 3, rng, rate=3.0e-3):
        ill_o = o.mechanical_response(F.reshape(n, 9), 0.5)
        P_o, _ = o.mechanical_response2(0.5, tangent=False)
        P = g.to_natural(g.constitutive_response(g.to_device(F), 0.5)).reshape(n, 3, 3)
        assert g.terminally_ill(reset=True) == ill_o
        if ill_o:
            o.terminally_ill = False
        quiet = (o.point_cutbacks == 0) & (g.get_state('cutbacks', 1)[:, 0] == 0)
        assert (g.get_state('lp_iter', 1)[quiet, 0] == o.point_lp_iter[quiet]).all()
        assert (_pointwise_rel(P, P_o)[quiet] < RTOL).all()
        n_clean += quiet.sum()
        g.materialpoint_forward(); o.forward()
    assert n_clean > n or integrator == 'AdaptiveEuler'


def test_iteration_counters_match_oracle():
    """identical inner iteration counts (Lp Newton, stress calls, state iterations, sub-steps) on a cut-back-free path"""
    g, o = _setup((8, 6, 5), 11, 'Al_pheno')
    rng = np.random.default_rng(5)
    n = 8 * 6 * 5
    tot = np.zeros(6, dtype=np.int64)
    for F in _F_sequence((8, 6, 5), 4, rng, rate=3.0e-3):
        o.mechanical_response(F.reshape(n, 9), 0.5)
        tot += o.counters
        g.constitutive_response(g.to_device(F), 0.5)
        g.materialpoint_forward(); o.forward()
    c = g.counters()
    # oracle counters: Lp iters, Li iters, stress calls, state iters, substeps, cutbacks
    assert (c['Lp_iter'], c['stress_calls'], c['state_iter'], c['substeps'], c['cutbacks']) == \
        (tot[0], tot[2], tot[3], tot[4], tot[5])


def test_g2_known_answer_through_the_c_abi(golden_dir):
    """output_damask_cpp.sh:290-303: P of all 27 cells for F = diag(1.0005,1,1), dt = 0.5 s, virgin state."""
    from solver_cpp_b200 import material as M
    g2 = np.load(os.path.join(golden_dir, 'g2_constitutive_known_answer.npz'))
    mat = yaml.safe_load(open(os.path.join(golden_dir, 'material_20grains_Al.yaml')))
    g = _ctx((3, 3, 3), (1e-5,) * 3)
    g.material_init(M.Material(mat, g2['grain_map']))
    F = torch.from_numpy(np.ascontiguousarray(g2['F'])).cuda().reshape(g.field_shape)     # already Fortran order
    P = g.constitutive_response(F, 0.5).cpu().numpy().reshape(-1)
    assert not g.terminally_ill()
    assert np.abs(P - g2['P']).max() / np.abs(g2['P']).max() < 1e-11
    P_av = P.reshape(27, 9).mean(0)
    assert np.abs(P_av - g2['P_av']).max() / np.abs(g2['P_av']).max() < 1e-11


@pytest.mark.parametrize('phase,rate,noise', [('Al_pheno', 3.0e-3, 2.0e-4), ('Al_iso', 1.0e-5, 1.0e-6), ('TWIP', 2.0e-3, 1.0e-4),
                                               ('Al_elastic', 3.0e-3, 2.0e-4)])
def test_tangent_and_homogenisation_match_oracle(phase, rate, noise):
    """dPdF per cell against the oracle (cells on a rounding-chaotic cut-back path excluded, see above); C_volAvg and the
    arg-max / arg-min selection with the shared tie rule against numpy on the device's own materialised dPdF."""
    cells = (8, 6, 5)
    g, o = _setup(cells, 11, phase)
    rng = np.random.default_rng(3)
    n = int(np.prod(cells))
    dt = 0.5
    clean = np.ones(n, bool)
    for F in _F_sequence(cells, 3, rng, rate=rate, noise=noise) + [_F_sequence(cells, 4, rng, rate=rate, noise=noise)[-1]]:
        o.mechanical_response(F.reshape(n, 9), dt)
        P = g.to_natural(g.constitutive_response(g.to_device(F), dt)).reshape(n, 3, 3)
        clean &= (o.point_cutbacks == 0) & (g.get_state('cutbacks', 1)[:, 0] == 0) & (_pointwise_rel(P, o.a['P'].reshape(n, 3, 3)) < RTOL)
        P_o, dPdF_o = o.mechanical_response2(dt, tangent=True)
        Cv, Cm, cellsel, d = g.tangent_response(dt, materialise=True)
        g.materialpoint_forward(); o.forward()
    assert clean.mean() > 0.5
    from solver_cpp_b200.grid import fortran_to_c3333
    dP = np.stack([fortran_to_c3333(r) for r in d.cpu().numpy()])
    assert _pointwise_rel(dP[clean], dPdF_o[clean]).max() < 1e-8
    grid = osp.Grid(cells, g.size)
    _, C_vol, C_mm, sel = osp.homogenise(P_o, dP, grid)              # numpy reduction of the DEVICE tangents
    assert _rel(Cv, C_vol) < 1e-12
    assert cellsel == sel
    assert _rel(Cm, C_mm) < 1e-12
    if clean.all():
        _, C_vol_o, C_mm_o, sel_o = osp.homogenise(P_o, dPdF_o, grid)
        assert cellsel == sel_o and _rel(Cv, C_vol_o) < 1e-9


def test_mass_cutback_evaluation_default_path():
    """Default path on the evaluation the reference-order kernel is slowest at: virgin soft Cu, a strain step that makes EVERY
    cell cut back.  More than 1/32 of the cells give up their first sub-step, so the list goes through the lane-scheduled
    in-lane cut-back kernel.  Such an evaluation is rounding-chaotic for any build of the reference algorithm (see
    test_constitutive_matches_oracle_over_increments): required here is what survives that -- no ill cell, every cell cut back
    on both sides, identical sub-step bookkeeping in the mean, the homogenised stress of THIS evaluation within 5e-3 of the
    oracle's (observed ~1e-3, the same as between the three device kernels and between the oracle and its FMA build), and --
    the property a solve needs -- the NEXT increment, started from either state, back inside 1e-9 per cell on the cells that
    agreed."""
    cells = (16, 16, 8)
    n = int(np.prod(cells))
    rng = np.random.default_rng(11)
    seq = _F_sequence(cells, 2, rng, rate=2.0e-3)
    g, o = _setup(cells, 20, 'Cu_pheno', mass_deferral=-1)
    ill_o = o.mechanical_response(seq[0].reshape(n, 9), 0.5)
    P_o, _ = o.mechanical_response2(0.5, tangent=False)
    c0 = g.counters()
    P = g.to_natural(g.constitutive_response(g.to_device(seq[0]), 0.5)).reshape(n, 3, 3)
    c1 = g.counters()
    assert not ill_o and not g.terminally_ill()
    assert (o.point_cutbacks > 0).all() and (g.get_state('cutbacks', 1)[:, 0] > 0).all()
    assert c1['deferred'] - c0['deferred'] == n                                  # every cell gave up its full first sub-step
    assert _rel(P.mean(0), P_o.mean(0)) < 5e-3
    err = _pointwise_rel(P, P_o)
    assert np.median(err) < 1e-2 and err.max() < 0.5
    same = err < RTOL
    g.materialpoint_forward(); o.forward()
    o.mechanical_response(seq[1].reshape(n, 9), 0.5)
    P_o, _ = o.mechanical_response2(0.5, tangent=False)
    P = g.to_natural(g.constitutive_response(g.to_device(seq[1]), 0.5)).reshape(n, 3, 3)
    quiet = (o.point_cutbacks == 0) & (g.get_state('cutbacks', 1)[:, 0] == 0) & (g.get_state('lp_iter', 1)[:, 0] == o.point_lp_iter)
    if (same & quiet).any():
        assert _pointwise_rel(P, P_o)[same & quiet].max() < RTOL
    assert _rel(P.mean(0), P_o.mean(0)) < 5e-3


def test_g3_homogenisation_literal_through_the_device_reductions(golden_dir):
    """the literal two-cell P / dPdF fields and expected P_av, C_volAvg, C_minMaxAvg of the reference's own unit test
    (test_spectral.cpp:58-216, golden set G3) through the DEVICE reductions (dmk_homogenise: record sum, dPdF sum, norm,
    arg-max / arg-min selection with the tie rule)"""
    import os
    g3 = np.load(os.path.join(golden_dir, 'g3_unit_test_literals.npz'))
    g = _ctx((2, 1, 1), (2e-5, 1e-5, 1e-5))
    P = g.to_device(np.transpose(g3['hom_P'], (2, 0, 1)).reshape(1, 1, 2, 3, 3))
    # (i, j, k, l, cell) -> [cell][81] with the first index fastest
    d = np.ascontiguousarray(np.transpose(g3['hom_dPdF'], (4, 3, 2, 1, 0))).reshape(2, 81)
    P_av, C_vol, C_mm, sel = g.homogenise(P, torch.from_numpy(d).cuda())
    for got, key in ((P_av, 'hom_P_av'), (C_vol, 'hom_C_vol_avg'), (C_mm, 'hom_C_minmax_avg')):
        assert np.abs(got - g3[key]).max() <= 1e-12 * np.abs(g3[key]).max(), key
    assert sorted(sel) == [0, 1]


def test_homogenise_fields_matches_numpy():
    """dmk_homogenise on a random field against the oracle's numpy reduction (ties: a constant-norm block)"""
    cells = (6, 7, 8)
    n = int(np.prod(cells))
    g = _ctx(cells, (6e-5, 7e-5, 8e-5))
    rng = np.random.default_rng(2)
    P = rng.standard_normal((n, 3, 3)) * 1e8
    dP = rng.standard_normal((n, 3, 3, 3, 3)) * 1e9 + 1e11
    dP[40:50] = dP[40]                                                        # ties at neither extreme
    big = np.argmax((dP.reshape(n, -1) ** 2).sum(1))
    dP[300] = dP[big]                                                         # a tie at the maximum: the lower index wins
    d = np.ascontiguousarray(np.transpose(dP, (0, 4, 3, 2, 1))).reshape(n, 81)
    P_av, C_vol, C_mm, sel = g.homogenise(g.to_device(P.reshape(8, 7, 6, 3, 3)), torch.from_numpy(d).cuda())
    P_av_o, C_vol_o, C_mm_o, sel_o = osp.homogenise(P, dP, osp.Grid(cells, g.size))
    assert sel == tuple(sel_o) and sel[0] == min(big, 300)
    assert _rel(P_av, P_av_o) < 1e-13 and _rel(C_vol, C_vol_o) < 1e-13 and _rel(C_mm, C_mm_o) < 1e-15


def test_forward_field_removes_the_mean():
    """MechBase::forward_field with dt = 0 and a zero aim: the field minus its mean (device record sum against numpy)"""
    cells = (6, 7, 8)
    g = _ctx(cells, (6e-5, 7e-5, 8e-5))
    rng = np.random.default_rng(9)
    P = rng.standard_normal((8, 7, 6, 3, 3)) * 1e8
    out = g.forward_field(0.0, g.to_device(P), g.to_device(np.zeros_like(P)), np.zeros((3, 3)))
    ref = P - P.mean(axis=(0, 1, 2))
    assert np.abs(g.to_natural(out) - ref).max() < 1e-6


def test_point_state_restart_round_trip(tmp_path):
    """save_point_state / load_point_state (dmk_get_state / dmk_set_state): a fresh context loaded with the saved per-point
    state continues bit-identically (restart of the constitutive half, phase_mechanical.f90:1264-1298)"""
    from solver_cpp_b200 import workloads as W, material as M
    cells = (8, 6, 5)
    material, material_at, size = W.polycrystal(cells, 11, 'Al_pheno', 0)
    rng = np.random.default_rng(5)
    seq = _F_sequence(cells, 3, rng, rate=3.0e-3)
    g = _ctx(cells, size)
    g.material_init(M.Material(material, material_at))
    for F in seq[:2]:
        g.constitutive_response(g.to_device(F), 0.5)
        g.materialpoint_forward()
    path = str(tmp_path / 'restart.npz')
    g.save_point_state(path)
    P_a = g.constitutive_response(g.to_device(seq[2]), 0.5).cpu().numpy().copy()
    g2 = _ctx(cells, size)
    g2.material_init(M.Material(material, material_at))
    g2.load_point_state(path)
    P_b = g2.constitutive_response(g2.to_device(seq[2]), 0.5).cpu().numpy()
    assert np.array_equal(P_a, P_b)
    assert np.array_equal(g.get_state('state', 24), g2.get_state('state', 24))


@pytest.mark.gpu
def test_fast_reciprocal_and_sqrt_are_ieee():
    """the per-point kernel's branch-free reciprocal / square root (pheno_fast_dev.cuh: the compiler's own fast-path sequences
    without the inline slow-path call) return the correctly rounded results of 1.0 / x and sqrt(x), bit for bit, over their
    whole admitted range -- so using them changes no result"""
    import ctypes as C
    from solver_cpp_b200 import lib as L
    lib = L.load()
    for seed, lim in ((1, 900), (2, 40), (3, 2)):
        br, bs = C.c_longlong(-1), C.c_longlong(-1)
        L.check(lib.dmk_selftest_special(C.c_longlong(200_000_000), C.c_ulonglong(seed), C.c_int(lim), C.byref(br), C.byref(bs)))
        assert (br.value, bs.value) == (0, 0), (seed, lim, br.value, bs.value)
