"""CPU tests: the product's per-point device functions (csrc/constitutive_dev.cuh, tangent_dev.cuh) compiled for the
host by tests/host_twin (test infrastructure) and compared with the oracle.  This is the same source the CUDA kernels
instantiate, so law / tangent logic errors show up here, on the GPU-less build box, before the -m gpu parity run."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'host_twin'))

from oracle.constitutive import ConstitutiveOracle   # noqa: E402

V6 = [0, 4, 8, 5, 2, 1]


def _path(cells, n_inc, rng, rate):
    nx, ny, nz = cells
    fluct = rng.standard_normal((nz * ny * nx, 3, 3))
    out = []
    for inc in range(n_inc):
        Fbar = np.eye(3) + np.diag([rate * 0.5 * (inc + 1), -0.3 * rate * 0.5 * (inc + 1), -0.3 * rate * 0.5 * (inc + 1)])
        out.append(Fbar + 2.0e-4 * (inc + 1) * fluct)
    return out


@pytest.mark.parametrize('phase,rate', [('Al_pheno', 3.0e-3), ('Al_iso', 2.0e-3), ('TWIP', 7.0e-3), ('Al_elastic', 2.0e-3)])
def test_twin_matches_oracle(phase, rate):
    import twin
    from solver_cpp_b200 import workloads as W, material as M
    lib = twin.build()
    cells = (4, 4, 3)
    material, material_at, size = W.polycrystal(cells, 7, phase, 0)
    o = ConstitutiveOracle(material, material_at)
    m = M.Material(material, material_at)
    num = M.make_numerics()
    n = int(np.prod(cells))
    ns = o.phases[0].n_state
    nd = max(1, 2 * o.phases[0].n_sl + o.phases[0].n_tw)
    dep = o.a['dep'][:, :nd].copy()
    Lp = np.zeros((n, 9))
    rng = np.random.default_rng(2)
    worst = 0.0
    for F in _path(cells, 4, rng, rate):
        snap = o.snapshot()
        o.mechanical_response(F.reshape(n, 9), 0.5)
        P_o, dPdF_o = o.mechanical_response2(0.5, tangent=True)
        for i in range(n):
            r = twin.run_point(lib, m.phases[0], num, snap['F0'][i], F[i].reshape(9), snap['Fp0'][i], snap['Lp0'][i], Lp[i],
                               snap['S0'][i][V6], snap['state0'][i][:max(ns, 1)], dep[i], 0.5, tangent=True)
            Lp[i] = r['Lp']; dep[i] = r['dep']
            if o.point_cutbacks[i]:
                continue                                   # rounding-chaotic path, see test_oracle_golden.py
            e = np.abs(r['P'] - o.a['P'][i]).max() / np.abs(o.a['P'][i]).max()
            e = max(e, np.abs(r['Fp'] - o.a['Fp'][i]).max())
            if ns:
                atol = np.ctypeslib.as_array(o.phases[0].atol)[:ns]            # the integrator's own absolute state tolerance
                e = max(e, (np.abs(r['state'][:ns] - o.a['state'][i, :ns]) / np.maximum(np.abs(o.a['state'][i, :ns]), atol)).max())
            et = np.abs(r['dPdF'] - dPdF_o[i].reshape(81)).max() / np.abs(dPdF_o[i]).max()
            assert e < 1e-9 and et < 1e-8, (phase, i, e, et)
            worst = max(worst, e)
        o.forward()
    if phase != 'Al_elastic':
        assert np.abs(o.a['Lp']).max() > 1e-6


def test_state_machine_is_bit_identical_to_nested_loops():
    """machine_dev.cuh (what the kernel runs) against the straight nested-loop formulation of constitutive_dev.cuh:
    same arithmetic in the same order -> identical bits, including cells that cut back."""
    import twin
    from solver_cpp_b200 import workloads as W, material as M
    lib = twin.build()
    for phase, rate in (('Cu_pheno', 2.0e-3), ('Al_iso', 2.0e-3), ('TWIP', 7.0e-3)):
        cells = (4, 4, 2)
        material, material_at, size = W.polycrystal(cells, 5, phase, 0)
        o = ConstitutiveOracle(material, material_at)
        m = M.Material(material, material_at)
        num = M.make_numerics()
        n = int(np.prod(cells))
        ns = max(1, o.phases[0].n_state)
        nd = max(1, 2 * o.phases[0].n_sl + o.phases[0].n_tw)
        rng = np.random.default_rng(4)
        F = _path(cells, 1, rng, rate)[0]
        cut = 0
        for i in range(n):
            args = (lib, m.phases[0], num, o.a['F0'][i], F[i].reshape(9), o.a['Fp0'][i], o.a['Lp0'][i], o.a['Lp'][i],
                    o.a['S0'][i][V6], o.a['state0'][i][:ns], o.a['dep'][i][:nd], 0.5)
            lib.twin_use_nested(1)
            a = twin.run_point(*args)
            lib.twin_use_nested(0)
            b = twin.run_point(*args)
            for k in ('P', 'Fp', 'Fe', 'S', 'state', 'Lp', 'dep', 'cnt'):
                assert np.array_equal(a[k], b[k]), (phase, i, k)
            assert a['ok'] == b['ok']
            cut += int(a['cnt'][5])
        if phase != 'TWIP':
            assert cut > 0


@pytest.mark.parametrize('phase,rate', [('Al_pheno', 3.0e-3), ('Cu_pheno', 2.0e-3)])
def test_register_resident_pheno_matches_oracle(phase, rate):
    """pheno_fast_dev.cuh (6x6 Woodbury Newton system, constant-bank tables; what the GPU runs for slip-only
    phenopowerlaw phases) against the oracle over a multi-increment path: P, Fp, state to 1e-9, identical inner
    iteration counters on cells that do not cut back."""
    import twin
    from solver_cpp_b200 import workloads as W, material as M
    lib = twin.build()
    cells = (4, 4, 3)
    material, material_at, size = W.polycrystal(cells, 7, phase, 0)
    o = ConstitutiveOracle(material, material_at)
    m = M.Material(material, material_at)
    num = M.make_numerics()
    n = int(np.prod(cells))
    ns = o.phases[0].n_state
    Lp = np.zeros((n, 9))
    rng = np.random.default_rng(2)
    lib.twin_use_fast(1)
    try:
        n_clean = n_flipped = 0
        for F in _path(cells, 4, rng, rate):
            snap = o.snapshot()
            o.mechanical_response(F.reshape(n, 9), 0.5)
            _, dPdF_o = o.mechanical_response2(0.5, tangent=True)
            cnt = np.zeros(6, dtype=np.int64)
            for i in range(n):
                r = twin.run_point(lib, m.phases[0], num, snap['F0'][i], F[i].reshape(9), snap['Fp0'][i], snap['Lp0'][i], Lp[i],
                                   snap['S0'][i][V6], snap['state0'][i][:ns], np.zeros(1), 0.5, tangent=True)
                Lp[i] = r['Lp']
                cnt += r['cnt']
                assert r['ok']
                if o.point_cutbacks[i] or r['cnt'][5]:
                    continue
                # a stopping decision of the Newton loop (residual within rounding of its tolerance) may flip: such a
                # cell ends one iteration apart, both answers inside the integrator's tolerance (1e-6); it must be rare
                flipped = int(r['cnt'][0]) != int(o.point_lp_iter[i])
                n_flipped += flipped
                n_clean += 1
                e = np.abs(r['P'] - o.a['P'][i]).max() / np.abs(o.a['P'][i]).max()
                e = max(e, np.abs(r['Fp'] - o.a['Fp'][i]).max(), np.abs(r['Fe'] - o.a['Fe'][i]).max())
                atol = np.ctypeslib.as_array(o.phases[0].atol)[:ns]
                e = max(e, (np.abs(r['state'][:ns] - o.a['state'][i, :ns]) / np.maximum(np.abs(o.a['state'][i, :ns]), atol)).max())
                et = np.abs(r['dPdF'] - dPdF_o[i].reshape(81)).max() / np.abs(dPdF_o[i]).max()
                assert e < (1e-6 if flipped else 1e-9) and et < (1e-5 if flipped else 1e-8), (phase, i, e, et, flipped)
            if not o.point_cutbacks.any():
                assert tuple(cnt[[2, 3, 4, 5]]) == tuple(o.counters[[2, 3, 4, 5]])
                assert abs(int(cnt[0]) - int(o.counters[0])) <= n_flipped
            o.forward()
        assert n_clean > 2 * n and n_flipped <= 0.01 * n_clean
        assert np.abs(o.a['Lp']).max() > 1e-6
    finally:
        lib.twin_use_fast(0)


def test_lane_schedule_is_bit_identical_to_nested_fast_formulation():
    """pheno_lanes_dev.cuh (the resumable pieces the lane-persistent kernel runs: sub-step set-up with cut-backs, state block,
    Newton residual / step with line search) against pf_point without deferral, the one-thread-walks-one-point formulation
    of the same arithmetic: identical bits and counters on every point, cut-back points included, over a multi-increment
    path of soft Cu (virgin Cu cuts back in every cell)."""
    import twin
    from solver_cpp_b200 import workloads as W, material as M
    lib = twin.build()
    cells = (4, 4, 3)
    material, material_at, size = W.polycrystal(cells, 7, 'Cu_pheno', 0)
    o = ConstitutiveOracle(material, material_at)
    m = M.Material(material, material_at)
    num = M.make_numerics()
    n = int(np.prod(cells))
    ns = o.phases[0].n_state
    rng = np.random.default_rng(5)
    lib.twin_set_defer(0)
    n_cut = n_plain = 0
    try:
        for F in _path(cells, 4, rng, 2.0e-3):
            snap = o.snapshot()
            for i in range(n):
                args = (lib, m.phases[0], num, snap['F0'][i], F[i].reshape(9), snap['Fp0'][i], snap['Lp0'][i], snap['Lp'][i],
                        snap['S0'][i][V6], snap['state0'][i][:ns], np.zeros(1), 0.5)
                lib.twin_use_fast(1)
                a = twin.run_point(*args)
                lib.twin_use_fast(2)
                b = twin.run_point(*args)
                assert a['ok'] == b['ok'], i
                assert np.array_equal(a['cnt'], b['cnt']), (i, a['cnt'], b['cnt'])
                if not a['ok']:
                    continue                           # terminally ill: the evaluation is discarded by the solver
                n_cut += int(a['cnt'][5] > 0)
                n_plain += int(a['cnt'][5] == 0)
                for k in ('P', 'Fp', 'Fe', 'S', 'state', 'Lp'):
                    assert np.array_equal(a[k], b[k]), (i, k, a['cnt'])
                # the default variant (CUTBACK = false) gives up exactly the points that cut back, and nothing else
                lib.twin_set_defer(8)
                c = twin.run_point(*args)
                gave_up = bool(lib.twin_last_deferred())
                lib.twin_set_defer(0)
                assert gave_up == (a['cnt'][5] > 0), i
                if not gave_up:
                    for k in ('P', 'Fp', 'Fe', 'S', 'state', 'Lp', 'cnt'):
                        assert np.array_equal(a[k], c[k]), (i, k)
            o.mechanical_response(F.reshape(n, 9), 0.5)
            o.forward()
    finally:
        lib.twin_use_fast(0)
    assert n_plain > n and n_cut > n // 2


@pytest.mark.parametrize('integrator', ['Euler', 'AdaptiveEuler', 'RK4', 'RKCK45'])
@pytest.mark.parametrize('phase,rate', [('Al_pheno', 3.0e-3), ('Al_iso', 2.0e-3)])
def test_other_state_integrators_match_oracle(integrator, phase, rate):
    """integrateStateEuler / AdaptiveEuler / RK4 / RKCK45 (phase_mechanical.f90:699-932) on the device code (host twin)
    against the oracle: same convergence verdict per point, P / Fp / state to 1e-9 on points that do not cut back."""
    import twin
    from solver_cpp_b200 import workloads as W, material as M
    lib = twin.build()
    cells = (4, 4, 3)
    material, material_at, size = W.polycrystal(cells, 7, phase, 0)
    nm = {'plastic': {'integrator_state': integrator}}
    o = ConstitutiveOracle(material, material_at, num_phase_mech=nm)
    m = M.Material(material, material_at)
    num = M.make_numerics(nm)
    n = int(np.prod(cells))
    ns = o.phases[0].n_state
    Lp = np.zeros((n, 9))
    rng = np.random.default_rng(2)
    n_clean = 0
    for F in _path(cells, 3, rng, rate):
        snap = o.snapshot()
        o.mechanical_response(F.reshape(n, 9), 0.5)
        for i in range(n):
            r = twin.run_point(lib, m.phases[0], num, snap['F0'][i], F[i].reshape(9), snap['Fp0'][i], snap['Lp0'][i], Lp[i],
                               snap['S0'][i][V6], snap['state0'][i][:max(ns, 1)], np.zeros(1), 0.5)
            Lp[i] = r['Lp']
            if o.point_cutbacks[i] or r['cnt'][5]:
                continue
            n_clean += 1
            assert int(r['cnt'][0]) == int(o.point_lp_iter[i])
            e = np.abs(r['P'] - o.a['P'][i]).max() / np.abs(o.a['P'][i]).max()
            e = max(e, np.abs(r['Fp'] - o.a['Fp'][i]).max())
            atol = np.ctypeslib.as_array(o.phases[0].atol)[:ns]
            e = max(e, (np.abs(r['state'][:ns] - o.a['state'][i, :ns]) / np.maximum(np.abs(o.a['state'][i, :ns]), atol)).max())
            assert e < 1e-9, (integrator, phase, i, e)
        o.forward()
    if phase == 'Al_pheno' and integrator != 'AdaptiveEuler':      # (soft isotropic Al cuts back on most points with the explicit schemes)
        assert n_clean > n


def test_twin_matches_oracle_kinehardening_with_load_reversal():
    """kinehardening (phase_mechanical_plastic_kinehardening.f90; parameters of kinehardening_317L.yaml) through the product's
    device functions against the oracle on a tension -> unloading -> compression path: the shear direction of the active
    systems flips on the way back, so the deltaState jump (sgn_gamma, chi_flip, gamma_flip) is exercised, not just dotState.
    The law has no fixture in the reference tree: oracle-vs-product parity only (unpinned)."""
    import twin
    from solver_cpp_b200 import workloads as W, material as M
    lib = twin.build()
    cells = (4, 3, 3)
    material, material_at, size = W.polycrystal(cells, 6, 'Steel_kine', 0)
    o = ConstitutiveOracle(material, material_at)
    m = M.Material(material, material_at)
    num = M.make_numerics()
    n = int(np.prod(cells))
    ns = o.phases[0].n_state
    nsl = o.phases[0].n_sl
    assert ns == 6 * nsl and m.phases[0].n_state == ns
    rng = np.random.default_rng(5)
    fluct = rng.standard_normal((n, 3, 3))
    Lp = np.zeros((n, 9))
    dep = np.zeros((n, 1))
    atol = np.ctypeslib.as_array(o.phases[0].atol)[:ns].copy()
    atol[atol == 0.0] = 1e-300                            # the deltaState part has no tolerance: exact zeros must stay exact zeros
    compared = 0
    for k, e in enumerate([1.5e-3, 3.0e-3, 4.5e-3, 3.0e-3, 1.5e-3, 0.0, -1.5e-3, -3.0e-3]):
        F = np.eye(3) + np.diag([e, -0.3 * e, -0.3 * e]) + 1.0e-4 * (k + 1) * fluct
        snap = o.snapshot()
        assert not o.mechanical_response(F.reshape(n, 9), 0.5)
        P_o, dPdF_o = o.mechanical_response2(0.5, tangent=True)
        for i in range(n):
            r = twin.run_point(lib, m.phases[0], num, snap['F0'][i], F[i].reshape(9), snap['Fp0'][i], snap['Lp0'][i], Lp[i],
                               snap['S0'][i][V6], snap['state0'][i][:ns], dep[i], 0.5, tangent=True)
            Lp[i] = r['Lp']
            assert r['ok']
            if o.point_cutbacks[i]:
                continue
            err = np.abs(r['P'] - o.a['P'][i]).max() / np.abs(o.a['P'][i]).max()
            err = max(err, np.abs(r['Fp'] - o.a['Fp'][i]).max())
            # state: 1e-9 relative, or -- chi runs through zero on the way back -- a thousandth of the integrator's own absolute tolerance
            ds = np.abs(r['state'][:ns] - o.a['state'][i, :ns])
            ok_state = (ds <= 1e-9 * np.maximum(np.abs(o.a['state'][i, :ns]), atol)) | (ds <= 1e-3 * atol)
            et = np.abs(r['dPdF'] - dPdF_o[i].reshape(81)).max() / np.abs(dPdF_o[i]).max()
            assert err < 1e-9 and ok_state.all() and et < 1e-8, (k, i, err, ds.max(), et)
            compared += 1
        o.forward()
    st = o.a['state'][:, :ns]
    sgn, chi_flip = st[:, 3 * nsl:4 * nsl], st[:, 4 * nsl:5 * nsl]
    assert compared > 0.8 * 8 * n
    assert (sgn > 0.5).any() and (sgn < -0.5).any() and (chi_flip > 0.0).any()      # directions were set and flipped
    assert np.abs(st[:, nsl:2 * nsl]).max() > 1e5                                   # a back stress developed


def test_twin_matches_oracle_dislotungsten():
    """dislotungsten (phase_mechanical_plastic_dislotungsten.f90; cI tungsten with the non-Schmid coefficients of
    dislotungsten_W.yaml) through the product's device functions against the oracle, from the elastic range through first yield
    (P11 about 2 GPa): kinetics with kink-pair nucleation and phonon drag, dislocation multiplication, the dependent state
    (Lambda_sl, tau_pass) carried from call to call, and the tangent.  No fixture in the reference tree: unpinned."""
    import twin
    from solver_cpp_b200 import workloads as W, material as M
    lib = twin.build()
    cells = (4, 3, 3)
    material, material_at, size = W.polycrystal(cells, 6, 'W_tungsten', 0)
    o = ConstitutiveOracle(material, material_at)
    m = M.Material(material, material_at)
    num = M.make_numerics()
    n = int(np.prod(cells))
    ns, nsl = o.phases[0].n_state, o.phases[0].n_sl
    assert ns == 3 * nsl and m.phases[0].n_state == ns
    nd = 2 * nsl
    dep = o.a['dep'][:, :nd].copy()
    assert (dep[:, :nsl] > 0.0).all() and (dep[:, nsl:] > 0.0).all()          # crystallite_init: dependent state of the initial state
    rng = np.random.default_rng(9)
    fluct = rng.standard_normal((n, 3, 3))
    Lp = np.zeros((n, 9))
    atol = np.ctypeslib.as_array(o.phases[0].atol)[:ns].copy()
    compared = 0
    for k, e in enumerate([1.5e-3, 3.0e-3, 4.0e-3, 4.4e-3, 4.8e-3, 5.1e-3, 5.4e-3]):   # (the W parameter set turns ill soon after yield)
        F = np.eye(3) + np.diag([e, -0.3 * e, -0.3 * e]) + 5.0e-5 * (k + 1) * fluct
        snap = o.snapshot()
        assert not o.mechanical_response(F.reshape(n, 9), 1.0)
        P_o, dPdF_o = o.mechanical_response2(1.0, tangent=True)
        for i in range(n):
            r = twin.run_point(lib, m.phases[0], num, snap['F0'][i], F[i].reshape(9), snap['Fp0'][i], snap['Lp0'][i], Lp[i],
                               snap['S0'][i][V6], snap['state0'][i][:ns], dep[i], 1.0, tangent=True)
            Lp[i] = r['Lp']; dep[i] = r['dep']
            assert r['ok']
            if o.point_cutbacks[i]:
                continue
            err = np.abs(r['P'] - o.a['P'][i]).max() / np.abs(o.a['P'][i]).max()
            err = max(err, np.abs(r['Fp'] - o.a['Fp'][i]).max())
            err = max(err, (np.abs(r['state'][:ns] - o.a['state'][i, :ns]) / np.maximum(np.abs(o.a['state'][i, :ns]), atol)).max())
            err = max(err, (np.abs(r['dep'][:nd] - o.a['dep'][i, :nd]) / np.abs(o.a['dep'][i, :nd])).max())
            et = np.abs(r['dPdF'] - dPdF_o[i].reshape(81)).max() / np.abs(dPdF_o[i]).max()
            assert err < 1e-9 and et < 1e-8, (k, i, err, et)
            compared += 1
        o.forward()
    assert compared > 0.7 * 7 * n
    assert np.abs(o.a['Lp']).max() > 1e-6 and o.a['state'][:, :nsl].max() > 10 * 1.0e9       # plastic flow, dislocations multiplied
