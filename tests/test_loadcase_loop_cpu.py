"""The increment loop's bookkeeping (mech.loop_over_loadcase: DamaskGrid::loop_over_loadcase, damask_grid.cpp:82-196 with the
upstream cut-back rules) driven by a scripted stand-in for the device solver: time stepping, cut-back and
recovery, when the material-point state is forwarded, which increments a load step's f_out selects for output.  No GPU."""
import numpy as np
import pytest

from solver_cpp_b200 import lib as L
from solver_cpp_b200 import mech

LOAD = {'loadstep': [
    {'boundary_conditions': {'mechanical': {'dot_F': [[1e-3, 0, 0], [0, 'x', 0], [0, 0, 'x']],
                                            'P': [['x', 'x', 'x'], ['x', 0, 'x'], ['x', 'x', 0]]}},
     'discretization': {'t': 4.0, 'N': 4}, 'f_out': 2},
    {'boundary_conditions': {'mechanical': {'dot_F': [[0, 0, 0], [0, 'x', 0], [0, 0, 'x']],
                                            'P': [['x', 'x', 'x'], ['x', 0, 'x'], ['x', 'x', 0]]}},
     'discretization': {'t': 3.0, 'N': 3}, 'f_out': 3, 'estimate_rate': False},
]}


class _Grid:
    def __init__(self):
        self.forwards = 0

    def materialpoint_forward(self):
        self.forwards += 1


class _Scripted:
    """answers calculate_solution() from a list of outcomes (True: converged)"""

    def __init__(self, outcomes):
        self.g = _Grid()
        self.outcomes = list(outcomes)
        self.calls = []

    def forward(self, cut_back, guess, dt, dt_prev, t_remaining, deformation, stress, R):
        self.calls.append(dict(cut_back=cut_back, guess=guess, dt=dt, dt_prev=dt_prev, t_remaining=t_remaining,
                               kind=deformation.kind if hasattr(deformation, 'kind') else None))

    def calculate_solution(self):
        ok = self.outcomes.pop(0) if self.outcomes else True
        return dict(converged=ok, stag_converged=True, terminally_ill=False, required_iterations=3 if ok else 40)

    def get33(self, name):
        return np.eye(3)

    def get(self, name, n=1):
        return np.arange(float(n))


def test_plain_run_times_and_output_selection():
    s = _Scripted([])
    seen = []
    recs = mech.loop_over_loadcase(s, mech.parse_load(LOAD), callback=lambda r, _: seen.append(r['inc']))
    assert [r['inc'] for r in recs] == [1, 2, 3, 4, 5, 6, 7] == seen
    assert np.allclose([r['t'] for r in recs], [1, 2, 3, 4, 5, 6, 7])
    assert [(r['step'], r['inc_in_step']) for r in recs] == [(0, 1), (0, 2), (0, 3), (0, 4), (1, 1), (1, 2), (1, 3)]
    # the reference's Fortran rule mod(inc, f_out) == 0 on the step's own 1-based counter: increments 2, 4 and 7
    assert [r['inc'] for r in recs if r['inc_in_step'] % r['f_out'] == 0] == [2, 4, 7]
    assert s.g.forwards == 7 and not any(c['cut_back'] for c in s.calls)
    assert [c['guess'] for c in s.calls] == [False, True, True, True, False, True, True]     # step 2: estimate_rate false
    assert np.allclose([c['t_remaining'] for c in s.calls], [4, 3, 2, 1, 3, 2, 1])
    assert recs[0]['F_av'].shape == (3, 3) and recs[0]['iterations'] == 3


def test_cut_back_halves_the_step_and_recovers():
    s = _Scripted([True, False, True, True])                   # increment 2 fails once
    recs = mech.loop_over_loadcase(s, mech.parse_load(LOAD), max_increments=4)
    assert np.allclose([r['t'] for r in recs], [1, 2, 3, 4])
    assert np.allclose([c['dt'] for c in s.calls], [1, 1, .5, .5, 1, 1])
    assert [c['cut_back'] for c in s.calls] == [False, False, True, False, False, False]
    assert s.g.forwards == 5                                   # not forwarded when the step is repeated after a cut-back
    assert np.allclose([c['t_remaining'] for c in s.calls], [4, 3, 3, 2.5, 2, 1])


def test_nested_cut_backs_and_giving_up():
    s = _Scripted([False, False, True, True, True, True])     # increment 1 fails twice: quarter steps, four of them
    recs = mech.loop_over_loadcase(s, mech.parse_load(LOAD), max_increments=2)
    assert np.allclose([c['dt'] for c in s.calls[:6]], [1, .5, .25, .25, .25, .25])
    assert np.allclose(recs[0]['t'], 1.0)
    assert np.allclose(s.calls[6]['dt'], 0.5) and np.allclose(recs[1]['t'], 2.0)      # one level is kept for the next increment
    with pytest.raises(L.DmkError):
        mech.loop_over_loadcase(_Scripted([False] * 8), mech.parse_load(LOAD), max_increments=1)


def test_reference_loop_test_one_increment_and_no_cut_back_allowed():
    """test_damask_grid.cpp:74-128 (TestDamaskGridLoopOverLoadcase): one load step with t = 1, N = 1 and max_cut_back = 0.
    A converged solve forwards the fields once and reports the result exactly once, for increment 1 at t = 1
    (EXPECT_CALL materialpoint_result(Eq(1), Eq(1))); a failed solve raises (EXPECT_THROW std::runtime_error)."""
    load = {'loadstep': [{'boundary_conditions': {'mechanical': {'dot_F': [[1e-3, 0, 0], [0, 'x', 0], [0, 0, 'x']],
                                                                  'P': [['x', 'x', 'x'], ['x', 0, 'x'], ['x', 'x', 0]],
                                                                  'R': [1.0, 0.0, 0.0, 0.0]}},
                          'discretization': {'t': 1, 'N': 1}}]}
    steps = mech.parse_load(load)
    assert np.array_equal(steps[0]['deformation'].mask.astype(bool), [[False, False, False], [False, True, False], [False, False, True]])
    assert np.array_equal(steps[0]['R'], np.eye(3))
    s = _Scripted([True])
    results = []
    recs = mech.loop_over_loadcase(s, steps, max_cut_back=0, callback=lambda r, _: results.append((r['inc'], r['t'])))
    assert results == [(1, 1.0)] and len(recs) == 1
    assert len(s.calls) == 1 and s.calls[0]['dt'] == 1.0 and not s.calls[0]['cut_back'] and not s.calls[0]['guess']
    s = _Scripted([False])
    results = []
    with pytest.raises(L.DmkError):
        mech.loop_over_loadcase(s, steps, max_cut_back=0, callback=lambda r, _: results.append(r['inc']))
    assert results == [] and len(s.calls) == 1


# ------------------------------------------------------------------------------------------------ the library's C++ twin
# dmk_load_from_yaml / dmk_loop_over_loadcase (csrc/loadcase.cpp: Config::parse_load, config.cpp:247-358, and
# DamaskGrid::loop_over_loadcase, damask_grid.cpp:82-196, compiled like the originals) against the Python host on the same decks
# and the same scripted solver outcomes
def _c_steps(load, tmp_path):
    import ctypes as C
    import yaml
    p = tmp_path / 'load.yaml'
    p.write_text(yaml.safe_dump(load, sort_keys=False))
    lib = L.load()
    steps = (L.LoadStep * 8)()
    n = C.c_int()
    name = C.create_string_buffer(32)
    rc = lib.dmk_load_from_yaml(str(p).encode(), steps, C.c_int(8), C.byref(n), name, C.c_int(32))
    assert rc == 0, lib.dmk_last_error()
    return steps, n.value, name.value.decode()


@pytest.mark.parametrize('deck', ['two_steps', 'tension_x', 'mixed_shear', 'reference_parser_test'])
def test_cpp_load_reader_equals_python_host(deck, tmp_path):
    from solver_cpp_b200 import workloads as W
    load = {'two_steps': dict(LOAD, solver={'mechanical': 'spectral_polarisation'}), 'tension_x': W.TENSION_X, 'mixed_shear': W.MIXED_SHEAR,
            'reference_parser_test': {'solver': {'mechanical': 'spectral_basic'}, 'loadstep': [
                {'boundary_conditions': {'mechanical': {'dot_F': [[1.0e-3, 1, 0], [0, 'x', 0], [2, 0, 'x']],
                                                        'P': [['x', 'x', 'x'], ['x', 0, 'x'], ['x', 'x', 0]], 'R': [1, 2, 3, 4]}},
                 'discretization': {'t': 60, 'N': 120, 'r': 2}, 'estimate_rate': False, 'f_out': 4, 'f_restart': 10}]}}[deck]
    steps, n, solver = _c_steps(load, tmp_path)
    ref = mech.parse_load(load)
    assert n == len(ref) and solver == (load.get('solver') or {}).get('mechanical', 'spectral_basic')
    for c, p in zip(steps, ref):
        assert c.deformation_type.decode() == p['deformation'].type and c.stress_type.decode() == p['stress'].type
        assert np.array_equal(np.array(c.deformation_values).reshape(3, 3), p['deformation'].values)
        assert np.array_equal(np.array(c.deformation_mask).reshape(3, 3).astype(bool), p['deformation'].mask.astype(bool))
        assert np.array_equal(np.array(c.stress_values).reshape(3, 3), p['stress'].values)
        assert np.array_equal(np.array(c.stress_mask).reshape(3, 3).astype(bool), p['stress'].mask.astype(bool))
        assert np.array_equal(np.array(c.R).reshape(3, 3), p['R'])
        assert (c.t, c.r, c.N, bool(c.estimate_rate), c.f_out, c.f_restart) == (p['t'], p['r'], p['N'], p['estimate_rate'], p['f_out'], p['f_restart'])


def test_cpp_load_reader_refusals(tmp_path):
    import ctypes as C
    import yaml
    lib = L.load()
    steps = (L.LoadStep * 2)()
    n = C.c_int()
    good = LOAD['loadstep'][0]

    def rc_of(load):
        p = tmp_path / 'bad.yaml'
        p.write_text(yaml.safe_dump(load, sort_keys=False))
        return lib.dmk_load_from_yaml(str(p).encode(), steps, C.c_int(2), C.byref(n), None, C.c_int(0))
    assert rc_of({'loadstep': [good]}) == 0
    assert rc_of({'solver': {'mechanical': 'spectral_basic'}}) != 0 and b'loadstep' in lib.dmk_last_error()
    both = {'boundary_conditions': {'mechanical': dict(good['boundary_conditions']['mechanical'], F=[[1, 0, 0], [0, 'x', 0], [0, 0, 'x']])},
            'discretization': good['discretization']}
    assert rc_of({'loadstep': [both]}) != 0 and b'Only one deformation' in lib.dmk_last_error()
    no_stress = {'boundary_conditions': {'mechanical': {'dot_F': good['boundary_conditions']['mechanical']['dot_F']}}, 'discretization': good['discretization']}
    # no stress entry = nothing stress controlled (upstream semantics): refused here because this dot_F leaves components open ...
    assert rc_of({'loadstep': [no_stress]}) != 0 and b'exactly one' in lib.dmk_last_error()
    with pytest.raises(L.DmkError):
        mech.parse_load({'loadstep': [no_stress]})
    # ... and accepted when the deformation prescribes all nine, as in the reference's examples/grid/shearXY.yaml
    full = {'boundary_conditions': {'mechanical': {'dot_F': [[0, 0, 0], [1.0e-3, 0, 0], [0, 0, 0]]}}, 'discretization': {'t': 60, 'N': 120}, 'f_out': 20}
    assert rc_of({'loadstep': [full]}) == 0, lib.dmk_last_error()
    assert list(steps[0].stress_mask) == [1] * 9 and steps[0].stress_type == b'P' and list(steps[0].deformation_mask) == [0] * 9
    py = mech.parse_load({'loadstep': [full]})[0]
    assert py['stress'].mask.all() and py['stress'].type == 'P' and not py['deformation'].mask.any() and py['deformation'].values[1, 0] == 1.0e-3
    overlap = {'boundary_conditions': {'mechanical': {'dot_F': [[1e-3, 0, 0], [0, 0, 0], [0, 0, 'x']], 'P': good['boundary_conditions']['mechanical']['P']}},
               'discretization': good['discretization']}
    assert rc_of({'loadstep': [overlap]}) != 0 and b'exactly one' in lib.dmk_last_error()
    assert rc_of({'loadstep': [good, good, good]}) != 0                     # more steps than the buffer holds
    assert lib.dmk_load_from_yaml(str(tmp_path / 'missing.yaml').encode(), steps, C.c_int(2), C.byref(n), None, C.c_int(0)) != 0


def _run_cpp_loop(outcomes, steps, n_steps, max_cut_back=3, max_increments=0):
    """the C++ loop above a scripted solver (ctypes callbacks): -> (rc, records, calls, forwards, results)"""
    import ctypes as C
    lib = L.load()
    outcomes = list(outcomes)
    calls, results = [], []
    forwards = [0]

    def forward(user, cut_back, guess, dt, dt_old, t_remaining, st):
        calls.append(dict(cut_back=bool(cut_back), guess=bool(guess), dt=dt, dt_prev=dt_old, t_remaining=t_remaining,
                          kind=st.contents.deformation_type.decode()))
        return 0

    def point_forward(user):
        forwards[0] += 1
        return 0

    def solve(user, out):
        ok = outcomes.pop(0) if outcomes else True
        out.contents.converged = int(ok); out.contents.stag_converged = 1; out.contents.terminally_ill = 0
        out.contents.required_iterations = 3 if ok else 40
        return 0

    def result(user, rec):
        results.append((rec.contents.inc, rec.contents.t))
        return 0
    ops = L.LoopOps(None, L.LOOP_FORWARD(forward), L.LOOP_POINT_FORWARD(point_forward), L.LOOP_SOLVE(solve), L.LOOP_RESULT(result))
    recs = (L.IncrementRecord * 32)()
    n = C.c_int()
    rc = lib.dmk_loop_over_loadcase(None, steps, C.c_int(n_steps), C.c_int(max_cut_back), C.c_int(max_increments), C.byref(ops),
                                    recs, C.c_int(32), C.byref(n))
    return rc, [recs[i] for i in range(n.value)], calls, forwards[0], results


@pytest.mark.parametrize('outcomes,max_increments', [([], None), ([True, False, True, True], 4), ([False, False, True, True, True, True], 2),
                                                     ([True, True, False, False, False, True] + [True] * 12, None)])
def test_cpp_loop_equals_python_loop_on_scripted_outcomes(outcomes, max_increments, tmp_path):
    steps, n_steps, _ = _c_steps(LOAD, tmp_path)
    s = _Scripted(outcomes)
    seen = []
    recs = mech.loop_over_loadcase(s, mech.parse_load(LOAD), max_increments=max_increments,
                                   callback=lambda r, _: seen.append((r['inc'], r['t'])) if r['inc_in_step'] % r['f_out'] == 0 else None)
    rc, crecs, calls, forwards, results = _run_cpp_loop(outcomes, steps, n_steps, max_increments=max_increments or 0)
    assert rc == 0, L.load().dmk_last_error()
    assert [(r.inc, r.step, r.inc_in_step, r.f_out, bool(r.converged), r.iterations) for r in crecs] == \
           [(r['inc'], r['step'], r['inc_in_step'], r['f_out'], r['converged'], r['iterations']) for r in recs]
    assert [r.t for r in crecs] == [r['t'] for r in recs]                    # the same floating-point bookkeeping
    assert forwards == s.g.forwards and len(calls) == len(s.calls)
    for a, b in zip(calls, s.calls):
        assert (a['cut_back'], a['guess'], a['dt'], a['dt_prev'], a['t_remaining']) == (b['cut_back'], b['guess'], b['dt'], b['dt_prev'], b['t_remaining'])
    assert results == seen


def test_cpp_loop_gives_up_like_the_reference_and_checks_its_arguments(tmp_path):
    lib = L.load()
    steps, n_steps, _ = _c_steps(LOAD, tmp_path)
    rc, recs, calls, forwards, results = _run_cpp_loop([False] * 8, steps, n_steps, max_cut_back=3, max_increments=1)
    assert rc != 0 and b'no solution found for increment 1' in lib.dmk_last_error() and recs == [] and results == []
    assert len(calls) == 4 and [c['dt'] for c in calls] == [1.0, 0.5, 0.25, 0.125]
    # test_damask_grid.cpp:74-128: max_cut_back = 0 -> one failed solve ends the run
    rc, _, calls, _, _ = _run_cpp_loop([False], steps, n_steps, max_cut_back=0)
    assert rc != 0 and len(calls) == 1
    # without a table of operations the loop drives the context's own solver: no context, no run
    import ctypes as C
    n = C.c_int()
    assert lib.dmk_loop_over_loadcase(None, steps, C.c_int(n_steps), C.c_int(3), C.c_int(0), None, None, C.c_int(0), C.byref(n)) != 0
    assert b'neither a context' in lib.dmk_last_error()


def test_example_load_decks_of_the_reference_through_all_parsers():
    """examples/grid/{tensionX,shearXY,shearZX}.yaml of the reference through the Python host, the library's C++ reader and the
    oracle's parser.  The two shear decks carry no stress entry: upstream DAMASK reads that as "nothing stress controlled";
    Config::parse_load (config.cpp:309-311) throws on them.  Build container only."""
    import ctypes as C
    import glob
    import yaml
    from oracle.solver import parse_load as oparse
    files = sorted(glob.glob('/root/reference/examples/grid/*.yaml'))
    decks = [f for f in files if 'loadstep' in (yaml.safe_load(open(f)) or {})]
    if not files:
        pytest.skip('needs the reference tree')
    assert [f.rsplit('/', 1)[-1] for f in decks] == ['shearXY.yaml', 'shearZX.yaml', 'tensionX.yaml']
    lib = L.load()
    for f in decks:
        d = yaml.safe_load(open(f))
        py, orc = mech.parse_load(d), oparse(d)
        steps = (L.LoadStep * 4)()
        n = C.c_int()
        name = C.create_string_buffer(32)
        assert lib.dmk_load_from_yaml(f.encode(), steps, C.c_int(4), C.byref(n), name, C.c_int(32)) == 0, lib.dmk_last_error()
        assert n.value == len(py) == len(orc) == 1 and name.value == b'spectral_basic'
        c, p, o = steps[0], py[0], orc[0]
        for side in (p, o):
            assert np.array_equal(np.array(c.deformation_values).reshape(3, 3), side['deformation'].values)
            assert np.array_equal(np.array(c.deformation_mask).reshape(3, 3).astype(bool), np.asarray(side['deformation'].mask, bool))
            assert np.array_equal(np.array(c.stress_mask).reshape(3, 3).astype(bool), np.asarray(side['stress'].mask, bool))
            assert (c.t, c.N, c.f_out) == (side['t'], side['N'], side['f_out'])
        if 'shear' in f:
            assert p['stress'].mask.all() and (c.t, c.N, c.f_out) == (60.0, 120, 20)


def test_cpp_loop_geometric_time_steps_equal_python_loop(tmp_path):
    """discretization with r != 1 (config.h:80: geometric series of time steps, damask_grid.cpp:95-101) incl. a cut-back"""
    load = {'loadstep': [{'boundary_conditions': LOAD['loadstep'][0]['boundary_conditions'], 'discretization': {'t': 7.0, 'N': 5, 'r': 1.7}, 'f_out': 2},
                         {'boundary_conditions': LOAD['loadstep'][1]['boundary_conditions'], 'discretization': {'t': 2.0, 'N': 3, 'r': 0.5}}]}
    steps, n_steps, _ = _c_steps(load, tmp_path)
    outcomes = [True, True, False, True, True]
    s = _Scripted(outcomes)
    recs = mech.loop_over_loadcase(s, mech.parse_load(load))
    rc, crecs, calls, forwards, _ = _run_cpp_loop(outcomes, steps, n_steps)
    assert rc == 0 and len(crecs) == len(recs) == 8
    assert [r.t for r in crecs] == [r['t'] for r in recs] and abs(recs[4]['t'] - 7.0) < 1e-14 and abs(recs[-1]['t'] - 9.0) < 1e-14
    assert [(c['dt'], c['dt_prev'], c['t_remaining'], c['cut_back'], c['guess']) for c in calls] == \
           [(c['dt'], c['dt_prev'], c['t_remaining'], c['cut_back'], c['guess']) for c in s.calls]
    dts = [c['dt'] for c in s.calls[:2]]
    assert abs(dts[1] / dts[0] - 1.0 / 1.7) < 1e-15 or abs(dts[1] / dts[0] - 1.7) < 1e-15      # a geometric series
    assert forwards == s.g.forwards


@pytest.mark.parametrize('deck', ['load_tensionX.yaml', '6grains6x7x8_tensionY.load.yaml', 'g9_irregular.load.yaml'])
def test_cpp_load_reader_on_the_golden_decks(deck, golden_dir):
    """the load files of the golden runs (G1's tensionX, G4's tensionY, G9's tensionX) through the C++ reader and the Python host"""
    import ctypes as C
    import os
    import yaml
    path = os.path.join(golden_dir, deck)
    ref = mech.parse_load(yaml.safe_load(open(path)))
    lib = L.load()
    steps = (L.LoadStep * 4)()
    n = C.c_int()
    assert lib.dmk_load_from_yaml(path.encode(), steps, C.c_int(4), C.byref(n), None, C.c_int(0)) == 0, lib.dmk_last_error()
    assert n.value == len(ref)
    for c, p in zip(steps, ref):
        assert np.array_equal(np.array(c.deformation_values).reshape(3, 3), p['deformation'].values)
        assert np.array_equal(np.array(c.deformation_mask).reshape(3, 3).astype(bool), p['deformation'].mask.astype(bool))
        assert np.array_equal(np.array(c.stress_values).reshape(3, 3), p['stress'].values)
        assert np.array_equal(np.array(c.stress_mask).reshape(3, 3).astype(bool), p['stress'].mask.astype(bool))
        assert (c.deformation_type.decode(), c.stress_type.decode(), c.t, c.N, c.r, c.f_out) == \
               (p['deformation'].type, p['stress'].type, p['t'], p['N'], p['r'], p['f_out'])
