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
