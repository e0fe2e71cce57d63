"""Experiment (GPU box): increment loop on the device against the free-running oracle for the materials and loads of
BASELINE.json's configs that tests/test_gpu_solver.py did not cover in round 1 (Cu phenopowerlaw, TWIP dislotwin, the mixed
shear load).  Prints iteration counts, flags and the relative differences of the homogenised quantities per increment.
Usage: python tests/experiments/exp_increment_parity.py [case ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from solver_cpp_b200 import mech, workloads as W                                 # noqa: E402
from oracle.solver import MechSolverBasicOracle, parse_load as oparse, run_loadcase   # noqa: E402

CASES = {
    'Cu_tension': ('Cu_pheno', (16, 16, 16), 20, W.TENSION_X, 2),
    'Cu_shear': ('Cu_pheno', (16, 16, 16), 20, W.MIXED_SHEAR, 2),
    'Al_shear': ('Al_pheno', (16, 16, 16), 20, W.MIXED_SHEAR, 3),
    'TWIP_tension': ('TWIP', (12, 10, 8), 12, W.TENSION_X, 2),
    'iso_64': ('Al_iso', (16, 16, 16), 20, W.TENSION_X, 3),
}


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def run(name):
    phase, cells, grains, load, n_inc = CASES[name]
    material, material_at, size = W.polycrystal(cells, grains, phase, 0)
    t0 = time.time()
    o = MechSolverBasicOracle(cells, size, material, material_at, tangent_every_eval=False)
    o.log_fields = False
    o.init()
    orecs = []

    class Stop(Exception):
        pass

    def cb(rec, solver):
        orecs.append(rec)
        if len(orecs) == n_inc:
            raise Stop()
    try:
        run_loadcase(o, oparse(load), callback=cb)
    except Stop:
        pass
    t_o = time.time() - t0
    t0 = time.time()
    s = mech.make_solver(cells, size, material, material_at)
    s.init()
    recs = mech.loop_over_loadcase(s, mech.parse_load(load), max_increments=n_inc)
    t_d = time.time() - t0
    F = s.g.to_natural(s.field('F'))
    print('%-13s oracle %.1f s, device %.1f s; iterations %s vs %s; evaluations %d vs %d' %
          (name, t_o, t_d, [r['iterations'] for r in recs], [r['iterations'] for r in orecs], len(s.eval_log()), len(o.log)))
    for a, b in zip(recs, orecs):
        print('    inc %d: P_av rel %.2e, F_av rel %.2e, F_aim rel %.2e, converged %s/%s' %
              (a['inc'], rel(a['P_av'], b['P_av']), rel(a['F_av'], b['F_av']), rel(a['F_aim'], b['F_aim']), a['converged'], b['converged']))
    print('    final F field: rel %.2e (fluctuation %.2e)' % (rel(F, o.F), rel(F - np.eye(3), o.F - np.eye(3))), flush=True)


if __name__ == '__main__':
    for n in (sys.argv[1:] or list(CASES)):
        run(n)
