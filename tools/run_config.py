"""Runs the first increments of a BASELINE.json configuration (solver_cpp_b200.workloads.CONFIGS) on one GPU and reports
wall time, solver iterations and residual evaluations per increment.  Usage: python tools/run_config.py C1 [increments] [nx ny nz]"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from solver_cpp_b200 import workloads as W, mech

name = sys.argv[1] if len(sys.argv) > 1 else 'C1'
n_inc = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cells = tuple(int(v) for v in sys.argv[3:6]) if len(sys.argv) > 5 else None
c = W.config(name, cells=cells)
s = mech.make_solver(c['cells'], c['size'], c['material'], c['material_at'])
t0 = time.time(); s.init(); torch.cuda.synchronize(); t_init = time.time() - t0
out = []
last = [time.time(), s.g.counters()]


def cb(rec, solver):
    torch.cuda.synchronize()
    now, cnt = time.time(), solver.g.counters()
    n = int(np.prod(c['cells']))
    evals = (cnt['substeps'] - last[1]['substeps']) / n if cnt['cutbacks'] == last[1]['cutbacks'] else None
    out.append(dict(inc=rec['inc'], seconds=round(now - last[0], 3), iterations=rec['iterations'], P11_MPa=round(float(rec['P_av'][0, 0]) / 1e6, 3),
                    evaluations=evals, Lp_iter_per_cell=round((cnt['Lp_iter'] - last[1]['Lp_iter']) / n, 1),
                    cutbacks_per_cell=round((cnt['cutbacks'] - last[1]['cutbacks']) / n, 3)))
    last[0], last[1] = now, cnt


recs = mech.loop_over_loadcase(s, mech.parse_load(c['load']), callback=cb, max_increments=n_inc)
print(json.dumps(dict(config=name, name=c['name'], cells=list(c['cells']), init_s=round(t_init, 2), increments=out)))
