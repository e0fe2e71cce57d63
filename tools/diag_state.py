"""Characterises the per-point work of bench.py's state: who is deferred, how many cut-backs, time per call.
Usage: python tools/diag_state.py N [relax]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
if len(sys.argv) > 2:
    bench.N_RELAX = int(sys.argv[2])
s, X, R, grains = bench.setup_state((n, n, n))
g = s.g
P = g.new_field()
out = {'n': n, 'relax': bench.N_RELAX}
for rep in range(3):
    c0 = g.counters(); s0 = g.schedule_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    g.constitutive_response(X, bench.DT, P)
    e1.record(); torch.cuda.synchronize()
    c1 = g.counters()
    out['ms'] = e0.elapsed_time(e1)
    out['per_cell'] = {k: (c1[k] - c0[k]) / n ** 3 for k in c1}
    s1 = g.schedule_stats()
    out['sched'] = {k: s1[k] - s0[k] for k in s1}
    out['lanes_per_trip'] = {b: out['sched'][b + '_lanes'] / max(out['sched'][b + '_trips'], 1) for b in ('lp', 'sb', 'refill')}
cb = g.get_state('cutbacks', 1)[:, 0]
out['cutback_hist'] = np.bincount(cb.astype(int), minlength=8)[:12].tolist()
F = g.to_natural(X).reshape(-1, 3, 3)
out['F_std'] = F.std(axis=0).round(6).tolist()
out['F_mean'] = F.mean(axis=0).round(6).tolist()
print(json.dumps(out))
