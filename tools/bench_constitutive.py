"""Times the constitutive kernel alone on a synthetic polycrystal. Usage: python tools/bench_constitutive.py N phase [reps]"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from solver_cpp_b200.grid import GridContext
from solver_cpp_b200 import workloads as W, material as M

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
phase = sys.argv[2] if len(sys.argv) > 2 else 'Cu_pheno'
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cells = (n, n, n)
t0 = time.time()
material, material_at, size = W.polycrystal(cells, max(20, n ** 3 // 33000), phase, 0)
g = GridContext(cells, size)
g.material_init(M.Material(material, material_at))
t_setup = time.time() - t0
rng = np.random.default_rng(0)
res = []
F = torch.zeros(g.field_shape, dtype=torch.float64, device='cuda')
F[..., 0] = 1; F[..., 4] = 1; F[..., 8] = 1
fl = torch.randn(g.field_shape, dtype=torch.float64, device='cuda') * 2e-4
P = g.new_field()
for inc in range(1, reps + 1):
    Fi = F.clone()
    Fi[..., 0] += 1.0e-3 * 0.5 * inc * 2
    Fi[..., 4] -= 0.3e-3 * 0.5 * inc * 2
    Fi[..., 8] -= 0.3e-3 * 0.5 * inc * 2
    Fi += fl * inc
    c0 = g.counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    g.constitutive_response(Fi, 0.5, P)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    c1 = g.counters()
    e0.record(); Cv, Cm, sel = g.tangent_response(0.5); e1.record(); torch.cuda.synchronize()
    ms_t = e0.elapsed_time(e1)
    g.materialpoint_forward()
    res.append(dict(inc=inc, ms=ms, Mpts_per_s=n ** 3 / ms / 1e3, Lp_iter_per_pt=(c1['Lp_iter'] - c0['Lp_iter']) / n ** 3,
                    stress_calls_per_pt=(c1['stress_calls'] - c0['stress_calls']) / n ** 3,
                    deferred_frac=(c1['deferred'] - c0['deferred']) / n ** 3, ms_tangent=ms_t,
                    ill=g.terminally_ill(), P11=float(P[..., 0].mean()) / 1e6))
print(json.dumps(dict(n=n, phase=phase, setup_s=t_setup, evals=res)))
