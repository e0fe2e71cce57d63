"""Times the spectral residual tail (x r2c, y, fused z+Gamma+div, y inverse, x c2r) on one GPU with CUDA events.
Usage: python tools/bench_spectral.py N [kt_y kt_z] [reps]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from solver_cpp_b200.grid import GridContext

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
kty = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ktz = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
g = GridContext((n, n, n), (1e-5,) * 3, kt_y=kty, kt_z=ktz)
C66 = np.zeros((6, 6)); C66[:3, :3] = 60e9
C66[np.arange(3), np.arange(3)] = 107e9; C66[np.arange(3, 6), np.arange(3, 6)] = 28e9
C = np.zeros((3, 3, 3, 3))
V = [(0, 0), (1, 1), (2, 2), (1, 2), (0, 2), (0, 1)]
for a, (i, j) in enumerate(V):
    for b, (k, l) in enumerate(V):
        C[i, j, k, l] = C[j, i, k, l] = C[i, j, l, k] = C[j, i, l, k] = C66[a, b]
g.update_gamma(C)
f = torch.randn(g.field_shape, dtype=torch.float64, device='cuda') * 1e8
aim = np.zeros((3, 3))
st = torch.cuda.current_stream()
for _ in range(3):
    g.divergence_and_gamma(f, aim)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
torch.cuda.synchronize()
ev[0].record()
for i in range(reps):
    g.divergence_and_gamma(f, aim)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
N = n ** 3
best, med = min(ms), sorted(ms)[len(ms) // 2]
print(json.dumps(dict(n=n, kt_y=kty, kt_z=ktz, ms_median=med, ms_best=best, GBs_median=720 * N / med / 1e6,
                      frac_of_6535=720 * N / med / 1e6 / 6535.1)))
