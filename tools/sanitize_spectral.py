"""Spectral kernels only (the shared-memory FFT / Gamma passes), small grids, for compute-sanitizer --tool racecheck."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from solver_cpp_b200.grid import GridContext

rng = np.random.default_rng(0)
for cells, size in (((16, 16, 16), (1e-5,) * 3), ((6, 7, 8), (6e-5, 7e-5, 8e-5))):
    g = GridContext(cells, size)
    C = np.zeros((3, 3, 3, 3))
    for i in range(3):
        for j in range(3):
            C[i, j, i, j] += 1e11; C[i, j, j, i] += 2e10; C[i, i, j, j] += 5e10
    g.update_gamma(C)
    f = g.to_device(1e8 * rng.standard_normal(cells[::-1] + (3, 3)))
    g.divergence_and_gamma(f, np.zeros((3, 3)))
torch.cuda.synchronize()
print('SANITIZE_SPECTRAL_DONE')
