"""Small invocation of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck):
16^3 and 32x16x16 spectral residual tail, 16^3 phenopowerlaw residual evaluations incl. a cut-back increment, tangent,
update_coords, isotropic and dislotwin evaluations.  Usage: compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from solver_cpp_b200 import workloads as W, mech, material as M
from solver_cpp_b200.grid import GridContext

rng = np.random.default_rng(0)
for cells, size in (((16, 16, 16), (1e-5,) * 3), ((32, 16, 16), (2e-5, 1e-5, 1e-5)), ((6, 7, 8), (6e-5, 7e-5, 8e-5))):
    g = GridContext(cells, size)
    C = np.zeros((3, 3, 3, 3))
    for i in range(3):
        for j in range(3):
            C[i, j, i, j] += 1e11; C[i, j, j, i] += 2e10; C[i, i, j, j] += 5e10
    g.update_gamma(C)
    f = g.to_device(1e8 * rng.standard_normal(cells[::-1] + (3, 3)))
    g.divergence_and_gamma(f, np.zeros((3, 3)))
    g.update_coords(g.to_device(np.eye(3) + 1e-3 * rng.standard_normal(cells[::-1] + (3, 3))))
for phase in ('Cu_pheno', 'Al_iso', 'TWIP'):
    cells = (16, 16, 16) if phase == 'Cu_pheno' else (8, 8, 8)
    material, material_at, size = W.polycrystal(cells, 9, phase, 0)
    s = mech.make_solver(cells, size, material, material_at)
    s.init()
    recs = mech.loop_over_loadcase(s, mech.parse_load(W.TENSION_X), max_increments=2)
    s.update_coords()
    print(phase, [r['iterations'] for r in recs])
torch.cuda.synchronize()
print('SANITIZE_SMOKE_DONE')
