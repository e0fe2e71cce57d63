"""Experiment (1 GPU): is the per-point update of a virgin soft-Cu evaluation (every cell cuts back) reproducible run to run, and
independent of which cells share a launch?  Whole grid twice, then the two z-halves in contexts of their own."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from solver_cpp_b200 import workloads as W, material as M          # noqa: E402
from solver_cpp_b200.grid import GridContext                       # noqa: E402

cells = (32, 32, 16)
n = int(np.prod(cells))
material, material_at, size = W.polycrystal(cells, 12, 'Cu_pheno', 0)
rng = np.random.default_rng(0)
F = np.broadcast_to(np.eye(3), (16, 32, 32, 3, 3)).copy()
F[..., 0, 0] += 5e-4
F += 1e-5 * rng.standard_normal(F.shape)


def run(sel, mass):
    g = GridContext((32, 32, sel.stop - sel.start), size)
    g.set_option('pheno_mass_deferral', mass)
    g.material_init(M.Material(material, material_at.reshape(16, 32, 32)[sel].reshape(-1)))
    P = g.to_natural(g.constitutive_response(g.to_device(F[sel]), 0.5)).copy()
    P2 = g.to_natural(g.constitutive_response(g.to_device(F[sel] * (1 + 1e-7)), 0.5)).copy()
    c = g.counters()
    g.close()
    return P, P2, c


for mass, name in ((2, 'lanes'), (0, 'nested')):
    a = run(slice(0, 16), mass)
    b = run(slice(0, 16), mass)
    lo = run(slice(0, 8), mass)
    hi = run(slice(8, 16), mass)
    halves = np.concatenate([lo[0], hi[0]]), np.concatenate([lo[1], hi[1]])
    for k in (0, 1):
        d1 = np.abs(a[k] - b[k]).max() / np.abs(a[k]).max()
        d2 = np.abs(a[k] - halves[k]).max() / np.abs(a[k]).max()
        nd = int((np.abs(a[k] - halves[k]).reshape(n, -1).max(1) > 0).sum())
        print('%s evaluation %d: run-to-run rel diff %.3e, whole vs halves %.3e (%d of %d cells differ); deferred %d' % (name, k, d1, d2, nd, n, a[2]['deferred']))
