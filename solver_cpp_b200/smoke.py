"""One small invocation of the hot path on cuda:0 through the C ABI, checked against the oracle (the oracle is the
checker only; it is never on the product path)."""
import numpy as np


def run(verbose=True):
    import torch
    from .grid import GridContext
    from oracle import spectral as osp                       # checker

    cells, size = (16, 16, 16), (1e-5, 1e-5, 1e-5)
    rng = np.random.default_rng(7)
    nx, ny, nz = cells
    P = 1e8 * rng.standard_normal((nz, ny, nx, 3, 3))
    aim = 1e-4 * rng.standard_normal((3, 3))
    C66 = np.zeros((6, 6)); C66[:3, :3] = 60e9
    C66[np.arange(3), np.arange(3)] = 107e9; C66[np.arange(3, 6), np.arange(3, 6)] = 28e9
    from oracle.crystal import voigt66_to_3333
    C = voigt66_to_3333(C66)
    grid = osp.Grid(cells, size)
    xi1, _ = osp.xi_tables(grid, 0)
    ref, hat = osp.gamma_convolution(P, aim, xi1, C / grid.wgt, grid, return_hat=True)
    rms_ref, _ = osp.divergence_rms(hat, xi1, grid, osp.scaled_geom_size(grid))
    g = GridContext(cells, size)
    g.update_gamma(C)
    d = g.to_device(P)
    rms = g.divergence_and_gamma(d, aim)
    out = g.to_natural(d)
    err = np.abs(out - ref).max() / np.abs(ref).max()
    assert abs(rms - rms_ref) <= 1e-12 * rms_ref, (rms, rms_ref)
    assert err <= 1e-11, err
    if verbose:
        print('smoke: spectral residual tail 16^3 ok, rel err %.2e, launches %d' % (err, g.kernel_launches()))
