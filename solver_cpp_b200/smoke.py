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
    run_residual(verbose)


def run_residual(verbose=True):
    """one residual evaluation (constitutive update + homogenisation + FFT + Gamma/div + inverse FFT) of the 16^3
    phenopowerlaw configuration through the C ABI, against the oracle's form_residual"""
    from . import mech, workloads as W
    from oracle.solver import MechSolverBasicOracle, parse_load as oparse      # checker
    from oracle import spectral as osp
    cells = (16, 16, 16)
    material, material_at, size = W.polycrystal(cells, 20, 'Al_pheno', 0)
    s = mech.make_solver(cells, size, material, material_at)
    s.init()
    o = MechSolverBasicOracle(cells, size, material, material_at)
    o.init()
    st, sto = mech.parse_load(W.TENSION_X)[0], oparse(W.TENSION_X)[0]
    s.forward(False, False, 0.5, 1.0, 5.0, st['deformation'], st['stress'], st['R'])
    s.g.materialpoint_forward()
    o.forward(False, False, 0.5, 1.0, 5.0, sto['deformation'], sto['stress'], sto['R'])
    o.con.forward()
    s.prepare_evaluation()
    o.S = osp.masked_compliance(o.C_volAvg, o.params['R'], o.params['stress_mask'])
    r = s.g.to_natural(s.form_residual(s.field('F')))
    r_o = o.form_residual(o.F.reshape(-1), 1, 1).reshape(o.shape)
    err = np.abs(r - r_o).max() / np.abs(r_o).max()
    errP = np.abs(s.get33('P_av') - o.P_av).max() / np.abs(o.P_av).max()
    assert err < 1e-9 and errP < 1e-9, (err, errP)
    if verbose:
        print('smoke: residual evaluation 16^3 phenopowerlaw ok, rel err r %.2e, P_av %.2e, launches %d'
              % (err, errP, s.g.kernel_launches()))
