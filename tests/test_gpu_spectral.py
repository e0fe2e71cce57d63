"""GPU parity tests of the spectral half (hand-written FFT passes, fused divergence + Gamma pass) against the
oracle, through the C ABI.  Tolerances: FP64 transforms of O(1) data -> 1e-12 relative to the field maximum."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from oracle import spectral as osp  # noqa: E402  (checker only)

GRIDS = [((3, 3, 3), (1e-5, 1e-5, 1e-5)), ((6, 7, 8), (6e-5, 7e-5, 8e-5)), ((16, 16, 16), (1e-5,) * 3),
         ((8, 4, 2), (4e-5, 2e-5, 1e-5)), ((5, 1, 1), (5e-5, 1e-5, 1e-5)), ((2, 1, 1), (2e-5, 1e-5, 1e-5)),
         ((64, 64, 64), (1e-5,) * 3), ((32, 16, 48), (2e-5, 1e-5, 3e-5))]


# thin grids that put every long-line plan of the register-resident passes (128: <16,8,1>, 256: <16,16,1> and the x c2r
# <8,8,4>, 512: <8,8,8> -- the plan of every 8-GPU number) on each axis in turn, small enough for the oracle
THIN = [(c, tuple(1e-5 * v / 16 for v in c)) for n in (128, 256, 512) for c in ((n, 16, 16), (16, n, 16), (16, 16, n))]


def _ctx(cells, size, **kw):
    from solver_cpp_b200.grid import GridContext
    return GridContext(cells, size, **kw)


def _stiffness(rng):
    """a positive definite, major+minor symmetric stiffness plus a small non-symmetric part (plastic tangent-like)"""
    C66 = np.zeros((6, 6))
    C66[:3, :3] = 60e9
    C66[np.arange(3), np.arange(3)] = 107e9
    C66[np.arange(3, 6), np.arange(3, 6)] = 28e9
    from oracle.crystal import voigt66_to_3333
    C = voigt66_to_3333(C66)
    return C + 1e8 * rng.standard_normal((3, 3, 3, 3))


@pytest.mark.parametrize('cells,size', GRIDS + THIN)
def test_fft_forward_matches_rfftn(cells, size):
    rng = np.random.default_rng(1)
    nx, ny, nz = cells
    f = rng.standard_normal((nz, ny, nx, 3, 3))
    g = _ctx(cells, size)
    hat = g.fft_forward(g.to_device(f)).cpu().numpy()
    hat = hat[..., 0] + 1j * hat[..., 1]                      # [kz, ky, kx, 9]
    ref = osp.fft_forward(f)                                  # [kz, ky, kx, 3, 3]
    ref = np.swapaxes(ref, -1, -2).reshape(hat.shape)
    assert np.abs(hat - ref).max() <= 1e-12 * np.abs(ref).max()


@pytest.mark.parametrize('cells,size', GRIDS + THIN)
def test_fft_roundtrip_unnormalised(cells, size):
    """test_fft.cpp:11-58 of the reference: forward then backward with wgt = 1/N is the identity."""
    rng = np.random.default_rng(2)
    nx, ny, nz = cells
    f = rng.standard_normal((nz, ny, nx, 3, 3))
    g = _ctx(cells, size)
    d = g.to_device(f)
    back = g.to_natural(g.fft_backward(g.fft_forward(d), 1.0 / (nx * ny * nz)))
    assert np.abs(back - f).max() < 1e-13


@pytest.mark.parametrize('derivative', [0, 1, 2])
@pytest.mark.parametrize('cells,size', GRIDS + THIN)
def test_divergence_and_gamma_match_oracle(cells, size, derivative):
    rng = np.random.default_rng(3)
    nx, ny, nz = cells
    P = 1e8 * rng.standard_normal((nz, ny, nx, 3, 3))
    aim = 1e-4 * rng.standard_normal((3, 3))
    C = _stiffness(rng)
    grid = osp.Grid(cells, size)
    xi1, _ = osp.xi_tables(grid, derivative)
    C_ref = C / grid.wgt
    ref, hat = osp.gamma_convolution(P, aim, xi1, C_ref, grid, return_hat=True)
    rms_ref, _ = osp.divergence_rms(hat, xi1, grid, osp.scaled_geom_size(grid))
    g = _ctx(cells, size, derivative=derivative)
    g.update_gamma(C)
    d = g.to_device(P)
    rms = g.divergence_and_gamma(d, aim)
    out = g.to_natural(d)
    assert abs(rms - rms_ref) <= 1e-12 * max(rms_ref, 1e-300)
    assert np.abs(out - ref).max() <= 1e-11 * max(np.abs(ref).max(), 1e-300)


def test_divergence_sum_is_reproducible():
    """the divergence sum decides `err_div / divTol < 1`, i.e. the iteration count: it must not depend on the order in which
    the CTAs of the fused pass finish (two-stage reduction in a fixed order, no floating-point atomics)"""
    cells, size = (64, 64, 64), (1e-5,) * 3
    rng = np.random.default_rng(11)
    P = 1e8 * rng.standard_normal((64, 64, 64, 3, 3))
    g = _ctx(cells, size)
    g.update_gamma(_stiffness(rng))
    vals = set()
    for _ in range(6):
        vals.add(g.divergence_and_gamma(g.to_device(P), np.zeros((3, 3))))
    assert len(vals) == 1, vals


def test_gamma_dc_only_broadcasts_aim():
    """test_mech_base.cpp:389-490: with a constant field only the DC coefficient survives -> output == aim."""
    cells, size = (2, 1, 1), (2e-5, 1e-5, 1e-5)
    g = _ctx(cells, size)
    rng = np.random.default_rng(4)
    g.update_gamma(_stiffness(rng))
    P = np.broadcast_to(rng.standard_normal((3, 3)), (1, 1, 2, 3, 3)).copy()
    aim = rng.standard_normal((3, 3))
    d = g.to_device(P)
    g.divergence_and_gamma(d, aim)
    out = g.to_natural(d)
    assert np.abs(out - aim).max() < 1e-14


def test_forward_field_and_rate():
    cells, size = (6, 7, 8), (6e-5, 7e-5, 8e-5)
    rng = np.random.default_rng(5)
    g = _ctx(cells, size)
    grid = osp.Grid(cells, size)
    f0 = rng.standard_normal((8, 7, 6, 3, 3)); f1 = rng.standard_normal((8, 7, 6, 3, 3)); aim = rng.standard_normal((3, 3))
    rate = g.calculate_rate(True, g.to_device(f0), g.to_device(f1), 0.37, np.zeros((3, 3)))
    ref_rate = osp.calculate_rate(True, f0, f1, 0.37, None)
    assert np.abs(g.to_natural(rate) - ref_rate).max() < 1e-14
    hom = g.calculate_rate(False, g.to_device(f0), g.to_device(f1), 0.37, aim)
    assert np.abs(g.to_natural(hom) - aim).max() == 0.0
    out = g.forward_field(0.5, g.to_device(f0), rate, aim)
    ref = osp.forward_field(0.5, f0, ref_rate, grid, aim)
    assert np.abs(g.to_natural(out) - ref).max() < 1e-13


def test_g3_divergence_rms_unit_test_literal(golden_dir):
    """the reference's own unit-test vector (test_mech_base.cpp:324-385): literal 4x1x1 field -> 1481.2323577332318"""
    import os
    g3 = np.load(os.path.join(golden_dir, 'g3_unit_test_literals.npz'))
    cells = tuple(int(c) for c in g3['div_cells'])
    g = _ctx(cells, tuple(1e-5 * c for c in cells), derivative=0)
    g.update_gamma(_stiffness(np.random.default_rng(0)))
    field = np.ascontiguousarray(np.transpose(g3['div_field'], (4, 3, 2, 0, 1)))
    rms = g.divergence_and_gamma(g.to_device(field), np.zeros((3, 3)))
    assert abs(rms - float(g3['div_rms'])) <= 1e-12 * float(g3['div_rms'])


@pytest.mark.parametrize('cells,size', [((6, 7, 8), (0.75, 0.875, 1.0)), ((16, 16, 16), (1e-5,) * 3), ((5, 1, 1), (5e-5, 1e-5, 1e-5)),
                                        ((32, 16, 48), (2e-5, 1e-5, 3e-5))])
def test_update_coords_matches_oracle(cells, size):
    """MechBase::base_update_coords: Fourier-space integration of F, nodal averaging (oracle pinned by the displacements of
    a real DAMASK result file, tests/test_oracle_golden.py)."""
    rng = np.random.default_rng(4)
    nx, ny, nz = cells
    F = np.eye(3) + 1e-2 * rng.standard_normal((nz, ny, nx, 3, 3))
    grid = osp.Grid(cells, size)
    x_n_o, x_p_o = osp.update_coords(F, grid)
    g = _ctx(cells, size)
    x_n, x_p = g.update_coords(g.to_device(F))
    scale = max(size)
    assert np.abs(x_p - x_p_o).max() < 1e-13 * scale
    assert np.abs(x_n - x_n_o).max() < 1e-13 * scale


def test_g4_displacements_of_the_real_damask_run(golden_dir):
    """cell-centre and nodal displacements u_p, u_n stored by DAMASK itself (golden set G4) from the stored F"""
    import os
    g4 = np.load(os.path.join(golden_dir, 'g4_damask_run.npz'))
    cells, size = (6, 7, 8), (0.75, 0.875, 1.0)
    g = _ctx(cells, size)
    x_n0, x_p0 = g.update_coords(g.to_device(np.broadcast_to(np.eye(3), (8, 7, 6, 3, 3)).copy()))
    for k in range(len(g4['increments'])):
        x_n, x_p = g.update_coords(g.to_device(g4['F'][k].reshape(8, 7, 6, 3, 3)))
        assert np.abs((x_p - x_p0).reshape(-1, 3) - g4['u_p'][k]).max() < 1e-14
        assert np.abs((x_n - x_n0).reshape(-1, 3) - g4['u_n'][k]).max() < 1e-14
