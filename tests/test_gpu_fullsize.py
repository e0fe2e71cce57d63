"""GPU parity at BASELINE.json's full single-GPU size (256^3), where the oracle cannot run the whole grid in seconds:
size-independent properties of the spectral half (round trip, Parseval, the Gamma operator as a projector onto compatible
fields, a one-mode divergence known answer) and a random sample of cells of the constitutive update against the oracle.
Through the C ABI, like every -m gpu test."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from oracle import spectral as osp                      # noqa: E402  (checker only)
from oracle.constitutive import ConstitutiveOracle      # noqa: E402  (checker only)

CELLS = (256, 256, 256)
SIZE = (1.0e-5, 1.0e-5, 1.0e-5)


def _ctx(**kw):
    from solver_cpp_b200.grid import GridContext
    return GridContext(CELLS, SIZE, **kw)


def _stiffness():
    from oracle.crystal import voigt66_to_3333
    C66 = np.zeros((6, 6))
    C66[:3, :3] = 121.0e9
    C66[np.arange(3), np.arange(3)] = 168.0e9
    C66[np.arange(3, 6), np.arange(3, 6)] = 75.0e9
    C = voigt66_to_3333(C66)
    # a rotated cubic stiffness: anisotropic, major and minor symmetric
    a = 0.37
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]]) @ \
        np.array([[1, 0, 0], [0, np.cos(0.61), -np.sin(0.61)], [0, np.sin(0.61), np.cos(0.61)]])
    return np.einsum('ia,jb,kc,ld,abcd->ijkl', R, R, R, R, C)


def _phase_field(modes, dev):
    """cos(k.x + phi) of each mode on the cell centres, device tensors [nz, ny, nx]; k as physical wave vectors"""
    nx, ny, nz = CELLS
    x = (torch.arange(nx, device=dev, dtype=torch.float64) + 0.5) * (SIZE[0] / nx)
    y = (torch.arange(ny, device=dev, dtype=torch.float64) + 0.5) * (SIZE[1] / ny)
    z = (torch.arange(nz, device=dev, dtype=torch.float64) + 0.5) * (SIZE[2] / nz)
    out = []
    for (n1, n2, n3), phi in modes:
        k = 2 * np.pi * np.array([n1 / SIZE[0], n2 / SIZE[1], n3 / SIZE[2]])
        arg = (k[0] * x)[None, None, :] + (k[1] * y)[None, :, None] + (k[2] * z)[:, None, None] + phi
        out.append((k, torch.cos(arg)))
    return out


def test_fft_round_trip_and_parseval_256():
    g = _ctx()
    gen = torch.Generator(device='cuda').manual_seed(7)
    f = torch.randn(g.field_shape, dtype=torch.float64, device='cuda', generator=gen)
    hat = g.fft_forward(f)
    n = float(np.prod(CELLS))
    back = g.fft_backward(hat, 1.0 / n)
    assert float((back - f).abs().max()) < 1e-13 * float(f.abs().max()) * 16
    # Parseval on the half spectrum: interior kx counted twice (the weights of calculate_divergence_rms)
    w = torch.full((g.nxr,), 2.0, dtype=torch.float64, device='cuda')
    w[0] = 1.0
    w[-1] = 1.0
    e_hat = float(((hat ** 2).sum(dim=(-1, -2, 0, 1)) * w).sum()) / n
    e = float((f ** 2).sum())
    assert abs(e_hat - e) <= 1e-12 * e


def test_gamma_projects_onto_compatible_fields_256():
    """For a periodic displacement field u, Gamma(C) * (C : grad u) = grad u (the basic scheme's fixed point), and the DC
    term is the aim.  Band-limited u, continuous derivative: exact up to rounding.  No oracle involved."""
    g = _ctx(derivative=0)
    C = _stiffness()
    g.update_gamma(C)
    rng = np.random.default_rng(5)
    modes = [((1, 0, 0), 0.3), ((0, 2, 0), 1.1), ((0, 0, 3), 2.0), ((1, 1, 0), 0.7), ((2, -1, 3), 1.9), ((5, 7, -4), 0.2),
             ((17, 3, 9), 2.6)]
    grad = torch.zeros((CELLS[2], CELLS[1], CELLS[0], 3, 3), dtype=torch.float64, device='cuda')
    for k, c in _phase_field(modes, 'cuda'):
        A = 1.0e-9 * rng.standard_normal(3)                               # u = A sin(k.x + phi): |grad u| ~ 1e-3
        grad += torch.from_numpy(np.outer(A, k)).cuda()[None, None, None] * c[..., None, None]
    tau = torch.einsum('ijkl,zyxkl->zyxij', torch.from_numpy(C).cuda(), grad)
    aim = 1.0e-4 * rng.standard_normal((3, 3))
    field = tau.transpose(-1, -2).reshape(g.field_shape).contiguous()      # record order c = i + 3 j
    g.divergence_and_gamma(field, aim)
    out = field.reshape(grad.shape).transpose(-1, -2)
    expect = grad + torch.from_numpy(aim).cuda()
    scale = float(grad.abs().max())
    assert float((out - expect).abs().max()) < 1e-10 * scale


def test_divergence_rms_one_mode_known_answer_256():
    """tau = T cos(k.x): RMS of div tau = |T k| / sqrt(2), times the reference's geometry rescaling of each component."""
    g = _ctx(derivative=0)
    g.update_gamma(_stiffness())
    rng = np.random.default_rng(9)
    T = 1.0e8 * rng.standard_normal((3, 3))
    (k, c), = _phase_field([((3, -2, 5), 0.4)], 'cuda')
    tau = torch.from_numpy(T).cuda()[None, None, None] * c[..., None, None]
    field = tau.transpose(-1, -2).reshape(g.field_shape).contiguous()
    rms = g.divergence_and_gamma(field, np.zeros((3, 3)))
    grid = osp.Grid(CELLS, SIZE)
    resc = np.asarray(SIZE) / np.asarray(osp.scaled_geom_size(grid))
    expect = np.linalg.norm(T @ (k * resc)) / np.sqrt(2.0)
    assert abs(rms - expect) <= 1e-10 * expect


def test_constitutive_sample_of_cells_matches_oracle_256():
    """Al phenopowerlaw polycrystal (500 grains) at 256^3: three increments of a smooth heterogeneous deformation on the GPU,
    4096 randomly chosen cells replayed by the oracle (the update is point-local): P to 1e-9, identical Lp iteration counts."""
    from solver_cpp_b200 import workloads as W, material as M
    material, material_at, size = W.polycrystal(CELLS, 500, 'Al_pheno', 0)
    g = _ctx()
    g.material_init(M.Material(material, material_at))
    rng = np.random.default_rng(13)
    n = int(np.prod(CELLS))
    idx = np.sort(rng.choice(n, 4096, replace=False))
    o = ConstitutiveOracle(material, np.asarray(material_at).reshape(-1)[idx])
    modes = [((1, 0, 0), 0.3), ((0, 1, 0), 1.1), ((0, 0, 1), 2.0), ((1, 2, 0), 0.7), ((2, 1, 3), 1.9), ((3, 5, 4), 0.2)]
    fl = torch.zeros((CELLS[2], CELLS[1], CELLS[0], 3, 3), dtype=torch.float64, device='cuda')
    for k, c in _phase_field(modes, 'cuda'):
        fl += torch.from_numpy(rng.standard_normal((3, 3))).cuda()[None, None, None] * c[..., None, None]
    eye = torch.eye(3, dtype=torch.float64, device='cuda')
    dt = 0.5
    P = g.new_field()
    for inc in (1, 2, 3):
        e = 1.5e-3 * inc
        F = eye + torch.diag(torch.tensor([e, -0.3 * e, -0.3 * e], dtype=torch.float64, device='cuda')) + 1.0e-4 * inc * fl
        Frec = F.transpose(-1, -2).reshape(g.field_shape).contiguous()
        g.constitutive_response(Frec, dt, P)
        assert not g.terminally_ill()
        Fs = F.reshape(n, 3, 3)[torch.from_numpy(idx).cuda()].cpu().numpy()
        assert not o.mechanical_response(Fs.reshape(-1, 9), dt)
        P_o, _ = o.mechanical_response2(dt, tangent=False)
        Ps = P.reshape(n, 3, 3)[torch.from_numpy(idx).cuda()].transpose(-1, -2).cpu().numpy()
        err = np.abs(Ps - P_o.reshape(-1, 3, 3)).reshape(len(idx), -1).max(1) / np.abs(P_o).reshape(len(idx), -1).max(1)
        lp = g.get_state('lp_iter', 1)[idx, 0]
        cb = g.get_state('cutbacks', 1)[idx, 0]
        quiet = (cb == 0) & (o.point_cutbacks == 0)
        flipped = quiet & (lp != o.point_lp_iter)
        assert quiet.mean() > 0.99 and flipped.mean() <= 0.005
        assert (err[quiet & ~flipped] < 1e-9).all() and (err[quiet] < 1e-6).all(), (inc, err[quiet].max())
        g.materialpoint_forward()
        o.forward()
    assert np.abs(o.a['Lp']).max() > 1e-6                 # plastic flow was active


def test_constitutive_sample_of_cells_matches_oracle_256_cu_bench_iterate():
    """C3's own material at C3's size: the 256^3 500-grain Cu phenopowerlaw polycrystal taken to the kind of iterate bench.py's
    evaluation protocol times (two homogeneous increments of tensionX, then a third with a smooth fluctuation inside the grains);
    4096 randomly chosen cells replayed by the oracle.  Virgin Cu cuts back in every cell in the first increment, and cells on
    such a path are rounding-chaotic for any build of the reference algorithm (tests/test_oracle_golden.py: both sides end
    within the integrator's tolerance of each other, not within 1e-9).  So the first two increments are compared at that
    tolerance, the oracle then takes over the device's per-point state of its cells, and in the third increment -- the bench-like
    iterate -- the cells that do not cut back must agree to 1e-9 where the Lp iteration counts are identical, 1e-6 where a stopping
    decision of the Newton loop flipped."""
    from solver_cpp_b200 import workloads as W, material as M
    material, material_at, size = W.polycrystal(CELLS, 500, 'Cu_pheno', 0)
    g = _ctx()
    g.material_init(M.Material(material, material_at))
    rng = np.random.default_rng(17)
    n = int(np.prod(CELLS))
    idx = np.sort(rng.choice(n, 4096, replace=False))
    o = ConstitutiveOracle(material, np.asarray(material_at).reshape(-1)[idx])
    modes = [((1, 0, 0), 0.3), ((0, 1, 0), 1.1), ((0, 0, 1), 2.0), ((1, 2, 0), 0.7), ((2, 1, 3), 1.9), ((3, 5, 4), 0.2)]
    fl = torch.zeros((CELLS[2], CELLS[1], CELLS[0], 3, 3), dtype=torch.float64, device='cuda')
    for k, c in _phase_field(modes, 'cuda'):
        fl += torch.from_numpy(rng.standard_normal((3, 3))).cuda()[None, None, None] * c[..., None, None]
    eye = torch.eye(3, dtype=torch.float64, device='cuda')
    dt = 0.5
    P = g.new_field()
    sel = torch.from_numpy(idx).cuda()
    for inc in (1, 2, 3):
        e = 1.0e-3 * dt * inc
        F = eye + torch.diag(torch.tensor([e, -0.5 * e, -0.5 * e], dtype=torch.float64, device='cuda')) + (4.0e-5 * fl if inc == 3 else 0.0)
        F = F.expand(CELLS[2], CELLS[1], CELLS[0], 3, 3)
        Frec = F.transpose(-1, -2).reshape(g.field_shape).contiguous()
        g.constitutive_response(Frec, dt, P)
        Fs = F.reshape(n, 3, 3)[sel].cpu().numpy()
        ill_o = o.mechanical_response(Fs.reshape(-1, 9), dt)
        assert not g.terminally_ill() and not ill_o
        P_o, _ = o.mechanical_response2(dt, tangent=False)
        Ps = P.reshape(n, 3, 3)[sel].transpose(-1, -2).cpu().numpy()
        err = np.abs(Ps - P_o.reshape(-1, 3, 3)).reshape(len(idx), -1).max(1) / np.abs(P_o).reshape(len(idx), -1).max(1)
        lp = g.get_state('lp_iter', 1)[idx, 0]
        quiet = (g.get_state('cutbacks', 1)[idx, 0] == 0) & (o.point_cutbacks == 0)
        flipped = quiet & (lp != o.point_lp_iter)
        # cut-back history: both sides inside the integrator's tolerance of each other (a few cells per 4096 land further apart)
        assert np.quantile(err, 0.99) < 0.05 and err.max() < 0.2, (inc, np.quantile(err, 0.99), err.max())
        if inc == 3:                                       # same start state on both sides: cells that do not cut back agree to 1e-9
            # (observed on B200: 81 % of the sampled cells quiet, 4 % of them one Newton iteration apart from the oracle)
            assert quiet.mean() > 0.5 and flipped.mean() <= 0.1, (quiet.mean(), flipped.mean())
            assert (err[quiet & ~flipped] < 1e-9).all() and (err[quiet] < 1e-6).all(), (inc, err[quiet & ~flipped].max(), err[quiet].max())
        g.materialpoint_forward()
        o.forward()
        if inc == 2:                                       # the oracle continues from the device's state of its cells
            ms = o.max_state
            for k in ('F', 'F0', 'Fp', 'Fp0', 'Fe', 'Lp', 'Lp0', 'S', 'S0'):
                o.a[k][:] = g.get_state(k, 9)[idx]
            for k in ('state', 'state0'):
                o.a[k][:] = g.get_state(k, int(g.lib.dmk_max_state(g.ctx)))[idx][:, :ms]
    assert np.abs(o.a['Lp']).max() > 1e-6                 # plastic flow was active
