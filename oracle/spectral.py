"""ORACLE (test infrastructure only -- never imported by the product path).

numpy restatement of the spectral half of the hot path, upstream-DAMASK semantics (SURVEY.md 0.1 fixes D1-D5
applied).  Follows, in /root/reference/src/grid:
  xi tables / frequency derivative     spectral.cpp:12-130      (Nyquist zeroing per component: D4)
  divergence-correction scaling        mech_base.cpp:11-43
  update_gamma / gamma_convolution     mech_base.cpp:156-211, 400-477   (Gamma-hat per k-point: D1/D2)
  calculate_divergence_rms             mech_base.cpp:302-398
  calculate_masked_compliance          mech_base.cpp:249-300
  forward_field / calculate_rate       mech_base.cpp:213-247, 479-497
  FFT layout / normalisation           fft.cpp:16-78, fft.h:37-60  (unnormalised r2c and c2r)
  homogenised P / C reductions         spectral.cpp:160-218      (tie rule: SURVEY.md 0.2 V7)
Third-party arithmetic: FFTW 3.3 (absent) is replaced by numpy/scipy pocketfft -- the DFT is mathematically
pinned, its rounding is not.

Array convention: fields are "natural" numpy arrays [z, y, x, i, j]; the half spectrum is [kz, ky, kx, i, j].
The reference's Fortran memory order F(3,3,x,y,z) equals natural.swapaxes(-1,-2) flattened in C order.
"""
import numpy as np

try:
    import scipy.fft as _fft
    _KW = dict(workers=-1)
except Exception:                                           # pragma: no cover
    import numpy.fft as _fft
    _KW = {}

TAU = 2.0 * np.pi
DERIVATIVE_CONTINUOUS, DERIVATIVE_CENTRAL_DIFF, DERIVATIVE_FWBW_DIFF = 0, 1, 2
DIVCORR_NONE, DIVCORR_SIZE, DIVCORR_SIZE_GRID = 0, 1, 2

# relative window inside which two ||dPdF||^2 values count as a tie; ties -> lowest cell index
TIE_RTOL = 1.0e-12


def fortran_to_natural(flat, cells):
    nx, ny, nz = cells
    return np.asarray(flat, float).reshape(nz, ny, nx, 3, 3).swapaxes(-1, -2)


def natural_to_fortran(field):
    return np.ascontiguousarray(field.swapaxes(-1, -2)).reshape(-1)


class Grid:
    def __init__(self, cells, size):
        self.cells = tuple(int(c) for c in cells)            # (nx, ny, nz)
        self.size = tuple(float(s) for s in size)
        self.n = int(np.prod(self.cells))
        self.wgt = 1.0 / self.n
        self.nxr = self.cells[0] // 2 + 1


def scaled_geom_size(grid, divergence_correction=DIVCORR_SIZE_GRID):
    size, cells = np.array(grid.size), np.array(grid.cells, float)
    if divergence_correction == DIVCORR_SIZE:
        mid = np.argsort(size, kind='stable')[1]
        return size / size[mid]
    if divergence_correction == DIVCORR_SIZE_GRID:
        mid = np.argsort(size / cells, kind='stable')[1]
        return size / size[mid] * cells[mid]
    return size.copy()


def freq_derivative(kind, cells, size, k):
    """k: integer wave numbers array [..., 3] -> complex xi [..., 3] (spectral.cpp:98-130)."""
    cells = np.array(cells, float); size = np.array(size, float)
    k = np.asarray(k, float)
    if kind == DERIVATIVE_CONTINUOUS:
        return 1j * TAU * k / size
    ang = TAU * k / cells
    if kind == DERIVATIVE_CENTRAL_DIFF:
        return 1j * np.sin(ang) / (2.0 * size / cells)
    if kind == DERIVATIVE_FWBW_DIFF:
        out = np.empty(k.shape, complex)
        for i in range(3):
            j, l = (i + 1) % 3, (i + 2) % 3
            own = (np.cos(ang[..., i]) - 1.0) + 1j * np.sin(ang[..., i])      # upstream; the rewrite has +1 for i>0 (D11)
            a = (np.cos(ang[..., j]) + 1.0) + 1j * np.sin(ang[..., j])
            b = (np.cos(ang[..., l]) + 1.0) + 1j * np.sin(ang[..., l])
            out[..., i] = own * a * b / (4.0 * size[i] / cells[i])
        return out
    raise ValueError('invalid spectral derivative id')


def xi_tables(grid, kind=DERIVATIVE_CONTINUOUS, ky_range=None):
    """-> xi1st, xi2nd of shape [kz, ky, kx, 3] (the reference stores (3,kx,kz,ky_local))."""
    nx, ny, nz = grid.cells
    ky_idx = np.arange(ny) if ky_range is None else np.arange(*ky_range)
    kx = np.arange(grid.nxr)
    ky = np.where(ky_idx > ny // 2, ky_idx - ny, ky_idx)
    kz_idx = np.arange(nz)
    kz = np.where(kz_idx > nz // 2, kz_idx - nz, kz_idx)
    K = np.stack(np.meshgrid(kz, ky, kx, indexing='ij')[::-1], axis=-1)      # [..., (kx,ky,kz)]
    xi2nd = freq_derivative(kind, grid.cells, grid.size, K)
    xi1st = xi2nd.copy()
    if kind == DERIVATIVE_CONTINUOUS:                                        # per-component Nyquist zeroing (D4)
        IDX = np.stack(np.meshgrid(kz_idx, ky_idx, kx, indexing='ij')[::-1], axis=-1)
        for p, c in enumerate(grid.cells):
            if c % 2 == 0:
                xi1st[..., p] = np.where(IDX[..., p] == c // 2, 0.0, xi1st[..., p])
    return xi1st, xi2nd


def fft_forward(field):
    """unnormalised batched r2c over (z,y,x) -> [kz,ky,kx,...]"""
    return _fft.rfftn(field, axes=(0, 1, 2), **_KW)


def fft_backward(field_hat, cells, wgt=1.0):
    """unnormalised c2r (FFTW convention) times wgt"""
    nx, ny, nz = cells
    out = _fft.irfftn(field_hat, s=(nz, ny, nx), axes=(0, 1, 2), **_KW)
    return out * (float(nx) * ny * nz * wgt)


def divergence_rms(field_hat, xi1st, grid, scaled_size):
    """mech_base.cpp:302-398 on an already transformed field [kz,ky,kx,3,3]."""
    rescale = np.array(grid.size) / np.array(scaled_size)
    v = np.conj(-xi1st) * rescale                                 # [kz,ky,kx,3]
    prod = np.einsum('zyxlm,zyxm->zyxl', field_hat, v)
    mag = (prod.real ** 2 + prod.imag ** 2).sum(-1)               # [kz,ky,kx]
    nxr = grid.nxr
    total = 2.0 * mag[:, :, 1:nxr - 1].sum() + mag[:, :, 0].sum() + mag[:, :, nxr - 1].sum()
    rms = np.sqrt(total) * grid.wgt
    if grid.cells[0] == 1:
        rms = rms / 2.0
    return rms, total


def gamma_hat_apply(field_hat, xi1st, C_ref):
    """tau_hat(k) <- Gamma_hat(k) : tau_hat(k) for every k (DC handled by the caller).
    A(l,m) = sum_{n,o} C_ref(l,n,m,o) D(n,o),  D(l,m) = conj(-xi_l) xi_m,
    Gamma(l,m,n,o) = A^-1(l,n) D(o,m)   =>   result = A^-1 . tau_hat . D ;  Gamma := 0 where |det Re A| <= 1e-16."""
    D = np.einsum('...l,...m->...lm', np.conj(-xi1st), xi1st)
    A = np.einsum('lnmo,...no->...lm', C_ref.astype(complex), D)
    det = np.linalg.det(A.real)
    ok = np.abs(det) > 1.0e-16
    Ainv = np.zeros_like(A)
    Ainv[ok] = np.linalg.inv(A[ok])
    out = np.einsum('...ln,...no,...om->...lm', Ainv, field_hat, D)
    return out


def gamma_convolution(field, field_aim, xi1st, C_ref, grid, return_hat=False):
    """mech_base.cpp:400-477: in: real field [z,y,x,3,3]; out: same shape."""
    hat = fft_forward(field)
    out = gamma_hat_apply(hat, xi1st, C_ref)
    out[0, 0, 0] = np.asarray(field_aim, float)
    res = fft_backward(out, grid.cells, 1.0)
    return (res, hat) if return_hat else res


# ---------------------------------------------------------------------------------------------
def to99(T):       # math_3333to99 with MAPPLAIN (row-major pairs) == C-order reshape
    return np.asarray(T).reshape(9, 9)


def to3333(M):
    return np.asarray(M).reshape(3, 3, 3, 3)


def rotate2(R, T, active=False):
    """rotations.f90:288-300 with R = asMatrix()."""
    return R.T @ T @ R if active else R @ T @ R.T


def rotate4(R, T, active=False):
    Q = R.T if active else R
    return np.einsum('im,jn,ko,lp,mnop->ijkl', Q, Q, Q, Q, T)


def masked_compliance(C, R, stress_mask):
    """mech_base.cpp:249-300. stress_mask[i,j] True = component NOT stress-controlled ('x' in the load file)."""
    m = ~np.asarray(stress_mask, bool).reshape(9)              # True where stress is prescribed (row-major pairs)
    S99 = np.zeros((9, 9))
    if m.any():
        C99 = to99(rotate4(R, C))
        idx = np.where(m)[0]
        red = C99[np.ix_(idx, idx)]
        S99[np.ix_(idx, idx)] = np.linalg.inv(red)
    return to3333(S99)


def forward_field(dt, field_last, rate, grid, aim=None):
    f = field_last + rate * dt
    if aim is not None:
        diff = f.sum(axis=(0, 1, 2)) * grid.wgt - aim
        f = f - diff
    return f


def calculate_rate(heterogeneous, field0, field, dt, av_rate):
    if heterogeneous:
        return (field - field0) / dt
    return np.broadcast_to(np.asarray(av_rate, float), field.shape).copy()


def homogenise(P, dPdF, grid, tie_rtol=TIE_RTOL):
    """spectral.cpp:160-218 -> P_av (3,3), C_volAvg, C_minMaxAvg (3,3,3,3), and the two selected cell indices.
    Selection rule (shared with the CUDA path): scanning cells in order, a cell replaces the running max (min)
    only if its norm exceeds (undercuts) it by more than tie_rtol relative -- so ties keep the lowest index."""
    n = grid.n
    P = P.reshape(n, 3, 3); d = dPdF.reshape(n, 81)
    P_av = P.sum(0) * grid.wgt
    C_vol = d.sum(0).reshape(3, 3, 3, 3) * grid.wgt
    nrm = (d * d).sum(1)
    imax, imin = select_minmax(nrm, tie_rtol)
    C_mm = 0.5 * (d[imax] + d[imin]).reshape(3, 3, 3, 3)
    return P_av, C_vol, C_mm, (imax, imin)


def select_minmax(nrm, tie_rtol=TIE_RTOL):
    """order-independent statement of the tie rule: winners are the lowest-index cells whose norm lies within
    tie_rtol (relative) of the global extreme."""
    vmax, vmin = nrm.max(), nrm.min()
    imax = int(np.argmax(nrm >= vmax * (1.0 - tie_rtol)))
    imin = int(np.argmax(nrm <= vmin * (1.0 + tie_rtol)))
    return imax, imin


def update_coords(F, grid):
    """MechBase::base_update_coords (mech_base.cpp:45-154) with upstream DAMASK's arithmetic (utilities_updateCoords):
    cell-centre displacement fluctuation by integration in Fourier space,
        u_hat(k) = F_hat(k) . xi2nd(k) / sum_l conj(-xi2nd_l) xi2nd_l      (k != 0; the C++ port negates the denominator
    and applies wgt twice -- both contradict the real DAMASK result file G4 and are not followed), u_hat(0) = 0,
    u_tilde = c2r(u_hat) * wgt; nodes = F_avg . X_node + mean of the 8 surrounding cells' u_tilde (periodic),
    cell centres = F_avg . X_centre + u_tilde.   F [nz,ny,nx,3,3] -> x_n [nz+1,ny+1,nx+1,3], x_p [nz,ny,nx,3]"""
    nx, ny, nz = grid.cells
    F = np.asarray(F, float).reshape(nz, ny, nx, 3, 3)
    hat = fft_forward(F)
    F_avg = hat[0, 0, 0].real * grid.wgt
    _, xi2 = xi_tables(grid, DERIVATIVE_CONTINUOUS)
    num = np.einsum('zyxlm,zyxm->zyxl', hat, xi2)
    den = (np.conj(-xi2) * xi2).sum(-1)
    den[0, 0, 0] = 1.0
    u_hat = num / den[..., None]
    u_hat[0, 0, 0] = 0.0
    u = fft_backward(u_hat, grid.cells, grid.wgt)                      # [nz,ny,nx,3]
    step = np.array(grid.size) / np.array(grid.cells, float)
    # nodes: average of the 8 cells around node (i,j,k): cells (i-1..i, j-1..j, k-1..k), periodic
    up = u[np.arange(-1, nz + 1) % nz][:, np.arange(-1, ny + 1) % ny][:, :, np.arange(-1, nx + 1) % nx]   # index 0 = cell -1
    acc = np.zeros((nz + 1, ny + 1, nx + 1, 3))
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                acc += up[dz:dz + nz + 1, dy:dy + ny + 1, dx:dx + nx + 1]
    kk, jj, ii = np.meshgrid(np.arange(nz + 1), np.arange(ny + 1), np.arange(nx + 1), indexing='ij')
    Xn = np.stack([ii * step[0], jj * step[1], kk * step[2]], axis=-1)
    x_n = np.einsum('lm,zyxm->zyxl', F_avg, Xn) + 0.125 * acc
    kk, jj, ii = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing='ij')
    Xp = np.stack([(ii + 0.5) * step[0], (jj + 0.5) * step[1], (kk + 0.5) * step[2]], axis=-1)
    x_p = np.einsum('lm,zyxm->zyxl', F_avg, Xp) + u
    return x_n, x_p


# --------------------------------------------------------------------------------------------- polarisation scheme helpers
# (upstream DAMASK only: the reference tree has no polarisation solver -- damask_grid.cpp:8-12 accepts `spectral_basic` alone;
#  its numerics parser does read alpha, beta and the curl tolerances, config.cpp:450-457, config.h:68-71)
def curl_rms(field_hat, xi1st, grid, scaled_size):
    """utilities_curlRMS of upstream DAMASK (grid/spectral_utilities.f90): RMS of curl(F) from the transformed field
    [kz,ky,kx,3,3], the Hermitian half counted like the divergence (interior kx twice, kx = 0 and kx = nx/2 once)."""
    rescale = np.array(grid.size) / np.array(scaled_size)
    x = xi1st * rescale                                           # [kz,ky,kx,3]
    T = field_hat
    c0 = T[..., 2] * x[..., None, 1] - T[..., 1] * x[..., None, 2]
    c1 = T[..., 0] * x[..., None, 2] - T[..., 2] * x[..., None, 0]
    c2 = T[..., 1] * x[..., None, 0] - T[..., 0] * x[..., None, 1]
    mag = sum((c.real ** 2 + c.imag ** 2).sum(-1) for c in (c0, c1, c2))          # [kz,ky,kx]
    nxr = grid.nxr
    total = 2.0 * mag[:, :, 1:nxr - 1].sum() + mag[:, :, 0].sum() + mag[:, :, nxr - 1].sum()
    rms = np.sqrt(total) * grid.wgt
    if grid.cells[0] == 1:
        rms = rms / 2.0
    return rms, total


_NYE = ((0, 0), (1, 1), (2, 2), (0, 1), (1, 2), (2, 0))          # math.f90 MAPNYE: 11 22 33 12 23 31
_NRM_MANDEL = np.array([1.0, 1.0, 1.0, np.sqrt(2.0), np.sqrt(2.0), np.sqrt(2.0)])


def sym3333to66(A):
    """math_sym3333to66 (weighted): picks the Nye components, no symmetrisation"""
    idx = np.array(_NYE)
    M = A[..., idx[:, 0][:, None], idx[:, 1][:, None], idx[:, 0][None, :], idx[:, 1][None, :]]
    return M * (_NRM_MANDEL[:, None] * _NRM_MANDEL[None, :])


def from66toSym3333(M):
    """math_66toSym3333 (weighted): fills all four index permutations of every Nye pair"""
    W = M / (_NRM_MANDEL[:, None] * _NRM_MANDEL[None, :])
    out = np.zeros(M.shape[:-2] + (3, 3, 3, 3))
    for a, (i, j) in enumerate(_NYE):
        for b, (k, l) in enumerate(_NYE):
            v = W[..., a, b]
            out[..., i, j, k, l] = v; out[..., j, i, k, l] = v; out[..., i, j, l, k] = v; out[..., j, i, l, k] = v
    return out


def invsym3333(A):
    """math_invSym3333: inverse of a 3x3x3x3 tensor through its weighted 6x6 Nye/Mandel form"""
    return from66toSym3333(np.linalg.inv(sym3333to66(A)))
