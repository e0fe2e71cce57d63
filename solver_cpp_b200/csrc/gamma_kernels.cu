// Fused z pass of the spectral residual: forward FFT along z  ->  divergence-RMS accumulation  ->  Gamma-operator
// contraction in place  ->  inverse FFT along z, one read and one write of the half spectrum.
//
// Replaces MechBase::calculate_divergence_rms (/root/reference/src/grid/mech_base.cpp:302-398) and
// MechBase::gamma_convolution (mech_base.cpp:400-477, memory-efficient branch) with the reference's unfinished
// Gamma operator completed as in upstream DAMASK:
//     D(l,m)  = conj(-xi_l) xi_m                     (xi = xi1st, Nyquist components zeroed)
//     A(l,m)  = sum_{n,o} C_ref(l,n,m,o) D(n,o)
//     tau_hat <- A^-1 . tau_hat . D                  ( == Gamma_hat : tau_hat, Gamma(l,m,n,o) = A^-1(l,n) D(o,m) )
//     Gamma := 0 where |det Re A| <= 1e-16;  tau_hat(k=0) <- field_aim
// and the single forward transform shared by both consumers (D9).  xi is rebuilt per k-point from three per-axis
// factor tables (a few KB, L1-resident) instead of the reference's 48 B/k-point xi1st array.
//
// The z transform is done in shared memory in place; between forward and inverse the line is in digit-reversed
// order, which is irrelevant to a pointwise operator once kz = perm[slot] is known.
#include "fft_host.h"

struct C3 { double2 v[3]; };

__device__ __forceinline__ void load_xi(const DmkGammaParams& gp, int kx, int ky, int kz, double2 (&xi)[3]) {
  double2 ax = __ldg(&gp.xi_x[kx]), ay = __ldg(&gp.xi_y[ky]), az = __ldg(&gp.xi_z[kz]);
  double2 ox = __ldg(&gp.xo_x[kx]), oy = __ldg(&gp.xo_y[ky]), oz = __ldg(&gp.xo_z[kz]);
  xi[0] = cmul(cmul(ax, oy), oz);
  xi[1] = cmul(cmul(ay, oz), ox);
  xi[2] = cmul(cmul(az, ox), oy);
  if (__ldg(&gp.nyq_x[kx])) xi[0] = make_double2(0.0, 0.0);
  if (__ldg(&gp.nyq_y[ky])) xi[1] = make_double2(0.0, 0.0);
  if (__ldg(&gp.nyq_z[kz])) xi[2] = make_double2(0.0, 0.0);
}

// One k-point, xi purely imaginary (continuous / central-difference stencils): D, A real.
// t = 9 complex in record order (component c = i + 3 j).  Returns the divergence contribution.
__device__ __forceinline__ double gamma_point_imag(double2* __restrict__ t, const double* __restrict__ C,
                                                   const double q[3], const double resc[3], bool is_dc,
                                                   const double* __restrict__ dc_value) {
  // divergence: prod_l = sum_m tau(l,m) * conj(-xi_m) * resc_m = i * sum_m tau(l,m) q_m resc_m
  double div = 0.0;
  const double v0 = q[0] * resc[0], v1 = q[1] * resc[1], v2 = q[2] * resc[2];
#pragma unroll
  for (int l = 0; l < 3; l++) {
    double pr = t[l].x * v0 + t[l + 3].x * v1 + t[l + 6].x * v2;
    double pi = t[l].y * v0 + t[l + 3].y * v1 + t[l + 6].y * v2;
    div += pr * pr + pi * pi;
  }
  if (is_dc) {
#pragma unroll
    for (int c = 0; c < 9; c++) t[c] = make_double2(dc_value[c], 0.0);
    return div;
  }
  // D(l,m) = conj(-i q_l) (i q_m) = -q_l q_m
  double D[3][3];
#pragma unroll
  for (int l = 0; l < 3; l++)
#pragma unroll
    for (int m = 0; m < 3; m++) D[l][m] = -q[l] * q[m];
  double A[3][3];
#pragma unroll
  for (int l = 0; l < 3; l++)
#pragma unroll
    for (int m = 0; m < 3; m++) {
      double acc = 0.0;
#pragma unroll
      for (int n = 0; n < 3; n++)
#pragma unroll
        for (int o = 0; o < 3; o++) acc += C[((l * 3 + n) * 3 + m) * 3 + o] * D[n][o];
      A[l][m] = acc;
    }
  double c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1];
  double c01 = A[1][2] * A[2][0] - A[1][0] * A[2][2];
  double c02 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
  double det = A[0][0] * c00 + A[0][1] * c01 + A[0][2] * c02;
  if (!(fabs(det) > 1.0e-16)) {
#pragma unroll
    for (int c = 0; c < 9; c++) t[c] = make_double2(0.0, 0.0);
    return div;
  }
  double id = 1.0 / det;
  double Ai[3][3];
  Ai[0][0] = c00 * id; Ai[1][0] = c01 * id; Ai[2][0] = c02 * id;
  Ai[0][1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) * id;
  Ai[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) * id;
  Ai[2][1] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) * id;
  Ai[0][2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) * id;
  Ai[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) * id;
  Ai[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * id;
  // out = Ai . tau . D, real and imaginary parts separately (Ai, D real)
#pragma unroll
  for (int part = 0; part < 2; part++) {
    double M[3][3];
#pragma unroll
    for (int n = 0; n < 3; n++)
#pragma unroll
      for (int m = 0; m < 3; m++) {
        double acc = 0.0;
#pragma unroll
        for (int o = 0; o < 3; o++) acc += (part ? t[n + 3 * o].y : t[n + 3 * o].x) * D[o][m];
        M[n][m] = acc;
      }
#pragma unroll
    for (int l = 0; l < 3; l++)
#pragma unroll
      for (int m = 0; m < 3; m++) {
        double acc = Ai[l][0] * M[0][m] + Ai[l][1] * M[1][m] + Ai[l][2] * M[2][m];
        if (part) t[l + 3 * m].y = acc; else t[l + 3 * m].x = acc;
      }
  }
  return div;
}

// general complex xi (forward-backward difference stencil)
__device__ __noinline__ double gamma_point_cplx(double2* __restrict__ t, const double* __restrict__ C,
                                                const double2 xi[3], const double resc[3], bool is_dc,
                                                const double* __restrict__ dc_value) {
  double div = 0.0;
  double2 v[3];
  for (int m = 0; m < 3; m++) v[m] = cscale(cconj(make_double2(-xi[m].x, -xi[m].y)), resc[m]);
  for (int l = 0; l < 3; l++) {
    double2 p = cadd(cadd(cmul(t[l], v[0]), cmul(t[l + 3], v[1])), cmul(t[l + 6], v[2]));
    div += p.x * p.x + p.y * p.y;
  }
  if (is_dc) { for (int c = 0; c < 9; c++) t[c] = make_double2(dc_value[c], 0.0); return div; }
  double2 D[3][3], A[3][3], Ai[3][3], M[3][3];
  for (int l = 0; l < 3; l++) for (int m = 0; m < 3; m++) D[l][m] = cmul(cconj(make_double2(-xi[l].x, -xi[l].y)), xi[m]);
  for (int l = 0; l < 3; l++) for (int m = 0; m < 3; m++) {
    double2 acc = make_double2(0.0, 0.0);
    for (int n = 0; n < 3; n++) for (int o = 0; o < 3; o++) acc = cadd(acc, cscale(D[n][o], C[((l * 3 + n) * 3 + m) * 3 + o]));
    A[l][m] = acc;
  }
  double detre = A[0][0].x * (A[1][1].x * A[2][2].x - A[1][2].x * A[2][1].x)
               - A[0][1].x * (A[1][0].x * A[2][2].x - A[1][2].x * A[2][0].x)
               + A[0][2].x * (A[1][0].x * A[2][1].x - A[1][1].x * A[2][0].x);
  if (!(fabs(detre) > 1.0e-16)) { for (int c = 0; c < 9; c++) t[c] = make_double2(0.0, 0.0); return div; }
  double2 c00 = csub(cmul(A[1][1], A[2][2]), cmul(A[1][2], A[2][1]));
  double2 c01 = csub(cmul(A[1][2], A[2][0]), cmul(A[1][0], A[2][2]));
  double2 c02 = csub(cmul(A[1][0], A[2][1]), cmul(A[1][1], A[2][0]));
  double2 det = cadd(cadd(cmul(A[0][0], c00), cmul(A[0][1], c01)), cmul(A[0][2], c02));
  double dn = det.x * det.x + det.y * det.y;
  double2 id = make_double2(det.x / dn, -det.y / dn);
  Ai[0][0] = cmul(c00, id); Ai[1][0] = cmul(c01, id); Ai[2][0] = cmul(c02, id);
  Ai[0][1] = cmul(csub(cmul(A[0][2], A[2][1]), cmul(A[0][1], A[2][2])), id);
  Ai[1][1] = cmul(csub(cmul(A[0][0], A[2][2]), cmul(A[0][2], A[2][0])), id);
  Ai[2][1] = cmul(csub(cmul(A[0][1], A[2][0]), cmul(A[0][0], A[2][1])), id);
  Ai[0][2] = cmul(csub(cmul(A[0][1], A[1][2]), cmul(A[0][2], A[1][1])), id);
  Ai[1][2] = cmul(csub(cmul(A[0][2], A[1][0]), cmul(A[0][0], A[1][2])), id);
  Ai[2][2] = cmul(csub(cmul(A[0][0], A[1][1]), cmul(A[0][1], A[1][0])), id);
  for (int n = 0; n < 3; n++) for (int m = 0; m < 3; m++) {
    double2 acc = make_double2(0.0, 0.0);
    for (int o = 0; o < 3; o++) acc = cadd(acc, cmul(t[n + 3 * o], D[o][m]));
    M[n][m] = acc;
  }
  for (int l = 0; l < 3; l++) for (int m = 0; m < 3; m++)
    t[l + 3 * m] = cadd(cadd(cmul(Ai[l][0], M[0][m]), cmul(Ai[l][1], M[1][m])), cmul(Ai[l][2], M[2][m]));
  return div;
}

template <bool IMAG_XI>
__global__ void __launch_bounds__(256, 2) dmk_gamma_zpass_kernel(double2* __restrict__ data, DmkFftPlan pl, DmkLineMap map,
                                                              int n0, int n_outer, int KT, DmkGammaParams gp) {
  extern __shared__ double2 s[];
  __shared__ double sC[81];
  __shared__ double sred[8];
  const int n = pl.n;
  const int tiles0 = (n0 + KT - 1) / KT;
  const long long ntiles = (long long)tiles0 * n_outer;
  const int nxr = n0;
  for (int i = threadIdx.x; i < 81; i += blockDim.x) sC[i] = gp.C_ref[i];
  const double resc[3] = {gp.rescale[0], gp.rescale[1], gp.rescale[2]};
  double div_local = 0.0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int o = (int)(tile / tiles0), i0 = (int)(tile - (long long)o * tiles0) * KT;
    const int kt = min(KT, n0 - i0);
    const int B = kt * DMK_NC;
    for (int u = threadIdx.x; u < n * B; u += blockDim.x) {
      int pos = u / B, w = u - pos * B;
      s[u] = data[dmk_line_rec(map, o, pos, i0) * DMK_NC + w];
    }
    __syncthreads();
    fft_forward_smem(pl, s, B);
    const int ky = o + gp.ky_offset;
    for (int pnt = threadIdx.x; pnt < n * kt; pnt += blockDim.x) {
      int slot = pnt / kt, j = pnt - slot * kt;
      int kz = __ldg(&pl.perm[slot]);
      int kx = i0 + j;
      double2 t[9];
      double2* ps = s + (size_t)pnt * DMK_NC;
#pragma unroll
      for (int c = 0; c < 9; c++) t[c] = ps[c];
      double2 xi[3];
      load_xi(gp, kx, ky, kz, xi);
      const bool is_dc = (kx == 0) && (ky == 0) && (kz == 0);
      double wgt = ((kx == 0) ? 1.0 : 0.0) + ((kx == nxr - 1) ? 1.0 : 0.0) + ((kx > 0 && kx < nxr - 1) ? 2.0 : 0.0);
      double d;
      if (IMAG_XI) {
        const double q[3] = {xi[0].y, xi[1].y, xi[2].y};
        d = gamma_point_imag(t, sC, q, resc, is_dc, gp.dc_value);
      } else {
        d = gamma_point_cplx(t, sC, xi, resc, is_dc, gp.dc_value);
      }
      div_local += wgt * d;
#pragma unroll
      for (int c = 0; c < 9; c++) ps[c] = t[c];
    }
    __syncthreads();
    fft_inverse_smem(pl, s, B);
    for (int u = threadIdx.x; u < n * B; u += blockDim.x) {
      int pos = u / B, w = u - pos * B;
      data[dmk_line_rec(map, o, pos, i0) * DMK_NC + w] = s[u];
    }
    __syncthreads();
  }
  // block reduction of the divergence sum, one atomic per CTA
  for (int off = 16; off > 0; off >>= 1) div_local += __shfl_down_sync(0xffffffffu, div_local, off);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = div_local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); w++) tot += sred[w];
    atomicAdd(gp.div_accum, tot);
  }
}

int dmk_launch_gamma_zpass(double2* data, const DmkFftPlan& plz, DmkLineMap map, int n0, int n_outer, int KT,
                           const DmkGammaParams& gp, cudaStream_t st, bool imag_xi);

template <typename K>
static int grid_for_g(K kernel, int threads, size_t smem, long long work, int* grid) {
  int dev = 0, sms = 0, occ = 0;
  DMK_CUDA_CHECK(cudaGetDevice(&dev));
  DMK_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (smem > 48 * 1024) DMK_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  DMK_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem));
  if (occ < 1) { dmk_set_error("gamma kernel does not fit: %zu B shared memory", smem); return DMK_ERR_UNSUPPORTED; }
  long long g = (long long)sms * occ;
  *grid = (int)(work < g ? work : g);
  if (*grid < 1) *grid = 1;
  return DMK_OK;
}

int dmk_launch_gamma_zpass(double2* data, const DmkFftPlan& plz, DmkLineMap map, int n0, int n_outer, int KT,
                           const DmkGammaParams& gp, cudaStream_t st, bool imag_xi) {
  if (KT < 1) KT = 1;
  if (KT > n0) KT = n0;
  size_t smem = (size_t)plz.n * KT * DMK_NC * sizeof(double2);
  long long tiles = (long long)((n0 + KT - 1) / KT) * n_outer;
  int grid, rc;
  if (imag_xi) {
    rc = grid_for_g(dmk_gamma_zpass_kernel<true>, 256, smem, tiles, &grid); if (rc) return rc;
    dmk_gamma_zpass_kernel<true><<<grid, 256, smem, st>>>(data, plz, map, n0, n_outer, KT, gp);
  } else {
    rc = grid_for_g(dmk_gamma_zpass_kernel<false>, 256, smem, tiles, &grid); if (rc) return rc;
    dmk_gamma_zpass_kernel<false><<<grid, 256, smem, st>>>(data, plz, map, n0, n_outer, KT, gp);
  }
  DMK_CUDA_CHECK(cudaGetLastError());
  return DMK_OK;
}
