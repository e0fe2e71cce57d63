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
#include "gamma_point.cuh"
#include "fft_fast.cuh"

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
  // imaginary xi: the 54-entry symmetrised table behind the 81 stiffness entries (dmk_update_gamma)
  for (int i = threadIdx.x; i < (IMAG_XI ? 54 : 81); i += blockDim.x) sC[i] = gp.C_ref[IMAG_XI ? 81 + i : i];
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
    gp.div_partial[blockIdx.x] = tot;
  }
}

__global__ void __launch_bounds__(1024) dmk_sum_div_partials_kernel(const double* __restrict__ partial, int n, double* __restrict__ accum) {
  __shared__ double sh[1024];
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) a += partial[i];
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int w = 512; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) *accum += sh[0];
}
int dmk_sum_div_partials(const double* partial, int n, double* div_accum, cudaStream_t st) {
  dmk_sum_div_partials_kernel<<<1, 1024, 0, st>>>(partial, n, div_accum);
  DMK_CUDA_CHECK(cudaGetLastError());
  return DMK_OK;
}


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
                           const DmkGammaParams& gp, cudaStream_t st, bool imag_xi, const DmkPeerStore* ps) {
  {
    bool done = false;
    int rc = dmk_fast_gamma_zpass(data, plz, map, n0, n_outer, gp, st, imag_xi, &done, ps);
    if (rc || done) return rc;
  }
  if (ps && ps->remote) { dmk_set_error("the fused exchange needs a power-of-two line length (16..512), got %d", plz.n); return DMK_ERR_UNSUPPORTED; }
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
  return dmk_sum_div_partials(gp.div_partial, grid, gp.div_accum, st);
}
