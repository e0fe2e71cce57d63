// Batched 3-D r2c / c2r FFT passes for the 9-component tensor field, hand-written for sm_100a.
//
// Replaces the FFTW-MPI plans of the reference (/root/reference/src/grid/fft.cpp:48-58; FFT<5>::forward/backward
// fft.h:37-60).  Data stays array-of-records: a record is the 9 components of one point (72 B real, 144 B complex),
// so every global access of every pass is a run of whole records -- contiguous rows in the x pass, runs of KT
// neighbouring kx records in the strided (y / z) passes -- and all 9 components of a k-point sit together for the
// Gamma operator.  Each pass reads and writes the field exactly once (144 B per point): x r2c, y, z (fused with
// Gamma, gamma_kernels.cu), y inverse, x c2r.
//
// x pass: two real rows are packed as re/im of one complex line (any nx, odd included), transformed in shared
// memory and separated with the Hermitian identities A_k = (Z_k + conj Z_{n-k})/2, B_k = (Z_k - conj Z_{n-k})/(2i).
#include <vector>
#include <cmath>
#include "fft.cuh"
#include "fft_host.h"
#include "fft_fast.cuh"

// ------------------------------------------------------------------------------------------------ x pass, r2c
__global__ void __launch_bounds__(256, 2) dmk_xpass_r2c_kernel(const double* __restrict__ in, double2* __restrict__ out,
                                                            DmkFftPlan pl, int nrows) {
  extern __shared__ double2 s[];
  const int nx = pl.n, nxr = nx / 2 + 1, ne = nx * DMK_NC;
  const int npairs = (nrows + 1) / 2;
  for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
    const int ra = 2 * pair;
    const bool hasb = (ra + 1) < nrows;
    const double* a = in + (size_t)ra * ne;
    const double* b = a + ne;
    if ((ne & 1) == 0) {
      const double2* a2 = reinterpret_cast<const double2*>(a);
      const double2* b2 = reinterpret_cast<const double2*>(b);
      for (int e = threadIdx.x; e < ne / 2; e += blockDim.x) {
        double2 va = __ldg(a2 + e);
        double2 vb = hasb ? __ldg(b2 + e) : make_double2(0.0, 0.0);
        s[2 * e] = make_double2(va.x, vb.x);
        s[2 * e + 1] = make_double2(va.y, vb.y);
      }
    } else {
      for (int e = threadIdx.x; e < ne; e += blockDim.x) s[e] = make_double2(a[e], hasb ? b[e] : 0.0);
    }
    __syncthreads();
    fft_forward_smem(pl, s, DMK_NC);
    double2* oa = out + (size_t)ra * nxr * DMK_NC;
    double2* ob = oa + (size_t)nxr * DMK_NC;
    for (int t = threadIdx.x; t < nxr * DMK_NC; t += blockDim.x) {
      int k = t / DMK_NC, c = t - k * DMK_NC;
      int kn = k ? nx - k : 0;
      double2 zk = s[__ldg(&pl.iperm[k]) * DMK_NC + c];
      double2 zn = s[__ldg(&pl.iperm[kn]) * DMK_NC + c];
      oa[t] = make_double2(0.5 * (zk.x + zn.x), 0.5 * (zk.y - zn.y));
      if (hasb) ob[t] = make_double2(0.5 * (zk.y + zn.y), -0.5 * (zk.x - zn.x));
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ x pass, c2r
__global__ void __launch_bounds__(256, 2) dmk_xpass_c2r_kernel(const double2* __restrict__ in, double* __restrict__ out,
                                                            DmkFftPlan pl, int nrows, double scale) {
  extern __shared__ double2 s[];
  const int nx = pl.n, nxr = nx / 2 + 1, ne = nx * DMK_NC;
  const int npairs = (nrows + 1) / 2;
  for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
    const int ra = 2 * pair;
    const bool hasb = (ra + 1) < nrows;
    const double2* ia = in + (size_t)ra * nxr * DMK_NC;
    const double2* ib = ia + (size_t)nxr * DMK_NC;
    for (int t = threadIdx.x; t < nxr * DMK_NC; t += blockDim.x) {
      int k = t / DMK_NC, c = t - k * DMK_NC;
      double2 A = __ldg(ia + t);
      double2 Bv = hasb ? __ldg(ib + t) : make_double2(0.0, 0.0);
      const bool selfconj = (k == 0) || (2 * k == nx);
      if (selfconj) { A.y = 0.0; Bv.y = 0.0; }             // FFTW c2r ignores these imaginary parts
      s[__ldg(&pl.iperm[k]) * DMK_NC + c] = make_double2(A.x - Bv.y, A.y + Bv.x);
      if (!selfconj) s[__ldg(&pl.iperm[nx - k]) * DMK_NC + c] = make_double2(A.x + Bv.y, Bv.x - A.y);
    }
    __syncthreads();
    fft_inverse_smem(pl, s, DMK_NC);
    double* a = out + (size_t)ra * ne;
    double* b = a + ne;
    if ((ne & 1) == 0) {
      double2* a2 = reinterpret_cast<double2*>(a);
      double2* b2 = reinterpret_cast<double2*>(b);
      for (int e = threadIdx.x; e < ne / 2; e += blockDim.x) {
        double2 z0 = s[2 * e], z1 = s[2 * e + 1];
        a2[e] = make_double2(z0.x * scale, z1.x * scale);
        if (hasb) b2[e] = make_double2(z0.y * scale, z1.y * scale);
      }
    } else {
      for (int e = threadIdx.x; e < ne; e += blockDim.x) {
        double2 z = s[e];
        a[e] = z.x * scale;
        if (hasb) b[e] = z.y * scale;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ strided pass (y, z)
// A tile = all n positions of a line x kt neighbouring records along the contiguous axis.
template <bool INV>
__global__ void __launch_bounds__(256, 2) dmk_line_pass_kernel(const double2* __restrict__ src, double2* __restrict__ dst,
                                                            DmkFftPlan pl, DmkLineMap ms, DmkLineMap md, int n0,
                                                            int n_outer, int KT) {
  extern __shared__ double2 s[];
  const int n = pl.n;
  const int tiles0 = (n0 + KT - 1) / KT;
  const long long ntiles = (long long)tiles0 * n_outer;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int o = (int)(tile / tiles0), i0 = (int)(tile - (long long)o * tiles0) * KT;
    const int kt = min(KT, n0 - i0);
    const int B = kt * DMK_NC;
    for (int u = threadIdx.x; u < n * B; u += blockDim.x) {
      int pos = u / B, w = u - pos * B;
      int line = INV ? __ldg(&pl.perm[pos]) : pos;        // inverse: slot `pos` must receive frequency perm[pos]
      s[u] = __ldg(src + dmk_line_rec(ms, o, line, i0) * DMK_NC + w);
    }
    __syncthreads();
    if (INV) fft_inverse_smem(pl, s, B); else fft_forward_smem(pl, s, B);
    for (int u = threadIdx.x; u < n * B; u += blockDim.x) {
      int pos = u / B, w = u - pos * B;
      int line = INV ? pos : __ldg(&pl.perm[pos]);
      dst[dmk_line_rec(md, o, line, i0) * DMK_NC + w] = s[u];
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ host side
static int factorize(int n, int* radix) {
  static const int cand[] = {8, 4, 2, 3, 5, 7, 11, 13};
  int np = 0;
  for (int c : cand) {
    while (n % c == 0 && n > 1) {
      if (np >= DMK_FFT_MAX_PASSES) return -1;
      radix[np++] = c; n /= c;
    }
  }
  if (n != 1) return -1;
  return np;
}

int dmk_fft_plan_create(DmkFftPlan* pl, int n) {
  if (n < 1) { dmk_set_error("fft length %d", n); return DMK_ERR_ARG; }
  pl->n = n;
  pl->npass = (n == 1) ? 0 : factorize(n, pl->radix);
  if (pl->npass < 0) { dmk_set_error("fft length %d has a prime factor > 13 (unsupported)", n); return DMK_ERR_UNSUPPORTED; }
  std::vector<int> perm(n), iperm(n);
  for (int s = 0; s < n; s++) {          // slot s = q1*(n/r1) + q2*(n/(r1 r2)) + ...  ->  k = q1 + r1*(q2 + r2*(...))
    int rem = s, sub = n, k = 0, mult = 1;
    for (int p = 0; p < pl->npass; p++) {
      sub /= pl->radix[p];
      int q = rem / sub; rem -= q * sub;
      k += q * mult; mult *= pl->radix[p];
    }
    perm[s] = k; iperm[k] = s;
  }
  std::vector<double2> tw(n);
  for (int k = 0; k < n; k++) {
    // exact octant symmetry is not needed; long double keeps the table correctly rounded
    long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)n;
    tw[k] = make_double2((double)cosl(a), (double)sinl(a));
  }
  int *dperm = nullptr, *diperm = nullptr; double2* dtw = nullptr;
  DMK_CUDA_CHECK(cudaMalloc(&dperm, n * sizeof(int)));
  DMK_CUDA_CHECK(cudaMalloc(&diperm, n * sizeof(int)));
  DMK_CUDA_CHECK(cudaMalloc(&dtw, n * sizeof(double2)));
  DMK_CUDA_CHECK(cudaMemcpy(dperm, perm.data(), n * sizeof(int), cudaMemcpyHostToDevice));
  DMK_CUDA_CHECK(cudaMemcpy(diperm, iperm.data(), n * sizeof(int), cudaMemcpyHostToDevice));
  DMK_CUDA_CHECK(cudaMemcpy(dtw, tw.data(), n * sizeof(double2), cudaMemcpyHostToDevice));
  pl->perm = dperm; pl->iperm = diperm; pl->tw = dtw;
  return DMK_OK;
}

void dmk_fft_plan_destroy(DmkFftPlan* pl) {
  if (pl->perm) cudaFree((void*)pl->perm);
  if (pl->iperm) cudaFree((void*)pl->iperm);
  if (pl->tw) cudaFree((void*)pl->tw);
  pl->perm = pl->iperm = nullptr; pl->tw = nullptr;
}

template <typename K>
static int grid_for(K kernel, int threads, size_t smem, long long work, int* grid) {
  int dev = 0, sms = 0, occ = 0;
  DMK_CUDA_CHECK(cudaGetDevice(&dev));
  DMK_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (smem > 48 * 1024) DMK_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  DMK_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem));
  if (occ < 1) { dmk_set_error("kernel does not fit: %zu B shared memory", smem); return DMK_ERR_UNSUPPORTED; }
  long long g = (long long)sms * occ;
  *grid = (int)(work < g ? work : g);
  if (*grid < 1) *grid = 1;
  return DMK_OK;
}

int dmk_launch_xpass_r2c(const double* in, double2* out, const DmkFftPlan& pl, int nrows, cudaStream_t st) {
  { bool done = false; int rc = dmk_fast_xpass_r2c(in, out, pl, nrows, st, &done); if (rc || done) return rc; }
  size_t smem = (size_t)pl.n * DMK_NC * sizeof(double2);
  int grid; int rc = grid_for(dmk_xpass_r2c_kernel, 256, smem, (nrows + 1) / 2, &grid); if (rc) return rc;
  dmk_xpass_r2c_kernel<<<grid, 256, smem, st>>>(in, out, pl, nrows);
  DMK_CUDA_CHECK(cudaGetLastError());
  return DMK_OK;
}

int dmk_launch_xpass_c2r(const double2* in, double* out, const DmkFftPlan& pl, int nrows, double scale, cudaStream_t st) {
  { bool done = false; int rc = dmk_fast_xpass_c2r(in, out, pl, nrows, scale, st, &done); if (rc || done) return rc; }
  size_t smem = (size_t)pl.n * DMK_NC * sizeof(double2);
  int grid; int rc = grid_for(dmk_xpass_c2r_kernel, 256, smem, (nrows + 1) / 2, &grid); if (rc) return rc;
  dmk_xpass_c2r_kernel<<<grid, 256, smem, st>>>(in, out, pl, nrows, scale);
  DMK_CUDA_CHECK(cudaGetLastError());
  return DMK_OK;
}

int dmk_launch_line_pass(bool inverse, const double2* src, double2* dst, const DmkFftPlan& pl, DmkLineMap ms,
                         DmkLineMap md, int n0, int n_outer, int KT, cudaStream_t st, const DmkPeerStore* ps) {
  { bool done = false; int rc = dmk_fast_line_pass(inverse, src, dst, pl, ms, md, n0, n_outer, st, &done, ps); if (rc || done) return rc; }
  if (ps && ps->remote) { dmk_set_error("the fused exchange needs a power-of-two line length (16..512), got %d", pl.n); return DMK_ERR_UNSUPPORTED; }
  if (KT < 1) KT = 1;
  if (KT > n0) KT = n0;
  size_t smem = (size_t)pl.n * KT * DMK_NC * sizeof(double2);
  long long tiles = (long long)((n0 + KT - 1) / KT) * n_outer;
  int grid, rc;
  if (inverse) {
    rc = grid_for(dmk_line_pass_kernel<true>, 256, smem, tiles, &grid); if (rc) return rc;
    dmk_line_pass_kernel<true><<<grid, 256, smem, st>>>(src, dst, pl, ms, md, n0, n_outer, KT);
  } else {
    rc = grid_for(dmk_line_pass_kernel<false>, 256, smem, tiles, &grid); if (rc) return rc;
    dmk_line_pass_kernel<false><<<grid, 256, smem, st>>>(src, dst, pl, ms, md, n0, n_outer, KT);
  }
  DMK_CUDA_CHECK(cudaGetLastError());
  return DMK_OK;
}
