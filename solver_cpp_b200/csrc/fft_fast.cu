// Register-resident batched FFT passes for power-of-two line lengths (see fft_fast.cuh for the scheme).
// Replaces FFT<5>::forward / backward of the reference (/root/reference/src/grid/fft.h:37-60, plans fft.cpp:48-58).
//
//   x r2c : rows of nx reals x 9 components; two rows packed as re/im of one complex line; HBM -> registers (first
//           stage) -> shared -> Hermitian separation -> HBM.  Loads and stores are whole contiguous rows.
//   y / z : tile = all N positions of a line x KT neighbouring kx records (KT*9 columns of double2); HBM -> registers
//           -> shared -> registers -> HBM; every global access is a run of KT*144 bytes.
//   x c2r : the reverse of r2c.
// One tile (or NP row pairs) per CTA, 2-3 CTAs per SM, so that one CTA's loads overlap another's butterflies/stores.
#include <cstdlib>
#include <cstring>
#include "fft_fast.cuh"

bool dmk_fast_disabled() {
  static const bool v = [] { const char* e = getenv("DMK_FFT"); return e && !strcmp(e, "generic"); }();
  return v;
}

// ------------------------------------------------------------------------------------------------ strided pass (y, z)
template <class PL, int KT, bool INV, int NT, int MINB, bool REMOTE>
__global__ void __launch_bounds__(NT, MINB) dmk_fast_line_kernel(const double2* __restrict__ src, double2* __restrict__ dst,
                                                           const double2* __restrict__ tw, DmkLineMap ms, DmkLineMap md,
                                                           int n0, int tiles0, const __grid_constant__ DmkPeerStore ps) {
  constexpr int N = PL::N, R1 = PL::R1, R2 = PL::R2, R3 = PL::R3, B = KT * DMK_NC;
  extern __shared__ double2 s[];
  const int o = blockIdx.x / tiles0, i0 = (blockIdx.x - o * tiles0) * KT;
  const int bact = min(KT, n0 - i0) * DMK_NC;
  const bool ns_s = ms.split >= N, ns_d = md.split >= N;
  auto ldg = [&](int, int p, int col) { return __ldg(src + dmk_line_rec_fast(ms, ns_s, o, p, i0) * DMK_NC + col); };
  auto lds = [&](int, int p, int col) { return s[p * B + col]; };
  auto sts = [&](int, int p, int col, double2 v) { s[p * B + col] = v; };
  auto stg = [&](int, int p, int col, double2 v) { dmk_line_store_ptr<REMOTE>(md, ps, dst, ns_d, o, PL::freq(p), i0)[col] = v; };
  fstage<N, R1, R2 * R3, B, INV, false>(1, bact, tw, ldg, sts);
  __syncthreads();
  if constexpr (R3 > 1) {
    fstage<N, R2, R3, B, INV, false>(1, bact, tw, lds, sts);
    __syncthreads();
    fstage<N, R3, 1, B, INV, false>(1, bact, tw, lds, stg);
  } else {
    fstage<N, R2, 1, B, INV, false>(1, bact, tw, lds, stg);
  }
}

// ------------------------------------------------------------------------------------------------ x pass, r2c
template <class PL, int NP, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) dmk_fast_xr2c_kernel(const double* __restrict__ in, double2* __restrict__ out,
                                                           const double2* __restrict__ tw, int nrows) {
  constexpr int N = PL::N, R1 = PL::R1, R2 = PL::R2, R3 = PL::R3, B = DMK_NC, NXR = N / 2 + 1, NE = N * DMK_NC;
  extern __shared__ double2 s[];
  const int pair0 = blockIdx.x * NP;
  const int npairs = (nrows + 1) / 2;
  const int np = min(NP, npairs - pair0);
  auto ldg = [&](int sub, int p, int col) {
    const int ra = 2 * (pair0 + sub);
    const double* a = in + (size_t)ra * NE + p * DMK_NC + col;
    return make_double2(__ldg(a), (ra + 1 < nrows) ? __ldg(a + NE) : 0.0);
  };
  auto lds = [&](int sub, int p, int col) { return s[(sub * N + p) * B + col]; };
  auto sts = [&](int sub, int p, int col, double2 v) { s[(sub * N + p) * B + col] = v; };
  fstage<N, R1, R2 * R3, B, false, false>(np, B, tw, ldg, sts);
  __syncthreads();
  if constexpr (R3 > 1) {
    fstage<N, R2, R3, B, false, false>(np, B, tw, lds, sts);
    __syncthreads();
    fstage<N, R3, 1, B, false, false>(np, B, tw, lds, sts);
  } else {
    fstage<N, R2, 1, B, false, false>(np, B, tw, lds, sts);
  }
  __syncthreads();
  // Hermitian separation of the two packed rows: A_k = (Z_k + conj Z_{N-k}) / 2, B_k = (Z_k - conj Z_{N-k}) / (2i)
  constexpr int PER = NXR * DMK_NC;
  for (int t = threadIdx.x; t < np * PER; t += blockDim.x) {
    const int sub = t / PER, r = t - sub * PER;
    const int k = r / DMK_NC, c = r - k * DMK_NC;
    const int kn = (N - k) & (N - 1);
    const double2 zk = s[(sub * N + PL::slot(k)) * B + c], zn = s[(sub * N + PL::slot(kn)) * B + c];
    const int ra = 2 * (pair0 + sub);
    double2* oa = out + (size_t)ra * PER;
    oa[r] = make_double2(0.5 * (zk.x + zn.x), 0.5 * (zk.y - zn.y));
    if (ra + 1 < nrows) oa[PER + r] = make_double2(0.5 * (zk.y + zn.y), -0.5 * (zk.x - zn.x));
  }
}

// ------------------------------------------------------------------------------------------------ x pass, c2r
template <class PL, int NP, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) dmk_fast_xc2r_kernel(const double2* __restrict__ in, double* __restrict__ out,
                                                           const double2* __restrict__ tw, int nrows, double scale) {
  constexpr int N = PL::N, R1 = PL::R1, R2 = PL::R2, R3 = PL::R3, B = DMK_NC, NXR = N / 2 + 1, NE = N * DMK_NC;
  extern __shared__ double2 s[];
  const int pair0 = blockIdx.x * NP;
  const int npairs = (nrows + 1) / 2;
  const int np = min(NP, npairs - pair0);
  constexpr int PER = NXR * DMK_NC;
  // Z_k = A_k + i B_k, Z_{N-k} = conj(A_k) + i conj(B_k), natural order
  // UL elements per thread and trip, their 2 UL loads issued together: with one element per trip a thread had two loads in flight
  // and the pass was bound by load latency (ncu: long_scoreboard the top stall, 51 % of DRAM throughput with 27 warps per SM)
  constexpr int UL = 4;
  for (int t0 = threadIdx.x; t0 < np * PER; t0 += UL * blockDim.x) {
    double2 A[UL], Bv[UL];
#pragma unroll
    for (int u = 0; u < UL; u++) {
      const int t = t0 + u * blockDim.x;
      if (t < np * PER) {
        const int sub = t / PER, r = t - sub * PER;
        const int ra = 2 * (pair0 + sub);
        const double2* ia = in + (size_t)ra * PER;
        A[u] = __ldg(ia + r);
        Bv[u] = (ra + 1 < nrows) ? __ldg(ia + PER + r) : make_double2(0.0, 0.0);
      }
    }
#pragma unroll
    for (int u = 0; u < UL; u++) {
      const int t = t0 + u * blockDim.x;
      if (t < np * PER) {
        const int sub = t / PER, r = t - sub * PER;
        const int k = r / DMK_NC, c = r - k * DMK_NC;
        const bool selfconj = (k == 0) || (2 * k == N);
        if (selfconj) { A[u].y = 0.0; Bv[u].y = 0.0; }   // FFTW c2r ignores these imaginary parts
        s[(sub * N + k) * B + c] = make_double2(A[u].x - Bv[u].y, A[u].y + Bv[u].x);
        if (!selfconj) s[(sub * N + N - k) * B + c] = make_double2(A[u].x + Bv[u].y, Bv[u].x - A[u].y);
      }
    }
  }
  __syncthreads();
  auto lds = [&](int sub, int p, int col) { return s[(sub * N + p) * B + col]; };
  auto sts = [&](int sub, int p, int col, double2 v) { s[(sub * N + p) * B + col] = v; };
  auto stg = [&](int sub, int p, int col, double2 v) {
    const int ra = 2 * (pair0 + sub);
    double* a = out + (size_t)ra * NE + PL::freq(p) * DMK_NC + col;
    a[0] = v.x * scale;
    if (ra + 1 < nrows) a[NE] = v.y * scale;
  };
  fstage<N, R1, R2 * R3, B, true, false>(np, B, tw, lds, sts);
  __syncthreads();
  if constexpr (R3 > 1) {
    fstage<N, R2, R3, B, true, false>(np, B, tw, lds, sts);
    __syncthreads();
    // Last stage of a three-stage network: butterfly b holds the frequencies b / R2 + R1 (b % R2) + R1 R2 q, so neighbouring
    // butterflies stored 72-byte runs R1 * 72 bytes apart (ncu: mio_throttle the second stall reason, 51 % of DRAM throughput).
    // Item g of the stage therefore takes butterfly R2 (g % R1) + g / R1, whose outputs are the frequencies g + (N / R3) q:
    // a warp's stores are one contiguous run again, at the price of a two-way bank conflict where two butterflies' reads meet.
    auto lds3 = [&](int sub, int p, int col) {
      const int g = p / R3, m = p - g * R3;
      return s[(sub * N + (R2 * (g % R1) + g / R1) * R3 + m) * B + col];
    };
    auto stg3 = [&](int sub, int p, int col, double2 v) {
      const int g = p / R3, q = p - g * R3;
      const int ra = 2 * (pair0 + sub);
      double* a = out + (size_t)ra * NE + (g + (N / R3) * q) * DMK_NC + col;
      a[0] = v.x * scale;
      if (ra + 1 < nrows) a[NE] = v.y * scale;
    };
    fstage<N, R3, 1, B, true, false>(np, B, tw, lds3, stg3);
  } else {
    fstage<N, R2, 1, B, true, false>(np, B, tw, lds, stg);
  }
}

// ------------------------------------------------------------------------------------------------ host side
template <typename K>
static int set_smem(K kernel, size_t smem) {
  if (smem > 48 * 1024) DMK_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return DMK_OK;
}

template <class PL, int KT, int NT, int MINB = 1>
static int launch_line(bool inverse, const double2* src, double2* dst, const DmkFftPlan& pl, DmkLineMap ms, DmkLineMap md, int n0,
                       int n_outer, cudaStream_t st, const DmkPeerStore& ps) {
  const size_t smem = (size_t)PL::N * KT * DMK_NC * sizeof(double2);
  const int tiles0 = (n0 + KT - 1) / KT;
  const long long grid = (long long)tiles0 * n_outer;
  if (grid > 0x7fffffffLL) { dmk_set_error("line pass grid too large"); return DMK_ERR_UNSUPPORTED; }
  int rc;
  // REMOTE (fused exchange over peer memory) is a compile-time variant: the single-GPU passes must not pay for the branch
#define DMK_LINE_LAUNCH(INV_, REM_)                                                                                          \
  do {                                                                                                                       \
    if ((rc = set_smem(dmk_fast_line_kernel<PL, KT, INV_, NT, MINB, REM_>, smem))) return rc;                               \
    dmk_fast_line_kernel<PL, KT, INV_, NT, MINB, REM_><<<(int)grid, NT, smem, st>>>(src, dst, pl.tw, ms, md, n0, tiles0, ps); \
  } while (0)
  if (inverse) { if (ps.remote) DMK_LINE_LAUNCH(true, true); else DMK_LINE_LAUNCH(true, false); }
  else { if (ps.remote) DMK_LINE_LAUNCH(false, true); else DMK_LINE_LAUNCH(false, false); }
#undef DMK_LINE_LAUNCH
  DMK_CUDA_CHECK(cudaGetLastError());
  return DMK_OK;
}

int dmk_fast_line_pass(bool inverse, const double2* src, double2* dst, const DmkFftPlan& pl, DmkLineMap ms, DmkLineMap md,
                       int n0, int n_outer, cudaStream_t st, bool* done, const DmkPeerStore* psp) {
  *done = true;
  if (dmk_fast_disabled()) { *done = false; return DMK_OK; }
  DmkPeerStore ps;
  if (psp) ps = *psp; else memset(&ps, 0, sizeof ps);
  // tile width, threads and CTAs/SM per length: tools/fft_tune.cu sweeps on B200 (profiles/r01_fft_tuning.md)
  switch (pl.n) {
    case 16: return launch_line<DmkFastPlan<4, 4, 1>, 4, 144, 4>(inverse, src, dst, pl, ms, md, n0, n_outer, st, ps);
    case 32: return launch_line<DmkFastPlan<8, 4, 1>, 4, 144, 4>(inverse, src, dst, pl, ms, md, n0, n_outer, st, ps);
    case 64: return launch_line<DmkFastPlan<8, 8, 1>, 2, 144, 4>(inverse, src, dst, pl, ms, md, n0, n_outer, st, ps);
    case 128: return launch_line<DmkFastPlan<16, 8, 1>, 2, 144, 3>(inverse, src, dst, pl, ms, md, n0, n_outer, st, ps);
    case 256: return launch_line<DmkFastPlan<16, 16, 1>, 1, 144, 3>(inverse, src, dst, pl, ms, md, n0, n_outer, st, ps);
    case 512: return launch_line<DmkFastPlan<8, 8, 8>, 1, 288, 2>(inverse, src, dst, pl, ms, md, n0, n_outer, st, ps);
    default: *done = false; return DMK_OK;
  }
}

template <class PL, int NP, int NT, int MINB = 1>
static int launch_xr2c(const double* in, double2* out, const DmkFftPlan& pl, int nrows, cudaStream_t st) {
  const size_t smem = (size_t)PL::N * NP * DMK_NC * sizeof(double2);
  const int npairs = (nrows + 1) / 2, grid = (npairs + NP - 1) / NP;
  int rc = set_smem(dmk_fast_xr2c_kernel<PL, NP, NT, MINB>, smem); if (rc) return rc;
  dmk_fast_xr2c_kernel<PL, NP, NT, MINB><<<grid, NT, smem, st>>>(in, out, pl.tw, nrows);
  DMK_CUDA_CHECK(cudaGetLastError());
  return DMK_OK;
}
template <class PL, int NP, int NT, int MINB = 1>
static int launch_xc2r(const double2* in, double* out, const DmkFftPlan& pl, int nrows, double scale, cudaStream_t st) {
  const size_t smem = (size_t)PL::N * NP * DMK_NC * sizeof(double2);
  const int npairs = (nrows + 1) / 2, grid = (npairs + NP - 1) / NP;
  int rc = set_smem(dmk_fast_xc2r_kernel<PL, NP, NT, MINB>, smem); if (rc) return rc;
  dmk_fast_xc2r_kernel<PL, NP, NT, MINB><<<grid, NT, smem, st>>>(in, out, pl.tw, nrows, scale);
  DMK_CUDA_CHECK(cudaGetLastError());
  return DMK_OK;
}

int dmk_fast_xpass_r2c(const double* in, double2* out, const DmkFftPlan& pl, int nrows, cudaStream_t st, bool* done) {
  *done = true;
  if (dmk_fast_disabled()) { *done = false; return DMK_OK; }
  switch (pl.n) {
    case 16: return launch_xr2c<DmkFastPlan<4, 4, 1>, 4, 144, 4>(in, out, pl, nrows, st);
    case 32: return launch_xr2c<DmkFastPlan<8, 4, 1>, 4, 144, 4>(in, out, pl, nrows, st);
    case 64: return launch_xr2c<DmkFastPlan<8, 8, 1>, 2, 144, 4>(in, out, pl, nrows, st);
    case 128: return launch_xr2c<DmkFastPlan<16, 8, 1>, 2, 144, 3>(in, out, pl, nrows, st);
    case 256: return launch_xr2c<DmkFastPlan<16, 16, 1>, 1, 144, 4>(in, out, pl, nrows, st);
    case 512: return launch_xr2c<DmkFastPlan<8, 8, 8>, 1, 288, 2>(in, out, pl, nrows, st);
    default: *done = false; return DMK_OK;
  }
}
int dmk_fast_xpass_c2r(const double2* in, double* out, const DmkFftPlan& pl, int nrows, double scale, cudaStream_t st, bool* done) {
  *done = true;
  if (dmk_fast_disabled()) { *done = false; return DMK_OK; }
  switch (pl.n) {
    case 16: return launch_xc2r<DmkFastPlan<4, 4, 1>, 4, 144, 4>(in, out, pl, nrows, scale, st);
    case 32: return launch_xc2r<DmkFastPlan<8, 4, 1>, 4, 144, 4>(in, out, pl, nrows, scale, st);
    case 64: return launch_xc2r<DmkFastPlan<8, 8, 1>, 2, 144, 4>(in, out, pl, nrows, scale, st);
    case 128: return launch_xc2r<DmkFastPlan<16, 8, 1>, 2, 144, 4>(in, out, pl, nrows, scale, st);
    case 256: {
      static const int v = [] { const char* e = getenv("DMK_XC2R_PLAN"); return e ? atoi(e) : 16; }();    // comparison runs (tools/gpu_call26.sh: 0.505 ms <8,8,4> x 6 CTAs, 0.44 x 5, 0.43 <16,16,1> x 4)
      if (v == 16) return launch_xc2r<DmkFastPlan<16, 16, 1>, 1, 144, 4>(in, out, pl, nrows, scale, st);
      if (v == 163) return launch_xc2r<DmkFastPlan<16, 16, 1>, 1, 144, 3>(in, out, pl, nrows, scale, st);
      if (v == 5) return launch_xc2r<DmkFastPlan<8, 8, 4>, 1, 144, 5>(in, out, pl, nrows, scale, st);
      return launch_xc2r<DmkFastPlan<8, 8, 4>, 1, 144, 6>(in, out, pl, nrows, scale, st);
    }
    case 512: return launch_xc2r<DmkFastPlan<8, 8, 8>, 1, 288, 2>(in, out, pl, nrows, scale, st);
    default: *done = false; return DMK_OK;
  }
}
