// Register-resident FFT stages for power-of-two line lengths (16 ... 512), the sizes of every benchmark grid.
//
// Replaces FFTW's many-dft r2c/c2r as called by the reference (/root/reference/src/grid/fft.cpp:48-58,
// fft.h:37-60), like fft.cuh, but organised around what the first profile showed (profiles/r01_summary.md: the
// shared-memory-only passes reached 25 % of HBM peak because each CTA loaded a tile, synchronised, ran three
// shared-memory passes and stored it, with nothing overlapping):
//
//  * a line of N = R1*R2(*R3) points is transformed in two (three) stages of radix 16/8/4; a thread holds one whole
//    butterfly (up to 16 complex values) in registers;
//  * the FIRST stage loads its operands straight from HBM (16 independent 16-byte loads in flight per thread) and the
//    LAST stage stores straight to HBM, so a tile makes one shared-memory round trip per intermediate stage instead of
//    one per radix pass plus two for staging;
//  * the digit reversal is folded into the address of the last stage's stores (a thread owns all R outputs of a
//    butterfly, so any output order is free): every pass reads and writes natural order, exactly once (144 B/point).
//
// Decimation in frequency, forward sign exp(-i...).  The unnormalised inverse of the plain passes is the same
// network with conjugated roots; the fused Gamma pass (gamma_fast.cu) instead runs the exact adjoint (decimation in
// time) from the digit-reversed order the forward network leaves in shared memory, as fft.cuh does.
#pragma once
#include "fft.cuh"
#include "fft_host.h"

// ---------------------------------------------------------------- radix-16 butterfly (4 x 4, natural order out)
template <bool INV> __device__ __forceinline__ double2 cmul_w(double2 a, double wr, double wi) {   // a * (wr + i wi), conj if INV
  return INV ? make_double2(a.x * wr + a.y * wi, a.y * wr - a.x * wi) : make_double2(a.x * wr - a.y * wi, a.y * wr + a.x * wi);
}
template <bool INV> __device__ __forceinline__ void bfly16(double2 (&v)[16]) {
  const double c = 0.92387953251128675613, s = 0.38268343236508977173, h = 0.70710678118654752440;
#pragma unroll
  for (int n2 = 0; n2 < 4; n2++) bfly4<INV>(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);      // v[4 k1 + n2] = y[n2][k1]
  // v[4 k1 + n2] *= w16^(n2 k1), w16 = exp(-i pi/8)
  v[5] = cmul_w<INV>(v[5], c, -s);                                   // w^1
  v[6] = cmul_w<INV>(v[6], h, -h);                                   // w^2
  v[7] = cmul_w<INV>(v[7], s, -c);                                   // w^3
  v[9] = cmul_w<INV>(v[9], h, -h);                                   // w^2
  v[10] = cmul_mi<INV>(v[10]);                                       // w^4 = -i
  v[11] = cmul_w<INV>(v[11], -h, -h);                                // w^6
  v[13] = cmul_w<INV>(v[13], s, -c);                                 // w^3
  v[14] = cmul_w<INV>(v[14], -h, -h);                                // w^6
  v[15] = cmul_w<INV>(v[15], -c, s);                                 // w^9
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++) bfly4<INV>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);   // v[4 k1 + k2] = X[k1 + 4 k2]
  double2 t;
#define DMK_SWAP16(a, b) t = v[a]; v[a] = v[b]; v[b] = t;
  DMK_SWAP16(1, 4) DMK_SWAP16(2, 8) DMK_SWAP16(3, 12) DMK_SWAP16(6, 9) DMK_SWAP16(7, 13) DMK_SWAP16(11, 14)
#undef DMK_SWAP16
}
template <int R, bool INV> __device__ __forceinline__ void bfly_p2(double2 (&v)[R]) {
  if constexpr (R == 2) bfly2<INV>(v[0], v[1]);
  else if constexpr (R == 4) bfly4<INV>(v[0], v[1], v[2], v[3]);
  else if constexpr (R == 8) bfly8<INV>(v);
  else bfly16<INV>(v);
}

// ---------------------------------------------------------------- the stage network of one line length
template <int R1_, int R2_, int R3_>
struct DmkFastPlan {
  static constexpr int R1 = R1_, R2 = R2_, R3 = R3_, N = R1_ * R2_ * R3_;
  // slot after the forward network -> frequency, and back
  __host__ __device__ static constexpr int freq(int slot) { return slot / (R2 * R3) + R1 * ((slot / R3) % R2) + R1 * R2 * (slot % R3); }
  __host__ __device__ static constexpr int slot(int k) { return (k % R1) * (R2 * R3) + ((k / R1) % R2) * R3 + k / (R1 * R2); }
};

// One stage over NSUB independent tiles of N positions x B columns: item = (sub, butterfly, column), one butterfly of
// radix R with operand distance ST per item.  B, R, ST are compile-time, so the index split costs shifts and one
// multiply-high.  `bact` <= B columns are live (ragged last tile).  ld(sub, pos, col) / st(sub, pos, col, value).
//   DIT = false: butterfly, then multiply output q by w_L^(q j)          (decimation in frequency, forward network)
//   DIT = true : multiply input q by conj-or-not w_L^(q j), then butterfly (its adjoint; used with INV = true)
// Twiddles w_L^(q j), q = 1..R-1, of one butterfly.  Only the powers of two (w^1, w^2, w^4, w^8) are loaded from the
// length-N table; the others are products of two of them (one rounding each, ~1e-16).  Loading all R-1 made the L1/LSU
// data pipe, which also carries every shared-memory access of these kernels, the limiter (ncu: 73-86 % of its peak
// with DRAM at 35-50 %, profiles/r01_summary.md).
template <int R> __device__ __forceinline__ void load_twiddles(const double2* __restrict__ tw, int base, double2 (&w)[R]) {
  w[1] = __ldg(&tw[base]);
  if constexpr (R > 2) { w[2] = __ldg(&tw[2 * base]); w[3] = cmul(w[1], w[2]); }
  if constexpr (R > 4) { w[4] = __ldg(&tw[4 * base]); w[5] = cmul(w[1], w[4]); w[6] = cmul(w[2], w[4]); w[7] = cmul(w[3], w[4]); }
  if constexpr (R > 8) {
    w[8] = __ldg(&tw[8 * base]);
#pragma unroll
    for (int q = 9; q < 16; q++) w[q] = cmul(w[q - 8], w[8]);
  }
}

template <int N, int R, int ST, int B, bool INV, bool DIT, class LD, class SR>
__device__ __forceinline__ void fstage(int nsub, int bact, const double2* __restrict__ tw, LD ld, SR sr) {
  constexpr int L = R * ST, TSTEP = N / L, PER = (N / R) * B;
  const int items = nsub * PER;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int sub = it / PER, r0 = it - sub * PER;
    const int bf = r0 / B, col = r0 - bf * B;
    if (col >= bact) continue;
    const int blk = bf / ST, j = bf - blk * ST;
    const int p0 = blk * L + j;
    double2 v[R];
#pragma unroll
    for (int m = 0; m < R; m++) v[m] = ld(sub, p0 + m * ST, col);
    if constexpr (DIT && ST > 1) {
      double2 w[R];
      load_twiddles<R>(tw, j * TSTEP, w);
#pragma unroll
      for (int q = 1; q < R; q++) v[q] = INV ? cmulc(v[q], w[q]) : cmul(v[q], w[q]);
    }
    bfly_p2<R, INV>(v);
    if constexpr (!DIT && ST > 1) {
      double2 w[R];
      load_twiddles<R>(tw, j * TSTEP, w);
#pragma unroll
      for (int q = 1; q < R; q++) v[q] = INV ? cmulc(v[q], w[q]) : cmul(v[q], w[q]);
    }
#pragma unroll
    for (int q = 0; q < R; q++) sr(sub, p0 + q * ST, col, v[q]);
  }
}

// record address of a strided line (DmkLineMap) without the division when the line never crosses a peer block
__device__ __forceinline__ long long dmk_line_rec_fast(const DmkLineMap& m, bool nosplit, int o, int line, int i0) {
  if (nosplit) return o * m.o_stride + line * m.e_stride + i0;
  return dmk_line_rec(m, o, line, i0);
}

// where the last stage of a strided pass stores line position `line`: own buffer, or a peer's (DmkPeerStore)
template <bool REMOTE>
__device__ __forceinline__ double2* dmk_line_store_ptr(const DmkLineMap& m, const DmkPeerStore& ps, double2* local, bool nosplit, int o,
                                                       int line, int i0) {
  if constexpr (REMOTE) {
    const int hi = line / ps.split, lo = line - hi * ps.split;
    return ps.peer[hi] + ((long long)ps.me * ps.hi_stride + o * ps.o_stride + lo * ps.e_stride + i0) * DMK_NC;
  }
  return local + dmk_line_rec_fast(m, nosplit, o, line, i0) * DMK_NC;
}

// launchers (fft_fast.cu, gamma_fast.cu): return DMK_OK and set *done = true if the length is covered
int dmk_fast_xpass_r2c(const double* in, double2* out, const DmkFftPlan& pl, int nrows, cudaStream_t st, bool* done);
int dmk_fast_xpass_c2r(const double2* in, double* out, const DmkFftPlan& pl, int nrows, double scale, cudaStream_t st, bool* done);
int dmk_fast_line_pass(bool inverse, const double2* src, double2* dst, const DmkFftPlan& pl, DmkLineMap ms, DmkLineMap md,
                       int n0, int n_outer, cudaStream_t st, bool* done, const DmkPeerStore* ps = nullptr);
int dmk_fast_gamma_zpass(double2* data, const DmkFftPlan& plz, DmkLineMap map, int n0, int n_outer, const DmkGammaParams& gp,
                         cudaStream_t st, bool imag_xi, bool* done, const DmkPeerStore* ps = nullptr);
bool dmk_fast_disabled();      // DMK_FFT=generic forces the shared-memory-only kernels of fft_kernels.cu (comparison runs)
