// In-shared-memory mixed-radix FFT building blocks (FP64, complex interleaved), used by every pass kernel.
//
// Replaces FFTW's many-dft r2c/c2r as called by the reference (/root/reference/src/grid/fft.cpp:48-58,
// fft.h:37-60): unnormalised forward (exp(-i...)) and unnormalised backward transforms, "howmany" = 9
// interleaved tensor components.
//
// Scheme: decimation-in-frequency, IN PLACE, no reordering pass.  After the forward passes slot s of a line holds
// frequency perm[s] (mixed-radix digit reversal); the inverse is the exact adjoint (decimation in time, conjugate
// twiddles applied before each butterfly), taking digit-reversed input back to natural order.  Kernels fold the
// permutation into the address of their global loads/stores, so it costs no traffic; the fused Gamma kernel
// (gamma_kernels.cu) never leaves the permuted order at all.
#pragma once
#include "dmk_common.cuh"

#define DMK_FFT_MAX_PASSES 8

struct DmkFftPlan {          // one axis length
  int n;
  int npass;
  int radix[DMK_FFT_MAX_PASSES];
  const int* perm;           // device: perm[slot] = frequency index after forward passes
  const int* iperm;          // device: iperm[frequency] = slot
  const double2* tw;         // device: tw[k] = exp(-2 pi i k / n), k in [0,n)
};

// ---------------------------------------------------------------- radix butterflies (in registers)
template <bool INV> __device__ __forceinline__ void bfly2(double2& a, double2& b) {
  double2 t = csub(a, b); a = cadd(a, b); b = t;
}
template <bool INV> __device__ __forceinline__ void bfly4(double2& v0, double2& v1, double2& v2, double2& v3) {
  double2 p0 = cadd(v0, v2), p2 = csub(v0, v2), p1 = cadd(v1, v3), p3 = cmul_mi<INV>(csub(v1, v3));
  v0 = cadd(p0, p1); v2 = csub(p0, p1); v1 = cadd(p2, p3); v3 = csub(p2, p3);
}
template <bool INV> __device__ __forceinline__ void bfly8(double2 (&v)[8]) {
  const double h = 0.70710678118654752440;
  double2 u0 = cadd(v[0], v[4]), u1 = cadd(v[1], v[5]), u2 = cadd(v[2], v[6]), u3 = cadd(v[3], v[7]);
  double2 d0 = csub(v[0], v[4]), d1 = csub(v[1], v[5]), d2 = csub(v[2], v[6]), d3 = csub(v[3], v[7]);
  // d_k *= w8^k  (forward: w8 = exp(-i pi/4))
  if (!INV) {
    d1 = make_double2(h * (d1.x + d1.y), h * (d1.y - d1.x));
    d3 = make_double2(h * (d3.y - d3.x), -h * (d3.x + d3.y));
  } else {
    d1 = make_double2(h * (d1.x - d1.y), h * (d1.x + d1.y));
    d3 = make_double2(-h * (d3.x + d3.y), h * (d3.x - d3.y));
  }
  d2 = cmul_mi<INV>(d2);
  bfly4<INV>(u0, u1, u2, u3);      // -> Y0, Y2, Y4, Y6
  bfly4<INV>(d0, d1, d2, d3);      // -> Y1, Y3, Y5, Y7
  v[0] = u0; v[2] = u1; v[4] = u2; v[6] = u3;
  v[1] = d0; v[3] = d1; v[5] = d2; v[7] = d3;
}
// generic small prime radix by direct summation; roots taken from the length-n table: w_R^m = tw[m * n / R]
template <int R, bool INV> __device__ __forceinline__ void bfly_generic(double2 (&v)[R], const double2* __restrict__ tw, int n) {
  double2 o[R];
  const int st = n / R;
#pragma unroll
  for (int q = 0; q < R; q++) {
    double2 acc = v[0];
#pragma unroll
    for (int k = 1; k < R; k++) {
      double2 w = __ldg(&tw[((q * k) % R) * st]);
      acc = cadd(acc, INV ? cmulc(v[k], w) : cmul(v[k], w));
    }
    o[q] = acc;
  }
#pragma unroll
  for (int q = 0; q < R; q++) v[q] = o[q];
}
template <int R, bool INV> __device__ __forceinline__ void bfly(double2 (&v)[R], const double2* __restrict__ tw, int n) {
  if constexpr (R == 2) bfly2<INV>(v[0], v[1]);
  else if constexpr (R == 4) bfly4<INV>(v[0], v[1], v[2], v[3]);
  else if constexpr (R == 8) bfly8<INV>(v);
  else bfly_generic<R, INV>(v, tw, n);
}

// ---------------------------------------------------------------- one pass over a tile held in shared memory
// s[(line position) * B + b]: B independent transforms ("batch", contiguous), line length n.
// L = current sub-transform length of this pass (n for the first forward pass).
template <int R, bool INV>
__device__ __forceinline__ void fft_pass(double2* __restrict__ s, int n, int L, int B, const double2* __restrict__ tw) {
  const int Ls = L / R;
  const int tstep = n / L;
  const int total = (n / R) * B;
  for (int t = threadIdx.x; t < total; t += blockDim.x) {
    int bf = t / B, c = t - bf * B;
    int blk = bf / Ls, j = bf - blk * Ls;
    double2* p = s + (size_t)(blk * L + j) * B + c;
    const int es = Ls * B;
    double2 v[R];
#pragma unroll
    for (int k = 0; k < R; k++) v[k] = p[k * es];
    if (INV) {
      if (j != 0) {
#pragma unroll
        for (int q = 1; q < R; q++) v[q] = cmulc(v[q], __ldg(&tw[q * j * tstep]));
      }
      bfly<R, true>(v, tw, n);
    } else {
      bfly<R, false>(v, tw, n);
      if (j != 0) {
#pragma unroll
        for (int q = 1; q < R; q++) v[q] = cmul(v[q], __ldg(&tw[q * j * tstep]));
      }
    }
#pragma unroll
    for (int k = 0; k < R; k++) p[k * es] = v[k];
  }
}

// odd radices (parity-test grids only: 3, 6, 7, ...) are kept out of line so that they do not inflate the register
// allocation of the power-of-two path the benchmark sizes use
template <bool INV>
__device__ __noinline__ void fft_pass_odd(int radix, double2* s, int n, int L, int B, const double2* tw) {
  switch (radix) {
    case 3: fft_pass<3, INV>(s, n, L, B, tw); break;
    case 5: fft_pass<5, INV>(s, n, L, B, tw); break;
    case 7: fft_pass<7, INV>(s, n, L, B, tw); break;
    case 11: fft_pass<11, INV>(s, n, L, B, tw); break;
    case 13: fft_pass<13, INV>(s, n, L, B, tw); break;
    default: break;
  }
}
template <bool INV>
__device__ __forceinline__ void fft_pass_dispatch(int radix, double2* s, int n, int L, int B, const double2* tw) {
  switch (radix) {
    case 8: fft_pass<8, INV>(s, n, L, B, tw); break;
    case 4: fft_pass<4, INV>(s, n, L, B, tw); break;
    case 2: fft_pass<2, INV>(s, n, L, B, tw); break;
    default: fft_pass_odd<INV>(radix, s, n, L, B, tw); break;
  }
}

// forward: natural order in, digit-reversed (perm) order out.  Block-wide; ends with a barrier.
__device__ __forceinline__ void fft_forward_smem(const DmkFftPlan& pl, double2* s, int B) {
  int L = pl.n;
  for (int p = 0; p < pl.npass; p++) {
    fft_pass_dispatch<false>(pl.radix[p], s, pl.n, L, B, pl.tw);
    L /= pl.radix[p];
    __syncthreads();
  }
}
// inverse: digit-reversed order in, natural order out (unnormalised).  Block-wide; ends with a barrier.
__device__ __forceinline__ void fft_inverse_smem(const DmkFftPlan& pl, double2* s, int B) {
  int L = 1;
  for (int p = pl.npass - 1; p >= 0; p--) {
    L *= pl.radix[p];
    fft_pass_dispatch<true>(pl.radix[p], s, pl.n, L, B, pl.tw);
    __syncthreads();
  }
}
