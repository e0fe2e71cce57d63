// Host-visible declarations of the FFT pass launchers (fft_kernels.cu) and the fused Gamma pass (gamma_kernels.cu).
#pragma once
#include "fft.cuh"

// Address map of one strided pass over a [..][..][n0] record array:
//   record(o, line, i0) = (line / split) * hi_stride + o * o_stride + (line % split) * e_stride + i0
// single GPU  y pass on [z][ky][kx]             : split = ny,     o = z,  o_stride = ny*nxr,     e_stride = nxr
// multi  GPU  y pass on [peer][z_loc][ky_loc][kx]: split = ky_loc, hi_stride = nz_loc*ky_loc*nxr, o_stride = ky_loc*nxr
//             z pass on [z][ky_loc][kx]         : split = nz,     o = ky, o_stride = nxr,        e_stride = ky_loc*nxr
struct DmkLineMap {
  long long hi_stride, o_stride, e_stride;
  int split;
};
__host__ __device__ __forceinline__ long long dmk_line_rec(const DmkLineMap& m, int o, int line, int i0) {
  int hi = line / m.split, lo = line - hi * m.split;
  return hi * m.hi_stride + o * m.o_stride + lo * m.e_stride + i0;
}

// Fused exchange over peer memory (NVLink): when `remote` is set, the LAST stage of a strided pass stores line position
// `line` not into its own buffer but into block `me` of the buffer of peer line / split -- the transpose of the
// distributed FFT happens inside the FFT pass, tile by tile, instead of in a separate all-to-all.
#define DMK_MAX_PEERS 8
struct DmkPeerStore {
  int remote, me, split;
  long long hi_stride, o_stride, e_stride;      // record = me * hi_stride + o * o_stride + (line % split) * e_stride + i0
  double2* peer[DMK_MAX_PEERS];
};

int dmk_fft_plan_create(DmkFftPlan* pl, int n);
void dmk_fft_plan_destroy(DmkFftPlan* pl);
int dmk_launch_xpass_r2c(const double* in, double2* out, const DmkFftPlan& pl, int nrows, cudaStream_t st);
int dmk_launch_xpass_c2r(const double2* in, double* out, const DmkFftPlan& pl, int nrows, double scale, cudaStream_t st);
int dmk_launch_line_pass(bool inverse, const double2* src, double2* dst, const DmkFftPlan& pl, DmkLineMap ms,
                         DmkLineMap md, int n0, int n_outer, int KT, cudaStream_t st, const DmkPeerStore* ps = nullptr);

// ---- fused z pass: forward FFT along z, divergence-RMS accumulation, Gamma-operator contraction, inverse FFT
struct DmkGammaParams {
  const double* C_ref;         // device, 81 doubles, index ((l*3+n)*3+m)*3+o  (== C_ref(l,n,m,o), row-major 3333)
  const double2* xi_x;         // device per-axis derivative factor tables (see context.cu: build_xi_tables)
  const double2* xi_y;
  const double2* xi_z;
  const double2* xo_x;         // "other axis" factors (FWBW stencil), 1 otherwise
  const double2* xo_y;
  const double2* xo_z;
  const unsigned char* nyq_x;  // per-axis flag: xi1st component zeroed at this index (even-grid Nyquist, continuous)
  const unsigned char* nyq_y;
  const unsigned char* nyq_z;
  double rescale[3];           // geom_size / scaled_geom_size  (divergence correction)
  int nx;                      // real-space cells along x (for the kx weighting)
  int ky_offset;               // global ky index of local ky 0 (slab decomposition)
  const double* dc_value;      // device, 9 doubles in record order c = i + 3 j: value written to the k = 0 coefficient
  double* div_accum;           // device, 1 double: += sum_k w_k |P_hat(k) . conj(-xi1st) rescale|^2
  double* div_partial;         // device, one double per CTA of the pass: the CTA's share of that sum (see dmk_sum_div_partials)
};
// Adds the per-CTA divergence sums to *div_accum in a FIXED order (strided per-thread sums, then a fixed tree), so the sum --
// and with it the convergence test err_div / divTol < 1 and the iteration count -- does not depend on the order in which the
// CTAs finish (a floating-point atomicAdd per CTA did).
int dmk_sum_div_partials(const double* partial, int n, double* div_accum, cudaStream_t st);
int dmk_launch_gamma_zpass(double2* data, const DmkFftPlan& plz, DmkLineMap map, int n0, int n_outer, int KT,
                           const DmkGammaParams& gp, cudaStream_t st, bool imag_xi, const DmkPeerStore* ps = nullptr);
