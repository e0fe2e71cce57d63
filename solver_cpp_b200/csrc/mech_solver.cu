// mechanical_spectral_basic on the device: residual evaluation, convergence test, increment forwarding and the
// NGMRES-driven solve with the solution vector, its history and all work fields resident in HBM.
//
// Mirrors (does not copy) MechSolverBasic of the reference with upstream semantics (D3, D9 applied):
//   init / calculate_solution / forward / converged / form_residual
//       /root/reference/src/grid/mech_solver_basic.cpp:19-108, 110-161, 163-220, 242-284, 286-339
//   Spectral::constitutive_response reductions   /root/reference/src/grid/spectral.cpp:132-219
//   MechBase::calculate_masked_compliance        /root/reference/src/grid/mech_base.cpp:249-300
// PETSc's SNESNGMRES (third party, 3.18.6; not in /root/reference) is restated from its published algorithm with
// PETSc's defaults (m = 30, restart/select type "difference", gammaA = gammaC = 2, epsilonB = 0.1, deltaB = 0.9,
// restart_it = 2, no preconditioner so X^M = X - F(X)); the 108 logged iterates of the reference's Fortran run pin it.
//
// One residual evaluation is fully stream-ordered: constitutive kernel -> record sum -> (allreduce) -> one-thread
// "aim" kernel (P_av, Delta F_aim = S:(P_av - P_aim), err_BC, DC term) -> x r2c -> y -> fused z/Gamma/div -> y -> x c2r,
// with a single host synchronisation at the end to fetch 40 doubles.
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <algorithm>
#include <chrono>
#include "ctx.h"
#include "reduce.cuh"
#include "constitutive.cuh"
#include "polarisation.cuh"

#define NGMRES_M 30

struct DmkMech {
  long long n9;                                      // doubles in one tensor field of the slab
  long long nvec;                                    // doubles in the solution vector: n9 (basic), 2 n9 = (F, F_tau) (polarisation)
  int scheme;                                        // 0 spectral_basic, 1 spectral_polarisation
  // polarisation only: work field, materialised tangent, numerics, scaling tensors (row-major)
  double *W, *dPdF; int* d_singular;
  double alpha, beta, eps_abs_curl, eps_rel_curl, err_curl;
  double C_scale[81], S_scale[81], F_av[9];
  double *F, *F_lastInc, *F_dot, *R_field;          // device fields (record layout)
  double *XM, *FM, *XA, *FA, *Xcur, *Fcur;           // NGMRES work vectors
  std::vector<double*> Xdot, Fdot;                   // NGMRES history, allocated on demand
  double* d_blk;                                     // device scalar block (c->d_scratch + DMK_SCRATCH_MECH)
  // host state (row-major 3x3 / 3333)
  double F_aim[9], F_aim_lastInc[9], F_aim_dot[9], P_av[9], P_aim[9], S[81];
  double C_volAvg[81], C_minMaxAvg[81], C_volAvgLastInc[81], C_minMaxAvgLastInc[81];
  int stress_mask[9];                                // 1 = 'x' (component not stress controlled)
  double R[9];
  double dt;
  double err_div, err_BC, div_raw;
  int total_iter;
  bool tangent_stale;
  double tangent_dt;
  long long minmax_cells[2];
  std::vector<double> log;                           // 12 doubles per evaluation
  std::vector<double> reports;                       // 6 doubles per convergence check
  std::vector<double> stats;                         // 2 doubles per evaluation: cells of this slab whose first sub-step failed, 0
  int snes_its; int snes_reason;
  double t_solver[12];                               // wall seconds inside ngmres_solve: evaluations, dot products, least squares, combination, norms, copies, total, iterations
  bool capture;                                      // test aid: keep host copies of every evaluation's F and r
  std::vector<double> capF, capR;
};

// device scalar block layout (doubles)
enum { B_SUM = 0, B_PAV = 16, B_FAIM = 32, B_PAIM = 48, B_MASK = 64, B_R = 80, B_S = 96, B_ERRBC = 192, B_DIV = 193, B_OUT = 200 };

// ------------------------------------------------------------------------------------------------ small host tensor algebra
static void mm33(const double* A, const double* B, double* C) {
  double r[9];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += A[3 * i + k] * B[3 * k + j]; r[3 * i + j] = s; }
  memcpy(C, r, sizeof r);
}
static void tr33(const double* A, double* T) { double r[9]; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r[3 * i + j] = A[3 * j + i]; memcpy(T, r, sizeof r); }
// rotations.f90:288-300 with R given as a matrix: passive R T R^T, active R^T T R
static void rotate2(const double* R, const double* T, bool active, double* out) {
  double Rt[9], t[9]; tr33(R, Rt);
  if (active) { mm33(Rt, T, t); mm33(t, R, out); } else { mm33(R, T, t); mm33(t, Rt, out); }
}
static void rotate4(const double* R, const double* T, double* out) {      // passive
  double r[81];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) {
    double s = 0;
    for (int m = 0; m < 3; m++) for (int n = 0; n < 3; n++) for (int o = 0; o < 3; o++) for (int p = 0; p < 3; p++)
      s += R[3 * i + m] * R[3 * j + n] * R[3 * k + o] * R[3 * l + p] * T[((m * 3 + n) * 3 + o) * 3 + p];
    r[((i * 3 + j) * 3 + k) * 3 + l] = s;
  }
  memcpy(out, r, sizeof r);
}
static bool invert_n(std::vector<double>& A, int n) {       // Gauss-Jordan with partial pivoting, in place
  std::vector<double> I((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) I[(size_t)i * n + i] = 1.0;
  for (int k = 0; k < n; k++) {
    int p = k; double mx = fabs(A[(size_t)k * n + k]);
    for (int i = k + 1; i < n; i++) if (fabs(A[(size_t)i * n + k]) > mx) { mx = fabs(A[(size_t)i * n + k]); p = i; }
    if (mx == 0.0) return false;
    if (p != k) for (int j = 0; j < n; j++) { std::swap(A[(size_t)k * n + j], A[(size_t)p * n + j]); std::swap(I[(size_t)k * n + j], I[(size_t)p * n + j]); }
    const double d = A[(size_t)k * n + k];
    for (int j = 0; j < n; j++) { A[(size_t)k * n + j] /= d; I[(size_t)k * n + j] /= d; }
    for (int i = 0; i < n; i++) if (i != k) {
      const double f = A[(size_t)i * n + k];
      if (f != 0.0) for (int j = 0; j < n; j++) { A[(size_t)i * n + j] -= f * A[(size_t)k * n + j]; I[(size_t)i * n + j] -= f * I[(size_t)k * n + j]; }
    }
  }
  A = I;
  return true;
}

// MechBase::calculate_masked_compliance; all arguments row-major; mask 1 = 'x'
static int masked_compliance_rm(const double* C, const double* R, const int* mask, double* S) {
  double Cr[81];
  rotate4(R, C, Cr);
  std::vector<int> idx;
  for (int k = 0; k < 9; k++) if (!mask[k]) idx.push_back(k);
  for (int k = 0; k < 81; k++) S[k] = 0.0;
  const int n = (int)idx.size();
  if (n == 0) return DMK_OK;
  std::vector<double> red((size_t)n * n);
  for (int a = 0; a < n; a++) for (int b = 0; b < n; b++) red[(size_t)a * n + b] = Cr[9 * idx[a] + idx[b]];
  if (!invert_n(red, n)) { dmk_set_error("masked compliance: reduced stiffness is singular"); return DMK_ERR_STATE; }
  for (int a = 0; a < n; a++) for (int b = 0; b < n; b++) S[9 * idx[a] + idx[b]] = red[(size_t)a * n + b];
  return DMK_OK;
}

extern "C" int dmk_masked_compliance(const double* C, const double* R, const int* stress_mask, double* S) {
  double Cr[81], Sr[81];
  f2c3333(C, Cr);
  int rc = masked_compliance_rm(Cr, R, stress_mask, Sr);
  if (rc) return rc;
  c2f3333(Sr, S);
  return DMK_OK;
}

// ------------------------------------------------------------------------------------------------ device helpers
__global__ void dmk_fill_identity_kernel(double* f, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double* r = f + i * 9;
    r[0] = 1; r[1] = 0; r[2] = 0; r[3] = 0; r[4] = 1; r[5] = 0; r[6] = 0; r[7] = 0; r[8] = 1;
  }
}

// one thread: homogenised stress, aim update, boundary-condition error, DC term of the Gamma convolution
__global__ void dmk_mech_aim_kernel(double* __restrict__ blk, double wgt, double* __restrict__ dc_out, double* __restrict__ div_in) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double* R = blk + B_R; const double* S = blk + B_S; const double* mask = blk + B_MASK;
  double Pm[9], t[9], Pav[9];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Pm[3 * i + j] = blk[B_SUM + i + 3 * j] * wgt;     // record order -> row-major
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += R[3 * i + k] * Pm[3 * k + j]; t[3 * i + j] = s; }
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += t[3 * i + k] * R[3 * j + k]; Pav[3 * i + j] = s; }
  double dF[9], err = 0.0;
  for (int a = 0; a < 9; a++) {
    double s = 0;
    for (int b = 0; b < 9; b++) s += S[9 * a + b] * (Pav[b] - blk[B_PAIM + b]);
    dF[a] = s;
  }
  for (int a = 0; a < 9; a++) {
    blk[B_PAV + a] = Pav[a];
    blk[B_FAIM + a] -= dF[a];
    if (mask[a] == 0.0) err = fmax(err, fabs(Pav[a] - blk[B_PAIM + a]));
  }
  blk[B_ERRBC] = err;
  // active rotation R^T dF R, written in record order
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += R[3 * k + i] * dF[3 * k + j]; t[3 * i + j] = s; }
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += t[3 * i + k] * R[3 * k + j]; dc_out[i + 3 * j] = s; }
  *div_in = 0.0;
}
__global__ void dmk_mech_collect_kernel(double* __restrict__ blk, const double* __restrict__ div, const unsigned* __restrict__ given_up) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    blk[B_OUT + 20] = (double)*given_up;
    for (int a = 0; a < 9; a++) { blk[B_OUT + a] = blk[B_PAV + a]; blk[B_OUT + 9 + a] = blk[B_FAIM + a]; }
    blk[B_OUT + 18] = blk[B_ERRBC]; blk[B_OUT + 19] = *div;
  }
}

struct DmkVecs { const double* p[32]; double a[32]; };
// out = c0 * x + sum_k a_k v_k
__global__ void __launch_bounds__(256) dmk_lincomb_kernel(double* __restrict__ out, const double* __restrict__ x, double c0, DmkVecs v, int nv, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double s = c0 * x[i];
    for (int k = 0; k < nv; k++) s += v.a[k] * v.p[k][i];
    out[i] = s;
  }
}
// partial[b][k] = sum_i (x_i - v_k,i)^2
__global__ void __launch_bounds__(256) dmk_diffnorm_partial_kernel(const double* __restrict__ x, DmkVecs v, int nv, long long n, double* __restrict__ partial) {
  __shared__ double sh[8];
  for (int k = 0; k < nv; k++) {
    const double* __restrict__ b = v.p[k];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) { const double d = x[i] - b[i]; acc += d * d; }
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) { double s = 0.0; for (int w = 0; w < 8; w++) s += sh[w]; partial[(size_t)blockIdx.x * nv + k] = s; }
    __syncthreads();
  }
}

static int reduce_to_host(dmk_ctx* c, double* d_vals, int n, double* h_out) {
  if (c->P > 1) { if (c->comm.allreduce_sum(c->comm.user, d_vals, n)) { dmk_set_error("allreduce failed"); return DMK_ERR_STATE; } }
  DMK_CUDA_CHECK(cudaMemcpyAsync(c->h_pinned + 1024, d_vals, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  memcpy(h_out, c->h_pinned + 1024, n * sizeof(double));
  return DMK_OK;
}
// dots of a with up to 32 vectors -> host (global over ranks)
static int dots(dmk_ctx* c, const double* a, const std::vector<const double*>& bs, double* h_out) {
  if (bs.empty()) return DMK_OK;
  double* d_out = c->d_scratch + DMK_SCRATCH_TMP + 300;
  int rc = dmk_multi_dot(c, a, bs.data(), (int)bs.size(), c->mech->nvec, d_out);
  if (rc) return rc;
  return reduce_to_host(c, d_out, (int)bs.size(), h_out);
}
static int diffnorms2(dmk_ctx* c, const double* x, const std::vector<const double*>& vs, double* h_out) {
  if (vs.empty()) return DMK_OK;
  DmkVecs v; for (int k = 0; k < 32; k++) { v.p[k] = k < (int)vs.size() ? vs[k] : nullptr; v.a[k] = 0; }
  const long long n = c->mech->nvec;
  int blocks = (int)std::max(1LL, std::min((n + 255) / 256, (long long)DMK_REDUCE_BLOCKS));
  double* partial = c->d_scratch + DMK_SCRATCH_PARTIAL;
  double* d_out = c->d_scratch + DMK_SCRATCH_TMP + 300;
  dmk_diffnorm_partial_kernel<<<blocks, 256, 0, c->stream>>>(x, v, (int)vs.size(), n, partial);
  dmk_sum_partials_kernel<<<1, 32, 0, c->stream>>>(partial, blocks, (int)vs.size(), d_out);
  c->launches += 2;
  DMK_CUDA_CHECK(cudaGetLastError());
  return reduce_to_host(c, d_out, (int)vs.size(), h_out);
}
static int lincomb(dmk_ctx* c, double* out, const double* x, double c0, const std::vector<const double*>& vs, const double* a) {
  DmkVecs v; for (int k = 0; k < 32; k++) { v.p[k] = k < (int)vs.size() ? vs[k] : nullptr; v.a[k] = k < (int)vs.size() ? a[k] : 0.0; }
  dmk_lincomb_kernel<<<148 * 8, 256, 0, c->stream>>>(out, x, c0, v, (int)vs.size(), c->mech->nvec);
  c->launches++;
  DMK_CUDA_CHECK(cudaGetLastError());
  return DMK_OK;
}

// ---- fused NGMRES vector operations: ONE pass over the history per operation instead of one pass per history vector (the
// kernels above read `a` / `x` again for every v_k: 2 nv fields of traffic where nv + 1 suffice; at 256^3 a field is 1.2 GB and
// the history holds up to 30).  Launch geometry, per-thread summation order and the block / grid reduction trees are those of
// dmk_multidot_partial_kernel, dmk_lincomb_kernel and dmk_diffnorm_partial_kernel, so every result is bit-identical to theirs.
#define DMK_VECOP_BLOCKS 1184
// instantiated widths: the smallest one >= nv runs (at most 3 padded slots, which re-read v_0 from L1)
#define DMK_VECOP_DISPATCH(L) \
  if (nv <= 1) L(1) else if (nv <= 2) L(2) else if (nv <= 4) L(4) else if (nv <= 6) L(6) else if (nv <= 8) L(8) \
  else if (nv <= 12) L(12) else if (nv <= 16) L(16) else if (nv <= 20) L(20) else if (nv <= 24) L(24) \
  else if (nv <= 28) L(28) else L(32)
// Both kernels stage their operands through shared memory with cp.async: thread t copies element i of every vector into its own
// column of a [vector][256] tile (one 8-byte cp.async each, no register held per load), two tiles deep, and reads back only what it
// copied itself -- so no barrier is needed, and a whole tile (up to 67 kB per SM) is always in flight whatever the register budget;
// with plain loads ptxas interleaved each load with the FMA that consumes it (two loads in flight per thread).
__device__ __forceinline__ void dmk_cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void dmk_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void dmk_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// row b of partial: [k] = f . v_k, [nv + k] = v_piv . v_k (k < nv), [2 nv] = f . f
template <int NV>
__global__ void __launch_bounds__(256) dmk_ngmres_dots_kernel(const double* __restrict__ f, DmkVecs v, int nv, int piv, long long n,
                                                              double* __restrict__ partial) {
  extern __shared__ double dmk_vec_tile[];                      // [2][NV + 1][256]
  double acc[2 * NV + 1];
#pragma unroll
  for (int k = 0; k < 2 * NV + 1; k++) acc[k] = 0.0;
  const long long T = (long long)gridDim.x * 256;
  long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  auto issue = [&](int st, long long ii) {
    if (ii < n) {
      double* d = dmk_vec_tile + (size_t)st * (NV + 1) * 256 + threadIdx.x;
      dmk_cp_async8(d, f + ii);
#pragma unroll
      for (int k = 0; k < NV; k++) dmk_cp_async8(d + (k + 1) * 256, v.p[k] + ii);     // slots >= nv alias v_0 (the host pads the table)
    }
    dmk_cp_async_commit();
  };
  issue(0, i);
  int st = 0;
  for (; i < n; i += T) {
    issue(st ^ 1, i + T);
    dmk_cp_async_wait<1>();
    const double* d = dmk_vec_tile + (size_t)st * (NV + 1) * 256 + threadIdx.x;
    const double fi = d[0], pi = d[(piv + 1) * 256];
#pragma unroll
    for (int k = 0; k < NV; k++) { const double b = d[(k + 1) * 256]; acc[k] += fi * b; acc[NV + k] += pi * b; }
    acc[2 * NV] += fi * fi;
    st ^= 1;
  }
  dmk_cp_async_wait<0>();
  __syncthreads();
  double (*sh)[8] = reinterpret_cast<double (*)[8]>(dmk_vec_tile);                      // [2 NV + 1][8] warp partials (the tiles are dead)
#pragma unroll
  for (int j = 0; j < 2 * NV + 1; j++) {
    double a = acc[j];
    for (int off = 16; off > 0; off >>= 1) a += __shfl_down_sync(0xffffffffu, a, off);
    if ((threadIdx.x & 31) == 0) sh[j][threadIdx.x >> 5] = a;
  }
  __syncthreads();
  double* __restrict__ row = partial + (size_t)blockIdx.x * (2 * nv + 1);
  for (int j = threadIdx.x; j < 2 * nv + 1; j += blockDim.x) {
    const int src = j < nv ? j : (j < 2 * nv ? NV + (j - nv) : 2 * NV);
    double t = 0.0;
    for (int w = 0; w < 8; w++) t += sh[src][w];
    row[j] = t;
  }
}
// out = c0 x + sum_k a_k v_k; row b of partial: [k] = |out - v_k|^2 (k < nv), [nv] = |out - x|^2
template <int NV>
__global__ void __launch_bounds__(256) dmk_ngmres_combine_kernel(double* __restrict__ out, const double* __restrict__ x, double c0, DmkVecs v,
                                                                 int nv, long long n, double* __restrict__ partial) {
  extern __shared__ double dmk_vec_tile[];                      // [2][NV + 1][256]
  double acc[NV + 1];
#pragma unroll
  for (int k = 0; k < NV + 1; k++) acc[k] = 0.0;
  const long long T = (long long)gridDim.x * 256;
  long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  auto issue = [&](int st, long long ii) {
    if (ii < n) {
      double* d = dmk_vec_tile + (size_t)st * (NV + 1) * 256 + threadIdx.x;
      dmk_cp_async8(d, x + ii);
#pragma unroll
      for (int k = 0; k < NV; k++) dmk_cp_async8(d + (k + 1) * 256, v.p[k] + ii);
    }
    dmk_cp_async_commit();
  };
  issue(0, i);
  int st = 0;
  for (; i < n; i += T) {
    issue(st ^ 1, i + T);
    dmk_cp_async_wait<1>();
    const double* d = dmk_vec_tile + (size_t)st * (NV + 1) * 256 + threadIdx.x;
    const double xi = d[0];
    double s = c0 * xi;
#pragma unroll
    for (int k = 0; k < NV; k++) if (k < nv) s += v.a[k] * d[(k + 1) * 256];
    out[i] = s;
#pragma unroll
    for (int k = 0; k < NV; k++) { const double e = s - d[(k + 1) * 256]; acc[k] += e * e; }
    { const double e = s - xi; acc[NV] += e * e; }
    st ^= 1;
  }
  dmk_cp_async_wait<0>();
  __syncthreads();
  double (*sh)[8] = reinterpret_cast<double (*)[8]>(dmk_vec_tile);
#pragma unroll
  for (int j = 0; j < NV + 1; j++) {
    double a = acc[j];
    for (int off = 16; off > 0; off >>= 1) a += __shfl_down_sync(0xffffffffu, a, off);
    if ((threadIdx.x & 31) == 0) sh[j][threadIdx.x >> 5] = a;
  }
  __syncthreads();
  double* __restrict__ row = partial + (size_t)blockIdx.x * (nv + 1);
  for (int j = threadIdx.x; j < nv + 1; j += blockDim.x) {
    const int src = j < nv ? j : NV;
    double t = 0.0;
    for (int w = 0; w < 8; w++) t += sh[src][w];
    row[j] = t;
  }
}
// out = x - f (the un-preconditioned NGMRES candidate X^M); dmk_lincomb_kernel's rounding for c0 = 1, a_0 = -1 (exact either way)
__global__ void __launch_bounds__(256) dmk_ngmres_xm_kernel(double* __restrict__ out, const double* __restrict__ x, const double* __restrict__ f, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = x[i] - f[i];
}
template <class K> static int vecop_smem(K kernel, int nvt) {
  const int bytes = 2 * (nvt + 1) * 256 * (int)sizeof(double);
  if (bytes > 48 * 1024) DMK_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return DMK_OK;
}
static bool vecops_split() { static const int v = [] { const char* e = getenv("DMK_NGMRES_VECOPS"); return (e && !strcmp(e, "split")) ? 1 : 0; }(); return v != 0; }
static int vecop_blocks(long long n) { return (int)std::max(1LL, std::min((n + 255) / 256, (long long)DMK_VECOP_BLOCKS)); }
// xi_k = f . v_k, beta_k = v_piv . v_k (k < nv), nu = f . f -> host (global over ranks), one pass over f and the v_k
static int ngmres_dots(dmk_ctx* c, const double* f, const std::vector<const double*>& vs, int piv, double* xi, double* beta, double* nu) {
  const int nv = (int)vs.size();
  int rc;
  if (nv < 1 || nv > 32) { dmk_set_error("ngmres_dots: 1..32 vectors"); return DMK_ERR_ARG; }
  DmkVecs v; for (int k = 0; k < 32; k++) { v.p[k] = vs[k < nv ? k : 0]; v.a[k] = 0; }
  const long long n = c->mech->nvec;
  const int blocks = vecop_blocks(n);
  double* partial = c->d_scratch + DMK_SCRATCH_PARTIAL;
  double* d_out = c->d_scratch + DMK_SCRATCH_TMP + 300;
#define DMK_DOTS(NV) { if ((rc = vecop_smem(dmk_ngmres_dots_kernel<NV>, NV))) return rc; \
    dmk_ngmres_dots_kernel<NV><<<blocks, 256, 2 * (NV + 1) * 256 * sizeof(double), c->stream>>>(f, v, nv, piv, n, partial); }
  DMK_VECOP_DISPATCH(DMK_DOTS)
#undef DMK_DOTS
  dmk_sum_partials_kernel<<<1, 96, 0, c->stream>>>(partial, blocks, 2 * nv + 1, d_out);
  c->launches += 2;
  DMK_CUDA_CHECK(cudaGetLastError());
  double h[65];
  if ((rc = reduce_to_host(c, d_out, 2 * nv + 1, h))) return rc;
  for (int k = 0; k < nv; k++) { xi[k] = h[k]; beta[k] = h[nv + k]; }
  *nu = h[2 * nv];
  return DMK_OK;
}
// out = c0 x + sum a_k v_k and dn_k = |out - v_k|^2 (k < nv), dn_nv = |out - x|^2 -> host, one pass
static int ngmres_combine(dmk_ctx* c, double* out, const double* x, double c0, const std::vector<const double*>& vs, const double* a, double* dn) {
  const int nv = (int)vs.size();
  int rc;
  if (nv < 1 || nv > 32) { dmk_set_error("ngmres_combine: 1..32 vectors"); return DMK_ERR_ARG; }
  DmkVecs v; for (int k = 0; k < 32; k++) { v.p[k] = vs[k < nv ? k : 0]; v.a[k] = k < nv ? a[k] : 0.0; }
  const long long n = c->mech->nvec;
  const int blocks = vecop_blocks(n);
  double* partial = c->d_scratch + DMK_SCRATCH_PARTIAL;
  double* d_out = c->d_scratch + DMK_SCRATCH_TMP + 300;
#define DMK_COMB(NV) { if ((rc = vecop_smem(dmk_ngmres_combine_kernel<NV>, NV))) return rc; \
    dmk_ngmres_combine_kernel<NV><<<blocks, 256, 2 * (NV + 1) * 256 * sizeof(double), c->stream>>>(out, x, c0, v, nv, n, partial); }
  DMK_VECOP_DISPATCH(DMK_COMB)
#undef DMK_COMB
  dmk_sum_partials_kernel<<<1, 96, 0, c->stream>>>(partial, blocks, nv + 1, d_out);
  c->launches += 2;
  DMK_CUDA_CHECK(cudaGetLastError());
  return reduce_to_host(c, d_out, nv + 1, dn);
}

// ------------------------------------------------------------------------------------------------ tangent bookkeeping
static int ensure_tangent(dmk_ctx* c) {
  DmkMech* m = c->mech;
  if (!m || !m->tangent_stale) return DMK_OK;
  double Cv[81], Cm[81];
  int rc = dmk_tangent_response(c, m->tangent_dt, Cv, Cm, nullptr, m->minmax_cells);
  if (rc) return rc;
  f2c3333(Cv, m->C_volAvg); f2c3333(Cm, m->C_minMaxAvg);
  m->tangent_stale = false;
  return DMK_OK;
}
int dmk_mech_before_point_forward(dmk_ctx* c) { return ensure_tangent(c); }

void dmk_mech_teardown(dmk_ctx* c) {
  DmkMech* m = c->mech;
  if (!m) return;
  cudaFree(m->F); cudaFree(m->F_lastInc); cudaFree(m->F_dot); cudaFree(m->R_field);
  if (m->W) cudaFree(m->W); if (m->dPdF) cudaFree(m->dPdF); if (m->d_singular) cudaFree(m->d_singular);
  cudaFree(m->XM); cudaFree(m->FM); cudaFree(m->XA); cudaFree(m->FA); cudaFree(m->Xcur); cudaFree(m->Fcur);
  for (double* p : m->Xdot) cudaFree(p);
  for (double* p : m->Fdot) cudaFree(p);
  delete m;
  c->mech = nullptr;
}

static int need_mech(dmk_ctx* c) {
  if (!c || !c->pts) { dmk_set_error("material not initialised (dmk_material_init)"); return DMK_ERR_STATE; }
  if (!c->mech) { dmk_set_error("mechanical solver not initialised (dmk_mech_init)"); return DMK_ERR_STATE; }
  return DMK_OK;
}

// ------------------------------------------------------------------------------------------------ residual evaluation
static int upload_block(dmk_ctx* c) {
  DmkMech* m = c->mech;
  double* h = c->h_pinned + 2048;
  memset(h, 0, 200 * sizeof(double));
  for (int a = 0; a < 9; a++) { h[B_FAIM + a] = m->F_aim[a]; h[B_PAIM + a] = m->P_aim[a]; h[B_MASK + a] = (double)m->stress_mask[a]; h[B_R + a] = m->R[a]; }
  for (int a = 0; a < 81; a++) h[B_S + a] = m->S[a];
  DMK_CUDA_CHECK(cudaMemcpyAsync(m->d_blk, h, 200 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream));           // h_pinned staging area may be reused
  return DMK_OK;
}

// evaluates r = Gamma * P(F) with all side effects of form_residual except the iteration bookkeeping
#define PROF(i) do { if (c->profiling) cudaEventRecord(c->ev[i], c->stream); } while (0)
static int evaluate(dmk_ctx* c, const double* d_F, double* d_r) {
  DmkMech* m = c->mech;
  int rc;
  PROF(0);
  if ((rc = dmk_pipe_begin(c, d_r))) return rc;                 // multi-GPU: forward x / y passes chunk by chunk under the per-point kernel
  if ((rc = dmk_points_constitutive(c, d_F, d_r, m->dt))) return rc;
  bool forward_done = false;
  if ((rc = dmk_pipe_end(c, &forward_done))) return rc;
  if (forward_done) {
    // cells integrated AFTER their chunk's launch (given up by the register-resident kernel, then cut back by another one)
    // make the pipelined transform stale: the host learns the count here (one small read behind the per-point kernel)
    DMK_CUDA_CHECK(cudaMemcpyAsync(c->h_pinned + 3072, dmk_points_given_up_total(c), sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (c->h_pinned[3072] > 0.0) { if ((rc = dmk_fft_forward_redo_pipelined(c, d_r))) return rc; }
  }
  PROF(1);
  m->tangent_stale = true; m->tangent_dt = m->dt;
  if ((rc = dmk_sum_records(c, d_r, c->n_local, m->d_blk + B_SUM))) return rc;
  if (c->P > 1) { if (c->comm.allreduce_sum(c->comm.user, m->d_blk + B_SUM, 9)) { dmk_set_error("allreduce failed"); return DMK_ERR_STATE; } }
  dmk_mech_aim_kernel<<<1, 32, 0, c->stream>>>(m->d_blk, c->wgt, c->d_dc, c->d_div);
  c->launches++;
  PROF(2);
  if (forward_done) { if ((rc = dmk_fft_forward_finish_pipelined(c))) return rc; }
  else if ((rc = dmk_fft_forward_internal(c, d_r))) return rc;
  PROF(3);
  if ((rc = dmk_gamma_internal(c))) return rc;
  PROF(4);
  if (c->P > 1) { if (c->comm.allreduce_sum(c->comm.user, c->d_div, 1)) { dmk_set_error("allreduce failed"); return DMK_ERR_STATE; } }
  if ((rc = dmk_fft_backward_internal(c, d_r, 1.0))) return rc;
  PROF(5);
  dmk_mech_collect_kernel<<<1, 32, 0, c->stream>>>(m->d_blk, c->d_div, dmk_points_given_up_counter(c));
  c->launches++;
  DMK_CUDA_CHECK(cudaMemcpyAsync(c->h_pinned, m->d_blk + B_OUT, 21 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  { const double st[2] = {c->h_pinned[20], 0.0}; m->stats.insert(m->stats.end(), st, st + 2); }
  for (int a = 0; a < 9; a++) { m->P_av[a] = c->h_pinned[a]; m->F_aim[a] = c->h_pinned[9 + a]; }
  m->err_BC = c->h_pinned[18]; m->div_raw = c->h_pinned[19];
  if (c->profiling) {      // sections: constitutive | P sum + aim | forward FFT (x, y) | fused z + Gamma + div | inverse FFT (y, x)
    for (int i = 0; i < 5; i++) { float ms = 0.f; cudaEventElapsedTime(&ms, c->ev[i], c->ev[i + 1]); c->prof_ms[i] += ms; }
    c->prof_n++;
  }
  m->err_div = dmk_finish_div(c, m->div_raw);
  return DMK_OK;
}

// ---- polarisation scheme: one evaluation of (r_F, r_F_tau) at (F, F_tau); algorithm and provenance: oracle/polarisation.py,
// polarisation.cuh (upstream DAMASK's grid_mechanical_spectral_polarisation; the reference tree holds its numerics keys only)
static int evaluate_polarisation(dmk_ctx* c, const double* d_X, double* d_R) {
  DmkMech* m = c->mech;
  const long long n9 = m->n9, n = c->n_local;
  const double *F = d_X, *Ft = d_X + n9;
  double *rF = d_R, *rFt = d_R + n9;
  int rc;
  double s9[9];
  // F_av
  if ((rc = dmk_sum_records(c, F, n, c->d_scratch + DMK_SCRATCH_SUM9))) return rc;
  if ((rc = reduce_to_host(c, c->d_scratch + DMK_SCRATCH_SUM9, 9, s9))) return rc;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m->F_av[3 * i + j] = s9[i + 3 * j] * c->wgt;
  // T = beta C:(F - I) - alpha F.(C:(F_tau - F - I));  T <- Gamma * T with the DC term rot(beta F_aim)
  DmkPolParams pp;
  for (int k = 0; k < 81; k++) pp.C[k] = m->C_scale[k];
  pp.alpha = m->alpha; pp.beta = m->beta;
  if ((rc = dmk_pol_pre(F, Ft, m->W, n, pp, c->stream))) return rc;
  {
    double aim[9], rot[9];
    for (int a = 0; a < 9; a++) aim[a] = m->beta * m->F_aim[a];
    rotate2(m->R, aim, true, rot);
    double* h = c->h_pinned + 2560;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) h[i + 3 * j] = rot[3 * i + j];
    DMK_CUDA_CHECK(cudaMemcpyAsync(c->d_dc, h, 9 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    DMK_CUDA_CHECK(cudaMemsetAsync(c->d_div, 0, sizeof(double), c->stream));
  }
  if ((rc = dmk_fft_forward_internal(c, m->W))) return rc;
  if ((rc = dmk_gamma_internal(c))) return rc;
  if ((rc = dmk_fft_backward_internal(c, m->W, 1.0))) return rc;
  // r_F_tau = beta F - T;  constitutive response at F - r_F_tau / beta (P lands in r_F)
  if ((rc = dmk_pol_rtau(F, m->W, rFt, m->W, n, m->beta, c->stream))) return rc;
  c->launches += 2;
  if ((rc = dmk_points_constitutive(c, m->W, rF, m->dt))) return rc;
  m->tangent_stale = true; m->tangent_dt = m->dt;
  if ((rc = dmk_sum_records(c, rF, n, c->d_scratch + DMK_SCRATCH_SUM9))) return rc;
  if ((rc = reduce_to_host(c, c->d_scratch + DMK_SCRATCH_SUM9, 9, s9))) return rc;
  {
    double Pm[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Pm[3 * i + j] = s9[i + 3 * j] * c->wgt;
    rotate2(m->R, Pm, false, m->P_av);
  }
  if ((rc = dmk_points_tangent(c, m->dt, c->d_scratch + DMK_SCRATCH_TMP, m->dPdF))) return rc;
  // stress boundary condition: F_aim -= S:(P_av - P_aim); err_BC = max |mask ? C:(F_aim - rot(F_av)) : P_av - P_aim|
  {
    double dP[9], Fr[9], dFa[9];
    for (int a = 0; a < 9; a++) dP[a] = m->P_av[a] - m->P_aim[a];
    for (int a = 0; a < 9; a++) { double s = 0.0; for (int b = 0; b < 9; b++) s += m->S[9 * a + b] * dP[b]; m->F_aim[a] -= s; }
    rotate2(m->R, m->F_av, false, Fr);
    for (int a = 0; a < 9; a++) dFa[a] = m->F_aim[a] - Fr[a];
    double err = 0.0;
    for (int a = 0; a < 9; a++) {
      double v = dP[a];
      if (m->stress_mask[a]) { v = 0.0; for (int b = 0; b < 9; b++) v += m->C_scale[9 * a + b] * dFa[b]; }
      err = std::max(err, fabs(v));
    }
    m->err_BC = err;
  }
  // err_div of the stress field (still in r_F), then r_F = (dPdF + C)^-1:(P - F.(C:(F_tau - F - I))) + r_F_tau, then err_curl of F
  double* d_two = c->d_scratch + DMK_SCRATCH_TMP + 320;
  DMK_CUDA_CHECK(cudaMemsetAsync(d_two, 0, 2 * sizeof(double), c->stream));
  if ((rc = dmk_spectrum_rms(c, rF, 0, d_two))) return rc;
  if ((rc = dmk_pol_rF(F, Ft, m->dPdF, rFt, rF, n, pp, m->d_singular, c->stream))) return rc;
  c->launches++;
  if ((rc = dmk_spectrum_rms(c, F, 1, d_two + 1))) return rc;
  double two[2];
  if ((rc = reduce_to_host(c, d_two, 2, two))) return rc;
  m->div_raw = two[0];
  m->err_div = dmk_finish_div(c, two[0]);
  m->err_curl = dmk_finish_div(c, two[1]);
  int sing = 0;
  DMK_CUDA_CHECK(cudaMemcpy(&sing, m->d_singular, sizeof(int), cudaMemcpyDeviceToHost));
  if (sing) { dmk_set_error("polarisation residual: dPdF + C_scale is singular in a cell"); return DMK_ERR_STATE; }
  return DMK_OK;
}

static int form_residual(dmk_ctx* c, const double* d_F, double* d_r, int nfuncs, int petsc_iter) {
  DmkMech* m = c->mech;
  if (nfuncs == 0 && petsc_iter == 0) m->total_iter = -1;
  if (m->total_iter <= petsc_iter) m->total_iter += 1;
  int rc = m->scheme == 1 ? evaluate_polarisation(c, d_F, d_r) : evaluate(c, d_F, d_r);
  if (rc) return rc;
  double rec[12];
  for (int a = 0; a < 9; a++) rec[a] = m->P_av[a];
  rec[9] = m->err_div; rec[10] = m->err_BC; rec[11] = (double)m->total_iter;
  m->log.insert(m->log.end(), rec, rec + 12);
  if (m->capture && m->scheme == 0) {
    const size_t o = m->capF.size();
    m->capF.resize(o + m->n9); m->capR.resize(o + m->n9);
    DMK_CUDA_CHECK(cudaMemcpy(m->capF.data() + o, d_F, m->n9 * sizeof(double), cudaMemcpyDeviceToHost));
    DMK_CUDA_CHECK(cudaMemcpy(m->capR.data() + o, d_r, m->n9 * sizeof(double), cudaMemcpyDeviceToHost));
  }
  return DMK_OK;
}

// MechSolverBasic::converged -> SNES reason (0 = keep iterating)
static int converged(dmk_ctx* c, int* reason) {
  DmkMech* m = c->mech;
  const dmk_numerics& n = c->num;
  double maxP = m->P_av[0];
  for (int a = 1; a < 9; a++) maxP = std::max(maxP, m->P_av[a]);            // signed maximum, as in the reference
  const double divTol = std::max(maxP * n.eps_rel_div, n.eps_abs_div);
  const double BCTol = std::max(maxP * n.eps_rel_P, n.eps_abs_P);
  int ill = 0;
  int rc = dmk_terminally_ill(c, &ill, 0); if (rc) return rc;
  if (m->scheme == 1) {
    // upstream's polarisation test: three criteria, tolerances from max |P_av| and max |F_aim - I|
    double aP = 0.0, aF = 0.0;
    for (int a = 0; a < 9; a++) { aP = std::max(aP, fabs(m->P_av[a])); aF = std::max(aF, fabs(m->F_aim[a] - ((a % 4 == 0) ? 1.0 : 0.0))); }
    const double dTol = std::max(aP * n.eps_rel_div, n.eps_abs_div), bTol = std::max(aP * n.eps_rel_P, n.eps_abs_P);
    const double cTol = std::max(aF * m->eps_rel_curl, m->eps_abs_curl);
    const bool ok = std::max(std::max(m->err_div / dTol, m->err_curl / cTol), m->err_BC / bTol) < 1.0;
    if ((m->total_iter >= n.itmin && ok) || ill) *reason = 6;
    else if (m->total_iter >= n.itmax) *reason = -5;
    else *reason = 0;
    const double rep[6] = {(double)m->total_iter, m->err_div, dTol, m->err_BC, bTol, (double)*reason};
    m->reports.insert(m->reports.end(), rep, rep + 6);
    return DMK_OK;
  }
  if ((m->total_iter >= n.itmin && std::max(m->err_div / divTol, m->err_BC / BCTol) < 1.0) || ill) *reason = 6;   // SNES_BREAKOUT_INNER_ITER
  else if (m->total_iter >= n.itmax) *reason = -5;                                                                  // SNES_DIVERGED_MAX_IT
  else *reason = 0;
  const double rep[6] = {(double)m->total_iter, m->err_div, divTol, m->err_BC, BCTol, (double)*reason};
  m->reports.insert(m->reports.end(), rep, rep + 6);
  return DMK_OK;
}

// ------------------------------------------------------------------------------------------------ small dense least squares
// minimum-norm least-squares solution of the symmetric system H a = b through the eigen-decomposition of H
// (cyclic Jacobi), dropping |lambda| <= eps * max|lambda| -- the same solution LAPACK dgelss returns with rcond = -1.
static void sym_lstsq(std::vector<double> H, int n, const double* b, double* x) {
  std::vector<double> V((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) V[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { if (i != j) off += H[(size_t)i * n + j] * H[(size_t)i * n + j]; else diag += H[(size_t)i * n + i] * H[(size_t)i * n + i]; }
    if (off <= 1e-32 * diag || off == 0.0) break;
    for (int p = 0; p < n - 1; p++) for (int q = p + 1; q < n; q++) {
      const double apq = H[(size_t)p * n + q];
      if (apq == 0.0) continue;
      const double theta = (H[(size_t)q * n + q] - H[(size_t)p * n + p]) / (2.0 * apq);
      const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
      for (int k = 0; k < n; k++) { const double akp = H[(size_t)k * n + p], akq = H[(size_t)k * n + q]; H[(size_t)k * n + p] = cs * akp - sn * akq; H[(size_t)k * n + q] = sn * akp + cs * akq; }
      for (int k = 0; k < n; k++) { const double apk = H[(size_t)p * n + k], aqk = H[(size_t)q * n + k]; H[(size_t)p * n + k] = cs * apk - sn * aqk; H[(size_t)q * n + k] = sn * apk + cs * aqk; }
      for (int k = 0; k < n; k++) { const double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q]; V[(size_t)k * n + p] = cs * vkp - sn * vkq; V[(size_t)k * n + q] = sn * vkp + cs * vkq; }
    }
  }
  double lmax = 0.0;
  for (int i = 0; i < n; i++) lmax = std::max(lmax, fabs(H[(size_t)i * n + i]));
  const double cut = 2.220446049250313e-16 * lmax;
  for (int i = 0; i < n; i++) x[i] = 0.0;
  for (int k = 0; k < n; k++) {
    const double lam = H[(size_t)k * n + k];
    if (!(fabs(lam) > cut)) continue;
    double vb = 0.0;
    for (int i = 0; i < n; i++) vb += V[(size_t)i * n + k] * b[i];
    for (int i = 0; i < n; i++) x[i] += V[(size_t)i * n + k] * vb / lam;
  }
}

// ------------------------------------------------------------------------------------------------ NGMRES
static int hist_vec(dmk_ctx* c, std::vector<double*>& h, int i, double** out) {
  while ((int)h.size() <= i) { double* p = nullptr; DMK_CUDA_CHECK(cudaMalloc((void**)&p, c->mech->nvec * sizeof(double))); h.push_back(p); }
  *out = h[i];
  return DMK_OK;
}
static int copy_field(dmk_ctx* c, double* dst, const double* src) {
  DMK_CUDA_CHECK(cudaMemcpyAsync(dst, src, c->mech->nvec * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  return DMK_OK;
}

struct DmkLap {                                      // every operation timed here ends in a host synchronisation (reduce_to_host / evaluate)
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  double lap() { auto t1 = std::chrono::steady_clock::now(); const double d = std::chrono::duration<double>(t1 - t0).count(); t0 = t1; return d; }
};
static int ngmres_solve(dmk_ctx* c) {
  DmkMech* m = c->mech;
  DmkLap lapse, whole;
  struct Total { DmkMech* m; DmkLap* w; ~Total() { m->t_solver[6] += w->lap(); } } total_guard{m, &whole};
  const double gammaA = 2.0, gammaC = 2.0, epsilonB = 0.1, deltaB = 0.9;
  const int restart_it = 2, max_its = 10000, M = NGMRES_M;
  std::vector<double> Q((size_t)M * M, 0.0);
  int rc, reason = 0, nfuncs = 0, it = 0;
  double *X = m->Xcur, *Fv = m->Fcur;
  if ((rc = copy_field(c, X, m->F))) return rc;
  if ((rc = form_residual(c, X, Fv, nfuncs, it))) return rc; nfuncs++;
  double tmp[64];
  { std::vector<const double*> b = {Fv}; if ((rc = dots(c, Fv, b, tmp))) return rc; }
  double fnorm = sqrt(tmp[0]), fminnorm = fnorm;
  if ((rc = converged(c, &reason))) return rc;
  m->snes_its = 0; m->snes_reason = reason;
  if (reason) return copy_field(c, m->F, X);
  double *x0, *f0;
  if ((rc = hist_vec(c, m->Xdot, 0, &x0)) || (rc = hist_vec(c, m->Fdot, 0, &f0))) return rc;
  if ((rc = copy_field(c, x0, X)) || (rc = copy_field(c, f0, Fv))) return rc;
  int k_restart = 1, l = 1, ivec = 0, restart_count = 0;
  DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream)); m->t_solver[11] += lapse.lap();           // first evaluation and set-up of the history
  for (int k = 1; k <= max_its; k++) {
    // X^M = X - F(X); F^M
    if (vecops_split()) { std::vector<const double*> v = {Fv}; const double a[1] = {-1.0}; if ((rc = lincomb(c, m->XM, X, 1.0, v, a))) return rc; }
    else { dmk_ngmres_xm_kernel<<<DMK_VECOP_BLOCKS, 256, 0, c->stream>>>(m->XM, X, Fv, m->nvec); c->launches++; DMK_CUDA_CHECK(cudaGetLastError()); }
    DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream)); m->t_solver[5] += lapse.lap();
    if ((rc = form_residual(c, m->XM, m->FM, nfuncs, it))) return rc; nfuncs++;
    m->t_solver[0] += lapse.lap();
    // nu, xi_i = <F^M, Fdot_i>, beta_i = <Fdot_ivec, Fdot_i>
    std::vector<const double*> fd(l);
    for (int i = 0; i < l; i++) fd[i] = m->Fdot[i];
    std::vector<const double*> b1 = fd; b1.push_back(m->FM);
    std::vector<double> xi(l + 1), beta(l);
    if (vecops_split()) {
      if ((rc = dots(c, m->FM, b1, xi.data()))) return rc;
      if ((rc = dots(c, m->Fdot[ivec], fd, beta.data()))) return rc;
    } else if ((rc = ngmres_dots(c, m->FM, fd, ivec, xi.data(), beta.data(), &xi[l]))) return rc;
    m->t_solver[1] += lapse.lap();
    const double nu = xi[l], fMnorm = sqrt(nu);
    for (int i = 0; i < l; i++) { Q[(size_t)i * M + ivec] = beta[i]; Q[(size_t)ivec * M + i] = beta[i]; }
    std::vector<double> H((size_t)l * l), rhs(l), alpha(l);
    for (int i = 0; i < l; i++) { rhs[i] = nu - xi[i]; for (int j = 0; j < l; j++) H[(size_t)i * l + j] = Q[(size_t)i * M + j] - xi[i] - xi[j] + nu; }
    if (l == 1) alpha[0] = (H[0] != 0.0) ? rhs[0] / H[0] : 0.0;
    else sym_lstsq(H, l, rhs.data(), alpha.data());
    double asum = 0.0; for (int i = 0; i < l; i++) asum += alpha[i];
    m->t_solver[2] += lapse.lap();
    // X^A = (1 - sum alpha) X^M + sum alpha_i Xdot_i ; F^A
    std::vector<double> dn(l + 1);
    { std::vector<const double*> xd(l); for (int i = 0; i < l; i++) xd[i] = m->Xdot[i];
      if (vecops_split()) { if ((rc = lincomb(c, m->XA, m->XM, 1.0 - asum, xd, alpha.data()))) return rc; }
      else if ((rc = ngmres_combine(c, m->XA, m->XM, 1.0 - asum, xd, alpha.data(), dn.data()))) return rc; }   // + the distances used below
    DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream)); m->t_solver[3] += lapse.lap();
    if ((rc = form_residual(c, m->XA, m->FA, nfuncs, it))) return rc; nfuncs++;
    m->t_solver[0] += lapse.lap();
    if (fminnorm > fMnorm) fminnorm = fMnorm;
    // norms for selection and restart
    { std::vector<const double*> b = {m->FA}; if ((rc = dots(c, m->FA, b, tmp))) return rc; }
    const double fAnorm = sqrt(tmp[0]);
    if (vecops_split()) {
      std::vector<const double*> dv(l + 1);
      for (int i = 0; i < l; i++) dv[i] = m->Xdot[i];
      dv[l] = m->XM;
      if ((rc = diffnorms2(c, m->XA, dv, dn.data()))) return rc;
    }
    m->t_solver[4] += lapse.lap();
    const double dnorm = sqrt(dn[l]);
    double dminnorm = sqrt(dn[0]);
    for (int i = 1; i < l; i++) dminnorm = std::min(dminnorm, sqrt(dn[i]));
    bool selectA = true;
    if (fAnorm >= gammaA * fminnorm) selectA = false;
    if (!(epsilonB * dnorm < dminnorm || sqrt(fnorm) < deltaB * sqrt(fminnorm))) selectA = false;
    if (selectA) { std::swap(m->XA, X); std::swap(m->FA, Fv); fnorm = fAnorm; }
    else { std::swap(m->XM, X); std::swap(m->FM, Fv); fnorm = fMnorm; }
    m->Xcur = X; m->Fcur = Fv;
    bool selectRestart = false;
    if (epsilonB * dnorm > dminnorm && sqrt(fAnorm) > deltaB * sqrt(fminnorm) && l > 0) selectRestart = true;
    if (sqrt(fAnorm) > gammaC * sqrt(fminnorm)) selectRestart = true;
    restart_count = selectRestart ? restart_count + 1 : 0;
    ivec = k_restart % M;
    if (restart_count >= restart_it) {
      restart_count = 0; k_restart = 1; l = 1; ivec = 0;
      if ((rc = copy_field(c, m->Xdot[0], X)) || (rc = copy_field(c, m->Fdot[0], Fv))) return rc;
    } else {
      if (l < M) l++;
      k_restart++;
      if (fminnorm > fnorm) fminnorm = fnorm;
      double *xs, *fs;
      if ((rc = hist_vec(c, m->Xdot, ivec, &xs)) || (rc = hist_vec(c, m->Fdot, ivec, &fs))) return rc;
      DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream)); m->t_solver[8] += lapse.lap();      // selection + allocation of history slots
      if ((rc = copy_field(c, xs, X)) || (rc = copy_field(c, fs, Fv))) return rc;
    }
    DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream)); m->t_solver[9] += lapse.lap();        // history copies
    it = k;
    m->t_solver[7] += 1.0;
    if ((rc = converged(c, &reason))) return rc;
    m->t_solver[10] += lapse.lap();                                                          // convergence test
    if (reason) break;
  }
  if (!reason) reason = -5;
  m->snes_its = it; m->snes_reason = reason;
  return copy_field(c, m->F, X);
}

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" int dmk_update_gamma(dmk_ctx* c, const double* C);

static int mech_init(dmk_ctx* c, const long long* gamma_pair_cells, int scheme) {
  if (!c || !c->pts) { dmk_set_error("material not initialised (dmk_material_init)"); return DMK_ERR_STATE; }
  if (c->P > 1 && (!c->has_comm || !c->comm.allreduce_sum || !c->comm.allreduce_max || !c->comm.alltoall)) {
    dmk_set_error("%d ranks but no collective callbacks: call dmk_set_comm before dmk_mech_init", c->P); return DMK_ERR_STATE;
  }
  dmk_mech_teardown(c);
  DmkMech* m = new DmkMech();
  c->mech = m;
  m->n9 = c->n_local * 9;
  m->scheme = scheme;
  m->nvec = (scheme == 1 ? 2 : 1) * m->n9;               // polarisation: every solution-sized field is (F | F_tau)
  m->W = m->dPdF = nullptr; m->d_singular = nullptr;
  m->d_blk = c->d_scratch + DMK_SCRATCH_MECH;
  double** fields[] = {&m->F, &m->F_lastInc, &m->F_dot, &m->R_field, &m->XM, &m->FM, &m->XA, &m->FA, &m->Xcur, &m->Fcur};
  for (double** f : fields) { DMK_CUDA_CHECK(cudaMalloc((void**)f, m->nvec * sizeof(double))); }
  dmk_fill_identity_kernel<<<148 * 4, 256, 0, c->stream>>>(m->F, c->n_local);
  dmk_fill_identity_kernel<<<148 * 4, 256, 0, c->stream>>>(m->F_lastInc, c->n_local);
  c->launches += 2;
  DMK_CUDA_CHECK(cudaMemsetAsync(m->F_dot, 0, m->nvec * sizeof(double), c->stream));
  const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  memcpy(m->F_aim, I3, sizeof I3); memcpy(m->F_aim_lastInc, I3, sizeof I3); memcpy(m->R, I3, sizeof I3);
  memset(m->F_aim_dot, 0, sizeof m->F_aim_dot); memset(m->P_av, 0, sizeof m->P_av); memset(m->P_aim, 0, sizeof m->P_aim);
  memset(m->S, 0, sizeof m->S);
  for (int a = 0; a < 9; a++) m->stress_mask[a] = 1;
  m->dt = 0.0; m->err_div = m->err_BC = m->div_raw = 0.0; m->total_iter = 0; m->tangent_stale = false; m->tangent_dt = 0.0;
  m->snes_its = 0; m->snes_reason = 0; m->minmax_cells[0] = m->minmax_cells[1] = -1; m->capture = false;
  // constitutive_response at F = I, dt = 0 (mech_solver_basic.cpp:96-101): P_av and the initial tangents
  int rc;
  if ((rc = dmk_points_constitutive(c, m->F, m->R_field, 0.0))) return rc;
  if ((rc = dmk_sum_records(c, m->R_field, c->n_local, m->d_blk + B_SUM))) return rc;
  double s9[9];
  if ((rc = reduce_to_host(c, m->d_blk + B_SUM, 9, s9))) return rc;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m->P_av[3 * i + j] = s9[i + 3 * j] * c->wgt;
  m->tangent_stale = true; m->tangent_dt = 0.0;
  if ((rc = ensure_tangent(c))) return rc;
  if (gamma_pair_cells && !dmk_points_single_constituent(c)) { dmk_set_error("forced reference-stiffness cells with N_constituents > 1 are not offered"); return DMK_ERR_UNSUPPORTED; }
  if (gamma_pair_cells) {
    // replay aid: force the two cells whose tangents define the reference stiffness
    double* d_two = c->d_scratch + DMK_SCRATCH_TMP + 112;
    DMK_CUDA_CHECK(cudaMemsetAsync(d_two, 0, 162 * sizeof(double), c->stream));
    const long long offset = (long long)c->z_off * c->nx * c->ny;
    for (int k = 0; k < 2; k++) {
      const long long loc = gamma_pair_cells[k] - offset;
      if (loc >= 0 && loc < c->n_local) { if ((rc = dmk_points_tangent_cell(c, 0.0, loc, d_two + 81 * k))) return rc; }
    }
    double two[162];
    if ((rc = reduce_to_host(c, d_two, 162, two))) return rc;
    for (int k = 0; k < 81; k++) m->C_minMaxAvg[k] = 0.5 * (two[k] + two[81 + k]);
    m->minmax_cells[0] = gamma_pair_cells[0]; m->minmax_cells[1] = gamma_pair_cells[1];
  }
  memcpy(m->C_volAvgLastInc, m->C_volAvg, sizeof m->C_volAvg); memcpy(m->C_minMaxAvgLastInc, m->C_minMaxAvg, sizeof m->C_minMaxAvg);
  double Cf[81];
  c2f3333(m->C_minMaxAvg, Cf);
  return dmk_update_gamma(c, Cf);
}

extern "C" int dmk_mech_init(dmk_ctx* c, const long long* gamma_pair_cells) { return mech_init(c, gamma_pair_cells, 0); }

// the polarisation variant (solver: mechanical: spectral_polarisation of the load file): alpha, beta and the curl tolerances
// are the numerics keys the reference parses (config.cpp:450-457; defaults 1, 1, 1e-10, 5e-4); F_tau = 2 F initially,
// C_scale = C_minMaxAvg, S_scale = its inverse
static __global__ void dmk_scale_kernel(double* __restrict__ out, const double* __restrict__ in, double s, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = s * in[i];
}
extern "C" int dmk_mech_init_polarisation(dmk_ctx* c, double alpha, double beta, double eps_abs_curl, double eps_rel_curl) {
  if (!(beta != 0.0)) { dmk_set_error("polarisation: beta must not be 0"); return DMK_ERR_ARG; }
  if (c && !dmk_points_single_constituent(c)) { dmk_set_error("polarisation scheme with N_constituents > 1 is not offered"); return DMK_ERR_UNSUPPORTED; }
  int rc = mech_init(c, nullptr, 1);
  if (rc) return rc;
  DmkMech* m = c->mech;
  m->alpha = alpha; m->beta = beta; m->eps_abs_curl = eps_abs_curl; m->eps_rel_curl = eps_rel_curl; m->err_curl = 0.0;
  DMK_CUDA_CHECK(cudaMalloc((void**)&m->W, m->n9 * sizeof(double)));
  DMK_CUDA_CHECK(cudaMalloc((void**)&m->dPdF, (size_t)c->n_local * 81 * sizeof(double)));
  DMK_CUDA_CHECK(cudaMalloc((void**)&m->d_singular, sizeof(int)));
  DMK_CUDA_CHECK(cudaMemsetAsync(m->d_singular, 0, sizeof(int), c->stream));
  dmk_scale_kernel<<<148 * 8, 256, 0, c->stream>>>(m->F + m->n9, m->F, 2.0, m->n9);
  dmk_scale_kernel<<<148 * 8, 256, 0, c->stream>>>(m->F_lastInc + m->n9, m->F_lastInc, 2.0, m->n9);
  c->launches += 2;
  DMK_CUDA_CHECK(cudaGetLastError());
  memcpy(m->C_scale, m->C_minMaxAvg, sizeof m->C_scale);
  if (!dmk_invsym3333(m->C_minMaxAvg, m->S_scale)) { dmk_set_error("polarisation: C_minMaxAvg is singular"); return DMK_ERR_STATE; }
  memset(m->F_av, 0, sizeof m->F_av); m->F_av[0] = m->F_av[4] = m->F_av[8] = 1.0;
  return DMK_OK;
}

// deformation_values / masks / R: row-major 3x3; mask 1 = 'x'
extern "C" int dmk_mech_forward(dmk_ctx* c, int cut_back, int guess, double dt, double dt_old, double t_remaining,
                                const char* deformation_type, const double* deformation_values, const int* deformation_mask,
                                const char* stress_type, const double* stress_values, const int* stress_mask, const double* R) {
  int rc = need_mech(c); if (rc) return rc;
  DmkMech* m = c->mech;
  if ((rc = ensure_tangent(c))) return rc;
  const std::string dtype(deformation_type ? deformation_type : ""), stype(stress_type ? stress_type : "");
  if (cut_back) {
    memcpy(m->C_volAvg, m->C_volAvgLastInc, sizeof m->C_volAvg); memcpy(m->C_minMaxAvg, m->C_minMaxAvgLastInc, sizeof m->C_minMaxAvg);
  } else {
    memcpy(m->C_volAvgLastInc, m->C_volAvg, sizeof m->C_volAvg); memcpy(m->C_minMaxAvgLastInc, m->C_minMaxAvg, sizeof m->C_minMaxAvg);
    for (int a = 0; a < 9; a++) m->F_aim_dot[a] = (guess && !stress_mask[a]) ? (m->F_aim[a] - m->F_aim_lastInc[a]) / dt_old : 0.0;
    memcpy(m->F_aim_lastInc, m->F_aim, sizeof m->F_aim);
    double step[9];
    if (dtype == "L") mm33(deformation_values, m->F_aim_lastInc, step);            // upstream: L . F_aim_lastInc
    else if (dtype == "dot_F") memcpy(step, deformation_values, sizeof step);
    else if (dtype == "F") { for (int a = 0; a < 9; a++) step[a] = (deformation_values[a] - m->F_aim_lastInc[a]) / t_remaining; }
    else { dmk_set_error("unknown deformation BC type '%s'", dtype.c_str()); return DMK_ERR_ARG; }
    for (int a = 0; a < 9; a++) if (!deformation_mask[a]) m->F_aim_dot[a] += step[a];
    double rot[9], rec[9];
    rotate2(R, m->F_aim_dot, true, rot);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) rec[i + 3 * j] = rot[3 * i + j];
    if ((rc = dmk_calculate_rate(c, guess, m->F_lastInc, m->F, dt_old, rec, m->F_dot))) return rc;
    if (m->scheme == 1) { if ((rc = dmk_calculate_rate(c, guess, m->F_lastInc + m->n9, m->F + m->n9, dt_old, rec, m->F_dot + m->n9))) return rc; }
    DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if ((rc = copy_field(c, m->F_lastInc, m->F))) return rc;                       // (F | F_tau) in one copy
  }
  for (int a = 0; a < 9; a++) m->F_aim[a] = m->F_aim_lastInc[a] + m->F_aim_dot[a] * dt;
  if (stype == "P") { for (int a = 0; a < 9; a++) if (!stress_mask[a]) m->P_aim[a] += (stress_values[a] - m->P_aim[a]) / t_remaining * dt; }
  else if (stype == "dot_P") { for (int a = 0; a < 9; a++) if (!stress_mask[a]) m->P_aim[a] += stress_values[a] * dt; }
  else { dmk_set_error("unknown stress BC type '%s'", stype.c_str()); return DMK_ERR_ARG; }
  double rot[9], rec[9];
  rotate2(R, m->F_aim, true, rot);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) rec[i + 3 * j] = rot[3 * i + j];
  if ((rc = dmk_forward_field(c, dt, m->F_lastInc, m->F_dot, rec, m->F))) return rc;
  if (m->scheme == 1) {
    if (guess) { if ((rc = dmk_forward_field(c, dt, m->F_lastInc + m->n9, m->F_dot + m->n9, nullptr, m->F + m->n9))) return rc; }
    else { if ((rc = dmk_pol_forward_noguess(m->F, m->F + m->n9, c->n_local, m->C_scale, m->S_scale, c->stream))) return rc; c->launches++; }
  }
  DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  for (int a = 0; a < 9; a++) { m->stress_mask[a] = stress_mask[a]; m->R[a] = R[a]; }
  m->dt = dt;
  return DMK_OK;
}

extern "C" int dmk_mech_calculate_solution(dmk_ctx* c, dmk_solution_state* out) {
  int rc = need_mech(c); if (rc) return rc;
  DmkMech* m = c->mech;
  if ((rc = ensure_tangent(c))) return rc;
  if ((rc = masked_compliance_rm(m->C_volAvg, m->R, m->stress_mask, m->S))) return rc;
  if (c->num.update_gamma) {
    double Cf[81]; c2f3333(m->C_minMaxAvg, Cf); if ((rc = dmk_update_gamma(c, Cf))) return rc;
    if (m->scheme == 1) {
      memcpy(m->C_scale, m->C_minMaxAvg, sizeof m->C_scale);
      if (!dmk_invsym3333(m->C_minMaxAvg, m->S_scale)) { dmk_set_error("polarisation: C_minMaxAvg is singular"); return DMK_ERR_STATE; }
    }
  }
  if ((rc = upload_block(c))) return rc;
  if ((rc = ngmres_solve(c))) return rc;
  int ill = 0;
  if ((rc = dmk_terminally_ill(c, &ill, 1))) return rc;
  if (out) { out->converged = m->snes_reason > 0; out->required_iterations = m->total_iter; out->terminally_ill = ill; out->stag_converged = 1; }
  for (int a = 0; a < 9; a++) if (m->stress_mask[a]) m->P_aim[a] = m->P_av[a];
  return DMK_OK;
}

// what calculate_solution does before SNESSolve: tangent reductions, masked compliance, optional Gamma update
extern "C" int dmk_mech_prepare(dmk_ctx* c) {
  int rc = need_mech(c); if (rc) return rc;
  DmkMech* m = c->mech;
  if ((rc = ensure_tangent(c))) return rc;
  if ((rc = masked_compliance_rm(m->C_volAvg, m->R, m->stress_mask, m->S))) return rc;
  if (c->num.update_gamma) {
    double Cf[81]; c2f3333(m->C_minMaxAvg, Cf); if ((rc = dmk_update_gamma(c, Cf))) return rc;
    if (m->scheme == 1) {
      memcpy(m->C_scale, m->C_minMaxAvg, sizeof m->C_scale);
      if (!dmk_invsym3333(m->C_minMaxAvg, m->S_scale)) { dmk_set_error("polarisation: C_minMaxAvg is singular"); return DMK_ERR_STATE; }
    }
  }
  return DMK_OK;
}
// 3x3 host values in Fortran order: "F_aim", "F_aim_lastInc", "P_aim"
extern "C" int dmk_mech_set(dmk_ctx* c, const char* what, const double* values) {
  int rc = need_mech(c); if (rc) return rc;
  DmkMech* m = c->mech;
  const std::string w(what ? what : "");
  double* dst = w == "F_aim" ? m->F_aim : w == "F_aim_lastInc" ? m->F_aim_lastInc : w == "P_aim" ? m->P_aim : nullptr;
  if (!dst) { dmk_set_error("dmk_mech_set: unknown quantity '%s'", w.c_str()); return DMK_ERR_ARG; }
  f2c33(values, dst);
  return DMK_OK;
}
extern "C" int dmk_mech_set_field_device(dmk_ctx* c, const char* what, const double* d_src) {
  int rc = need_mech(c); if (rc) return rc;
  double* f = nullptr;
  const std::string w(what ? what : "");
  if (w == "F") f = c->mech->F; else if (w == "F_lastInc") f = c->mech->F_lastInc; else if (w == "F_dot") f = c->mech->F_dot;
  if (!f) { dmk_set_error("dmk_mech_set_field_device: unknown field '%s'", w.c_str()); return DMK_ERR_ARG; }
  DMK_CUDA_CHECK(cudaMemcpyAsync(f, d_src, c->mech->n9 * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  return DMK_OK;
}

extern "C" int dmk_mech_form_residual(dmk_ctx* c, const double* d_F, double* d_r) {
  int rc = need_mech(c); if (rc) return rc;
  if (c->mech->scheme == 1) {                               // (F | F_tau) in, (r_F | r_F_tau) out: 2 x 9 doubles per cell
    if ((rc = evaluate_polarisation(c, d_F, d_r))) return rc;
    DmkMech* mp = c->mech;
    double rec[12];
    for (int a = 0; a < 9; a++) rec[a] = mp->P_av[a];
    rec[9] = mp->err_div; rec[10] = mp->err_BC; rec[11] = (double)mp->total_iter;
    mp->log.insert(mp->log.end(), rec, rec + 12);
    return DMK_OK;
  }
  if ((rc = upload_block(c))) return rc;
  if ((rc = evaluate(c, d_F, d_r))) return rc;
  DmkMech* m = c->mech;
  double rec[12];
  for (int a = 0; a < 9; a++) rec[a] = m->P_av[a];
  rec[9] = m->err_div; rec[10] = m->err_BC; rec[11] = (double)m->total_iter;
  m->log.insert(m->log.end(), rec, rec + 12);
  return DMK_OK;
}

extern "C" int dmk_mech_form_residual_host(dmk_ctx* c, const double* F, double* r) {
  int rc = need_mech(c); if (rc) return rc;
  DmkMech* m = c->mech;
  if (m->scheme == 1) { dmk_set_error("dmk_mech_form_residual_host: basic scheme only (the polarisation residual is evaluated on device vectors)"); return DMK_ERR_UNSUPPORTED; }
  // F goes up in z-chunks on a second stream; the per-point kernel of chunk k starts when chunk k has arrived, so the
  // host-to-device copy (72 B/cell over PCIe) hides behind the constitutive update of the chunks before it
  static const int chunks_setting = [] { const char* e = getenv("DMK_STAGE_CHUNKS"); int v = e ? atoi(e) : 8; return v < 1 ? 1 : (v > 8 ? 8 : v); }();
  const int chunks = (c->nz_loc >= chunks_setting) ? chunks_setting : 1;
  if (chunks > 1) {
    if (!c->stage_ready) {
      DMK_CUDA_CHECK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
      for (int k = 0; k < 8; k++) DMK_CUDA_CHECK(cudaEventCreateWithFlags(&c->stage_ev[k], cudaEventDisableTiming));
      DMK_CUDA_CHECK(cudaEventCreateWithFlags(&c->stage_begin, cudaEventDisableTiming));
      c->stage_ready = true;
    }
    DMK_CUDA_CHECK(cudaEventRecord(c->stage_begin, c->stream));             // XM may still be read by earlier work
    DMK_CUDA_CHECK(cudaStreamWaitEvent(c->copy_stream, c->stage_begin, 0));
    const long long plane9 = 9LL * c->nx * c->ny;
    for (int k = 0; k < chunks; k++) {
      const long long z0 = (long long)c->nz_loc * k / chunks, z1 = (long long)c->nz_loc * (k + 1) / chunks;
      DMK_CUDA_CHECK(cudaMemcpyAsync(m->XM + z0 * plane9, F + z0 * plane9, (size_t)((z1 - z0) * plane9) * sizeof(double),
                                     cudaMemcpyHostToDevice, c->copy_stream));
      DMK_CUDA_CHECK(cudaEventRecord(c->stage_ev[k], c->copy_stream));
    }
    c->stage_chunks = chunks; c->stage_dst = m->XM;
  } else {
    DMK_CUDA_CHECK(cudaMemcpyAsync(m->XM, F, m->n9 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  }
  rc = dmk_mech_form_residual(c, m->XM, m->FM);
  if (chunks > 1) {
    c->stage_chunks = 0; c->stage_dst = nullptr;
  }
  if (rc) return rc;
  DMK_CUDA_CHECK(cudaMemcpyAsync(r, m->FM, m->n9 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  return DMK_OK;
}

// named host scalars; 3x3 and 3x3x3x3 results in Fortran order
extern "C" int dmk_mech_get(dmk_ctx* c, const char* what, double* out, size_t n) {
  int rc = need_mech(c); if (rc) return rc;
  DmkMech* m = c->mech;
  const std::string w(what ? what : "");
  auto put33 = [&](const double* a) -> int { if (n < 9) { dmk_set_error("need 9 doubles"); return DMK_ERR_ARG; } c2f33(a, out); return DMK_OK; };
  auto put3333 = [&](const double* a) -> int { if (n < 81) { dmk_set_error("need 81 doubles"); return DMK_ERR_ARG; } c2f3333(a, out); return DMK_OK; };
  if (w == "P_av") return put33(m->P_av);
  if (w == "F_aim") return put33(m->F_aim);
  if (w == "P_aim") return put33(m->P_aim);
  if (w == "F_aim_dot") return put33(m->F_aim_dot);
  if (w == "C_volAvg" || w == "C_minMaxAvg") { if ((rc = ensure_tangent(c))) return rc; return put3333(w == "C_volAvg" ? m->C_volAvg : m->C_minMaxAvg); }
  if (w == "S") return put3333(m->S);
  if (n < 1) { dmk_set_error("need 1 double"); return DMK_ERR_ARG; }
  if (w == "err_div") { out[0] = m->err_div; return DMK_OK; }
  if (w == "err_BC") { out[0] = m->err_BC; return DMK_OK; }
  if (w == "div_raw") { out[0] = m->div_raw; return DMK_OK; }
  if (w == "err_curl") { out[0] = m->scheme == 1 ? m->err_curl : 0.0; return DMK_OK; }
  if (w == "scheme") { out[0] = m->scheme; return DMK_OK; }
  if (w == "total_iter") { out[0] = m->total_iter; return DMK_OK; }
  if (w == "snes_its") { out[0] = m->snes_its; return DMK_OK; }
  if (w == "solver_seconds") { if (n < 12) { dmk_set_error("need 12 doubles"); return DMK_ERR_ARG; } for (int k = 0; k < 12; k++) out[k] = m->t_solver[k]; return DMK_OK; }
  if (w == "snes_reason") { out[0] = m->snes_reason; return DMK_OK; }
  if (w == "minmax_cells") { if (n < 2) { dmk_set_error("need 2 doubles"); return DMK_ERR_ARG; } if ((rc = ensure_tangent(c))) return rc; out[0] = (double)m->minmax_cells[0]; out[1] = (double)m->minmax_cells[1]; return DMK_OK; }
  if (w == "F_av") {
    if (n < 9) { dmk_set_error("need 9 doubles"); return DMK_ERR_ARG; }
    double s9[9];
    if ((rc = dmk_sum_records(c, m->F, c->n_local, c->d_scratch + DMK_SCRATCH_SUM9))) return rc;
    if ((rc = reduce_to_host(c, c->d_scratch + DMK_SCRATCH_SUM9, 9, s9))) return rc;
    for (int a = 0; a < 9; a++) out[a] = s9[a] * c->wgt;
    return DMK_OK;
  }
  dmk_set_error("dmk_mech_get: unknown quantity '%s'", w.c_str());
  return DMK_ERR_ARG;
}

static double* field_by_name(DmkMech* m, const std::string& w) {
  if (w == "F") return m->F;
  if (w == "F_lastInc") return m->F_lastInc;
  if (w == "F_dot") return m->F_dot;
  if (w == "r") return m->Fcur;
  if (m->scheme == 1) {
    if (w == "F_tau") return m->F + m->n9;
    if (w == "F_tau_lastInc") return m->F_lastInc + m->n9;
  }
  return nullptr;
}
extern "C" int dmk_mech_get_field(dmk_ctx* c, const char* what, double* host_out) {
  int rc = need_mech(c); if (rc) return rc;
  double* f = field_by_name(c->mech, what ? what : "");
  if (!f) { dmk_set_error("dmk_mech_get_field: unknown field '%s'", what ? what : ""); return DMK_ERR_ARG; }
  DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  DMK_CUDA_CHECK(cudaMemcpy(host_out, f, c->mech->n9 * sizeof(double), cudaMemcpyDeviceToHost));
  return DMK_OK;
}
extern "C" int dmk_mech_set_field(dmk_ctx* c, const char* what, const double* host_in) {
  int rc = need_mech(c); if (rc) return rc;
  double* f = field_by_name(c->mech, what ? what : "");
  if (!f) { dmk_set_error("dmk_mech_set_field: unknown field '%s'", what ? what : ""); return DMK_ERR_ARG; }
  DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  DMK_CUDA_CHECK(cudaMemcpy(f, host_in, c->mech->n9 * sizeof(double), cudaMemcpyHostToDevice));
  return DMK_OK;
}
extern "C" int dmk_mech_device_field(dmk_ctx* c, const char* what, double** d_ptr) {
  int rc = need_mech(c); if (rc) return rc;
  double* f = field_by_name(c->mech, what ? what : "");
  if (!f) { dmk_set_error("dmk_mech_device_field: unknown field '%s'", what ? what : ""); return DMK_ERR_ARG; }
  *d_ptr = f;
  return DMK_OK;
}
extern "C" int dmk_mech_copy_field(dmk_ctx* c, const char* what, double* d_dst) {
  int rc = need_mech(c); if (rc) return rc;
  double* f = field_by_name(c->mech, what ? what : "");
  if (!f) { dmk_set_error("dmk_mech_copy_field: unknown field '%s'", what ? what : ""); return DMK_ERR_ARG; }
  DMK_CUDA_CHECK(cudaMemcpyAsync(d_dst, f, c->mech->n9 * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  return DMK_OK;
}
extern "C" int dmk_mech_capture(dmk_ctx* c, int enable) {
  int rc = need_mech(c); if (rc) return rc;
  c->mech->capture = enable != 0;
  if (!enable) { c->mech->capF.clear(); c->mech->capR.clear(); }
  return DMK_OK;
}
extern "C" int dmk_mech_get_capture(dmk_ctx* c, double* F_out, double* r_out, size_t max_evals, size_t* n_evals) {
  int rc = need_mech(c); if (rc) return rc;
  DmkMech* m = c->mech;
  const size_t n = m->capF.size() / (size_t)m->n9;
  if (n_evals) *n_evals = n;
  const size_t k = std::min(n, max_evals);
  if (F_out) memcpy(F_out, m->capF.data(), k * m->n9 * sizeof(double));
  if (r_out) memcpy(r_out, m->capR.data(), k * m->n9 * sizeof(double));
  return DMK_OK;
}
extern "C" int dmk_mech_eval_log(dmk_ctx* c, double* out, size_t max_rows, size_t* n_rows) {
  int rc = need_mech(c); if (rc) return rc;
  const size_t rows = c->mech->log.size() / 12;
  if (n_rows) *n_rows = rows;
  if (out) memcpy(out, c->mech->log.data(), std::min(rows, max_rows) * 12 * sizeof(double));
  return DMK_OK;
}
extern "C" int dmk_mech_eval_stats(dmk_ctx* c, double* out, size_t max_rows, size_t* n_rows) {
  int rc = need_mech(c); if (rc) return rc;
  const size_t rows = c->mech->stats.size() / 2;
  if (n_rows) *n_rows = rows;
  if (out) memcpy(out, c->mech->stats.data(), std::min(rows, max_rows) * 2 * sizeof(double));
  return DMK_OK;
}
extern "C" int dmk_mech_reports(dmk_ctx* c, double* out, size_t max_rows, size_t* n_rows) {
  int rc = need_mech(c); if (rc) return rc;
  const size_t rows = c->mech->reports.size() / 6;
  if (n_rows) *n_rows = rows;
  if (out) memcpy(out, c->mech->reports.data(), std::min(rows, max_rows) * 6 * sizeof(double));
  return DMK_OK;
}
