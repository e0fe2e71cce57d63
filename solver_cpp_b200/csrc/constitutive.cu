// Constitutive half of the C ABI: per-point state in HBM, the integration kernel (one thread per material point),
// the consistent-tangent kernel with on-the-fly homogenisation, and f_materialpoint_forward.
//
// Replaces f_homogenization_mechanical_response / _response2 / f_materialpoint_forward
// (/root/reference/src/homogenization.f90:219-324, materialpoint.f90:208-218) and the reductions of
// Spectral::constitutive_response (/root/reference/src/grid/spectral.cpp:160-218).
#include <cstring>
#include <cstdlib>
#include <string>
#include <cmath>
#include <vector>
#include <algorithm>
#include "constitutive_dev.cuh"
#include "machine_dev.cuh"
#include "tangent_dev.cuh"
#include "reduce.cuh"

#define DMK_PT_THREADS 128

// ------------------------------------------------------------------------------------------------ kernels
// Lane-persistent integration kernel: every lane owns one material point at a time and walks it through the state
// machine of machine_dev.cuh; one trip of the outer loop = one Lp Newton iteration for all lanes inside an Lp loop.
template <int LAW, int NS, int ND>
__global__ void __launch_bounds__(DMK_PT_THREADS) dmk_constitutive_kernel(DmkPointArrays pa, int phase, DmkNumDev num,
                                                                          const double* __restrict__ Fin, double* __restrict__ Pout,
                                                                          double Dt, const int* __restrict__ list,
                                                                          const unsigned* __restrict__ list_n,
                                                                          const double* __restrict__ list_total, double list_min) {
  // list == nullptr: every cell of `phase`; else the *list_n cells of list[] (given up by the register-resident kernel), and
  // only if the lists of all ranks together hold more than list_min (fewer go through the nested kernel; read on the device)
  const long long n = pa.n;
  const long long total = list ? (long long)*list_n : n;
  if (list && !(*list_total > list_min)) return;
  const long long stride = (long long)gridDim.x * blockDim.x;
  DmkCnt cnt = {0, 0, 0, 0, 0, 0};
  unsigned ill = 0, cb0 = 0, lp0 = 0;
  DmkMachine<LAW, NS, ND> m;
  DmkMachineStore<LAW, NS, ND> store;
  machine_bind(m, store);
  m.state = M_FETCH;
  m.p.ph = pa.d_phases + phase;
  m.p.C66 = m.p.ph->p.C66;
  const int ns = m.p.ph->p.n_state;
  long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x - stride, i = 0;
  bool have = false;
  for (;;) {
    while (m.state != M_LP && m.state != M_EXIT) {
      if (m.state == M_FETCH) {
        if (have) {                                                      // store the finished point
          if (!m.converged) ill = 1;
          pa.cutbacks[i] = (int)(cnt.cutbacks - cb0);
          pa.lp_iter[i] = (int)(cnt.Lp_iter - lp0);
#pragma unroll
          for (int a = 0; a < 9; a++) {
            pa.F[a * n + i] = m.Ft[a]; pa.Fp[a * n + i] = m.p.Fp[a]; pa.Lp[a * n + i] = m.p.Lp[a]; pa.Fe[a * n + i] = m.p.Fe[a];
          }
#pragma unroll
          for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) Pout[i * 9 + a + 3 * b] = m.p.P[3 * a + b];
#pragma unroll
          for (int a = 0; a < 6; a++) pa.S[a * n + i] = m.p.S[a];
          for (int k = 0; k < ns; k++) pa.state[k * n + i] = m.p.st[k];
          if constexpr (ND > 0) { const int nd = m.p.ph->n_dep; for (int k = 0; k < nd; k++) pa.dep[k * n + i] = m.p.dep[k]; }
          have = false;
        }
        do { j += stride; i = (list && j < total) ? (long long)list[j] : j; } while (j < total && pa.phase[i] != phase);
        if (j >= total) { m.state = M_EXIT; break; }
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
          for (int b = 0; b < 3; b++) m.Ft[3 * a + b] = Fin[i * 9 + a + 3 * b];     // boundary record: c = i + 3 j
#pragma unroll
        for (int a = 0; a < 9; a++) {
          m.F0in[a] = pa.F0[a * n + i]; m.Fp0[a] = pa.Fp0[a * n + i]; m.Lp0[a] = pa.Lp0[a * n + i];
          m.p.Lp[a] = pa.Lp[a * n + i];            // phase_restore(includeL = .false.): Lp keeps the last guess
          m.p.Fe[a] = pa.Fe[a * n + i];
        }
#pragma unroll
        for (int a = 0; a < 6; a++) m.S0in[a] = pa.S0[a * n + i];
        for (int k = 0; k < ns; k++) m.state0[k] = pa.state0[k * n + i];
        if constexpr (ND > 0) { const int nd = m.p.ph->n_dep; for (int k = 0; k < nd; k++) m.p.dep[k] = pa.dep[k * n + i]; }
        if constexpr (LAW == DMK_LAW_DISLOTWIN) m.p.C66 = m.p.ph->p.C66;
        cb0 = cnt.cutbacks; lp0 = cnt.Lp_iter;
        have = true;
        machine_begin(m, num);
      } else if (m.state == M_SUBSTEP) machine_substep(m, num, Dt, cnt);
      else if (m.state == M_STRESS_PRO) machine_stress_prologue(m, num, cnt);
      else machine_stress_epilogue(m, num);
    }
    if (!__any_sync(0xffffffffu, m.state != M_EXIT)) break;
    if (m.state == M_LP) machine_lp_iteration(m, num, cnt);
  }
  // counters: warp reduce, one atomic per warp and counter
  unsigned v[7] = {cnt.Lp_iter, cnt.Li_iter, cnt.stress_calls, cnt.state_iter, cnt.substeps, cnt.cutbacks, ill};
#pragma unroll
  for (int k = 0; k < 7; k++) {
    unsigned s = v[k];
    for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(&pa.d_counters[k], (unsigned long long)s);
  }
  if (ill) atomicOr(pa.d_ill, 1);
}

// Straight formulation: one thread per point, nested sub-step / state / Lp loops (lanes idle while their neighbours iterate).
template <int LAW, int NS, int ND>
__global__ void __launch_bounds__(DMK_PT_THREADS) dmk_constitutive_nested_kernel(DmkPointArrays pa, int phase, DmkNumDev num,
                                                                          const double* __restrict__ Fin, double* __restrict__ Pout,
                                                                          double Dt, const int* __restrict__ list,
                                                                          const unsigned* __restrict__ list_n,
                                                                          const double* __restrict__ list_total, double list_max) {
  // list == nullptr: every cell of `phase`; else the *list_n cells of list[] (deferred by the register-resident kernel),
  // unless list_max >= 0 and the lists of all ranks together hold more than list_max (then a lane-scheduled kernel has
  // taken the list, pheno_fast.cu)
  if (list && list_max >= 0.0 && *list_total > list_max) return;
  const long long n = pa.n;
  const long long total = list ? (long long)*list_n : n;
  DmkCnt cnt = {0, 0, 0, 0, 0, 0};
  unsigned ill = 0;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (long long)gridDim.x * blockDim.x) {
    const long long i = list ? (long long)list[j] : j;
    if (pa.phase[i] != phase) continue;
    DmkPt<LAW, NS, ND> p;
    DmkPtStore<LAW, NS, ND> pstore;
    dmk_pt_bind(p, pstore);
    p.ph = pa.d_phases + phase;
    p.C66 = p.ph->p.C66;
    const int ns = p.ph->p.n_state;
    double F0[9], Ft[9], Fp0[9], Lp0[9], S0[6];
    double state0[NS > 0 ? NS : 1];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int b = 0; b < 3; b++) Ft[3 * a + b] = Fin[i * 9 + a + 3 * b];            // boundary record: c = i + 3 j
#pragma unroll
    for (int a = 0; a < 9; a++) {
      F0[a] = pa.F0[a * n + i]; Fp0[a] = pa.Fp0[a * n + i]; Lp0[a] = pa.Lp0[a * n + i];
      p.Lp[a] = pa.Lp[a * n + i];                  // phase_restore(includeL = .false.): Lp keeps the last guess
      p.Fp[a] = Fp0[a]; p.Fe[a] = pa.Fe[a * n + i]; p.P[a] = 0.0;
    }
#pragma unroll
    for (int a = 0; a < 6; a++) { S0[a] = pa.S0[a * n + i]; p.S[a] = S0[a]; }
    for (int k = 0; k < ns; k++) { state0[k] = pa.state0[k * n + i]; p.st[k] = state0[k]; }
    if constexpr (ND > 0) { const int nd = p.ph->n_dep; for (int k = 0; k < nd; k++) p.dep[k] = pa.dep[k * n + i]; }
    const unsigned cb0 = cnt.cutbacks, lp0 = cnt.Lp_iter;
    const bool ok = constitutive_point(p, num, F0, Ft, Fp0, Lp0, S0, state0, Dt, cnt);
    if (!ok) ill = 1;
    pa.cutbacks[i] = (int)(cnt.cutbacks - cb0);
    pa.lp_iter[i] = (int)(cnt.Lp_iter - lp0);
#pragma unroll
    for (int a = 0; a < 9; a++) {
      pa.F[a * n + i] = Ft[a]; pa.Fp[a * n + i] = p.Fp[a]; pa.Lp[a * n + i] = p.Lp[a]; pa.Fe[a * n + i] = p.Fe[a];
    }
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int b = 0; b < 3; b++) Pout[i * 9 + a + 3 * b] = p.P[3 * a + b];
#pragma unroll
    for (int a = 0; a < 6; a++) pa.S[a * n + i] = p.S[a];
    for (int k = 0; k < ns; k++) pa.state[k * n + i] = p.st[k];
    if constexpr (ND > 0) { const int nd = p.ph->n_dep; for (int k = 0; k < nd; k++) pa.dep[k * n + i] = p.dep[k]; }
  }
  // counters: warp reduce, one atomic per warp and counter
  unsigned v[7] = {cnt.Lp_iter, cnt.Li_iter, cnt.stress_calls, cnt.state_iter, cnt.substeps, cnt.cutbacks, ill};
#pragma unroll
  for (int k = 0; k < 7; k++) {
    unsigned s = v[k];
    for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(&pa.d_counters[k], (unsigned long long)s);
  }
  if (ill) atomicOr(pa.d_ill, 1);
}

// tangent: dPdF per point (row-major 3333), summed per block into partials; optional materialisation; ||.||^2 per point
template <int LAW, int NS, int ND>
__global__ void __launch_bounds__(DMK_PT_THREADS) dmk_tangent_kernel(DmkPointArrays pa, int phase, double Dt, long long first,
                                                                     long long count, double* __restrict__ partial,
                                                                     double* __restrict__ dPdF_out, double* __restrict__ single_out) {
  __shared__ double sh[DMK_PT_THREADS / 32];
  const long long n = pa.n;
  double acc[81];
  for (int k = 0; k < 81; k++) acc[k] = 0.0;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += (long long)gridDim.x * blockDim.x) {
    const long long i = first + j;
    if (pa.phase[i] != phase) continue;
    DmkPt<LAW, NS, ND> p;
    DmkPtStore<LAW, NS, ND> pstore;
    dmk_pt_bind(p, pstore);
    p.ph = pa.d_phases + phase;
    p.C66 = p.ph->p.C66;
    const int ns = p.ph->p.n_state;
    double F[9], Fp0[9];
#pragma unroll
    for (int a = 0; a < 9; a++) {
      F[a] = pa.F[a * n + i]; Fp0[a] = pa.Fp0[a * n + i];
      p.Fp[a] = pa.Fp[a * n + i]; p.Fe[a] = pa.Fe[a * n + i]; p.Lp[a] = 0.0; p.P[a] = 0.0;
    }
#pragma unroll
    for (int a = 0; a < 6; a++) p.S[a] = pa.S[a * n + i];
    for (int k = 0; k < ns; k++) p.st[k] = pa.state[k * n + i];
    if constexpr (ND > 0) { const int nd = p.ph->n_dep; for (int k = 0; k < nd; k++) p.dep[k] = pa.dep[k * n + i]; }
    double dPdF[81];
    tangent_point(p, F, Fp0, Dt, dPdF);
    double nrm = 0.0;
    for (int k = 0; k < 81; k++) { acc[k] += dPdF[k]; nrm += dPdF[k] * dPdF[k]; }
    if (pa.nrm && !single_out) pa.nrm[i] = nrm;
    if (single_out) for (int k = 0; k < 81; k++) single_out[k] = dPdF[k];
    if (dPdF_out) {
      // Fortran order of dPdF(3,3,3,3,cell): i fastest
      for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) for (int cidx = 0; cidx < 3; cidx++) for (int d = 0; d < 3; d++)
        dPdF_out[i * 81 + a + 3 * b + 9 * cidx + 27 * d] = dPdF[((a * 3 + b) * 3 + cidx) * 3 + d];
    }
  }
  if (!partial) return;
  for (int k = 0; k < 81; k++) {
    double s = acc[k];
    for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < DMK_PT_THREADS / 32; w++) t += sh[w];
      partial[(size_t)blockIdx.x * 81 + k] += t;          // partials accumulate over the per-phase launches
    }
    __syncthreads();
  }
}

__global__ void dmk_forward_copy_kernel(DmkPointArrays pa) {
  const long long n = pa.n;
  const long long tot9 = 9 * n, tot6 = 6 * n, tots = (long long)pa.max_state * n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tots || i < tot9; i += (long long)gridDim.x * blockDim.x) {
    if (i < tot9) { pa.F0[i] = pa.F[i]; pa.Fp0[i] = pa.Fp[i]; pa.Lp0[i] = pa.Lp[i]; }
    if (i < tot6) pa.S0[i] = pa.S[i];
    if (i < tots) pa.state0[i] = pa.state[i];
  }
}

// ------------------------------------------------------------------------------------------------ launch dispatch
static DmkNumDev num_dev(const dmk_numerics& n) {
  DmkNumDev d;
  d.r_cutback_min = n.r_cutback_min; d.r_cutback = n.r_cutback; d.r_increase = n.r_increase; d.eps_rel_state = n.eps_rel_state;
  d.N_iter_state_max = n.N_iter_state_max; d.N_iter_Lp_max = n.N_iter_Lp_max; d.r_linesearch_Lp = n.r_linesearch_Lp;
  d.eps_rel_Lp = n.eps_rel_Lp; d.eps_abs_Lp = n.eps_abs_Lp; d.f_update_jacobi_Lp = n.f_update_jacobi_Lp > 0 ? n.f_update_jacobi_Lp : 1;
  d.integrator = n.integrator_state;
  return d;
}

static int point_grid(long long n) {
  long long b = (n + DMK_PT_THREADS - 1) / DMK_PT_THREADS;
  const long long cap = 148LL * 64;
  return (int)std::max(1LL, std::min(b, cap));
}
// persistent launch: one resident wave (SM count x blocks per SM), fewer if there are fewer points
template <typename K>
static int persistent_grid(K kernel, long long n) {
  int dev = 0, sms = 148, occ = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, DMK_PT_THREADS, 0) != cudaSuccess || occ < 1) occ = 1;
  const long long b = (n + DMK_PT_THREADS - 1) / DMK_PT_THREADS;
  return (int)std::max(1LL, std::min(b, (long long)sms * occ));
}
#define DMK_LAUNCH_PERSISTENT(KERNEL, law, nstate, n, st, ...)                                                                      \
  do {                                                                                                                             \
    if (law == DMK_LAW_NONE) KERNEL<DMK_LAW_NONE, 0, 0><<<persistent_grid(KERNEL<DMK_LAW_NONE, 0, 0>, n), DMK_PT_THREADS, 0, st>>> __VA_ARGS__; \
    else if (law == DMK_LAW_ISOTROPIC) KERNEL<DMK_LAW_ISOTROPIC, 1, 0><<<persistent_grid(KERNEL<DMK_LAW_ISOTROPIC, 1, 0>, n), DMK_PT_THREADS, 0, st>>> __VA_ARGS__; \
    else if (law == DMK_LAW_PHENOPOWERLAW && nstate <= 24) KERNEL<DMK_LAW_PHENOPOWERLAW, 24, 0><<<persistent_grid(KERNEL<DMK_LAW_PHENOPOWERLAW, 24, 0>, n), DMK_PT_THREADS, 0, st>>> __VA_ARGS__; \
    else if (law == DMK_LAW_PHENOPOWERLAW && nstate <= 48) KERNEL<DMK_LAW_PHENOPOWERLAW, 48, 0><<<persistent_grid(KERNEL<DMK_LAW_PHENOPOWERLAW, 48, 0>, n), DMK_PT_THREADS, 0, st>>> __VA_ARGS__; \
    else if (law == DMK_LAW_PHENOPOWERLAW) KERNEL<DMK_LAW_PHENOPOWERLAW, 2 * DMK_MAX_SYS, 0><<<persistent_grid(KERNEL<DMK_LAW_PHENOPOWERLAW, 2 * DMK_MAX_SYS, 0>, n), DMK_PT_THREADS, 0, st>>> __VA_ARGS__; \
    else if (law == DMK_LAW_KINEHARDENING) KERNEL<DMK_LAW_KINEHARDENING, DMK_MAX_STATE, 0><<<persistent_grid(KERNEL<DMK_LAW_KINEHARDENING, DMK_MAX_STATE, 0>, n), DMK_PT_THREADS, 0, st>>> __VA_ARGS__; \
    else if (law == DMK_LAW_DISLOTUNGSTEN) KERNEL<DMK_LAW_DISLOTUNGSTEN, 3 * DMK_MAX_SL, 2 * DMK_MAX_SL><<<persistent_grid(KERNEL<DMK_LAW_DISLOTUNGSTEN, 3 * DMK_MAX_SL, 2 * DMK_MAX_SL>, n), DMK_PT_THREADS, 0, st>>> __VA_ARGS__; \
    else if (law == DMK_LAW_DISLOTWIN && nstate <= 48) KERNEL<DMK_LAW_DISLOTWIN, 48, 36><<<persistent_grid(KERNEL<DMK_LAW_DISLOTWIN, 48, 36>, n), DMK_PT_THREADS, 0, st>>> __VA_ARGS__; \
    else KERNEL<DMK_LAW_DISLOTWIN, DMK_MAX_STATE, DMK_MAX_DEP><<<persistent_grid(KERNEL<DMK_LAW_DISLOTWIN, DMK_MAX_STATE, DMK_MAX_DEP>, n), DMK_PT_THREADS, 0, st>>> __VA_ARGS__; \
  } while (0)

#define DMK_DISPATCH(KERNEL, law, nstate, ...)                                                                  \
  do {                                                                                                          \
    if (law == DMK_LAW_NONE) KERNEL<DMK_LAW_NONE, 0, 0> __VA_ARGS__;                                            \
    else if (law == DMK_LAW_ISOTROPIC) KERNEL<DMK_LAW_ISOTROPIC, 1, 0> __VA_ARGS__;                             \
    else if (law == DMK_LAW_PHENOPOWERLAW && nstate <= 24) KERNEL<DMK_LAW_PHENOPOWERLAW, 24, 0> __VA_ARGS__;    \
    else if (law == DMK_LAW_PHENOPOWERLAW && nstate <= 48) KERNEL<DMK_LAW_PHENOPOWERLAW, 48, 0> __VA_ARGS__;    \
    else if (law == DMK_LAW_PHENOPOWERLAW) KERNEL<DMK_LAW_PHENOPOWERLAW, 2 * DMK_MAX_SYS, 0> __VA_ARGS__;       \
    else if (law == DMK_LAW_KINEHARDENING) KERNEL<DMK_LAW_KINEHARDENING, DMK_MAX_STATE, 0> __VA_ARGS__;         \
    else if (law == DMK_LAW_DISLOTUNGSTEN) KERNEL<DMK_LAW_DISLOTUNGSTEN, 3 * DMK_MAX_SL, 2 * DMK_MAX_SL> __VA_ARGS__; \
    else if (law == DMK_LAW_DISLOTWIN && nstate <= 48) KERNEL<DMK_LAW_DISLOTWIN, 48, 36> __VA_ARGS__;           \
    else KERNEL<DMK_LAW_DISLOTWIN, DMK_MAX_STATE, DMK_MAX_DEP> __VA_ARGS__;                                     \
  } while (0)

// device counter of the cells the register-resident kernel of the last evaluation gave up (their full first sub-step failed:
// what the reference answers with a cut-back); 0 for phases that run in the reference-order kernel
const unsigned* dmk_points_given_up_counter(dmk_ctx* c) { return c->pts->d_defer_count; }
bool dmk_points_single_constituent(const dmk_ctx* c) { return !c->pts || c->pts->n_const <= 1; }
const double* dmk_points_given_up_total(dmk_ctx* c) { return c->pts->d_defer_total; }

// DMK_PHENO_MASS: which kernel takes the deferred list when it is long (see cutback_setting() in pheno_fast.cu):
//   lanes (default)    the register-resident 6x6 formulation with in-lane cut-backs (pheno_lanes_dev.cuh, CUTBACK = true)
//   nested             no special treatment: the one-thread-per-point reference-order kernel whatever the length of the list
//   machine            the reference-order arithmetic on the lane-persistent state machine (machine_dev.cuh)
// Measured on a virgin-state evaluation of the 128^3 Cu polycrystal, where every cell cuts back (profiles/r02_summary.md):
// lanes 227 ms, nested 1138 ms, machine 5173 ms (the state machine's per-thread arrays live in local memory).  Such evaluations
// are rounding-chaotic for ANY build of the reference algorithm -- the homogenised stress of that one evaluation differs by
// ~1e-3 between any two of the three kernels, and between the oracle and its own FMA build (tests/test_oracle_golden.py) --
// so the choice is made on speed; dmk_set_option("pheno_mass_deferral") / DMK_PHENO_MASS select the others.
int dmk_mass_deferral_setting(const dmk_ctx* c) {
  static const int env = [] { const char* e = getenv("DMK_PHENO_MASS"); return !e ? 2 : std::string(e) == "machine" ? 1 : std::string(e) == "nested" ? 0 : 2; }();
  return (c && c->opt_mass >= 0) ? c->opt_mass : env;
}

int dmk_wait_staged_input(dmk_ctx* c, const double* d_F) {
  if (c->stage_chunks > 1 && d_F == c->stage_dst) DMK_CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->stage_ev[c->stage_chunks - 1], 0));
  return DMK_OK;
}

// ---- isostrain homogenisation (N_constituents > 1): homogenization_mechanical_isostrain.f90:36-45 (every constituent sees the
// cell's F) and mechanical_homogenize (homogenization_mechanical.f90:125-145: P = sum_co v_co P_co, dPdF alike)
__global__ void dmk_isostrain_partition_kernel(const double* __restrict__ F, double* __restrict__ Fpt, long long n_cells9, int n_const) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_cells9; i += (long long)gridDim.x * blockDim.x) {
    const double f = F[i];
    for (int k = 0; k < n_const; k++) Fpt[k * n_cells9 + i] = f;
  }
}
// out[cell][w] = sum_k v[k n_cells + cell] * in[k n_cells + cell][w]   (w < width: 9 for P, 81 for dPdF), constituents in order
__global__ void dmk_isostrain_average_kernel(const double* __restrict__ in, const double* __restrict__ v, double* __restrict__ out,
                                             long long n_cells, int n_const, int width) {
  const long long total = n_cells * width;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long cell = i / width;
    // the reference's sequence, each product and each sum rounded on its own (no contraction into FMAs)
    double s = __dmul_rn(in[i], v[cell]);
    for (int k = 1; k < n_const; k++) s = __dadd_rn(s, __dmul_rn(in[k * total + i], v[k * n_cells + cell]));
    out[i] = s;
  }
}

static int points_constitutive_core(dmk_ctx* c, const double* d_F, double* d_P, double dt);
int dmk_points_constitutive(dmk_ctx* c, const double* d_F, double* d_P, double dt) {
  DmkPointArrays& pa = *c->pts;
  if (pa.n_const <= 1) return points_constitutive_core(c, d_F, d_P, dt);
  int rc = dmk_wait_staged_input(c, d_F); if (rc) return rc;         // the whole field must have arrived before it is partitioned
  dmk_isostrain_partition_kernel<<<148 * 8, 256, 0, c->stream>>>(d_F, pa.Fpt, pa.n_cells * 9, pa.n_const);
  c->launches++;
  if ((rc = points_constitutive_core(c, pa.Fpt, pa.Ppt, dt))) return rc;
  dmk_isostrain_average_kernel<<<148 * 8, 256, 0, c->stream>>>(pa.Ppt, pa.v, d_P, pa.n_cells, pa.n_const, 9);
  c->launches++;
  DMK_CUDA_CHECK(cudaGetLastError());
  return DMK_OK;
}
static int points_constitutive_core(dmk_ctx* c, const double* d_F, double* d_P, double dt) {
  DmkPointArrays& pa = *c->pts;
  const DmkNumDev nd = num_dev(c->num);
  // two formulations of the same arithmetic (bit-identical, tests/host_twin): lane-persistent state machine | nested loops
  static const bool want_machine = [] { const char* e = getenv("DMK_CONSTITUTIVE"); return e && std::string(e) == "machine"; }();
  const bool use_machine = want_machine && nd.integrator == DMK_INTEG_FPI;     // the lane machine walks the FPI loop only
  static const bool use_generic = [] { const char* e = getenv("DMK_CONSTITUTIVE"); return e && (std::string(e) == "generic" || std::string(e) == "machine"); }();
  for (int ph = 0; ph < pa.n_phases; ph++) {
    if (pa.n_cells_of_phase[ph] == 0) continue;
    const int law = pa.law_of_phase[ph], nst = pa.nstate_of_phase[ph];
    bool handled = false, deferred = false;
    double defer_max = -1.0;
    if (!use_generic && nd.integrator == DMK_INTEG_FPI) { int rc = dmk_points_pheno_fast(c, ph, nd, d_F, d_P, dt, &handled, &deferred, &defer_max); if (rc) return rc; }
    if (handled) {
      c->launches++;
      if (deferred) {          // second pass over the deferred cells only (count read on the device: no host round trip)
        if (defer_max >= 0.0 && dmk_mass_deferral_setting(c) == 1 && nd.integrator == DMK_INTEG_FPI) {
          // mass deferral (more than defer_max cells): the reference-order arithmetic on the lane-persistent state machine
          DMK_LAUNCH_PERSISTENT(dmk_constitutive_kernel, law, nst, pa.n, c->stream, (pa, ph, nd, d_F, d_P, dt, pa.defer_list, pa.d_defer_count, pa.d_defer_total, defer_max));
          c->launches++;
        }
        DMK_DISPATCH(dmk_constitutive_nested_kernel, law, nst, <<<148 * 4, DMK_PT_THREADS, 0, c->stream>>>(pa, ph, nd, d_F, d_P, dt, pa.defer_list, pa.d_defer_count, pa.d_defer_total, defer_max));
        c->launches++;
      }
      continue;
    }
    { int rc = dmk_wait_staged_input(c, d_F); if (rc) return rc; }
    if (use_machine) DMK_LAUNCH_PERSISTENT(dmk_constitutive_kernel, law, nst, pa.n, c->stream, (pa, ph, nd, d_F, d_P, dt, nullptr, nullptr, nullptr, 0.0));
    else DMK_DISPATCH(dmk_constitutive_nested_kernel, law, nst, <<<point_grid(pa.n), DMK_PT_THREADS, 0, c->stream>>>(pa, ph, nd, d_F, d_P, dt, nullptr, nullptr, nullptr, -1.0));
    c->launches++;
  }
  DMK_CUDA_CHECK(cudaGetLastError());
  return DMK_OK;
}

#define DMK_TANGENT_BLOCKS (148 * 8)

int dmk_points_tangent(dmk_ctx* c, double dt, double* d_Csum81, double* d_dPdF) {
  DmkPointArrays& pa = *c->pts;
  double* partial = c->d_scratch + DMK_SCRATCH_PARTIAL;            // DMK_TANGENT_BLOCKS * 81 doubles
  const int grid = (int)std::max(1LL, std::min((pa.n + DMK_PT_THREADS - 1) / DMK_PT_THREADS, (long long)DMK_TANGENT_BLOCKS));
  DMK_CUDA_CHECK(cudaMemsetAsync(partial, 0, (size_t)grid * 81 * sizeof(double), c->stream));
  for (int ph = 0; ph < pa.n_phases; ph++) {
    if (pa.n_cells_of_phase[ph] == 0) continue;
    const int law = pa.law_of_phase[ph], nst = pa.nstate_of_phase[ph];
    DMK_DISPATCH(dmk_tangent_kernel, law, nst, <<<grid, DMK_PT_THREADS, 0, c->stream>>>(pa, ph, dt, 0LL, pa.n, partial, d_dPdF, nullptr));
    c->launches++;
  }
  dmk_sum_partials_kernel<<<1, 96, 0, c->stream>>>(partial, grid, 81, d_Csum81);
  c->launches++;
  DMK_CUDA_CHECK(cudaGetLastError());
  return DMK_OK;
}

int dmk_points_tangent_cell(dmk_ctx* c, double dt, long long cell, double* d_out81) {
  DmkPointArrays& pa = *c->pts;
  int ph = 0;
  DMK_CUDA_CHECK(cudaMemcpyAsync(&ph, pa.phase + cell, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  const int law = pa.law_of_phase[ph], nst = pa.nstate_of_phase[ph];
  DMK_DISPATCH(dmk_tangent_kernel, law, nst, <<<1, DMK_PT_THREADS, 0, c->stream>>>(pa, ph, dt, cell, 1LL, nullptr, nullptr, d_out81));
  c->launches++;
  DMK_CUDA_CHECK(cudaGetLastError());
  return DMK_OK;
}

// ------------------------------------------------------------------------------------------------ set-up
static void voigt_strain_of(const double* P, double* pv) {
  pv[0] = P[0]; pv[1] = P[4]; pv[2] = P[8]; pv[3] = P[5] + P[7]; pv[4] = P[2] + P[6]; pv[5] = P[1] + P[3];
}
static void q_host(const double* C66, const double* pv, double* Q) {
  for (int b = 0; b < 6; b++) { double s = 0.0; for (int a = 0; a < 6; a++) s += pv[a] * C66[6 * a + b]; Q[b] = s; }
}

static int check_phase(const dmk_phase& p, int idx) {
  if (p.law < DMK_LAW_NONE || p.law > DMK_LAW_DISLOTUNGSTEN) { dmk_set_error("phase %d: unknown plastic law %d", idx, p.law); return DMK_ERR_ARG; }
  if (p.n_sl < 0 || p.n_sl > DMK_MAX_SL || p.n_tw < 0 || p.n_tw > DMK_MAX_TW) { dmk_set_error("phase %d: too many slip/twin systems", idx); return DMK_ERR_UNSUPPORTED; }
  if (p.law == DMK_LAW_ISOTROPIC && p.iso_dilatation) { dmk_set_error("phase %d: isotropic plasticity with dilatation (Li != 0) is not offered on the device path", idx); return DMK_ERR_UNSUPPORTED; }
  if (p.law == DMK_LAW_DISLOTWIN && p.dt_gamma_0_sb != 0.0) { dmk_set_error("phase %d: dislotwin shear banding is not offered on the device path", idx); return DMK_ERR_UNSUPPORTED; }
  int expect = 0;
  if (p.law == DMK_LAW_ISOTROPIC) expect = 1;
  else if (p.law == DMK_LAW_PHENOPOWERLAW) expect = 2 * (p.n_sl + p.n_tw);
  else if (p.law == DMK_LAW_DISLOTWIN) expect = 3 * p.n_sl + p.n_tw;
  else if (p.law == DMK_LAW_DISLOTUNGSTEN) {
    if (p.n_sl < 1 || p.n_tw != 0) { dmk_set_error("phase %d: dislotungsten needs slip systems and no twin systems", idx); return DMK_ERR_UNSUPPORTED; }
    expect = 3 * p.n_sl;
  }
  else if (p.law == DMK_LAW_KINEHARDENING) {
    if (p.n_sl < 1 || 6 * p.n_sl > DMK_MAX_STATE || p.n_tw != 0) { dmk_set_error("phase %d: kinehardening needs 1..%d slip systems and no twin systems", idx, DMK_MAX_STATE / 6); return DMK_ERR_UNSUPPORTED; }
    expect = 6 * p.n_sl;
  }
  if (p.n_state != expect) { dmk_set_error("phase %d: n_state %d, expected %d", idx, p.n_state, expect); return DMK_ERR_ARG; }
  return DMK_OK;
}

void dmk_points_teardown(dmk_ctx* c) {
  if (!c->pts) return;
  DmkPointArrays* pa = c->pts;
  cudaFree(pa->phase); cudaFree(pa->F0); cudaFree(pa->F); cudaFree(pa->Fp0); cudaFree(pa->Fp); cudaFree(pa->Fe);
  cudaFree(pa->Lp0); cudaFree(pa->Lp); cudaFree(pa->S0); cudaFree(pa->S); cudaFree(pa->state0); cudaFree(pa->state);
  if (pa->dep) cudaFree(pa->dep);
  if (pa->v) cudaFree(pa->v);
  if (pa->Fpt) cudaFree(pa->Fpt);
  if (pa->Ppt) cudaFree(pa->Ppt);
  cudaFree(pa->nrm); cudaFree(pa->cutbacks); cudaFree(pa->lp_iter); cudaFree(pa->defer_list); cudaFree(pa->d_defer_count); cudaFree(pa->d_work); cudaFree(pa->d_defer_total); cudaFree(pa->d_phases); cudaFree(pa->d_counters); cudaFree(pa->d_ill);
  delete[] pa->h_phases;
  delete[] pa->h_dev_phases;
  delete pa;
  c->pts = nullptr;
}

template <int LAW, int NS, int ND>
__global__ void dmk_init_dependent_kernel(DmkPointArrays pa, int phase) {
  const long long n = pa.n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (pa.phase[i] != phase) continue;
    if constexpr (LAW == DMK_LAW_DISLOTWIN || LAW == DMK_LAW_DISLOTUNGSTEN) {
      DmkPt<LAW, NS, ND> p;
      DmkPtStore<LAW, NS, ND> pstore;
      dmk_pt_bind(p, pstore);
      p.ph = pa.d_phases + phase;
      const int ns = p.ph->p.n_state, nd = p.ph->n_dep;
      for (int k = 0; k < ns; k++) p.st[k] = pa.state[k * n + i];
      if constexpr (LAW == DMK_LAW_DISLOTWIN) dt_dependentState(p); else tu_dependentState(p);
      for (int k = 0; k < nd; k++) pa.dep[k * n + i] = p.dep[k];
    }
  }
}

static int material_init_impl(dmk_ctx* c, const dmk_phase* phases, int n_phases, int n_const, const int* phase_of_cell,
                              const double* O_of_cell, const double* v_of_point, const dmk_numerics* num);
extern "C" int dmk_material_init(dmk_ctx* c, const dmk_phase* phases, int n_phases, const int* phase_of_cell,
                                 const double* O_of_cell, const dmk_numerics* num) {
  return material_init_impl(c, phases, n_phases, 1, phase_of_cell, O_of_cell, nullptr, num);
}
extern "C" int dmk_material_init_constituents(dmk_ctx* c, const dmk_phase* phases, int n_phases, int n_constituents,
                                              const int* phase_of_point, const double* O_of_point, const double* v_of_point,
                                              const dmk_numerics* num) {
  if (n_constituents < 1 || n_constituents > 16) { dmk_set_error("1..16 constituents per cell supported"); return DMK_ERR_UNSUPPORTED; }
  if (n_constituents > 1 && !v_of_point) { dmk_set_error("null argument"); return DMK_ERR_ARG; }
  return material_init_impl(c, phases, n_phases, n_constituents, phase_of_point, O_of_point, v_of_point, num);
}
extern "C" int dmk_n_constituents(dmk_ctx* c) { return (c && c->pts) ? c->pts->n_const : 0; }
static int material_init_impl(dmk_ctx* c, const dmk_phase* phases, int n_phases, int n_const, const int* phase_of_cell,
                              const double* O_of_cell, const double* v_of_point, const dmk_numerics* num) {
  if (!c || !phases || !phase_of_cell || !O_of_cell || !num) { dmk_set_error("null argument"); return DMK_ERR_ARG; }
  if (n_phases < 1 || n_phases > DMK_MAX_PHASES) { dmk_set_error("1..%d phases supported", DMK_MAX_PHASES); return DMK_ERR_UNSUPPORTED; }
  for (int i = 0; i < n_phases; i++) { int rc = check_phase(phases[i], i); if (rc) return rc; }
  dmk_points_teardown(c);
  c->num = *num;
  DmkPointArrays* pa = new DmkPointArrays();
  memset(pa, 0, sizeof *pa);
  c->pts = pa;
  const long long n = c->n_local * n_const;           // material points of this slab
  pa->n = n; pa->n_phases = n_phases;
  pa->n_const = n_const; pa->n_cells = c->n_local;
  pa->h_phases = new dmk_phase[n_phases];
  memcpy(pa->h_phases, phases, sizeof(dmk_phase) * n_phases);
  std::vector<DmkDevPhase> dev(n_phases);
  int max_state = 1, max_dep = 0;
  for (int i = 0; i < n_phases; i++) {
    DmkDevPhase& d = dev[i];
    memset(&d, 0, sizeof d);
    d.p = phases[i];
    const dmk_phase& p = phases[i];
    bool nons = false;
    for (int a = 0; a < p.n_sl; a++) {
      const double* pos = (p.law == DMK_LAW_DISLOTWIN) ? p.P_sl[a] : p.P_nS_pos[a];
      voigt_strain_of(pos, d.pv_pos[a]); q_host(p.C66, d.pv_pos[a], d.Q_pos[a]);
      voigt_strain_of(p.P_nS_neg[a], d.pv_neg[a]); q_host(p.C66, d.pv_neg[a], d.Q_neg[a]);
      for (int k = 0; k < 9; k++) if (p.P_nS_neg[a][k] != -p.P_nS_pos[a][k]) nons = true;
    }
    for (int a = 0; a < p.n_tw; a++) { voigt_strain_of(p.P_tw[a], d.pv_tw[a]); q_host(p.C66, d.pv_tw[a], d.Q_tw[a]); }
    d.nonschmid = nons ? 1 : 0;
    auto small_int = [](double v) { return (v >= 1.0 && v <= 64.0 && v == std::floor(v)) ? (int)v : 0; };
    for (int a = 0; a < p.n_sl; a++) d.n_sl_int[a] = small_int(p.n_sl_exp[a]);
    for (int a = 0; a < p.n_tw; a++) d.n_tw_int[a] = small_int(p.n_tw_exp[a]);
    d.iso_n_int = small_int(p.iso_n);
    d.n_dep = (p.law == DMK_LAW_DISLOTWIN) ? 2 * p.n_sl + p.n_tw : (p.law == DMK_LAW_DISLOTUNGSTEN) ? 2 * p.n_sl : 0;
    pa->law_of_phase[i] = p.law; pa->nstate_of_phase[i] = p.n_state; pa->nsys_of_phase[i] = p.n_sl + p.n_tw;
    max_state = std::max(max_state, p.n_state); max_dep = std::max(max_dep, d.n_dep);
  }
  pa->max_state = max_state; pa->max_dep = max_dep;
  // host staging of the initial state (mechanical_init: F = Fe.Fp with Fp0 = O, Fe = O^T; phase_mechanical.f90:248-259)
  std::vector<int> ph(n);
  for (long long i = 0; i < n; i++) {
    ph[i] = phase_of_cell[i];
    if (ph[i] < 0 || ph[i] >= n_phases) { dmk_set_error("cell %lld: phase index %d out of range", i, ph[i]); return DMK_ERR_ARG; }
    pa->n_cells_of_phase[ph[i]]++;
  }
  auto dalloc = [&](double** p, size_t count) -> int { DMK_CUDA_CHECK(cudaMalloc((void**)p, count * sizeof(double))); return DMK_OK; };
  int rc;
  if ((rc = dalloc(&pa->F0, 9 * n)) || (rc = dalloc(&pa->F, 9 * n)) || (rc = dalloc(&pa->Fp0, 9 * n)) || (rc = dalloc(&pa->Fp, 9 * n)) ||
      (rc = dalloc(&pa->Fe, 9 * n)) || (rc = dalloc(&pa->Lp0, 9 * n)) || (rc = dalloc(&pa->Lp, 9 * n)) || (rc = dalloc(&pa->S0, 6 * n)) ||
      (rc = dalloc(&pa->S, 6 * n)) || (rc = dalloc(&pa->state0, (size_t)max_state * n)) || (rc = dalloc(&pa->state, (size_t)max_state * n)) ||
      (rc = dalloc(&pa->nrm, n))) return rc;
  if (max_dep > 0) { if ((rc = dalloc(&pa->dep, (size_t)max_dep * n))) return rc; DMK_CUDA_CHECK(cudaMemset(pa->dep, 0, (size_t)max_dep * n * sizeof(double))); }
  if (n_const > 1) {
    for (long long cell = 0; cell < c->n_local; cell++) {
      double sv = 0.0;
      for (int k = 0; k < n_const; k++) { const double vk = v_of_point[k * c->n_local + cell]; if (!(vk > 0.0)) { dmk_set_error("cell %lld: volume fraction %g of constituent %d", cell, vk, k); return DMK_ERR_ARG; } sv += vk; }
      if (fabs(sv - 1.0) > 1.0e-9) { dmk_set_error("cell %lld: volume fractions sum to %.12g", cell, sv); return DMK_ERR_ARG; }   // material.f90: sum(v) must be 1
    }
    if ((rc = dalloc(&pa->v, n)) || (rc = dalloc(&pa->Fpt, 9 * n)) || (rc = dalloc(&pa->Ppt, 9 * n))) return rc;
    DMK_CUDA_CHECK(cudaMemcpy(pa->v, v_of_point, n * sizeof(double), cudaMemcpyHostToDevice));
    DMK_CUDA_CHECK(cudaMemset(pa->Ppt, 0, 9 * n * sizeof(double)));
  }
  DMK_CUDA_CHECK(cudaMalloc((void**)&pa->phase, n * sizeof(int)));
  DMK_CUDA_CHECK(cudaMalloc((void**)&pa->cutbacks, n * sizeof(int)));
  DMK_CUDA_CHECK(cudaMemset(pa->cutbacks, 0, n * sizeof(int)));
  DMK_CUDA_CHECK(cudaMalloc((void**)&pa->lp_iter, n * sizeof(int)));
  DMK_CUDA_CHECK(cudaMemset(pa->lp_iter, 0, n * sizeof(int)));
  DMK_CUDA_CHECK(cudaMalloc((void**)&pa->defer_list, n * sizeof(int)));
  DMK_CUDA_CHECK(cudaMalloc((void**)&pa->d_defer_count, sizeof(unsigned)));
  DMK_CUDA_CHECK(cudaMemset(pa->d_defer_count, 0, sizeof(unsigned)));
  DMK_CUDA_CHECK(cudaMalloc((void**)&pa->d_work, 8 * sizeof(unsigned long long)));       // one queue per z-chunk launch
  DMK_CUDA_CHECK(cudaMalloc((void**)&pa->d_defer_total, sizeof(double)));
  DMK_CUDA_CHECK(cudaMemset(pa->d_defer_total, 0, sizeof(double)));
  DMK_CUDA_CHECK(cudaMalloc((void**)&pa->d_phases, n_phases * sizeof(DmkDevPhase)));
  DMK_CUDA_CHECK(cudaMalloc((void**)&pa->d_counters, 16 * sizeof(unsigned long long)));
  DMK_CUDA_CHECK(cudaMalloc((void**)&pa->d_ill, sizeof(int)));
  DMK_CUDA_CHECK(cudaMemset(pa->d_counters, 0, 16 * sizeof(unsigned long long)));
  DMK_CUDA_CHECK(cudaMemset(pa->d_ill, 0, sizeof(int)));
  DMK_CUDA_CHECK(cudaMemcpy(pa->phase, ph.data(), n * sizeof(int), cudaMemcpyHostToDevice));
  DMK_CUDA_CHECK(cudaMemcpy(pa->d_phases, dev.data(), n_phases * sizeof(DmkDevPhase), cudaMemcpyHostToDevice));
  pa->h_dev_phases = new DmkDevPhase[n_phases];
  memcpy(pa->h_dev_phases, dev.data(), n_phases * sizeof(DmkDevPhase));
  {
    std::vector<double> buf((size_t)9 * n);
    auto up = [&](double* d, size_t count) -> int { DMK_CUDA_CHECK(cudaMemcpy(d, buf.data(), count * sizeof(double), cudaMemcpyHostToDevice)); return DMK_OK; };
    for (int a = 0; a < 9; a++) for (long long i = 0; i < n; i++) buf[(size_t)a * n + i] = (a % 4 == 0) ? 1.0 : 0.0;      // F = I
    if ((rc = up(pa->F0, 9 * n)) || (rc = up(pa->F, 9 * n))) return rc;
    for (int a = 0; a < 9; a++) for (long long i = 0; i < n; i++) buf[(size_t)a * n + i] = O_of_cell[i * 9 + a];           // Fp = O
    if ((rc = up(pa->Fp0, 9 * n)) || (rc = up(pa->Fp, 9 * n))) return rc;
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) for (long long i = 0; i < n; i++) buf[(size_t)(3 * a + b) * n + i] = O_of_cell[i * 9 + 3 * b + a];
    if ((rc = up(pa->Fe, 9 * n))) return rc;                                                                              // Fe = O^T
    DMK_CUDA_CHECK(cudaMemset(pa->Lp0, 0, 9 * n * sizeof(double))); DMK_CUDA_CHECK(cudaMemset(pa->Lp, 0, 9 * n * sizeof(double)));
    DMK_CUDA_CHECK(cudaMemset(pa->S0, 0, 6 * n * sizeof(double))); DMK_CUDA_CHECK(cudaMemset(pa->S, 0, 6 * n * sizeof(double)));
    std::vector<double> st((size_t)max_state * n, 0.0);
    for (long long i = 0; i < n; i++) { const dmk_phase& p = phases[ph[i]]; for (int k = 0; k < p.n_state; k++) st[(size_t)k * n + i] = p.state_init[k]; }
    DMK_CUDA_CHECK(cudaMemcpy(pa->state0, st.data(), st.size() * sizeof(double), cudaMemcpyHostToDevice));
    DMK_CUDA_CHECK(cudaMemcpy(pa->state, st.data(), st.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  // crystallite_init: dependent state from the initial state (phase.f90:564)
  for (int i = 0; i < n_phases; i++) {
    if ((pa->law_of_phase[i] != DMK_LAW_DISLOTWIN && pa->law_of_phase[i] != DMK_LAW_DISLOTUNGSTEN) || pa->n_cells_of_phase[i] == 0) continue;
    const int grid = point_grid(n);
    if (pa->law_of_phase[i] == DMK_LAW_DISLOTUNGSTEN) dmk_init_dependent_kernel<DMK_LAW_DISLOTUNGSTEN, 3 * DMK_MAX_SL, 2 * DMK_MAX_SL><<<grid, DMK_PT_THREADS, 0, c->stream>>>(*pa, i);
    else if (pa->nstate_of_phase[i] <= 48) dmk_init_dependent_kernel<DMK_LAW_DISLOTWIN, 48, 36><<<grid, DMK_PT_THREADS, 0, c->stream>>>(*pa, i);
    else dmk_init_dependent_kernel<DMK_LAW_DISLOTWIN, DMK_MAX_STATE, DMK_MAX_DEP><<<grid, DMK_PT_THREADS, 0, c->stream>>>(*pa, i);
    c->launches++;
  }
  DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  DMK_CUDA_CHECK(cudaGetLastError());
  return DMK_OK;
}

// ------------------------------------------------------------------------------------------------ C ABI
static int need_points(dmk_ctx* c) {
  if (!c || !c->pts) { dmk_set_error("material not initialised (dmk_material_init)"); return DMK_ERR_STATE; }
  return DMK_OK;
}

extern "C" int dmk_constitutive_response(dmk_ctx* c, const double* d_F, double* d_P, double dt) {
  int rc = need_points(c); if (rc) return rc;
  return dmk_points_constitutive(c, d_F, d_P, dt);
}

extern "C" int dmk_materialpoint_forward(dmk_ctx* c) {
  int rc = need_points(c); if (rc) return rc;
  if (c->mech && (rc = dmk_mech_before_point_forward(c))) return rc;     // dPdF must see the pre-forward *0 state
  dmk_forward_copy_kernel<<<148 * 8, 256, 0, c->stream>>>(*c->pts);
  c->launches++;
  DMK_CUDA_CHECK(cudaGetLastError());
  return DMK_OK;
}

extern "C" int dmk_terminally_ill(dmk_ctx* c, int* flag, int reset) {
  int rc = need_points(c); if (rc) return rc;
  int h = 0;
  DMK_CUDA_CHECK(cudaMemcpyAsync(&h, c->pts->d_ill, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  if (reset) DMK_CUDA_CHECK(cudaMemsetAsync(c->pts->d_ill, 0, sizeof(int), c->stream));
  DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  if (c->P > 1 && c->has_comm && c->comm.allreduce_max) {
    double* d = c->d_scratch + DMK_SCRATCH_TMP;
    double v = (double)h;
    DMK_CUDA_CHECK(cudaMemcpyAsync(d, &v, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (c->comm.allreduce_max(c->comm.user, d, 1)) { dmk_set_error("allreduce failed"); return DMK_ERR_STATE; }
    DMK_CUDA_CHECK(cudaMemcpyAsync(&v, d, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    h = v != 0.0;
  }
  if (flag) *flag = h;
  return DMK_OK;
}

extern "C" int dmk_get_counters(dmk_ctx* c, long long counters[7]) {
  int rc = need_points(c); if (rc) return rc;
  unsigned long long h[8];
  DMK_CUDA_CHECK(cudaMemcpyAsync(h, c->pts->d_counters, sizeof h, cudaMemcpyDeviceToHost, c->stream));
  DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  counters[0] = (long long)h[0]; counters[1] = (long long)h[2]; counters[2] = (long long)h[3];
  counters[3] = (long long)h[4]; counters[4] = (long long)h[5]; counters[5] = (long long)h[6]; counters[6] = (long long)h[7];
  return DMK_OK;
}

// lane-schedule statistics of the lane-persistent kernel (pheno_fast.cu), cumulative: warp trips and lanes served of the
// Newton block, the state block and the refill; Newton steps solved (lanes that went past the residual)
extern "C" int dmk_get_schedule_stats(dmk_ctx* c, long long stats[8]) {
  int rc = need_points(c); if (rc) return rc;
  unsigned long long h[16];
  DMK_CUDA_CHECK(cudaMemcpyAsync(h, c->pts->d_counters, sizeof h, cudaMemcpyDeviceToHost, c->stream));
  DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  for (int k = 0; k < 8; k++) stats[k] = (long long)h[8 + k];
  return DMK_OK;
}

extern "C" int dmk_max_state(dmk_ctx* c) { return (c && c->pts) ? c->pts->max_state : 0; }
extern "C" int dmk_max_dep(dmk_ctx* c) { return (c && c->pts) ? c->pts->max_dep : 0; }

// host copy of a per-point array, cell-major: out[cell][component] (row-major tensors; S as full 3x3)
extern "C" int dmk_get_state(dmk_ctx* c, const char* name, double* out, size_t n_doubles) {
  int rc = need_points(c); if (rc) return rc;
  DmkPointArrays& pa = *c->pts;
  const long long n = pa.n;
  const double* src = nullptr; int comps = 9; bool voigt = false;
  std::string s(name ? name : "");
  if (s == "F") src = pa.F; else if (s == "F0") src = pa.F0; else if (s == "Fp") src = pa.Fp; else if (s == "Fp0") src = pa.Fp0;
  else if (s == "Fe") src = pa.Fe; else if (s == "Lp") src = pa.Lp; else if (s == "Lp0") src = pa.Lp0;
  else if (s == "S") { src = pa.S; comps = 6; voigt = true; } else if (s == "S0") { src = pa.S0; comps = 6; voigt = true; }
  else if (s == "state") { src = pa.state; comps = pa.max_state; } else if (s == "state0") { src = pa.state0; comps = pa.max_state; }
  else if (s == "dep") { src = pa.dep; comps = pa.max_dep; }
  if (s == "cutbacks" || s == "lp_iter") {
    if (n_doubles < (size_t)n) { dmk_set_error("dmk_get_state: buffer too small"); return DMK_ERR_ARG; }
    std::vector<int> ib(n);
    DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    DMK_CUDA_CHECK(cudaMemcpy(ib.data(), s == "cutbacks" ? pa.cutbacks : pa.lp_iter, n * sizeof(int), cudaMemcpyDeviceToHost));
    for (long long i = 0; i < n; i++) out[i] = (double)ib[i];
    return DMK_OK;
  }
  if (!src) { dmk_set_error("dmk_get_state: unknown or empty array '%s'", s.c_str()); return DMK_ERR_ARG; }
  const int out_comps = voigt ? 9 : comps;
  if (n_doubles < (size_t)out_comps * n) { dmk_set_error("dmk_get_state: buffer too small (%zu < %lld)", n_doubles, (long long)out_comps * n); return DMK_ERR_ARG; }
  std::vector<double> buf((size_t)comps * n);
  DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  DMK_CUDA_CHECK(cudaMemcpy(buf.data(), src, buf.size() * sizeof(double), cudaMemcpyDeviceToHost));
  static const int V[9] = {0, 5, 4, 5, 1, 3, 4, 3, 2};
  for (long long i = 0; i < n; i++)
    for (int k = 0; k < out_comps; k++) out[i * out_comps + k] = buf[(size_t)(voigt ? V[k] : k) * n + i];
  return DMK_OK;
}

// inverse of dmk_get_state for the arrays a restart needs (phase_mechanical.f90:1264-1298 writes / reads the same set:
// F, F0, Fp, Fp0, Lp, Lp0, S, S0, the plastic state and its increment-start copy; Fe is kept as well)
extern "C" int dmk_set_state(dmk_ctx* c, const char* name, const double* in, size_t n_doubles) {
  int rc = need_points(c); if (rc) return rc;
  DmkPointArrays& pa = *c->pts;
  const long long n = pa.n;
  double* dst = nullptr; int comps = 9; bool voigt = false;
  std::string s(name ? name : "");
  if (s == "F") dst = pa.F; else if (s == "F0") dst = pa.F0; else if (s == "Fp") dst = pa.Fp; else if (s == "Fp0") dst = pa.Fp0;
  else if (s == "Fe") dst = pa.Fe; else if (s == "Lp") dst = pa.Lp; else if (s == "Lp0") dst = pa.Lp0;
  else if (s == "S") { dst = pa.S; comps = 6; voigt = true; } else if (s == "S0") { dst = pa.S0; comps = 6; voigt = true; }
  else if (s == "state") { dst = pa.state; comps = pa.max_state; } else if (s == "state0") { dst = pa.state0; comps = pa.max_state; }
  else if (s == "dep") { dst = pa.dep; comps = pa.max_dep; }
  if (!dst) { dmk_set_error("dmk_set_state: unknown or empty array '%s'", s.c_str()); return DMK_ERR_ARG; }
  const int in_comps = voigt ? 9 : comps;
  if (n_doubles < (size_t)in_comps * n) { dmk_set_error("dmk_set_state: buffer too small (%zu < %lld)", n_doubles, (long long)in_comps * n); return DMK_ERR_ARG; }
  std::vector<double> buf((size_t)comps * n);
  static const int V6[6] = {0, 4, 8, 5, 2, 1};                       // Voigt (11,22,33,23,13,12) out of the full 3x3
  for (long long i = 0; i < n; i++)
    for (int k = 0; k < comps; k++) buf[(size_t)k * n + i] = in[i * in_comps + (voigt ? V6[k] : k)];
  DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  DMK_CUDA_CHECK(cudaMemcpy(dst, buf.data(), buf.size() * sizeof(double), cudaMemcpyHostToDevice));
  return DMK_OK;
}

// ---- min/max selection of ||dPdF||^2 with the shared tie rule (oracle/spectral.py select_minmax)
__global__ void dmk_minmax_value_kernel(const double* __restrict__ nrm, long long n, double* __restrict__ out2) {
  __shared__ double smx[8], smn[8];
  double mx = 0.0, mn = INFINITY;            // norms are >= 0; an idle block must not contribute -inf (bit pattern order)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double v = nrm[i]; mx = fmax(mx, v); mn = fmin(mn, v);
  }
  for (int off = 16; off > 0; off >>= 1) { mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, off)); mn = fmin(mn, __shfl_down_sync(0xffffffffu, mn, off)); }
  if ((threadIdx.x & 31) == 0) { smx[threadIdx.x >> 5] = mx; smn[threadIdx.x >> 5] = mn; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (blockDim.x >> 5); w++) { mx = fmax(mx, smx[w]); mn = fmin(mn, smn[w]); }
    // doubles >= 0: the IEEE bit pattern is monotone, so integer atomics give an exact max / min
    atomicMax((unsigned long long*)&out2[0], (unsigned long long)__double_as_longlong(mx));
    atomicMin((unsigned long long*)&out2[1], (unsigned long long)__double_as_longlong(mn));
  }
}
__global__ void dmk_minmax_index_kernel(const double* __restrict__ nrm, long long n, long long offset, double vmax, double vmin,
                                        double tol, unsigned long long* __restrict__ idx2) {
  unsigned long long imx = ~0ULL, imn = ~0ULL;
  const double tmax = vmax * (1.0 - tol), tmin = vmin * (1.0 + tol);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double v = nrm[i];
    if (v >= tmax && (unsigned long long)(i + offset) < imx) imx = (unsigned long long)(i + offset);
    if (v <= tmin && (unsigned long long)(i + offset) < imn) imn = (unsigned long long)(i + offset);
  }
  if (imx != ~0ULL) atomicMin(&idx2[0], imx);
  if (imn != ~0ULL) atomicMin(&idx2[1], imn);
}

// arg-max / arg-min of a per-cell norm with the shared tie rule, global over ranks: values within `tol` (relative) of the
// extreme are ties -> lowest global cell index (oracle/spectral.py select_minmax)
static int select_minmax(dmk_ctx* c, const double* d_nrm, long long n, long long gidx[2]) {
  double* d_mm = c->d_scratch + DMK_SCRATCH_TMP + 96;           // 2 values + 2 indices
  double init[2] = {0.0, INFINITY};
  DMK_CUDA_CHECK(cudaMemcpyAsync(d_mm, init, sizeof init, cudaMemcpyHostToDevice, c->stream));
  dmk_minmax_value_kernel<<<148 * 4, 256, 0, c->stream>>>(d_nrm, n, d_mm);
  c->launches++;
  if (c->P > 1) {
    // max over ranks of (vmax, -vmin)
    double h[2];
    DMK_CUDA_CHECK(cudaMemcpyAsync(h, d_mm, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    h[1] = -h[1];
    DMK_CUDA_CHECK(cudaMemcpyAsync(d_mm, h, sizeof h, cudaMemcpyHostToDevice, c->stream));
    if (c->comm.allreduce_max(c->comm.user, d_mm, 2)) { dmk_set_error("allreduce failed"); return DMK_ERR_STATE; }
    DMK_CUDA_CHECK(cudaMemcpyAsync(h, d_mm, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    h[1] = -h[1];
    DMK_CUDA_CHECK(cudaMemcpyAsync(d_mm, h, sizeof h, cudaMemcpyHostToDevice, c->stream));
  }
  double hv[2];
  DMK_CUDA_CHECK(cudaMemcpyAsync(hv, d_mm, sizeof hv, cudaMemcpyDeviceToHost, c->stream));
  DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  unsigned long long hidx[2] = {~0ULL, ~0ULL};
  unsigned long long* d_idx = (unsigned long long*)(d_mm + 2);
  DMK_CUDA_CHECK(cudaMemcpyAsync(d_idx, hidx, sizeof hidx, cudaMemcpyHostToDevice, c->stream));
  const long long offset = (long long)c->z_off * c->nx * c->ny;
  dmk_minmax_index_kernel<<<148 * 4, 256, 0, c->stream>>>(d_nrm, n, offset, hv[0], hv[1], 1.0e-12, d_idx);
  c->launches++;
  DMK_CUDA_CHECK(cudaMemcpyAsync(hidx, d_idx, sizeof hidx, cudaMemcpyDeviceToHost, c->stream));
  DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  gidx[0] = hidx[0] == ~0ULL ? -1 : (long long)hidx[0]; gidx[1] = hidx[1] == ~0ULL ? -1 : (long long)hidx[1];
  if (c->P > 1) {
    // lowest global index over ranks: max of the negated index (exact in a double below 2^53)
    double h[2] = {gidx[0] < 0 ? -9.0e15 : -(double)gidx[0], gidx[1] < 0 ? -9.0e15 : -(double)gidx[1]};
    DMK_CUDA_CHECK(cudaMemcpyAsync(d_mm, h, sizeof h, cudaMemcpyHostToDevice, c->stream));
    if (c->comm.allreduce_max(c->comm.user, d_mm, 2)) { dmk_set_error("allreduce failed"); return DMK_ERR_STATE; }
    DMK_CUDA_CHECK(cudaMemcpyAsync(h, d_mm, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    gidx[0] = (long long)(-h[0]); gidx[1] = (long long)(-h[1]);
  }
  return DMK_OK;
}

// dPdF of every CELL of the slab ([cell][81], the reference's Fortran order): the points' tangents, averaged over the constituents
// with their volume fractions under isostrain homogenisation.  d_Csum81 (may be null under isostrain) receives the plain sum over
// the material points in the kernel's own index order.
int dmk_cells_tangent(dmk_ctx* c, double dt, double* d_Csum81, double* d_cell_dPdF) {
  DmkPointArrays& pa = *c->pts;
  if (pa.n_const <= 1) return dmk_points_tangent(c, dt, d_Csum81, d_cell_dPdF);
  double* d_pt = nullptr;
  if (cudaMalloc((void**)&d_pt, (size_t)pa.n * 81 * sizeof(double)) != cudaSuccess) { cudaGetLastError(); dmk_set_error("out of device memory (point tangents)"); return DMK_ERR_CUDA; }
  int rc = dmk_points_tangent(c, dt, d_Csum81 ? d_Csum81 : c->d_scratch + DMK_SCRATCH_TMP, d_pt);
  if (!rc) {
    dmk_isostrain_average_kernel<<<148 * 8, 256, 0, c->stream>>>(d_pt, pa.v, d_cell_dPdF, pa.n_cells, pa.n_const, 81);
    c->launches++;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) { dmk_set_error("isostrain tangent average failed"); rc = DMK_ERR_CUDA; }
  }
  cudaFree(d_pt);
  return rc;
}

extern "C" int dmk_tangent_response(dmk_ctx* c, double dt, double* C_volAvg, double* C_minMaxAvg, double* d_dPdF,
                                    long long minmax_cells[2]) {
  int rc = need_points(c); if (rc) return rc;
  DmkPointArrays& pa = *c->pts;
  if (c->P > 1 && (!c->has_comm || !c->comm.allreduce_sum || !c->comm.allreduce_max)) { dmk_set_error("no allreduce callbacks (dmk_set_comm)"); return DMK_ERR_STATE; }
  double* d_sum = c->d_scratch + DMK_SCRATCH_TMP;               // 81
  if (pa.n_const > 1) {
    // isostrain: dPdF of a cell = sum_co v_co dPdF_co (homogenization_mechanical.f90:125-145).  The per-point tangents are
    // materialised, averaged per cell, and the reference's reductions run on the cell field (dmk_homogenise)
    double* d_cell = d_dPdF;
    if (!d_cell) { if (cudaMalloc((void**)&d_cell, (size_t)pa.n_cells * 81 * sizeof(double)) != cudaSuccess) { cudaGetLastError(); dmk_set_error("out of device memory (cell tangents)"); return DMK_ERR_CUDA; } }
    rc = dmk_cells_tangent(c, dt, d_sum, d_cell);
    if (!rc) rc = dmk_homogenise(c, pa.Ppt, d_cell, pa.n_cells, nullptr, C_volAvg, C_minMaxAvg, minmax_cells);
    if (d_cell != d_dPdF) cudaFree(d_cell);
    return rc;
  }
  if ((rc = dmk_points_tangent(c, dt, d_sum, d_dPdF))) return rc;
  if (c->P > 1) { if (c->comm.allreduce_sum(c->comm.user, d_sum, 81)) { dmk_set_error("allreduce failed"); return DMK_ERR_STATE; } }
  long long gidx[2];
  if ((rc = select_minmax(c, pa.nrm, pa.n, gidx))) return rc;
  const long long offset = (long long)c->z_off * c->nx * c->ny;
  if (minmax_cells) { minmax_cells[0] = gidx[0]; minmax_cells[1] = gidx[1]; }
  // dPdF of the two selected cells (owner computes, everyone receives through a sum)
  double* d_two = c->d_scratch + DMK_SCRATCH_TMP + 112;          // 2 x 81 -> up to TMP+274: keep inside the MECH gap
  DMK_CUDA_CHECK(cudaMemsetAsync(d_two, 0, 162 * sizeof(double), c->stream));
  for (int k = 0; k < 2; k++) {
    const long long loc = gidx[k] - offset;
    if (loc >= 0 && loc < pa.n) { if ((rc = dmk_points_tangent_cell(c, dt, loc, d_two + 81 * k))) return rc; }
  }
  if (c->P > 1) { if (c->comm.allreduce_sum(c->comm.user, d_two, 162)) { dmk_set_error("allreduce failed"); return DMK_ERR_STATE; } }
  double hs[81], ht[162], r[81];
  DMK_CUDA_CHECK(cudaMemcpyAsync(hs, d_sum, sizeof hs, cudaMemcpyDeviceToHost, c->stream));
  DMK_CUDA_CHECK(cudaMemcpyAsync(ht, d_two, sizeof ht, cudaMemcpyDeviceToHost, c->stream));
  DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  if (C_volAvg) { for (int k = 0; k < 81; k++) r[k] = hs[k] * c->wgt; c2f3333(r, C_volAvg); }
  if (C_minMaxAvg) { for (int k = 0; k < 81; k++) r[k] = 0.5 * (ht[k] + ht[81 + k]); c2f3333(r, C_minMaxAvg); }
  return DMK_OK;
}

// ---- the reductions of Spectral::constitutive_response (/root/reference/src/grid/spectral.cpp:160-218) on GIVEN fields, as
// the reference runs them over homogenization_P / homogenization_dPdF: P_av = wgt sum P, C_volAvg = wgt sum dPdF,
// C_minMaxAvg = (dPdF at arg-max ||dPdF||^2 + dPdF at arg-min) / 2.  d_P: records [cell][9] (c = i + 3 j); d_dPdF: [cell][81] in
// the reference's Fortran order (first index fastest).  Results in Fortran order; n_cells = this rank's cells.
__global__ void __launch_bounds__(128) dmk_dPdF_reduce_kernel(const double* __restrict__ dPdF, long long n, double* __restrict__ nrm,
                                                              double* __restrict__ partial) {
  __shared__ double sh[81];
  for (int k = threadIdx.x; k < 81; k += blockDim.x) sh[k] = 0.0;
  __syncthreads();
  // one warp per cell at a time: lanes over the 81 entries (coalesced), fixed order within a block's cells
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double acc[3] = {0.0, 0.0, 0.0};
  for (long long i = (long long)blockIdx.x * nw + warp; i < n; i += (long long)gridDim.x * nw) {
    double q = 0.0;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const int k = lane + 32 * j;
      const double v = k < 81 ? dPdF[i * 81 + k] : 0.0;
      acc[j] += v; q += v * v;
    }
    for (int off = 16; off > 0; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
    if (lane == 0) nrm[i] = q;
  }
  for (int w = 0; w < nw; w++) {                          // warps add their sums one after the other: deterministic
    if (warp == w) {
#pragma unroll
      for (int j = 0; j < 3; j++) { const int k = lane + 32 * j; if (k < 81) sh[k] += acc[j]; }
    }
    __syncthreads();
  }
  for (int k = threadIdx.x; k < 81; k += blockDim.x) partial[(size_t)blockIdx.x * 81 + k] = sh[k];
}

extern "C" int dmk_homogenise(dmk_ctx* c, const double* d_P, const double* d_dPdF, long long n_cells, double* P_av, double* C_volAvg,
                              double* C_minMaxAvg, long long minmax_cells[2]) {
  if (!c || !d_P || !d_dPdF || n_cells < 1) { dmk_set_error("dmk_homogenise: bad argument"); return DMK_ERR_ARG; }
  if (c->P > 1 && (!c->has_comm || !c->comm.allreduce_sum || !c->comm.allreduce_max)) { dmk_set_error("no allreduce callbacks (dmk_set_comm)"); return DMK_ERR_STATE; }
  int rc;
  double* d_sum9 = c->d_scratch + DMK_SCRATCH_SUM9;
  double* d_sum = c->d_scratch + DMK_SCRATCH_TMP;                // 81
  double* partial = c->d_scratch + DMK_SCRATCH_PARTIAL;
  double* d_nrm = nullptr;
  DMK_CUDA_CHECK(cudaMalloc((void**)&d_nrm, n_cells * sizeof(double)));
  if ((rc = dmk_sum_records(c, d_P, n_cells, d_sum9))) { cudaFree(d_nrm); return rc; }
  const int grid = (int)std::max(1LL, std::min((n_cells + 3) / 4, (long long)DMK_TANGENT_BLOCKS));
  dmk_dPdF_reduce_kernel<<<grid, 128, 0, c->stream>>>(d_dPdF, n_cells, d_nrm, partial);
  dmk_sum_partials_kernel<<<1, 96, 0, c->stream>>>(partial, grid, 81, d_sum);
  c->launches += 2;
  if (c->P > 1) {
    if (c->comm.allreduce_sum(c->comm.user, d_sum9, 9) || c->comm.allreduce_sum(c->comm.user, d_sum, 81)) { cudaFree(d_nrm); dmk_set_error("allreduce failed"); return DMK_ERR_STATE; }
  }
  long long gidx[2];
  if ((rc = select_minmax(c, d_nrm, n_cells, gidx))) { cudaFree(d_nrm); return rc; }
  cudaFree(d_nrm);
  if (minmax_cells) { minmax_cells[0] = gidx[0]; minmax_cells[1] = gidx[1]; }
  const long long offset = (long long)c->z_off * c->nx * c->ny;
  double* d_two = c->d_scratch + DMK_SCRATCH_TMP + 112;
  DMK_CUDA_CHECK(cudaMemsetAsync(d_two, 0, 162 * sizeof(double), c->stream));
  for (int k = 0; k < 2; k++) {
    const long long loc = gidx[k] - offset;
    if (loc >= 0 && loc < n_cells) DMK_CUDA_CHECK(cudaMemcpyAsync(d_two + 81 * k, d_dPdF + loc * 81, 81 * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  }
  if (c->P > 1) { if (c->comm.allreduce_sum(c->comm.user, d_two, 162)) { dmk_set_error("allreduce failed"); return DMK_ERR_STATE; } }
  double h9[9], hs[81], ht[162];
  DMK_CUDA_CHECK(cudaMemcpyAsync(h9, d_sum9, sizeof h9, cudaMemcpyDeviceToHost, c->stream));
  DMK_CUDA_CHECK(cudaMemcpyAsync(hs, d_sum, sizeof hs, cudaMemcpyDeviceToHost, c->stream));
  DMK_CUDA_CHECK(cudaMemcpyAsync(ht, d_two, sizeof ht, cudaMemcpyDeviceToHost, c->stream));
  DMK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  if (P_av) for (int k = 0; k < 9; k++) P_av[k] = h9[k] * c->wgt;                               // record order == Fortran order
  if (C_volAvg) for (int k = 0; k < 81; k++) C_volAvg[k] = hs[k] * c->wgt;                       // the field is in Fortran order already
  if (C_minMaxAvg) for (int k = 0; k < 81; k++) C_minMaxAvg[k] = 0.5 * (ht[k] + ht[81 + k]);
  return DMK_OK;
}
