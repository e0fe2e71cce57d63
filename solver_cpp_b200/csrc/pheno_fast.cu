// Kernel and launcher of the register-resident slip-only phenopowerlaw formulation (pheno_fast_dev.cuh); replaces
// f_homogenization_mechanical_response (/root/reference/src/homogenization.f90:219-266) for such phases.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include "pheno_fast_dev.cuh"

// One thread per point, one warp-sized pair of warps per block (DMK_PF_THREADS points), phase tables as a
// __grid_constant__ parameter, state-integrator arrays in shared memory ([k][thread]).  Blocks are small and end
// without a barrier: the per-point work varies by a factor of several (Newton iteration counts), and with 128-thread
// blocks 18 % of all warp samples were warps parked at the final __syncthreads waiting for the slowest one
// (profiles/r01_summary.md).  Points whose Newton path is ill-conditioned are appended to pa.defer_list (pf_stress).
#define DMK_PF_THREADS 64
#ifndef DMK_PF_MINBLOCKS
#define DMK_PF_MINBLOCKS 4
#endif
template <int NSL, int NEXP, int UNR>
__global__ void __launch_bounds__(DMK_PF_THREADS, DMK_PF_MINBLOCKS)
dmk_pheno_fast_kernel(const __grid_constant__ DmkPhenoK<NSL> K, DmkPointArrays pa, int phase, DmkNumDev num, int defer_nLp,
                      const double* __restrict__ Fin, double* __restrict__ Pout, double Dt) {
  extern __shared__ double pf_sm[];
  const DmkScr scr = {pf_sm + threadIdx.x, DMK_PF_THREADS};
  DmkCnt cnt = {0, 0, 0, 0, 0, 0};
  unsigned ill = 0, ndefer = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pa.n; i += (long long)gridDim.x * blockDim.x) {
    if (pa.phase[i] != phase) continue;
    const int rc = pf_point<NSL, NEXP, UNR>(K, num, defer_nLp, pa, i, Fin, Pout, Dt, scr, cnt);
    if (rc == PF_DEFER) { pa.defer_list[atomicAdd(pa.d_defer_count, 1u)] = (int)i; ndefer++; }
    else if (rc != PF_OK) ill = 1;
  }
  // counters: warp reduce, one atomic per warp and non-zero counter
  unsigned v[8] = {cnt.Lp_iter, cnt.Li_iter, cnt.stress_calls, cnt.state_iter, cnt.substeps, cnt.cutbacks, ill, ndefer};
#pragma unroll
  for (int k = 0; k < 8; k++) {
    unsigned s = v[k];
    for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(&pa.d_counters[k], (unsigned long long)s);
  }
  if (ill) atomicOr(pa.d_ill, 1);
}

// DMK_PHENO_DEFER=<n>: hand a point to the reference-order kernel when one stress integration needs more than n Newton
// iterations or a line-search step (0: never; default 8)
static int defer_setting() {
  static const int v = [] { const char* e = getenv("DMK_PHENO_DEFER"); return e ? atoi(e) : 8; }();
  return v < 0 ? 0 : v;
}

// DMK_PHENO_UNROLL=<1|4|12>: unroll factor of the slip-system loops inside the Newton iteration
static int unroll_setting() {
  static const int v = [] { const char* e = getenv("DMK_PHENO_UNROLL"); return e ? atoi(e) : 12; }();
  return v;
}

template <int NSL, int NEXP, int UNR>
static int launch_pheno_fast(dmk_ctx* c, DmkPointArrays& pa, int ph, const DmkDevPhase& d, const DmkNumDev& nd, const double* d_F,
                             double* d_P, double dt) {
  DmkPhenoK<NSL> K;
  dmk_pheno_fast_tables<NSL>(d, K);
  const size_t smem = (size_t)PfOff<NSL>::TOTAL * DMK_PF_THREADS * sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    DMK_CUDA_CHECK(cudaFuncSetAttribute(dmk_pheno_fast_kernel<NSL, NEXP, UNR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const long long blocks = (pa.n + DMK_PF_THREADS - 1) / DMK_PF_THREADS;
  const int grid = (int)std::max(1LL, std::min(blocks, (long long)INT32_MAX));
  DMK_CUDA_CHECK(cudaMemsetAsync(pa.d_defer_count, 0, sizeof(unsigned), c->stream));
  dmk_pheno_fast_kernel<NSL, NEXP, UNR><<<grid, DMK_PF_THREADS, smem, c->stream>>>(K, pa, ph, nd, defer_setting(), d_F, d_P, dt);
  return DMK_OK;
}

// Runs phase `ph` through the register-resident kernel if it qualifies (*handled), else leaves it to the generic kernel.
int dmk_points_pheno_fast(dmk_ctx* c, int ph, const DmkNumDev& nd, const double* d_F, double* d_P, double dt, bool* handled,
                          bool* deferred) {
  DmkPointArrays& pa = *c->pts;
  int nexp = 0;
  *handled = false; *deferred = false;
  if (!dmk_pheno_fast_eligible(pa.h_dev_phases[ph], 12, &nexp)) return DMK_OK;       // 12 slip systems: fcc, bcc {110}
  *handled = true; *deferred = defer_setting() > 0;
  const DmkDevPhase& d = pa.h_dev_phases[ph];
  const int u = unroll_setting();
  int rc;
  if (nexp == 20) rc = (u == 1) ? launch_pheno_fast<12, 20, 1>(c, pa, ph, d, nd, d_F, d_P, dt)
                 : (u == 4) ? launch_pheno_fast<12, 20, 4>(c, pa, ph, d, nd, d_F, d_P, dt)
                            : launch_pheno_fast<12, 20, 12>(c, pa, ph, d, nd, d_F, d_P, dt);
  else rc = (u == 1) ? launch_pheno_fast<12, 0, 1>(c, pa, ph, d, nd, d_F, d_P, dt)
          : (u == 4) ? launch_pheno_fast<12, 0, 4>(c, pa, ph, d, nd, d_F, d_P, dt)
                     : launch_pheno_fast<12, 0, 12>(c, pa, ph, d, nd, d_F, d_P, dt);
  if (rc) return rc;
  DMK_CUDA_CHECK(cudaGetLastError());
  return DMK_OK;
}
