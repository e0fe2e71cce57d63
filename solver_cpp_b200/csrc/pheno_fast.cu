// Kernel and launcher of the register-resident slip-only phenopowerlaw formulation (pheno_fast_dev.cuh); replaces
// f_homogenization_mechanical_response (/root/reference/src/homogenization.f90:219-266) for such phases.
#include <algorithm>
#include <cstdint>
#include "pheno_fast_dev.cuh"

// Register-resident slip-only phenopowerlaw kernel (pheno_fast_dev.cuh): one thread per point, one block per
// DMK_PF_THREADS points, phase tables as a __grid_constant__ parameter, state-integrator arrays in shared memory.
#define DMK_PF_THREADS 128
template <int NSL, int NEXP>
__global__ void __launch_bounds__(DMK_PF_THREADS, 2) dmk_pheno_fast_kernel(const __grid_constant__ DmkPhenoK<NSL> K, DmkPointArrays pa,
                                                                            int phase, DmkNumDev num, const double* __restrict__ Fin,
                                                                            double* __restrict__ Pout, double Dt) {
  extern __shared__ double pf_sm[];
  __shared__ unsigned pf_cnt[7];
  if (threadIdx.x < 7) pf_cnt[threadIdx.x] = 0;
  __syncthreads();
  const DmkScr scr = {pf_sm + threadIdx.x, DMK_PF_THREADS};
  DmkCnt cnt = {0, 0, 0, 0, 0, 0};
  unsigned ill = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pa.n; i += (long long)gridDim.x * blockDim.x) {
    if (pa.phase[i] != phase) continue;
    if (!pf_point<NSL, NEXP>(K, num, pa, i, Fin, Pout, Dt, scr, cnt)) ill = 1;
  }
  unsigned v[7] = {cnt.Lp_iter, cnt.Li_iter, cnt.stress_calls, cnt.state_iter, cnt.substeps, cnt.cutbacks, ill};
#pragma unroll
  for (int k = 0; k < 7; k++) {
    unsigned s = v[k];
    for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(&pf_cnt[k], s);
  }
  __syncthreads();
  if (threadIdx.x < 7 && pf_cnt[threadIdx.x]) atomicAdd(&pa.d_counters[threadIdx.x], (unsigned long long)pf_cnt[threadIdx.x]);
  if (threadIdx.x == 0 && pf_cnt[6]) atomicOr(pa.d_ill, 1);
}

template <int NSL, int NEXP>
static int launch_pheno_fast(dmk_ctx* c, DmkPointArrays& pa, int ph, const DmkDevPhase& d, const DmkNumDev& nd, const double* d_F,
                             double* d_P, double dt) {
  DmkPhenoK<NSL> K;
  dmk_pheno_fast_tables<NSL>(d, K);
  const size_t smem = (size_t)PfOff<NSL>::TOTAL * DMK_PF_THREADS * sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    DMK_CUDA_CHECK(cudaFuncSetAttribute(dmk_pheno_fast_kernel<NSL, NEXP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const long long blocks = (pa.n + DMK_PF_THREADS - 1) / DMK_PF_THREADS;
  const int grid = (int)std::max(1LL, std::min(blocks, (long long)INT32_MAX));
  dmk_pheno_fast_kernel<NSL, NEXP><<<grid, DMK_PF_THREADS, smem, c->stream>>>(K, pa, ph, nd, d_F, d_P, dt);
  return DMK_OK;
}

// Runs phase `ph` through the register-resident kernel if it qualifies (*handled), else leaves it to the generic kernel.
int dmk_points_pheno_fast(dmk_ctx* c, int ph, const DmkNumDev& nd, const double* d_F, double* d_P, double dt, bool* handled) {
  DmkPointArrays& pa = *c->pts;
  int nexp = 0;
  *handled = false;
  if (!dmk_pheno_fast_eligible(pa.h_dev_phases[ph], 12, &nexp)) return DMK_OK;       // 12 slip systems: fcc, bcc {110}
  *handled = true;
  int rc = (nexp == 20) ? launch_pheno_fast<12, 20>(c, pa, ph, pa.h_dev_phases[ph], nd, d_F, d_P, dt)
                        : launch_pheno_fast<12, 0>(c, pa, ph, pa.h_dev_phases[ph], nd, d_F, d_P, dt);
  if (rc) return rc;
  DMK_CUDA_CHECK(cudaGetLastError());
  return DMK_OK;
}
