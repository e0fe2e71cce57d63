// Deterministic block-partial reductions (sum of records, multi-vector dot products).
#include "reduce.cuh"

__global__ void __launch_bounds__(DMK_REDUCE_THREADS) dmk_sum9_partial_kernel(const double* __restrict__ f, long long n9,
                                                                              double* __restrict__ partial) {
  __shared__ double sh[DMK_REDUCE_THREADS];
  const long long T = (long long)gridDim.x * DMK_REDUCE_THREADS;     // multiple of 9
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * DMK_REDUCE_THREADS + threadIdx.x; i < n9; i += T) acc += f[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < 9) {
    // thread t of block b holds component (b*288 + t) % 9 == t % 9 (288 % 9 == 0)
    double s = 0.0;
    for (int k = 0; k < DMK_REDUCE_THREADS / 9; k++) s += sh[threadIdx.x + 9 * k];
    partial[(size_t)blockIdx.x * 9 + threadIdx.x] = s;
  }
}
__global__ void dmk_sum_partials_kernel(const double* __restrict__ partial, int nblocks, int width, double* __restrict__ out) {
  int c = threadIdx.x;
  if (c < width) {
    double s = 0.0;
    for (int b = 0; b < nblocks; b++) s += partial[(size_t)b * width + c];
    out[c] = s;
  }
}

int dmk_sum_records(dmk_ctx* c, const double* d_field, long long n_records, double* d_out9) {
  long long n9 = n_records * 9;
  int blocks = (int)((n9 + DMK_REDUCE_THREADS - 1) / DMK_REDUCE_THREADS);
  if (blocks > DMK_REDUCE_BLOCKS) blocks = DMK_REDUCE_BLOCKS;
  if (blocks < 1) blocks = 1;
  double* partial = c->d_scratch + DMK_SCRATCH_PARTIAL;
  dmk_sum9_partial_kernel<<<blocks, DMK_REDUCE_THREADS, 0, c->stream>>>(d_field, n9, partial);
  dmk_sum_partials_kernel<<<1, 32, 0, c->stream>>>(partial, blocks, 9, d_out9);
  c->launches += 2;
  DMK_CUDA_CHECK(cudaGetLastError());
  return DMK_OK;
}

// ---- multi-dot: out[k] = a . b_k for up to 32 vectors per launch
struct DmkPtrs { const double* p[32]; };
__global__ void __launch_bounds__(256) dmk_multidot_partial_kernel(const double* __restrict__ a, DmkPtrs bs, int nvec,
                                                                   long long n, double* __restrict__ partial) {
  __shared__ double sh[8];
  for (int k = 0; k < nvec; k++) {
    const double* __restrict__ b = bs.p[k];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      acc += a[i] * b[i];
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int w = 0; w < 8; w++) s += sh[w];
      partial[(size_t)blockIdx.x * nvec + k] = s;
    }
    __syncthreads();
  }
}

int dmk_multi_dot(dmk_ctx* c, const double* d_a, const double* const* bptrs, int nvec, long long n, double* d_out) {
  if (nvec < 1) return DMK_OK;
  if (nvec > 32) { dmk_set_error("dmk_multi_dot: at most 32 vectors"); return DMK_ERR_ARG; }
  DmkPtrs bs; for (int k = 0; k < 32; k++) bs.p[k] = (k < nvec) ? bptrs[k] : nullptr;
  int blocks = (int)((n + 255) / 256);
  if (blocks > DMK_REDUCE_BLOCKS) blocks = DMK_REDUCE_BLOCKS;
  if (blocks < 1) blocks = 1;
  double* partial = c->d_scratch + DMK_SCRATCH_PARTIAL;
  dmk_multidot_partial_kernel<<<blocks, 256, 0, c->stream>>>(d_a, bs, nvec, n, partial);
  dmk_sum_partials_kernel<<<1, 32, 0, c->stream>>>(partial, blocks, nvec, d_out);
  c->launches += 2;
  DMK_CUDA_CHECK(cudaGetLastError());
  return DMK_OK;
}
