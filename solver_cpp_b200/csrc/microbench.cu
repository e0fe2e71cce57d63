// Device micro-benchmark used as the roofline denominator of the FP64 constitutive kernel: independent DFMA chains
// held in registers, enough warps to fill the FP64 pipe of every SM.
#include "dmk_common.cuh"

__global__ void __launch_bounds__(256) dmk_dfma_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

extern "C" int dmk_fp64_peak(double* tflops) {
  int dev = 0, sms = 0;
  DMK_CUDA_CHECK(cudaGetDevice(&dev));
  DMK_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int blocks = sms * 8, threads = 256, iters = 4096;
  double* d = nullptr;
  DMK_CUDA_CHECK(cudaMalloc((void**)&d, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  DMK_CUDA_CHECK(cudaEventCreate(&e0)); DMK_CUDA_CHECK(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 5; rep++) {
    DMK_CUDA_CHECK(cudaEventRecord(e0));
    dmk_dfma_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
    DMK_CUDA_CHECK(cudaEventRecord(e1));
    DMK_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    DMK_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * threads;
    if (rep > 0) best = fmax(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  DMK_CUDA_CHECK(cudaGetLastError());
  *tflops = best;
  return DMK_OK;
}
