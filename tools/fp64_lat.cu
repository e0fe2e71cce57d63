// FP64 pipe of one SM sub-partition: cycles per warp-instruction as a function of the independent DFMA chains per warp (ILP) and
// the warps per scheduler.  Decides what 2 warps per scheduler (255 registers) can reach in the per-point kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/fp64_lat tools/fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, long long* cyc, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int j = 0; j < ILP; j++) x[j] = threadIdx.x * 1e-3 + j;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 16; r++)
#pragma unroll
      for (int j = 0; j < ILP; j++) x[j] = fma(x[j], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < ILP; j++) s += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP> void run(int warps_per_sm, double* d, long long* dc) {
  const int iters = 2000;
  k<ILP><<<148, warps_per_sm * 32>>>(d, dc, iters, 0.999999, 1e-9);
  k<ILP><<<148, warps_per_sm * 32>>>(d, dc, iters, 0.999999, 1e-9);
  long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  const double inst = (double)iters * 16 * ILP;                 // per warp
  printf("ILP %2d warps/SM %2d (per scheduler %.1f): %.2f cycles per DFMA per warp, %.3f DFMA/cycle/scheduler\n", ILP, warps_per_sm,
         warps_per_sm / 4.0, c / inst, inst * (warps_per_sm / 4.0) / c);
}
int main() {
  double* d; long long* dc;
  cudaMalloc(&d, 148 * 1024 * 8); cudaMalloc(&dc, 8);
  for (int w : {4, 8, 16, 32}) { run<1>(w, d, dc); run<2>(w, d, dc); run<4>(w, d, dc); run<8>(w, d, dc); run<16>(w, d, dc); }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
