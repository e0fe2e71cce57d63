// Shared device/host helpers for the DAMASK grid hot path on sm_100a.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

// return codes: the same values as include/dmk_grid.h publishes (guarded there, so that either order of inclusion works)
#ifndef DMK_OK
#define DMK_OK 0
#define DMK_ERR_CUDA 1
#define DMK_ERR_ARG 2
#define DMK_ERR_UNSUPPORTED 3
#define DMK_ERR_STATE 4
#endif

// record = the 9 tensor components of one grid point / k-point, component index c = i + 3*j
// (Fortran order of the reference's F(3,3,x,y,z) arrays: /root/reference/src/grid/spectral.cpp:143)
#define DMK_NC 9

#define DMK_CUDA_CHECK(call)                                                                   \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      dmk_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));    \
      return DMK_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

void dmk_set_error(const char* fmt, ...);

__host__ __device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ double2 cmulc(double2 a, double2 b) {   // a * conj(b)
  return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__host__ __device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }
__host__ __device__ __forceinline__ double2 cscale(double2 a, double s) { return make_double2(a.x * s, a.y * s); }
// multiply by -i (forward) / +i (inverse)
template <bool INV> __host__ __device__ __forceinline__ double2 cmul_mi(double2 a) {
  return INV ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}
