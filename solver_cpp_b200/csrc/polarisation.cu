// Point-wise kernels of the polarisation scheme (see polarisation.cuh).
#include <cmath>
#include "polarisation.cuh"

namespace {

__constant__ int NYE_I[6] = {0, 1, 2, 0, 1, 2};      // math.f90 MAPNYE: 11 22 33 12 23 31
__constant__ int NYE_J[6] = {0, 1, 2, 1, 2, 0};

// natural [i][j] from a record (c = i + 3 j)
__device__ __forceinline__ void load33(const double* __restrict__ rec, double (&A)[3][3]) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) A[i][j] = rec[i + 3 * j];
}
__device__ __forceinline__ void store33(double* __restrict__ rec, const double (&A)[3][3]) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) rec[i + 3 * j] = A[i][j];
}
__device__ __forceinline__ void ddot(const double* __restrict__ C, const double (&X)[3][3], double (&Y)[3][3]) {     // Y = C : X
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 3; k++)
#pragma unroll
        for (int l = 0; l < 3; l++) s += C[((i * 3 + j) * 3 + k) * 3 + l] * X[k][l];
      Y[i][j] = s;
    }
}
__device__ __forceinline__ void lambda_of(const double* __restrict__ C, const double (&F)[3][3], const double (&Ft)[3][3], double (&FL)[3][3]) {
  double D[3][3], L[3][3];                           // F . (C : (F_tau - F - I))
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) D[i][j] = Ft[i][j] - F[i][j] - (i == j ? 1.0 : 0.0);
  ddot(C, D, L);
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) FL[i][j] = F[i][0] * L[0][j] + F[i][1] * L[1][j] + F[i][2] * L[2][j];
}

__global__ void __launch_bounds__(128) pol_pre_kernel(const double* __restrict__ dF, const double* __restrict__ dFt, double* __restrict__ dT, long long n,
                                                     const __grid_constant__ DmkPolParams p) {
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
    double F[3][3], Ft[3][3], FL[3][3], E[3][3], CE[3][3], T[3][3];
    load33(dF + c * 9, F); load33(dFt + c * 9, Ft);
    lambda_of(p.C, F, Ft, FL);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) E[i][j] = F[i][j] - (i == j ? 1.0 : 0.0);
    ddot(p.C, E, CE);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) T[i][j] = p.beta * CE[i][j] - p.alpha * FL[i][j];
    store33(dT + c * 9, T);
  }
}

__global__ void pol_rtau_kernel(const double* __restrict__ F, const double* T, double* __restrict__ rFt, double* Fin, long long n9, double beta) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n9; i += (long long)gridDim.x * blockDim.x) {
    const double f = F[i], r = beta * f - T[i];
    rFt[i] = r;
    Fin[i] = f - r / beta;
  }
}

// in-place inverse of a 6 x 6 matrix, Gauss-Jordan with partial pivoting; false if singular
__device__ bool invert6(double (&A)[6][6]) {
  double B[6][6];
  for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) B[i][j] = (i == j) ? 1.0 : 0.0;
  for (int k = 0; k < 6; k++) {
    int piv = k; double mx = fabs(A[k][k]);
    for (int i = k + 1; i < 6; i++) if (fabs(A[i][k]) > mx) { mx = fabs(A[i][k]); piv = i; }
    if (mx == 0.0) return false;
    if (piv != k) for (int j = 0; j < 6; j++) { double t = A[k][j]; A[k][j] = A[piv][j]; A[piv][j] = t; t = B[k][j]; B[k][j] = B[piv][j]; B[piv][j] = t; }
    const double d = 1.0 / A[k][k];
    for (int j = 0; j < 6; j++) { A[k][j] *= d; B[k][j] *= d; }
    for (int i = 0; i < 6; i++) if (i != k) {
      const double f = A[i][k];
      for (int j = 0; j < 6; j++) { A[i][j] -= f * A[k][j]; B[i][j] -= f * B[k][j]; }
    }
  }
  for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) A[i][j] = B[i][j];
  return true;
}

__global__ void __launch_bounds__(128) pol_rF_kernel(const double* __restrict__ dF, const double* __restrict__ dFt, const double* __restrict__ dPdF,
                                                    const double* __restrict__ rFt, double* __restrict__ rF, long long n,
                                                    const __grid_constant__ DmkPolParams p, int* __restrict__ singular) {
  const double w[6] = {1.0, 1.0, 1.0, 1.4142135623730951, 1.4142135623730951, 1.4142135623730951};      // nrmMandel
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
    double F[3][3], Ft[3][3], FL[3][3], X[3][3], A[6][6];
    load33(dF + c * 9, F); load33(dFt + c * 9, Ft);
    lambda_of(p.C, F, Ft, FL);
    load33(rF + c * 9, X);                                                   // P
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) X[i][j] -= FL[i][j];
    const double* T = dPdF + c * 81;                                         // Fortran order: (i,j,k,l) at i + 3 j + 9 k + 27 l
    for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) {
      const int i = NYE_I[a], j = NYE_J[a], k = NYE_I[b], l = NYE_J[b];
      A[a][b] = w[a] * w[b] * (T[i + 3 * j + 9 * k + 27 * l] + p.C[((i * 3 + j) * 3 + k) * 3 + l]);
    }
    if (!invert6(A)) { atomicOr(singular, 1); continue; }
    // (A^-1 as a minor-symmetric 3333) : X  =  sum_b A66inv(a,b) / (w_a w_b) * (X_kl + X_lk, or X_kk)
    double y[6], r6[6];
    for (int b = 0; b < 6; b++) { const int k = NYE_I[b], l = NYE_J[b]; y[b] = (b < 3) ? X[k][l] : X[k][l] + X[l][k]; }
    for (int a = 0; a < 6; a++) { double s = 0.0; for (int b = 0; b < 6; b++) s += A[a][b] / (w[a] * w[b]) * y[b]; r6[a] = s; }
    double R[3][3];
    for (int a = 0; a < 6; a++) { const int i = NYE_I[a], j = NYE_J[a]; R[i][j] = r6[a]; R[j][i] = r6[a]; }
    const double* t = rFt + c * 9;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) R[i][j] += t[i + 3 * j];
    store33(rF + c * 9, R);
  }
}

struct DmkTwo81 { double C[81]; double S[81]; };
__global__ void __launch_bounds__(128) pol_forward_kernel(const double* __restrict__ dF, double* __restrict__ dFt, long long n,
                                                         const __grid_constant__ DmkTwo81 t) {
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
    double F[3][3], Ft[3][3], Fl[3][3], E[3][3], CE[3][3], M[3][3], SM[3][3];
    load33(dF + c * 9, F); load33(dFt + c * 9, Ft);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) Fl[i][j] = Ft[i][j] - F[i][j];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) E[i][j] = Fl[0][i] * Fl[0][j] + Fl[1][i] * Fl[1][j] + Fl[2][i] * Fl[2][j] - (i == j ? 1.0 : 0.0);
    ddot(t.C, E, CE);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) M[i][j] = 0.5 * (Fl[i][0] * CE[0][j] + Fl[i][1] * CE[1][j] + Fl[i][2] * CE[2][j]);
    ddot(t.S, M, SM);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) Ft[i][j] = (i == j ? 1.0 : 0.0) + SM[i][j] + F[i][j];
    store33(dFt + c * 9, Ft);
  }
}

}  // namespace

int dmk_pol_pre(const double* d_F, const double* d_Ftau, double* d_T, long long n, const DmkPolParams& p, cudaStream_t st) {
  pol_pre_kernel<<<148 * 8, 128, 0, st>>>(d_F, d_Ftau, d_T, n, p);
  DMK_CUDA_CHECK(cudaGetLastError());
  return DMK_OK;
}
int dmk_pol_rtau(const double* d_F, const double* d_T, double* d_rFtau, double* d_Fin, long long n, double beta, cudaStream_t st) {
  pol_rtau_kernel<<<148 * 8, 256, 0, st>>>(d_F, d_T, d_rFtau, d_Fin, n * 9, beta);
  DMK_CUDA_CHECK(cudaGetLastError());
  return DMK_OK;
}
int dmk_pol_rF(const double* d_F, const double* d_Ftau, const double* d_dPdF, const double* d_rFtau, double* d_rF, long long n,
               const DmkPolParams& p, int* d_singular, cudaStream_t st) {
  pol_rF_kernel<<<148 * 8, 128, 0, st>>>(d_F, d_Ftau, d_dPdF, d_rFtau, d_rF, n, p, d_singular);
  DMK_CUDA_CHECK(cudaGetLastError());
  return DMK_OK;
}
int dmk_pol_forward_noguess(const double* d_F, double* d_Ftau, long long n, const double* C, const double* S, cudaStream_t st) {
  DmkTwo81 t;
  for (int k = 0; k < 81; k++) { t.C[k] = C[k]; t.S[k] = S[k]; }
  pol_forward_kernel<<<148 * 8, 128, 0, st>>>(d_F, d_Ftau, n, t);
  DMK_CUDA_CHECK(cudaGetLastError());
  return DMK_OK;
}

bool dmk_invsym3333(const double* A, double* Ainv) {
  static const int NI[6] = {0, 1, 2, 0, 1, 2}, NJ[6] = {0, 1, 2, 1, 2, 0};
  const double w[6] = {1.0, 1.0, 1.0, std::sqrt(2.0), std::sqrt(2.0), std::sqrt(2.0)};
  double M[6][6], B[6][6];
  for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) {
    M[a][b] = w[a] * w[b] * A[((NI[a] * 3 + NJ[a]) * 3 + NI[b]) * 3 + NJ[b]];
    B[a][b] = a == b ? 1.0 : 0.0;
  }
  for (int k = 0; k < 6; k++) {
    int piv = k; double mx = std::fabs(M[k][k]);
    for (int i = k + 1; i < 6; i++) if (std::fabs(M[i][k]) > mx) { mx = std::fabs(M[i][k]); piv = i; }
    if (mx == 0.0) return false;
    if (piv != k) for (int j = 0; j < 6; j++) { std::swap(M[k][j], M[piv][j]); std::swap(B[k][j], B[piv][j]); }
    const double d = 1.0 / M[k][k];
    for (int j = 0; j < 6; j++) { M[k][j] *= d; B[k][j] *= d; }
    for (int i = 0; i < 6; i++) if (i != k) { const double f = M[i][k]; for (int j = 0; j < 6; j++) { M[i][j] -= f * M[k][j]; B[i][j] -= f * B[k][j]; } }
  }
  for (int k = 0; k < 81; k++) Ainv[k] = 0.0;
  for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) {
    const double v = B[a][b] / (w[a] * w[b]);
    const int i = NI[a], j = NJ[a], k = NI[b], l = NJ[b];
    Ainv[((i * 3 + j) * 3 + k) * 3 + l] = v; Ainv[((j * 3 + i) * 3 + k) * 3 + l] = v;
    Ainv[((i * 3 + j) * 3 + l) * 3 + k] = v; Ainv[((j * 3 + i) * 3 + l) * 3 + k] = v;
  }
  return true;
}
