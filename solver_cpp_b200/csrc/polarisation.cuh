// Point-wise kernels of the polarisation variant of the mechanical spectral solver (mech_solver.cu: evaluate_polarisation).
// The scheme is upstream DAMASK's grid_mechanical_spectral_polarisation; the reference tree has only its numerics keys
// (/root/reference/src/grid/config.cpp:450-457, config.h:68-71) -- see oracle/polarisation.py for the restated algorithm.
// Fields are 72-byte records, component c = i + 3 j; C_scale is row-major C(i,j,k,l) at ((i*3+j)*3+k)*3+l.
#pragma once
#include "dmk_common.cuh"

struct DmkPolParams { double C[81]; double alpha, beta; };

// T = beta C:(F - I) - alpha F.(C:(F_tau - F - I))
int dmk_pol_pre(const double* d_F, const double* d_Ftau, double* d_T, long long n, const DmkPolParams& p, cudaStream_t st);
// r_F_tau = beta F - T ;  F_in = F - r_F_tau / beta      (T and F_in may be the same buffer)
int dmk_pol_rtau(const double* d_F, const double* d_T, double* d_rFtau, double* d_Fin, long long n, double beta, cudaStream_t st);
// r_F = (dPdF + C)^-1 : (P - F.(C:(F_tau - F - I))) + r_F_tau, the inverse through the weighted 6x6 form of math_invSym3333;
// P comes in d_rF and is overwritten; dPdF: [cell][81] in the reference's Fortran order (first index fastest)
int dmk_pol_rF(const double* d_F, const double* d_Ftau, const double* d_dPdF, const double* d_rFtau, double* d_rF, long long n,
               const DmkPolParams& p, int* d_singular, cudaStream_t st);
// forward of F_tau without a rate guess: F_l = F_tau - F;  F_tau = I + S:(0.5 F_l.(C:(F_l^T F_l - I))) + F
int dmk_pol_forward_noguess(const double* d_F, double* d_Ftau, long long n, const double* C, const double* S, cudaStream_t st);
// host: math_invSym3333 of a row-major 3333 tensor; false if singular
bool dmk_invsym3333(const double* A, double* Ainv);
