// Consistent tangent dP/dF of one material point (phase_mechanical_dPdF, /root/reference/src/phase_mechanical.f90:
// 1137-1258) with Fi = I, Li = 0 (dFi/dS = 0, dLi/dS = 0).  Runs once per converged increment (only the tangent of
// the last evaluation is ever consumed: mech_solver_basic.cpp:116), so it trades registers for clarity: 3x3x3x3
// tensors live in per-thread local memory as row-major 9x9 matrices ((i,j) row, (k,l) column).
#pragma once
#include "constitutive_dev.cuh"

#define T4(t, i, j, k, l) (t)[(((i) * 3 + (j)) * 3 + (k)) * 3 + (l)]

DMK_HD __noinline__ void mul99(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C) {
  for (int i = 0; i < 9; i++)
    for (int j = 0; j < 9; j++) {
      double s = 0.0;
      for (int k = 0; k < 9; k++) s += A[9 * i + k] * B[9 * k + j];
      C[9 * i + j] = s;
    }
}
// math_invert (dgetrf + dgetri semantics): LU with partial pivoting; true on exact singularity
DMK_HD __noinline__ bool invert99(const double* __restrict__ A, double* __restrict__ inv) {
  double LU[81]; int piv[9];
  for (int k = 0; k < 81; k++) LU[k] = A[k];
  for (int k = 0; k < 9; k++) {
    int p = k; double mx = fabs(LU[9 * k + k]);
    for (int i = k + 1; i < 9; i++) if (fabs(LU[9 * i + k]) > mx) { mx = fabs(LU[9 * i + k]); p = i; }
    piv[k] = p;
    if (LU[9 * p + k] == 0.0) { for (int q = 0; q < 81; q++) inv[q] = 0.0; return true; }
    if (p != k) for (int j = 0; j < 9; j++) { double t = LU[9 * k + j]; LU[9 * k + j] = LU[9 * p + j]; LU[9 * p + j] = t; }
    for (int i = k + 1; i < 9; i++) {
      LU[9 * i + k] /= LU[9 * k + k];
      const double l = LU[9 * i + k];
      for (int j = k + 1; j < 9; j++) LU[9 * i + j] -= l * LU[9 * k + j];
    }
  }
  for (int cidx = 0; cidx < 9; cidx++) {
    double col[9];
    for (int i = 0; i < 9; i++) col[i] = (i == cidx) ? 1.0 : 0.0;
    for (int k = 0; k < 9; k++) { const int p = piv[k]; if (p != k) { double t = col[k]; col[k] = col[p]; col[p] = t; } }
    for (int i = 0; i < 9; i++) { double s = col[i]; for (int j = 0; j < i; j++) s -= LU[9 * i + j] * col[j]; col[i] = s; }
    for (int i = 8; i >= 0; i--) { double s = col[i]; for (int j = i + 1; j < 9; j++) s -= LU[9 * i + j] * col[j]; col[i] = s / LU[9 * i + i]; }
    for (int i = 0; i < 9; i++) inv[9 * i + cidx] = col[i];
  }
  return false;
}

DMK_HD __forceinline__ void inv33z(const double (&A)[9], double (&inv)[9]) { (void)invert33(A, inv); }   // math_inv33: 0 on failure

// dLp/dMp as a full 9x9 (row = Lp component, column = Mp component)
template <int LAW, int NS, int ND>
DMK_HD __noinline__ void full_dLp_dMp(DmkPt<LAW, NS, ND>& p, double* __restrict__ dM) {
  for (int k = 0; k < 81; k++) dM[k] = 0.0;
  const DmkDevPhase& d = *p.ph;
  if constexpr (LAW == DMK_LAW_ISOTROPIC) {
    double dev[9], sq;
    deviator(p.S, dev, sq);
    const double nrm = sqrt(sq);
    if (nrm > 0.0) {
      const double ratio = sqrt(1.5) * nrm / (d.p.iso_M * p.st[0]);
      const double dgam = d.p.iso_dot_gamma_0 * (d.iso_n_int ? powi(ratio, d.iso_n_int) : pow(ratio, d.p.iso_n));
      for (int k = 0; k < 9; k++) for (int m = 0; m < 9; m++) dM[9 * k + m] = (d.p.iso_n - 1.0) * dev[k] * dev[m] / sq;
      for (int k = 0; k < 9; k++) dM[10 * k] += 1.0;
      for (int k = 0; k < 3; k++) for (int m = 0; m < 3; m++) T4(dM, k, k, m, m) -= 1.0 / 3.0;
      for (int k = 0; k < 81; k++) dM[k] = dgam * dM[k] / nrm;
    }
  } else if constexpr (LAW == DMK_LAW_PHENOPOWERLAW) {
    const int ns = d.p.n_sl, nt = d.p.n_tw;
    for (int i = 0; i < ns; i++) {
      double dg, ddg;
      pheno_slip_rate(p, p.S, i, dg, ddg);
      const double* PnS = (dg > 0.0) ? d.p.P_nS_pos[i] : d.p.P_nS_neg[i];
      for (int a = 0; a < 9; a++) for (int b = 0; b < 9; b++) dM[9 * a + b] += ddg * d.p.P_sl[i][a] * PnS[b];
    }
    if (nt > 0) {
      const double sumF = pheno_sumF(p);
      for (int i = 0; i < nt; i++) {
        double dg, ddg;
        pheno_twin_rate(p, p.S, i, sumF, dg, ddg);
        for (int a = 0; a < 9; a++) for (int b = 0; b < 9; b++) dM[9 * a + b] += ddg * d.p.P_tw[i][a] * d.p.P_tw[i][b];
      }
    }
  } else if constexpr (LAW == DMK_LAW_KINEHARDENING) {           // phase_mechanical_plastic_kinehardening.f90:258-295
    for (int i = 0; i < d.p.n_sl; i++) {
      double dg, ddg;
      kh_slip_rate(p, p.S, i, dg, ddg);
      const double* PnS = (dg > 0.0) ? d.p.P_nS_pos[i] : d.p.P_nS_neg[i];
      for (int a = 0; a < 9; a++) for (int b = 0; b < 9; b++) dM[9 * a + b] += ddg * d.p.P_sl[i][a] * PnS[b];
    }
  } else if constexpr (LAW == DMK_LAW_DISLOTUNGSTEN) {           // phase_mechanical_plastic_dislotungsten.f90:270-310
    for (int i = 0; i < d.p.n_sl; i++) {
      double dg, ddg, te;
      tu_slip_rate(p, p.S, i, true, dg, ddg, te);
      const double* PnS = (dg > 0.0) ? d.p.P_nS_pos[i] : d.p.P_nS_neg[i];
      for (int a = 0; a < 9; a++) for (int b = 0; b < 9; b++) dM[9 * a + b] += ddg * d.p.P_sl[i][a] * PnS[b];
    }
  } else if constexpr (LAW == DMK_LAW_DISLOTWIN) {
    const int ns = d.p.n_sl, nt = d.p.n_tw;
    double dgs[DMK_MAX_SL];
    const double fm = dt_f_matrix(p);
    for (int i = 0; i < ns; i++) {
      double dg, ddg;
      dt_slip_rate(p, p.S, i, true, dg, ddg);
      dgs[i] = dg;
      for (int a = 0; a < 9; a++) for (int b = 0; b < 9; b++) dM[9 * a + b] += ddg * d.p.P_sl[i][a] * d.p.P_sl[i][b];
    }
    for (int i = 0; i < nt; i++) {
      double dg, ddg;
      dt_twin_rate(p, p.S, i, dgs, true, dg, ddg);
      for (int a = 0; a < 9; a++) for (int b = 0; b < 9; b++) dM[9 * a + b] += ddg * d.p.P_tw[i][a] * d.p.P_tw[i][b];
    }
    for (int k = 0; k < 81; k++) dM[k] *= fm;
  }
}

template <int LAW, int NS, int ND>
DMK_HD __noinline__ void tangent_point(DmkPt<LAW, NS, ND>& p, const double (&F)[9], const double (&Fp0)[9], double Dt,
                                           double* __restrict__ dPdF) {
  double dSdFe[81], dLpdS[81], rhs[81], tmp[81], lhs[81], t99[81], dSdF[81];
  if constexpr (LAW == DMK_LAW_DISLOTWIN) dt_homogenized_C66(p);
  const double* C66 = p.C66;
  const int V[9] = {0, 5, 4, 5, 1, 3, 4, 3, 2};
  // dS/dFe(i,j,k,l) = sum_m Fe(k,m) C(i,j,m,l)
  for (int ij = 0; ij < 9; ij++)
    for (int k = 0; k < 3; k++)
      for (int l = 0; l < 3; l++) {
        double s = 0.0;
        for (int m = 0; m < 3; m++) s += p.Fe[3 * k + m] * C66[6 * V[ij] + V[3 * m + l]];
        dSdFe[9 * ij + 3 * k + l] = s;
      }
  full_dLp_dMp(p, dLpdS);                     // FiTFi = I: dLp/dS = dLp/dMp
  double invFp[9], invSubFp0[9], invFpT[9], t2[9];
  inv33z(p.Fp, invFp); inv33z(Fp0, invSubFp0);
  tr3(invFp, invFpT);
  mm3(F, invSubFp0, t2);
  for (int q = 0; q < 3; q++) for (int o = 0; o < 3; o++) {
    double d[9], a[9];
    for (int k = 0; k < 9; k++) d[k] = dSdFe[9 * (3 * q + o) + k];
    mm3(d, invFpT, a);
    for (int k = 0; k < 9; k++) rhs[9 * (3 * q + o) + k] = a[k];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) d[3 * i + j] = T4(dLpdS, i, j, q, o);
    mm3(t2, d, a);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T4(tmp, i, j, q, o) = a[3 * i + j];
  }
  mul99(dSdFe, tmp, lhs);
  for (int k = 0; k < 81; k++) lhs[k] = lhs[k] * Dt;
  for (int k = 0; k < 9; k++) lhs[10 * k] += 1.0;
  if (invert99(lhs, t99)) { for (int k = 0; k < 81; k++) dSdF[k] = rhs[k]; }
  else mul99(t99, rhs, dSdF);
  // dFp^-1/dF (stored in lhs)
  mul99(dLpdS, dSdF, tmp);
  double* dFpinvdF = lhs;
  for (int o = 0; o < 3; o++) for (int q = 0; q < 3; q++) {
    double d[9], a[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) d[3 * i + j] = T4(tmp, i, j, q, o);
    mm3(invSubFp0, d, a);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T4(dFpinvdF, i, j, q, o) = -a[3 * i + j] * Dt;
  }
  // assemble
  double Sm[9], t1[9], t3[9], a0[9];
  Sm[0] = p.S[0]; Sm[1] = p.S[5]; Sm[2] = p.S[4]; Sm[3] = p.S[5]; Sm[4] = p.S[1]; Sm[5] = p.S[3]; Sm[6] = p.S[4]; Sm[7] = p.S[3]; Sm[8] = p.S[2];
  mm3(Sm, invFpT, t1);
  mm3(F, invFp, t2);
  mm3(t2, Sm, t3);
  for (int k = 0; k < 81; k++) dPdF[k] = 0.0;
  mm3(invFp, t1, a0);
  for (int q = 0; q < 3; q++) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T4(dPdF, q, i, q, j) = a0[3 * j + i];
  for (int o = 0; o < 3; o++) for (int q = 0; q < 3; q++) {
    double d[9], a[9], b[9], cc[9], u[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) d[3 * i + j] = T4(dFpinvdF, i, j, q, o);
    mm3(F, d, u); mm3(u, t1, a);
    mm3_nt(t3, d, cc);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) d[3 * i + j] = T4(dSdF, i, j, q, o);
    mm3(t2, d, u); mm3(u, invFpT, b);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T4(dPdF, i, j, q, o) += a[3 * i + j] + b[3 * i + j] + cc[3 * i + j];
  }
}
