// Device functions of the per-material-point FP64 constitutive update, one thread per point.
//
// Re-expresses (does not copy) the reference's constitutive library as called through
// f_homogenization_mechanical_response (/root/reference/src/homogenization.f90:219-266):
//   phase_mechanical_constitutive   phase_mechanical.f90:1013-1100   adaptive sub-stepping
//   integrateStateFPI               phase_mechanical.f90:611-693     damped fixed point on the state
//   integrateStress                 phase_mechanical.f90:380-604     Newton on Lp (Li loop degenerate: Li = 0, Fi = I)
//   phase_hooke_SandItsTangents     phase_mechanical_elastic.f90:157-195
//   phenopowerlaw / isotropic / dislotwin kinetics and dotState
//                                   phase_mechanical_plastic_{phenopowerlaw.f90:298-533,isotropic.f90:156-290,
//                                   dislotwin.f90:458-971}
//   phase_mechanical_dPdF           phase_mechanical.f90:1137-1258   (tangent kernel)
//
// B200 shape: everything a Newton iteration touches lives in registers.  The reference builds three 3x3x3x3
// tensors per iteration and multiplies them as 9x9 matrices (2 x 729 FMA + LAPACK dgesv); here the Jacobian
//   dR/dLp = I - dLp/dS : dS/dFe : dFe/dLp
// is assembled in closed form from the rank-one structure of crystal plasticity:
//   dLp/dS = sum_a  dgamma_a/dtau_a  P_a (x) Pn_a ,   dS/dFe(i,j,k,l) = [Fe C_ij](k,l) ,   dFe/dLp = -dt A (x) I
//   =>  dR/dLp = I + dt sum_a dgamma_a/dtau_a  P_a (x) (G Q_a) ,   G = A^T Fe ,  Q_a = C : Pn_a  (constant per system)
// i.e. 27 + 81 FMA per slip system, and solved by an unrolled, register-resident 9x9 LU with partial pivoting
// (predicated row swaps, no dynamic indexing).
#pragma once
#include <cfloat>
#include "constitutive.cuh"

#define DMK_HD __host__ __device__    // host instantiation exists only for the test-side twin (tests/host_twin)
#define DMK_TOL_MATH_CHECK 1.0e-8            // prec.f90:34
#define DMK_PI 3.14159265358979323846

DMK_HD __forceinline__ bool deq0(double a) { return fabs(a) <= 10.0 * DBL_MIN; }                // prec.f90:113-125
DMK_HD __forceinline__ bool deq(double a, double b) { return fabs(a - b) <= DBL_EPSILON * fmax(fabs(a), fabs(b)); }
DMK_HD __forceinline__ double fsign(double a, double b) { return copysign(fabs(a), b); }

// x^n for a small positive integer n by repeated squaring (a few ulp, like pow)
DMK_HD __forceinline__ double powi(double x, int n) {
  double r = 1.0, b = x;
  while (n) { if (n & 1) r *= b; b *= b; n >>= 1; }
  return r;
}

struct DmkCnt { unsigned Lp_iter, Li_iter, stress_calls, state_iter, substeps, cutbacks; };

// ------------------------------------------------------------------------------------------------ 3x3 helpers
DMK_HD __forceinline__ void mm3(const double (&A)[9], const double (&B)[9], double (&C)[9]) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
DMK_HD __forceinline__ void mm3_tn(const double (&A)[9], const double (&B)[9], double (&C)[9]) {     // C = A^T B
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
DMK_HD __forceinline__ void mm3_nt(const double (&A)[9], const double (&B)[9], double (&C)[9]) {     // C = A B^T
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
}
DMK_HD __forceinline__ void tr3(const double (&A)[9], double (&T)[9]) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) T[3 * i + j] = A[3 * j + i];
}
DMK_HD __forceinline__ double det3(const double (&A)[9]) {
  return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
}
// math_invert33 (math.f90:511): Cramer; true on error
DMK_HD __forceinline__ bool invert33(const double (&A)[9], double (&inv)[9]) {
  double c00 = A[4] * A[8] - A[5] * A[7];
  double c10 = -A[3] * A[8] + A[5] * A[6];
  double c20 = A[3] * A[7] - A[4] * A[6];
  double det = A[0] * c00 + A[1] * c10 + A[2] * c20;
  if (deq0(det)) {
#pragma unroll
    for (int i = 0; i < 9; i++) inv[i] = 0.0;
    return true;
  }
  // quotient, not reciprocal multiply, to stay on the oracle's rounding
  inv[0] = c00 / det; inv[3] = c10 / det; inv[6] = c20 / det;
  inv[1] = (-A[1] * A[8] + A[2] * A[7]) / det;
  inv[4] = (A[0] * A[8] - A[2] * A[6]) / det;
  inv[7] = (-A[0] * A[7] + A[1] * A[6]) / det;
  inv[2] = (A[1] * A[5] - A[2] * A[4]) / det;
  inv[5] = (-A[0] * A[5] + A[2] * A[3]) / det;
  inv[8] = (A[0] * A[4] - A[1] * A[3]) / det;
  return false;
}
DMK_HD __forceinline__ double norm9(const double (&A)[9]) {
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 9; i++) s += A[i] * A[i];
  return sqrt(s);
}
DMK_HD __forceinline__ bool any_nan9(const double (&A)[9]) {
  bool b = false;
#pragma unroll
  for (int i = 0; i < 9; i++) b = b || isnan(A[i]);
  return b;
}
// S (Voigt 6: 11,22,33,23,13,12) : P (row-major 9), summed in the order of a 9-term dot product
DMK_HD __forceinline__ double sdot(const double (&S)[6], const double* __restrict__ P) {
  double s = S[0] * P[0];
  s += S[5] * P[1]; s += S[4] * P[2];
  s += S[5] * P[3]; s += S[1] * P[4]; s += S[3] * P[5];
  s += S[4] * P[6]; s += S[3] * P[7]; s += S[2] * P[8];
  return s;
}

// ------------------------------------------------------------------------------------------------ 9x9 solve
// dgesv semantics (LU, partial pivoting, one right-hand side), fully unrolled so that J stays in registers.
// Returns true if a pivot is exactly zero.
DMK_HD __forceinline__ bool solve9(double (&J)[81], double (&b)[9]) {
  bool sing = false;
#pragma unroll
  for (int k = 0; k < 9; k++) {
    double mx = fabs(J[9 * k + k]);
    int p = k;
#pragma unroll
    for (int i = k + 1; i < 9; i++) {
      double v = fabs(J[9 * i + k]);
      if (v > mx) { mx = v; p = i; }
    }
    sing = sing || (mx == 0.0);
#pragma unroll
    for (int i = k + 1; i < 9; i++) {
      if (p == i) {
#pragma unroll
        for (int j = k; j < 9; j++) { double t = J[9 * k + j]; J[9 * k + j] = J[9 * i + j]; J[9 * i + j] = t; }
        double t = b[k]; b[k] = b[i]; b[i] = t;
      }
    }
    const double rpiv = 1.0 / J[9 * k + k];        // one reciprocal per pivot instead of 8 - k quotients (<= 1 ulp apart)
    J[9 * k + k] = rpiv;
#pragma unroll
    for (int i = k + 1; i < 9; i++) {
      const double l = J[9 * i + k] * rpiv;
#pragma unroll
      for (int j = k + 1; j < 9; j++) J[9 * i + j] -= l * J[9 * k + j];
      b[i] -= l * b[k];
    }
  }
#pragma unroll
  for (int i = 8; i >= 0; i--) {
    double s = b[i];
#pragma unroll
    for (int j = i + 1; j < 9; j++) s -= J[9 * i + j] * b[j];
    b[i] = s * J[9 * i + i];                        // diagonal holds the reciprocal pivot
  }
  return sing;
}

// J += coef * P (x) (G Qm)   with Qm symmetric given as Voigt 6
DMK_HD __forceinline__ void jac_add_system(double (&J)[81], double coef, const double* __restrict__ P,
                                               const double (&G)[9], const double (&Q)[6]) {
  double w[9];
#pragma unroll
  for (int q = 0; q < 3; q++) {
    w[3 * q + 0] = G[3 * q] * Q[0] + G[3 * q + 1] * Q[5] + G[3 * q + 2] * Q[4];
    w[3 * q + 1] = G[3 * q] * Q[5] + G[3 * q + 1] * Q[1] + G[3 * q + 2] * Q[3];
    w[3 * q + 2] = G[3 * q] * Q[4] + G[3 * q + 1] * Q[3] + G[3 * q + 2] * Q[2];
  }
#pragma unroll
  for (int r = 0; r < 9; r++) {
    const double cr = coef * P[r];
#pragma unroll
    for (int cidx = 0; cidx < 9; cidx++) J[9 * r + cidx] += cr * w[cidx];
  }
}
// Q(beta) = sum_alpha pv(alpha) C66(alpha, beta)
DMK_HD __forceinline__ void q_from_c66(const double* __restrict__ C66, const double* __restrict__ pv, double (&Q)[6]) {
#pragma unroll
  for (int b = 0; b < 6; b++) {
    double s = 0.0;
#pragma unroll
    for (int a = 0; a < 6; a++) s += pv[a] * C66[6 * a + b];
    Q[b] = s;
  }
}

// ------------------------------------------------------------------------------------------------ per-thread point
template <int LAW, int NS, int ND>
struct DmkPt {
  const DmkDevPhase* __restrict__ ph;
  const double* C66;          // 36 doubles: phase table (global) or the per-call homogenised copy (dislotwin, local)
  double Fp[9], Lp[9], Fe[9], P[9];
  double S[6];
  // dynamically indexed arrays live OUTSIDE this struct (DmkPtStore): one variable-index access would otherwise keep the
  // whole aggregate -- and with it every tensor above -- in local memory instead of registers
  double* st;
  double* dep;
  double* C66h;
};
template <int LAW, int NS, int ND>
struct DmkPtStore {
  double st[NS > 0 ? NS : 1];
  double dep[ND > 0 ? ND : 1];
  double C66h[LAW == DMK_LAW_DISLOTWIN ? 36 : 1];
};
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void dmk_pt_bind(DmkPt<LAW, NS, ND>& p, DmkPtStore<LAW, NS, ND>& s) { p.st = s.st; p.dep = s.dep; p.C66h = s.C66h; }

// Hooke with Fi = I: S = C66 . Voigt_strain(1/2 (Fe^T Fe - I))      (phase_mechanical_elastic.f90:157-195)
DMK_HD __forceinline__ void hooke_S(const double* __restrict__ C66, const double (&Fe)[9], double (&S)[6]) {
  double ev[6];
  ev[0] = 0.5 * (Fe[0] * Fe[0] + Fe[3] * Fe[3] + Fe[6] * Fe[6] - 1.0);
  ev[1] = 0.5 * (Fe[1] * Fe[1] + Fe[4] * Fe[4] + Fe[7] * Fe[7] - 1.0);
  ev[2] = 0.5 * (Fe[2] * Fe[2] + Fe[5] * Fe[5] + Fe[8] * Fe[8] - 1.0);
  ev[3] = 2.0 * (0.5 * (Fe[2] * Fe[1] + Fe[5] * Fe[4] + Fe[8] * Fe[7]));
  ev[4] = 2.0 * (0.5 * (Fe[2] * Fe[0] + Fe[5] * Fe[3] + Fe[8] * Fe[6]));
  ev[5] = 2.0 * (0.5 * (Fe[0] * Fe[1] + Fe[3] * Fe[4] + Fe[6] * Fe[7]));
#pragma unroll
  for (int a = 0; a < 6; a++) {
    double s = 0.0;
#pragma unroll
    for (int b = 0; b < 6; b++) s += C66[6 * a + b] * ev[b];
    S[a] = s;
  }
}

// ================================================================================================ phenopowerlaw
// state: xi_sl[ns] | xi_tw[nt] | gamma_sl[ns] | gamma_tw[nt]
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ double pheno_sumF(const DmkPt<LAW, NS, ND>& p) {
  const dmk_phase& q = p.ph->p;
  const int ns = q.n_sl, nt = q.n_tw;
  double sumF = 0.0;
  for (int i = 0; i < nt; i++) sumF += p.st[2 * ns + nt + i] / q.gamma_char_tw[i];
  return sumF;
}
// one slip system: shear rate and d(shear rate)/d(tau); returns which non-Schmid branch drives it
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void pheno_slip_rate(const DmkPt<LAW, NS, ND>& p, const double (&S)[6], int i, double& dg,
                                                double& ddg) {
  const DmkDevPhase& d = *p.ph;
  const double tp = sdot(S, d.p.P_nS_pos[i]);
  const double tn = d.nonschmid ? sdot(S, d.p.P_nS_neg[i]) : -tp;
  const double tmax = fmax(tp, tn);
  const double n = d.p.n_sl_exp[i];
  const int ni = d.n_sl_int[i];
  const double ratio = tmax / p.st[i];
  dg = ((tp > tn) ? 1.0 : -1.0) * d.p.dot_gamma_0_sl[i] * (ni ? powi(ratio, ni) : pow(ratio, n));
  ddg = deq0(dg) ? 0.0 : dg * n / tmax;
}
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void pheno_twin_rate(const DmkPt<LAW, NS, ND>& p, const double (&S)[6], int i, double sumF,
                                                double& dg, double& ddg) {
  const dmk_phase& q = p.ph->p;
  const double tau = sdot(S, q.P_tw[i]);
  const double n = q.n_tw_exp[i];
  const int ni = p.ph->n_tw_int[i];
  const double ratio = tau / p.st[q.n_sl + i];
  dg = (tau > 0.0) ? (1.0 - sumF) * q.dot_gamma_0_tw[i] * (ni ? powi(ratio, ni) : pow(ratio, n)) : 0.0;
  ddg = deq0(dg) ? 0.0 : dg * n / tau;
}
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void pheno_Lp(const DmkPt<LAW, NS, ND>& p, const double (&S)[6], double (&Lp)[9],
                                         double* __restrict__ ddg_out) {
  const dmk_phase& q = p.ph->p;
  const int ns = q.n_sl, nt = q.n_tw;
#pragma unroll
  for (int a = 0; a < 9; a++) Lp[a] = 0.0;
  for (int i = 0; i < ns; i++) {
    double dg, ddg;
    pheno_slip_rate(p, S, i, dg, ddg);
#pragma unroll
    for (int a = 0; a < 9; a++) Lp[a] += dg * q.P_sl[i][a];
    ddg_out[i] = ddg;
    ddg_out[DMK_MAX_SYS + i] = dg;                         // sign source for the non-Schmid branch
  }
  if (nt > 0) {
    const double sumF = pheno_sumF(p);
    for (int i = 0; i < nt; i++) {
      double dg, ddg;
      pheno_twin_rate(p, S, i, sumF, dg, ddg);
#pragma unroll
      for (int a = 0; a < 9; a++) Lp[a] += dg * q.P_tw[i][a];
      ddg_out[ns + i] = ddg;
    }
  }
}
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void pheno_jac(const DmkPt<LAW, NS, ND>& p, double dt, const double (&G)[9],
                                          const double* __restrict__ ddg, double (&J)[81]) {
  const DmkDevPhase& d = *p.ph;
  const int ns = d.p.n_sl, nt = d.p.n_tw;
  for (int i = 0; i < ns; i++) {
    const double c = dt * ddg[i];
    if (c != 0.0) {
      const double* Qv = (ddg[DMK_MAX_SYS + i] > 0.0) ? d.Q_pos[i] : d.Q_neg[i];
      double Q[6];
#pragma unroll
      for (int b = 0; b < 6; b++) Q[b] = Qv[b];
      jac_add_system(J, c, d.p.P_sl[i], G, Q);
    }
  }
  for (int i = 0; i < nt; i++) {
    const double c = dt * ddg[ns + i];
    if (c != 0.0) {
      double Q[6];
#pragma unroll
      for (int b = 0; b < 6; b++) Q[b] = d.Q_tw[i][b];
      jac_add_system(J, c, d.p.P_tw[i], G, Q);
    }
  }
}
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void pheno_dotState(const DmkPt<LAW, NS, ND>& p, const double (&S)[6], double* __restrict__ dot) {
  const dmk_phase& q = p.ph->p;
  const int ns = q.n_sl, nt = q.n_tw;
  double* dot_xi_sl = dot; double* dot_xi_tw = dot + ns; double* dot_g_sl = dot + ns + nt; double* dot_g_tw = dot + 2 * ns + nt;
  const double sumF = pheno_sumF(p);
  for (int i = 0; i < ns; i++) { double dg, ddg; pheno_slip_rate(p, S, i, dg, ddg); dot_g_sl[i] = fabs(dg); }
  for (int i = 0; i < nt; i++) { double dg, ddg; pheno_twin_rate(p, S, i, sumF, dg, ddg); dot_g_tw[i] = dg; }
  double sum_gamma_sl = 0.0;
  for (int i = 0; i < ns; i++) sum_gamma_sl += p.st[ns + nt + i];
  for (int i = 0; i < ns; i++) {
    const double off = q.f_sat_sl_tw[i] * sqrt(sumF);
    const double x = 1.0 - p.st[i] / (q.xi_inf_sl[i] + off);
    const double left = fsign(pow(fabs(x), q.a_sl[i]), x);
    double hs = 0.0; for (int j = 0; j < ns; j++) hs += q.h_sl_sl[i][j] * dot_g_sl[j];
    double ht = 0.0; for (int j = 0; j < nt; j++) ht += q.h_sl_tw[i][j] * dot_g_tw[j];
    dot_xi_sl[i] = q.h_0_sl_sl[i] * (1.0 + q.c_1[i] * pow(sumF, q.c_2[i])) * left * hs + ht;
  }
  for (int i = 0; i < nt; i++) {
    double hs = 0.0; for (int j = 0; j < ns; j++) hs += q.h_tw_sl[i][j] * dot_g_sl[j];
    double ht = 0.0; for (int j = 0; j < nt; j++) ht += q.h_tw_tw[i][j] * dot_g_tw[j];
    dot_xi_tw[i] = q.h_0_tw_sl[i] * pow(sum_gamma_sl, q.c_3[i]) * hs + q.h_0_tw_tw[i] * pow(sumF, q.c_4[i]) * ht;
  }
}

// ================================================================================================ kinehardening
// phase_mechanical_plastic_kinehardening.f90: kinetics :433-472, LpAndItsTangent :258-295, dotState :301-341, deltaState :347-379.
// state: xi[ns] | chi[ns] | gamma[ns]   (integrated: n_dot = 3 ns)   | sgn_gamma[ns] | chi_flip[ns] | gamma_flip[ns] (deltaState only)
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void kh_slip_rate(const DmkPt<LAW, NS, ND>& p, const double (&S)[6], int i, double& dg, double& ddg) {
  const DmkDevPhase& d = *p.ph;
  const double chi = p.st[d.p.n_sl + i];
  const double ts = sdot(S, d.p.P_nS_pos[i]);
  const double tp = ts - chi;
  const double tn = (d.nonschmid ? sdot(S, d.p.P_nS_neg[i]) : -ts) + chi;
  const double tmax = fmax(tp, tn);
  const double n = d.p.n_sl_exp[i];
  const int ni = d.n_sl_int[i];
  const double ratio = tmax / p.st[i];
  dg = ((tp > tn) ? 1.0 : -1.0) * d.p.dot_gamma_0_sl[i] * (ni ? powi(ratio, ni) : pow(ratio, n));
  ddg = deq0(dg) ? 0.0 : dg * n / tmax;
}
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void kh_Lp(const DmkPt<LAW, NS, ND>& p, const double (&S)[6], double (&Lp)[9], double* __restrict__ ddg_out) {
  const dmk_phase& q = p.ph->p;
#pragma unroll
  for (int a = 0; a < 9; a++) Lp[a] = 0.0;
  for (int i = 0; i < q.n_sl; i++) {
    double dg, ddg;
    kh_slip_rate(p, S, i, dg, ddg);
#pragma unroll
    for (int a = 0; a < 9; a++) Lp[a] += dg * q.P_sl[i][a];
    ddg_out[i] = ddg;
    ddg_out[DMK_MAX_SYS + i] = dg;                           // its sign selects the non-Schmid projection of the tangent (:283-287)
  }
}
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void kh_jac(const DmkPt<LAW, NS, ND>& p, double dt, const double (&G)[9], const double* __restrict__ ddg,
                                   double (&J)[81]) {
  const DmkDevPhase& d = *p.ph;
  for (int i = 0; i < d.p.n_sl; i++) {
    const double c = dt * ddg[i];
    if (c != 0.0) {
      const double* Qv = (ddg[DMK_MAX_SYS + i] > 0.0) ? d.Q_pos[i] : d.Q_neg[i];
      double Q[6];
#pragma unroll
      for (int b = 0; b < 6; b++) Q[b] = Qv[b];
      jac_add_system(J, c, d.p.P_sl[i], G, Q);
    }
  }
}
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void kh_dotState(const DmkPt<LAW, NS, ND>& p, const double (&S)[6], double* __restrict__ dot) {
  const dmk_phase& q = p.ph->p;
  const int n = q.n_sl;
  const double *gamma = p.st + 2 * n, *sgn = p.st + 3 * n, *chi_flip = p.st + 4 * n, *gamma_flip = p.st + 5 * n;
  double *dot_xi = dot, *dot_chi = dot + n, *dot_gamma = dot + 2 * n;
  for (int i = 0; i < n; i++) { double dg, ddg; kh_slip_rate(p, S, i, dg, ddg); dot_gamma[i] = fabs(dg); }
  double sumGamma = 0.0;
  for (int i = 0; i < n; i++) sumGamma += gamma[i];
  for (int i = 0; i < n; i++) {
    double hs = 0.0;
    for (int j = 0; j < n; j++) hs += q.h_sl_sl[i][j] * dot_gamma[j];
    const double x = sumGamma * q.kh_h_0_xi[i] / q.xi_inf_sl[i];
    dot_xi[i] = hs * (q.kh_h_inf_xi[i] + (q.kh_h_0_xi[i] - q.kh_h_inf_xi[i] * (1.0 - x)) * exp(-x));
    const double y = (gamma[i] - gamma_flip[i]) * q.kh_h_0_chi[i] / (q.kh_chi_inf[i] + chi_flip[i]);
    dot_chi[i] = sgn[i] * dot_gamma[i] * (q.kh_h_inf_chi[i] + (q.kh_h_0_chi[i] - q.kh_h_inf_chi[i] * (1.0 - y)) * exp(-y));
  }
}
// plastic_deltaState (phase_mechanical_plastic.f90:372-413) -> plastic_kinehardening_deltaState: where the shear direction of a
// system flips, its sense and the back stress / shear reached at the flip are recorded.  Returns broken (a NaN jump).
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ bool kh_deltaState(DmkPt<LAW, NS, ND>& p) {
  const int n = p.ph->p.n_sl;
  double *chi = p.st + n, *gamma = p.st + 2 * n, *sgn = p.st + 3 * n, *chi_flip = p.st + 4 * n, *gamma_flip = p.st + 5 * n;
  bool nan = false;
  double d0[DMK_MAX_STATE / 6], d1[DMK_MAX_STATE / 6], d2[DMK_MAX_STATE / 6];
  for (int i = 0; i < n; i++) {
    double dg, ddg;
    kh_slip_rate(p, p.S, i, dg, ddg);
    const double s = (fabs(dg) <= 1.0e-10) ? sgn[i] : ((dg >= 0.0 && !signbit(dg)) ? 1.0 : -1.0);     // merge(state, sign(1, dg), dEq0(dg, 1e-10))
    if (!(fabs(s - sgn[i]) <= 0.1)) { d0[i] = s - sgn[i]; d1[i] = fabs(chi[i]) - chi_flip[i]; d2[i] = gamma[i] - gamma_flip[i]; }
    else { d0[i] = 0.0; d1[i] = 0.0; d2[i] = 0.0; }
    nan = nan || isnan(d0[i]) || isnan(d1[i]) || isnan(d2[i]);
  }
  if (nan) return true;
  for (int i = 0; i < n; i++) { sgn[i] += d0[i]; chi_flip[i] += d1[i]; gamma_flip[i] += d2[i]; }
  return false;
}

// ================================================================================================ isotropic
DMK_HD __forceinline__ void deviator(const double (&S)[6], double (&dev)[9], double& sq) {
  const double tr = (S[0] + S[1] + S[2]) / 3.0;
  dev[0] = S[0] - tr; dev[1] = S[5]; dev[2] = S[4];
  dev[3] = S[5]; dev[4] = S[1] - tr; dev[5] = S[3];
  dev[6] = S[4]; dev[7] = S[3]; dev[8] = S[2] - tr;
  sq = 0.0;
#pragma unroll
  for (int a = 0; a < 9; a++) sq += dev[a] * dev[a];
}
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void iso_Lp(const DmkPt<LAW, NS, ND>& p, const double (&S)[6], double (&Lp)[9], double& dgam_over_nrm,
                                       double (&dev)[9], double& sq) {
  const dmk_phase& q = p.ph->p;
  deviator(S, dev, sq);
  const double nrm = sqrt(sq);
  if (nrm > 0.0) {
    const double ratio = sqrt(1.5) * nrm / (q.iso_M * p.st[0]);
    const double dgam = q.iso_dot_gamma_0 * (p.ph->iso_n_int ? powi(ratio, p.ph->iso_n_int) : pow(ratio, q.iso_n));
#pragma unroll
    for (int a = 0; a < 9; a++) Lp[a] = dgam * dev[a] / nrm;
    dgam_over_nrm = dgam / nrm;
  } else {
#pragma unroll
    for (int a = 0; a < 9; a++) Lp[a] = 0.0;
    dgam_over_nrm = 0.0;
  }
}
// J += dt * dLp_dMp . K,  dLp_dMp = c [ (n-1)/sq dev(x)dev + I9 - 1/3 1(x)1 ],  K(ij,qj') = sum_m G(q,m) C(i,j,m,j')
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void iso_jac(const DmkPt<LAW, NS, ND>& p, double dt, const double (&G)[9], double c,
                                        const double (&dev)[9], double sq, double (&J)[81]) {
  if (c == 0.0) return;
  const double* __restrict__ C66 = p.C66;
  const double n = p.ph->p.iso_n;
  const double pvd[6] = {dev[0], dev[4], dev[8], dev[5] + dev[7], dev[2] + dev[6], dev[1] + dev[3]};
  const double pvi[6] = {1.0, 1.0, 1.0, 0.0, 0.0, 0.0};
  double Qd[6], Qi[6];
  q_from_c66(C66, pvd, Qd);
  q_from_c66(C66, pvi, Qi);
  double devs[9], eye[9];
#pragma unroll
  for (int a = 0; a < 9; a++) { devs[a] = dev[a]; eye[a] = (a % 4 == 0) ? 1.0 : 0.0; }
  jac_add_system(J, dt * c * (n - 1.0) / sq, devs, G, Qd);
  jac_add_system(J, -dt * c / 3.0, eye, G, Qi);
  // + dt c K
  const int V[9] = {0, 5, 4, 5, 1, 3, 4, 3, 2};      // Voigt index of (i,j)
#pragma unroll
  for (int ij = 0; ij < 9; ij++)
#pragma unroll
    for (int q = 0; q < 3; q++)
#pragma unroll
      for (int j2 = 0; j2 < 3; j2++) {
        double s = 0.0;
#pragma unroll
        for (int m = 0; m < 3; m++) s += G[3 * q + m] * C66[6 * V[ij] + V[3 * m + j2]];
        J[9 * ij + 3 * q + j2] += dt * c * s;
      }
}
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void iso_dotState(const DmkPt<LAW, NS, ND>& p, const double (&S)[6], double* __restrict__ dot) {
  const dmk_phase& q = p.ph->p;
  const double xi = p.st[0];
  double dev[9], sq;
  deviator(S, dev, sq);
  const double nrm = sqrt(sq);
  const double ratio = sqrt(1.5) * nrm / (q.iso_M * xi);
  const double dgam = q.iso_dot_gamma_0 * (p.ph->iso_n_int ? powi(ratio, p.ph->iso_n_int) : pow(ratio, q.iso_n));
  if (dgam > 1e-12) {
    double xi_inf_star;
    if (deq0(q.iso_c_1)) xi_inf_star = q.iso_xi_inf;
    else xi_inf_star = q.iso_xi_inf + pow(asinh(pow(dgam / q.iso_c_1, 1.0 / q.iso_c_2)), 1.0 / q.iso_c_3) / q.iso_c_4
                                      * pow(dgam / q.iso_dot_gamma_0, 1.0 / q.iso_n);
    const double x = 1.0 - xi / xi_inf_star;
    dot[0] = dgam * (q.iso_h_0 + q.iso_h_ln * log(dgam)) * fsign(pow(fabs(x), q.iso_a) * q.iso_h, x);
  } else dot[0] = 0.0;
}

// ================================================================================================ dislotwin
// state: rho_mob[ns] | rho_dip[ns] | gamma_sl[ns] | f_tw[nt] ;  dep: Lambda_sl[ns] | Lambda_tw[nt] | tau_pass[ns]
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ double dt_f_matrix(const DmkPt<LAW, NS, ND>& p) {
  const dmk_phase& q = p.ph->p;
  double s = 0.0;
  for (int i = 0; i < q.n_tw; i++) s += p.st[3 * q.n_sl + i];
  return 1.0 - s;
}
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void dt_homogenized_C66(DmkPt<LAW, NS, ND>& p) {
  if constexpr (LAW == DMK_LAW_DISLOTWIN) {
    const dmk_phase& q = p.ph->p;
    const double fm = dt_f_matrix(p);
    for (int k = 0; k < 36; k++) p.C66h[k] = fm * q.C66[k];
    for (int i = 0; i < q.n_tw; i++) {
      const double f = p.st[3 * q.n_sl + i];
      for (int k = 0; k < 36; k++) p.C66h[k] += f * q.dt_C66_tw[i][k];
    }
    p.C66 = p.C66h;
  }
}
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void dt_dependentState(DmkPt<LAW, NS, ND>& p) {
  const dmk_phase& q = p.ph->p;
  const int ns = q.n_sl, nt = q.n_tw;
  double* Lambda_sl = p.dep; double* Lambda_tw = p.dep + ns; double* tau_pass = p.dep + ns + nt;
  double sumf = 0.0;
  for (int i = 0; i < nt; i++) sumf += p.st[3 * ns + i];
  for (int i = 0; i < ns; i++) {
    double s = 0.0; for (int j = 0; j < ns; j++) s += q.dt_forest[i][j] * (p.st[j] + p.st[ns + j]);
    double inv = sqrt(s) / q.dt_i_sl[i];
    if (nt > 0) { double h = 0.0; for (int j = 0; j < nt; j++) h += q.h_sl_tw[i][j] * (p.st[3 * ns + j] / q.dt_t_tw[j]); inv += h / (1.0 - sumf); }
    Lambda_sl[i] = q.dt_D / (1.0 + q.dt_D * inv);
    double hp = 0.0; for (int j = 0; j < ns; j++) hp += q.h_sl_sl[i][j] * (p.st[j] + p.st[ns + j]);
    tau_pass[i] = q.dt_mu * q.dt_b_sl[i] * sqrt(hp);
  }
  for (int i = 0; i < nt; i++) {
    double h = 0.0; for (int j = 0; j < nt; j++) h += q.h_tw_tw[i][j] * (p.st[3 * ns + j] / q.dt_t_tw[j]);
    const double inv = h / (1.0 - sumf);
    Lambda_tw[i] = q.dt_i_tw * q.dt_D / (1.0 + q.dt_D * inv);
  }
}
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void dt_slip_rate(const DmkPt<LAW, NS, ND>& p, const double (&S)[6], int i, bool want_d, double& dg,
                                             double& ddg) {
  const dmk_phase& q = p.ph->p;
  const double tau = sdot(S, q.P_sl[i]);
  const double tau_eff = fabs(tau) - p.dep[q.n_sl + q.n_tw + i];
  dg = 0.0; ddg = 0.0;
  if (tau_eff > DMK_TOL_MATH_CHECK) {
    const double sr = tau_eff / q.dt_tau_0[i], srp = pow(sr, q.dt_p[i]);
    const double QkT = q.dt_Q_sl[i] / q.dt_kB_T;
    const double om = 1.0 - srp, omq = pow(om, q.dt_q[i]);
    const double vw = exp(QkT * omq) / q.dt_v_0[i];
    const double vr = q.dt_B[i] / (tau_eff * q.dt_b_sl[i]);
    dg = fsign(p.st[i] * q.dt_b_sl[i] / (vw + vr), tau);
    if (want_d) {
      // sr^(p-1) = sr^p / sr and (1 - sr^p)^(q-1) = (1 - sr^p)^q / (1 - sr^p): two of the four FP64 pow() calls (each a few
      // hundred instructions) of a slip system, a few 1e-16 from the reference's separate pow() calls
      const double dvw = -1.0 * vw * q.dt_p[i] * q.dt_q[i] * QkT * (srp / sr) * (omq / om) / q.dt_tau_0[i];
      const double dvr = -1.0 * vr / tau_eff;
      const double dV = -1.0 * (dvw + dvr) / ((vw + vr) * (vw + vr));
      ddg = dV * p.st[i] * q.dt_b_sl[i];
    }
  }
}
// dg_sl: slip rates handed to the nucleation term (signed in Lp, absolute in dotState -- reference quirk,
// phase_mechanical_plastic_dislotwin.f90:563 vs :660)
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void dt_twin_rate(const DmkPt<LAW, NS, ND>& p, const double (&S)[6], int i, const double* __restrict__ dg_sl,
                                             bool want_d, double& dg, double& ddg) {
  const dmk_phase& q = p.ph->p;
  const int ns = q.n_sl;
  const double mu = q.dt_mu, nu = q.dt_nu, Gsf = q.dt_Gamma_sf;
  const double tau_hat = 3.0 * q.dt_b_tw[0] * mu / q.dt_L_tw + Gsf / (3.0 * q.dt_b_tw[0]);
  const double x0 = mu * q.dt_b_sl[0] * q.dt_b_sl[0] * (2.0 + nu) / (Gsf * 8.0 * DMK_PI * (1.0 - nu));
  const double tau_r = mu * q.dt_b_sl[0] / (2.0 * DMK_PI) * (1.0 / (x0 + q.dt_x_c) + cos(DMK_PI / 3.0) / x0);
  const double tau = sdot(S, q.P_tw[i]);
  dg = 0.0; ddg = 0.0;
  if (tau > DMK_TOL_MATH_CHECK && tau < tau_r) {
    const double r = q.dt_r[i];
    const double pw = pow(tau_hat / tau, r);
    const double Pn = exp(-pw);
    const double dP = r * pw / tau * Pn;
    const int s1 = q.dt_pair[i][0], s2 = q.dt_pair[i][1];
    const double dotN0 = (dg_sl[s2] * (p.st[s1] + p.st[ns + s1]) + dg_sl[s1] * (p.st[s2] + p.st[ns + s2])) / (q.dt_L_tw * 3.0);
    const double Pncs = 1.0 - exp(-q.dt_V_cs / q.dt_kB_T * (tau_r - tau));
    const double dPncs = q.dt_V_cs / q.dt_kB_T * (Pncs - 1.0);
    const double L = p.dep[ns + i];
    const double V = DMK_PI / 4.0 * L * L * q.dt_t_tw[i];
    dg = V * dotN0 * Pncs * Pn * q.gamma_char_tw[i];
    if (want_d) ddg = V * dotN0 * (Pn * dPncs + Pncs * dP) * q.gamma_char_tw[i];
  }
}
// Lp and per-system tangent factors (ddg[a], already scaled by f_matrix); dgs = scratch [DMK_MAX_SL]
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void dt_Lp(const DmkPt<LAW, NS, ND>& p, const double (&S)[6], double (&Lp)[9], double* __restrict__ ddg_out) {
  const dmk_phase& q = p.ph->p;
  const int ns = q.n_sl, nt = q.n_tw;
  double* dgs = ddg_out + DMK_MAX_SYS;
  const double fm = dt_f_matrix(p);
#pragma unroll
  for (int a = 0; a < 9; a++) Lp[a] = 0.0;
  for (int i = 0; i < ns; i++) {
    double dg, ddg;
    dt_slip_rate(p, S, i, true, dg, ddg);
#pragma unroll
    for (int a = 0; a < 9; a++) Lp[a] += dg * q.P_sl[i][a];
    dgs[i] = dg; ddg_out[i] = ddg * fm;
  }
  for (int i = 0; i < nt; i++) {
    double dg, ddg;
    dt_twin_rate(p, S, i, dgs, true, dg, ddg);
#pragma unroll
    for (int a = 0; a < 9; a++) Lp[a] += dg * q.P_tw[i][a];
    ddg_out[ns + i] = ddg * fm;
  }
#pragma unroll
  for (int a = 0; a < 9; a++) Lp[a] *= fm;
}
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void dt_jac(const DmkPt<LAW, NS, ND>& p, double dt, const double (&G)[9], const double* __restrict__ ddg,
                                       double (&J)[81]) {
  const DmkDevPhase& d = *p.ph;
  const int ns = d.p.n_sl, nt = d.p.n_tw;
  for (int i = 0; i < ns + nt; i++) {
    const double c = dt * ddg[i];
    if (c != 0.0) {
      double Q[6];
      q_from_c66(p.C66, (i < ns) ? d.pv_pos[i] : d.pv_tw[i - ns], Q);
      jac_add_system(J, c, (i < ns) ? d.p.P_sl[i] : d.p.P_tw[i - ns], G, Q);
    }
  }
}
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void dt_dotState(const DmkPt<LAW, NS, ND>& p, const double (&S)[6], double* __restrict__ dot) {
  const dmk_phase& q = p.ph->p;
  const int ns = q.n_sl, nt = q.n_tw;
  double* dot_mob = dot; double* dot_dip = dot + ns; double* adg = dot + 2 * ns; double* dot_f_tw = dot + 3 * ns;
  const double mu = q.dt_mu, nu = q.dt_nu, kT = q.dt_kB_T;
  const double fm = dt_f_matrix(p);
  for (int i = 0; i < ns; i++) { double dg, ddg; dt_slip_rate(p, S, i, false, dg, ddg); adg[i] = fabs(dg); }
  for (int i = 0; i < ns; i++) {
    const double rho_mob = p.st[i], rho_dip = p.st[ns + i];
    const double tau = sdot(S, q.P_sl[i]);
    double form = 0.0, climb = 0.0;
    if (!(deq0(tau) || q.dt_omit_dipoles)) {
      double d_hat = 3.0 * mu * q.dt_b_sl[i] / (16.0 * DMK_PI * fabs(tau));
      d_hat = fmin(d_hat, p.dep[i]);
      d_hat = fmax(d_hat, q.dt_d_caron[i]);
      form = 2.0 * (d_hat - q.dt_d_caron[i]) / q.dt_b_sl[i] * rho_mob * adg[i];
      if (!deq(d_hat, q.dt_d_caron[i])) {
        const double* n = q.dt_n0_sl[i];
        // sigma_cl = n . Mp . n with Mp = S symmetric
        const double Mn0 = S[0] * n[0] + S[5] * n[1] + S[4] * n[2];
        const double Mn1 = S[5] * n[0] + S[1] * n[1] + S[3] * n[2];
        const double Mn2 = S[4] * n[0] + S[3] * n[1] + S[2] * n[2];
        const double sigma_cl = n[0] * Mn0 + n[1] * Mn1 + n[2] * Mn2;
        const double b_d = q.dt_extended ? 24.0 * DMK_PI * (1.0 - nu) / (2.0 + nu) * q.dt_Gamma_sf / (mu * q.dt_b_sl[i]) : 1.0;
        const double b3 = q.dt_b_sl[i] * q.dt_b_sl[i] * q.dt_b_sl[i];
        const double v_cl = 2.0 * q.dt_omega * b_d * b_d * exp(-q.dt_Q_cl / kT) * (exp(fabs(sigma_cl) * b3 / kT) - 1.0);
        climb = 4.0 * v_cl * rho_dip / (d_hat - q.dt_d_caron[i]);
      }
    }
    dot_mob[i] = adg[i] / (q.dt_b_sl[i] * p.dep[i]) - form - 2.0 * q.dt_d_caron[i] / q.dt_b_sl[i] * rho_mob * adg[i];
    dot_dip[i] = form - 2.0 * q.dt_d_caron[i] / q.dt_b_sl[i] * rho_dip * adg[i] - climb;
  }
  for (int i = 0; i < nt; i++) {
    double dg, ddg;
    dt_twin_rate(p, S, i, adg, false, dg, ddg);
    dot_f_tw[i] = fm * dg / q.gamma_char_tw[i];
  }
}

// ================================================================================================ dislotungsten
// phase_mechanical_plastic_dislotungsten.f90: kinetics :444-528, LpAndItsTangent :270-310, dotState :316-378, dependentState :384-405.
// state: rho_mob[ns] | rho_dip[ns] | gamma_sl[ns] ;  dep: Lambda_sl[ns] | tau_pass[ns]
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void tu_dependentState(DmkPt<LAW, NS, ND>& p) {
  const dmk_phase& q = p.ph->p;
  const int ns = q.n_sl;
  double* Lambda_sl = p.dep; double* tau_pass = p.dep + ns;
  for (int i = 0; i < ns; i++) {
    double hp = 0.0, s = 0.0;
    for (int j = 0; j < ns; j++) { const double rho = p.st[j] + p.st[ns + j]; hp += q.h_sl_sl[i][j] * rho; s += q.dt_forest[i][j] * rho; }
    tau_pass[i] = q.dt_mu * q.dt_b_sl[i] * sqrt(hp);
    Lambda_sl[i] = 1.0 / (1.0 / q.dt_D + sqrt(s) / q.dt_i_sl[i]);
  }
}
// one slip system: shear rate, d(shear rate)/d(tau) (want_d) and the effective resolved stress
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void tu_slip_rate(const DmkPt<LAW, NS, ND>& p, const double (&S)[6], int i, bool want_d, double& dg, double& ddg,
                                         double& tau_eff) {
  const DmkDevPhase& d = *p.ph;
  const dmk_phase& q = d.p;
  const double tp = sdot(S, q.P_nS_pos[i]);
  const double tn = d.nonschmid ? sdot(S, q.P_nS_neg[i]) : -tp;
  tau_eff = fmax(fmax(tp, tn) - p.dep[q.n_sl + i], 0.0);               // math_clip(left = 0)
  dg = 0.0; ddg = 0.0;
  if (tau_eff > DMK_TOL_MATH_CHECK) {
    const double sr = tau_eff / q.dt_tau_0[i], srp = pow(sr, q.dt_p[i]);
    const double BR = q.dt_Q_sl[i] / q.dt_kB_T, effL = p.dep[i] - q.tu_w[i];
    const double om = 1.0 - srp, omq = pow(om, q.dt_q[i]);
    const double t_n = q.dt_b_sl[i] * exp(BR * omq) / (q.tu_omega[i] * effL);
    const double t_k = effL * q.dt_B[i] / (2.0 * q.dt_b_sl[i] * tau_eff);
    dg = p.st[i] * q.dt_b_sl[i] * q.tu_h[i] / (t_n + t_k) * ((tp > tn) ? 1.0 : -1.0);
    if (want_d) {
      // (1 - sr^p)^(q-1) = (1 - sr^p)^q / (1 - sr^p) and sr^(p-1) = sr^p / sr, as in dt_slip_rate
      const double dtn = -1.0 * t_n * BR * q.dt_p[i] * q.dt_q[i] * (omq / om) * (srp / sr) / q.dt_tau_0[i];
      const double dtk = -1.0 * t_k / tau_eff;
      ddg = -1.0 * dg * (dtn + dtk) / (t_n + t_k);
    }
  }
}
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void tu_Lp(const DmkPt<LAW, NS, ND>& p, const double (&S)[6], double (&Lp)[9], double* __restrict__ ddg_out) {
  const dmk_phase& q = p.ph->p;
#pragma unroll
  for (int a = 0; a < 9; a++) Lp[a] = 0.0;
  for (int i = 0; i < q.n_sl; i++) {
    double dg, ddg, te;
    tu_slip_rate(p, S, i, true, dg, ddg, te);
#pragma unroll
    for (int a = 0; a < 9; a++) Lp[a] += dg * q.P_sl[i][a];
    ddg_out[i] = ddg;
    ddg_out[DMK_MAX_SYS + i] = dg;                           // its sign selects the non-Schmid projection of the tangent (:297)
  }
}
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void tu_dotState(const DmkPt<LAW, NS, ND>& p, const double (&S)[6], double* __restrict__ dot) {
  const dmk_phase& q = p.ph->p;
  const int ns = q.n_sl;
  const double mu = q.dt_mu, nu = q.dt_nu, kT = q.dt_kB_T;
  for (int i = 0; i < ns; i++) {
    double dg, ddg, tau_eff;
    tu_slip_rate(p, S, i, false, dg, ddg, tau_eff);
    const double adg = fabs(dg);
    const double b = q.dt_b_sl[i], dc = q.dt_d_caron[i], Lambda = p.dep[i];
    double form = 0.0, climb = 0.0;
    if (!deq0(adg)) {
      double d_hat = mu * b / (8.0 * DMK_PI * (1.0 - nu) * tau_eff);
      d_hat = fmin(fmax(d_hat, dc), Lambda);                 // math_clip(left = d_caron, right = Lambda_sl)
      form = q.dt_omit_dipoles ? 0.0 : adg * 2.0 * (d_hat - dc) / b * p.st[i];
      const double v_cl = (3.0 * mu * q.tu_D_0 * exp(-q.dt_Q_cl / kT) * q.tu_f_at[i] / (2.0 * DMK_PI * kT)) * (1.0 / (d_hat + dc));
      climb = (4.0 * v_cl * p.st[ns + i]) / (d_hat - dc);
    }
    dot[i] = adg / (b * Lambda) - form - adg * 2.0 * dc / b * p.st[i];
    dot[ns + i] = form - climb - adg * 2.0 * dc / b * p.st[ns + i];
    dot[2 * ns + i] = adg;
  }
}

// ================================================================================================ dispatch
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void dot_state(const DmkPt<LAW, NS, ND>& p, double* __restrict__ dot) {
  if constexpr (LAW == DMK_LAW_ISOTROPIC) iso_dotState(p, p.S, dot);
  else if constexpr (LAW == DMK_LAW_PHENOPOWERLAW) pheno_dotState(p, p.S, dot);
  else if constexpr (LAW == DMK_LAW_DISLOTWIN) dt_dotState(p, p.S, dot);
  else if constexpr (LAW == DMK_LAW_KINEHARDENING) kh_dotState(p, p.S, dot);
  else if constexpr (LAW == DMK_LAW_DISLOTUNGSTEN) tu_dotState(p, p.S, dot);
}
// plastic_dependentState (phase_mechanical_plastic.f90): derived quantities of the laws that have them, before every stress integration
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void dependent_state(DmkPt<LAW, NS, ND>& p) {
  if constexpr (LAW == DMK_LAW_DISLOTWIN) { dt_dependentState(p); dt_homogenized_C66(p); }
  else if constexpr (LAW == DMK_LAW_DISLOTUNGSTEN) tu_dependentState(p);
}
// instantaneous state jump after a converged state integration (phase_mechanical_plastic.f90:372-413): kinehardening only; broken?
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ bool delta_state(DmkPt<LAW, NS, ND>& p) {
  if constexpr (LAW == DMK_LAW_KINEHARDENING) return kh_deltaState(p);
  else return false;
}
// entries of the state vector the integrators move (sizeDotState); the rest changes by delta_state only
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ int n_dot_state(const DmkPt<LAW, NS, ND>& p) {
  if constexpr (LAW == DMK_LAW_KINEHARDENING) return 3 * p.ph->p.n_sl;
  else return p.ph->p.n_state;
}

// ------------------------------------------------------------------------------------------------ integrateStress
// returns true when broken.  On success p.{P,S,Lp,Fp,Fe} hold the new values.
template <int LAW, int NS, int ND>
DMK_HD __noinline__ bool integrate_stress(DmkPt<LAW, NS, ND>& p, const DmkNumDev& num, const double (&F)[9],
                                              const double (&Fp0)[9], double dt, DmkCnt& cnt) {
  cnt.stress_calls++;
  dependent_state(p);
  double Lpg[9], invFp[9], A[9];
#pragma unroll
  for (int a = 0; a < 9; a++) Lpg[a] = p.Lp[a];
  if (invert33(Fp0, invFp)) return true;
  mm3(F, invFp, A);
  cnt.Li_iter++;                                    // the Li loop runs exactly once with Li = 0
  double Lpg_old[9], delta[9], B[9], Fe[9], S[6];
#pragma unroll
  for (int a = 0; a < 9; a++) { Lpg_old[a] = Lpg[a]; delta[a] = 0.0; }
  double res_old_norm = 0.0, step = 1.0;
  int jaco = 0, nLp = 0;
  double ddg[2 * DMK_MAX_SYS];
  for (;;) {
    nLp++; cnt.Lp_iter++;
    if (nLp > num.N_iter_Lp_max) return true;
#pragma unroll
    for (int a = 0; a < 9; a++) B[a] = ((a % 4 == 0) ? 1.0 : 0.0) - dt * Lpg[a];
    mm3(A, B, Fe);
    hooke_S(p.C66, Fe, S);
    double Lp_c[9], res[9];
    double iso_c = 0.0, iso_sq = 1.0, iso_dev[9];
    if constexpr (LAW == DMK_LAW_NONE) {
#pragma unroll
      for (int a = 0; a < 9; a++) Lp_c[a] = 0.0;
    } else if constexpr (LAW == DMK_LAW_ISOTROPIC) iso_Lp(p, S, Lp_c, iso_c, iso_dev, iso_sq);
    else if constexpr (LAW == DMK_LAW_PHENOPOWERLAW) pheno_Lp(p, S, Lp_c, ddg);
    else if constexpr (LAW == DMK_LAW_KINEHARDENING) kh_Lp(p, S, Lp_c, ddg);
    else if constexpr (LAW == DMK_LAW_DISLOTUNGSTEN) tu_Lp(p, S, Lp_c, ddg);
    else dt_Lp(p, S, Lp_c, ddg);
    const double atol = fmax(num.eps_rel_Lp * fmax(norm9(Lpg), norm9(Lp_c)), num.eps_abs_Lp);
#pragma unroll
    for (int a = 0; a < 9; a++) res[a] = Lpg[a] - Lp_c[a];
    const double rn = norm9(res);
    if (any_nan9(res)) return true;
    else if (rn < atol) break;
    else if (nLp == 1 || rn < res_old_norm) {
      res_old_norm = rn;
#pragma unroll
      for (int a = 0; a < 9; a++) Lpg_old[a] = Lpg[a];
      step = 1.0;
    } else {
      step = num.r_linesearch_Lp * step;
#pragma unroll
      for (int a = 0; a < 9; a++) Lpg[a] = Lpg_old[a] + delta[a] * step;
      continue;
    }
    if (jaco % num.f_update_jacobi_Lp == 0) {
      jaco++;
      double G[9], J[81];
      mm3_tn(A, Fe, G);
#pragma unroll
      for (int a = 0; a < 81; a++) J[a] = (a % 10 == 0) ? 1.0 : 0.0;
      if constexpr (LAW == DMK_LAW_ISOTROPIC) iso_jac(p, dt, G, iso_c, iso_dev, iso_sq, J);
      else if constexpr (LAW == DMK_LAW_PHENOPOWERLAW) pheno_jac(p, dt, G, ddg, J);
      else if constexpr (LAW == DMK_LAW_DISLOTWIN) dt_jac(p, dt, G, ddg, J);
      else if constexpr (LAW == DMK_LAW_KINEHARDENING || LAW == DMK_LAW_DISLOTUNGSTEN) kh_jac(p, dt, G, ddg, J);   // same rank-one form
      if (solve9(J, res)) return true;
#pragma unroll
      for (int a = 0; a < 9; a++) delta[a] = -res[a];
    }
#pragma unroll
    for (int a = 0; a < 9; a++) Lpg[a] += delta[a] * step;
  }
  // Fp update, P, Fe (phase_mechanical.f90:586-604)
  double invFp_new[9], Fp_new[9], t1[9], Sm[9], t2[9];
  mm3(invFp, B, invFp_new);
  if (invert33(invFp_new, Fp_new)) return true;
  mm3(F, invFp_new, t1);
  Sm[0] = S[0]; Sm[1] = S[5]; Sm[2] = S[4]; Sm[3] = S[5]; Sm[4] = S[1]; Sm[5] = S[3]; Sm[6] = S[4]; Sm[7] = S[3]; Sm[8] = S[2];
  mm3_nt(Sm, invFp_new, t2);
  mm3(t1, t2, p.P);
  const double sc = pow(det3(Fp_new), 1.0 / 3.0);
#pragma unroll
  for (int a = 0; a < 9; a++) { p.Fp[a] = Fp_new[a] / sc; p.Lp[a] = Lpg[a]; p.Fe[a] = t1[a]; }
#pragma unroll
  for (int a = 0; a < 6; a++) p.S[a] = S[a];
  return false;
}

// ------------------------------------------------------------------------------------------------ integrateStateFPI
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ bool integrate_state_FPI(DmkPt<LAW, NS, ND>& p, const DmkNumDev& num, const double (&F)[9],
                                                    const double (&Fp0)[9], const double* __restrict__ state0, double dt, DmkCnt& cnt) {
  const int n = n_dot_state(p);                       // 1:sizeDotState
  double dot[NS > 0 ? NS : 1], last1[NS > 0 ? NS : 1], last2[NS > 0 ? NS : 1];
  bool broken = true;
  for (int i = 0; i < n; i++) { dot[i] = 0.0; last1[i] = 0.0; last2[i] = 0.0; }
  dot_state(p, dot);
  for (int i = 0; i < n; i++) if (isnan(dot[i])) return true;
  for (int i = 0; i < n; i++) p.st[i] = state0[i] + dot[i] * dt;
  for (int it = 1; it <= num.N_iter_state_max; it++) {
    cnt.state_iter++;
    for (int i = 0; i < n; i++) { last2[i] = (it > 1) ? last1[i] : 0.0; last1[i] = dot[i]; }
    broken = integrate_stress(p, num, F, Fp0, dt, cnt);
    if (broken) break;
    dot_state(p, dot);
    bool nan = false;
    for (int i = 0; i < n; i++) nan = nan || isnan(dot[i]);
    if (nan) break;                                   // broken stays false here, as in the reference (:660)
    double d12 = 0.0, d22 = 0.0, d01 = 0.0;           // damper (:677-690)
    for (int i = 0; i < n; i++) {
      d12 += (dot[i] - last1[i]) * (last1[i] - last2[i]);
      d22 += (last1[i] - last2[i]) * (last1[i] - last2[i]);
      d01 += dot[i] * last1[i];
    }
    double zeta = 1.0;
    if (fmin(d01, d12) < 0.0 && d22 > 0.0) zeta = 0.75 + 0.25 * tanh(2.0 + 4.0 * d12 / d22);
    bool conv = true;
    for (int i = 0; i < n; i++) {
      dot[i] = dot[i] * zeta + last1[i] * (1.0 - zeta);
      const double r = p.st[i] - state0[i] - dot[i] * dt;
      p.st[i] -= r;
      if (!(fabs(r) <= fmax(p.ph->p.atol[i], num.eps_rel_state * fabs(p.st[i])))) conv = false;
    }
    if (conv) { broken = delta_state(p); break; }     // :661-664
  }
  return broken;
}

// ------------------------------------------------------------------------------------------------ the other state integrators
// integrateStateEuler / AdaptiveEuler / RK4 / RKCK45 (phase_mechanical.f90:699-932; numerics `integrator_state`).  Kept out
// of line: FPI is the default and the only integrator of the benchmark configurations.  Returns broken.
#define DMK_INTEG_FPI 0
#define DMK_INTEG_EULER 1
#define DMK_INTEG_ADAPTIVE_EULER 2
#define DMK_INTEG_RK4 3
#define DMK_INTEG_RKCK45 4
template <int LAW, int NS, int ND>
DMK_HD __noinline__ bool integrate_state_other(DmkPt<LAW, NS, ND>& p, const DmkNumDev& num, const double (&F_0)[9], const double (&F)[9],
                                               const double (&Fp0)[9], const double* __restrict__ state0, double dt, DmkCnt& cnt) {
  constexpr int NSX = NS > 0 ? NS : 1;
  const int n = n_dot_state(p);
  double dot[NSX];
  for (int i = 0; i < n; i++) dot[i] = 0.0;
  dot_state(p, dot);
  for (int i = 0; i < n; i++) if (isnan(dot[i])) return true;
  auto converged = [&](const double* r) {                                    // phase.f90:626-638
    for (int i = 0; i < n; i++) if (!(fabs(r[i]) <= fmax(p.ph->p.atol[i], num.eps_rel_state * fabs(p.st[i])))) return false;
    return true;
  };
  if (num.integrator == DMK_INTEG_EULER || num.integrator == DMK_INTEG_ADAPTIVE_EULER) {
    double r[NSX];
    for (int i = 0; i < n; i++) { r[i] = -dot[i] * 0.5 * dt; p.st[i] = state0[i] + dot[i] * dt; }
    if (delta_state(p)) return true;                                         // :717-719, :765-767
    if (integrate_stress(p, num, F, Fp0, dt, cnt)) return true;
    if (num.integrator == DMK_INTEG_EULER) return false;
    dot_state(p, dot);
    for (int i = 0; i < n; i++) if (isnan(dot[i])) return true;
    for (int i = 0; i < n; i++) r[i] = r[i] + 0.5 * dot[i] * dt;
    return !converged(r);
  }
  // explicit Runge-Kutta (integrateStateRK, :862-932); A[stage][m], nodes C, weights B, embedded error weights DB (RKCK45)
  const double A4[9] = {0.5, 0, 0, 0, 0.5, 0, 0, 0, 1.0}, C4[3] = {0.5, 0.5, 1.0}, B4[4] = {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0};
  const double A5[25] = {1. / 5., 0, 0, 0, 0, 3. / 40., 9. / 40., 0, 0, 0, 3. / 10., -9. / 10., 6. / 5., 0, 0,
                         -11. / 54., 5. / 2., -70. / 27., 35. / 27., 0,
                         1631. / 55296., 175. / 512., 575. / 13824., 44275. / 110592., 253. / 4096.};
  const double C5[5] = {0.2, 0.3, 0.6, 1.0, 0.875};
  const double B5[6] = {37. / 378., 0, 250. / 621., 125. / 594., 0, 512. / 1771.};
  const double DB5[6] = {37. / 378. - 2825. / 27648., 0, 250. / 621. - 18575. / 48384., 125. / 594. - 13525. / 55296.,
                         0 - 277. / 14336., 512. / 1771. - 1. / 4.};
  const bool ck = num.integrator == DMK_INTEG_RKCK45;
  const int ns = ck ? 5 : 3;
  const double* A = ck ? A5 : A4; const double* C = ck ? C5 : C4; const double* B = ck ? B5 : B4;
  double rk[6][NSX];
  bool broken = true;
  for (int stage = 0; stage < ns; stage++) {
    for (int i = 0; i < n; i++) { rk[stage][i] = dot[i]; dot[i] = A[stage * ns] * rk[0][i]; }
    for (int m = 1; m <= stage; m++) for (int i = 0; i < n; i++) dot[i] += A[stage * ns + m] * rk[m][i];
    for (int i = 0; i < n; i++) p.st[i] = state0[i] + dot[i] * dt;
    double Fs[9];
#pragma unroll
    for (int a = 0; a < 9; a++) Fs[a] = F_0[a] + (F[a] - F_0[a]) * dt * C[stage];      // times dt, as the reference has it (:902)
    broken = integrate_stress(p, num, Fs, Fp0, dt * C[stage], cnt);
    if (broken) break;
    dot_state(p, dot);
    bool nan = false;
    for (int i = 0; i < n; i++) nan = nan || isnan(dot[i]);
    if (nan) break;
  }
  if (broken) return true;
  for (int i = 0; i < n; i++) rk[ns][i] = dot[i];
  for (int i = 0; i < n; i++) { double s = 0.0; for (int m = 0; m <= ns; m++) s += rk[m][i] * B[m]; dot[i] = s; }
  for (int i = 0; i < n; i++) p.st[i] = state0[i] + dot[i] * dt;
  if (ck) {
    double r[NSX];
    for (int i = 0; i < n; i++) { double s = 0.0; for (int m = 0; m <= ns; m++) s += rk[m][i] * DB5[m]; r[i] = s * dt; }
    if (!converged(r)) return true;
  }
  if (delta_state(p)) return true;                                           // :924-926
  return integrate_stress(p, num, F, Fp0, dt, cnt);
}

// ------------------------------------------------------------------------------------------------ constitutive
// Integrates one point from its increment-start state (arguments *0) to the target F.  Returns converged.
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ bool constitutive_point(DmkPt<LAW, NS, ND>& p, const DmkNumDev& num, const double (&F0in)[9],
                                                   const double (&Ftarget)[9], const double (&Fp0in)[9], const double (&Lp0in)[9],
                                                   const double (&S0in)[6], const double* __restrict__ state0in, double Dt, DmkCnt& cnt) {
  const int ns = p.ph->p.n_state;
  double state0[NS > 0 ? NS : 1], Lp0[9], Fp0[9], F0[9], F[9];
  for (int i = 0; i < ns; i++) state0[i] = state0in[i];
#pragma unroll
  for (int a = 0; a < 9; a++) { Lp0[a] = Lp0in[a]; Fp0[a] = Fp0in[a]; F0[a] = F0in[a]; F[a] = F0in[a]; }
  double stepFrac = 0.0, step = 1.0 / num.r_cutback;
  bool converged = false, todo = true;
  while (todo) {
    if (converged) {
      stepFrac += step;
      step = fmin(1.0 - stepFrac, num.r_increase * step);
      todo = step > 0.0;
      if (todo) {
#pragma unroll
        for (int a = 0; a < 9; a++) { F0[a] = F[a]; Lp0[a] = p.Lp[a]; Fp0[a] = p.Fp[a]; }
        for (int i = 0; i < ns; i++) state0[i] = p.st[i];
      }
    } else {
      step = num.r_cutback * step;
      if (step < 1.0) cnt.cutbacks++;
#pragma unroll
      for (int a = 0; a < 9; a++) p.Fp[a] = Fp0[a];
#pragma unroll
      for (int a = 0; a < 6; a++) p.S[a] = S0in[a];
      if (step < 1.0) {
#pragma unroll
        for (int a = 0; a < 9; a++) p.Lp[a] = Lp0[a];
      }
      for (int i = 0; i < ns; i++) p.st[i] = state0[i];
      todo = step > num.r_cutback_min;
    }
    if (todo) {
      cnt.substeps++;
#pragma unroll
      for (int a = 0; a < 9; a++) F[a] = F0[a] + step * (Ftarget[a] - F0in[a]);
      converged = (num.integrator == DMK_INTEG_FPI) ? !integrate_state_FPI(p, num, F, Fp0, state0, step * Dt, cnt)
                                                     : !integrate_state_other(p, num, F0, F, Fp0, state0, step * Dt, cnt);
    }
  }
  return converged;
}
