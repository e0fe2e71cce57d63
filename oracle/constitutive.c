/* ORACLE -- test infrastructure only; the product path never links or calls this file.
 *
 * Plain-C CPU restatement of the DAMASK per-material-point constitutive update, upstream semantics
 * (the Fortran library the reference's C++ driver calls through six bind(C) symbols).
 * One cell = one constituent (homogenization type "pass").
 *
 * Follows (file:line in /root/reference/src):
 *   homogenization_mechanical_response   homogenization.f90:219-266
 *   mechanical_restore(includeL=.false.) phase_mechanical.f90:1106-1131
 *   phase_mechanical_constitutive        phase_mechanical.f90:1013-1100  (sub-step / cut-back loop)
 *   integrateStateFPI (+ damper)         phase_mechanical.f90:611-693
 *   integrateStress                      phase_mechanical.f90:380-604
 *   phase_hooke_SandItsTangents          phase_mechanical_elastic.f90:157-195
 *   plastic_LpAndItsTangents             phase_mechanical_plastic.f90:233-292
 *   phase_LiAndItsTangents               phase_mechanical_eigen.f90:102-174
 *   phenopowerlaw kinetics/Lp/dotState   phase_mechanical_plastic_phenopowerlaw.f90:298-394,444-533
 *   isotropic Lp/Li/dotState             phase_mechanical_plastic_isotropic.f90:156-290
 *   dislotwin C66/Lp/dotState/dependent  phase_mechanical_plastic_dislotwin.f90:458-774,837-971
 *   phase_mechanical_dPdF                phase_mechanical.f90:1137-1258
 *   mechanical_forward                   phase_mechanical.f90:992-1007
 *   converged                            phase.f90:626-638;  dEq0 prec.f90:113-125
 *   math helpers                         math.f90:73-84,408-460,491-560,578-595,800-1017
 *
 * Conventions: 3x3 tensors are row-major a[3*i+j]; 3x3x3x3 tensors t[((i*3+j)*3+k)*3+l]; with the
 * reference's MAPPLAIN pair order (row-major) math_3333to99 is then the identity on memory.
 * Transformation (trans) systems of dislotwin are not restated (n_tr must be 0).
 */
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <float.h>
#include "_gen_layout.h"

#define T4(t,i,j,k,l) (t)[(((i)*3+(j))*3+(k))*3+(l)]
static const double TOL_MATH_CHECK = 1.0e-8;             /* prec.f90:34 */
static const double PREAL_MIN10 = 10.0 * DBL_MIN;        /* dEq0 */
static const double PI_ = 3.14159265358979323846;

static int deq0(double a) { return fabs(a) <= PREAL_MIN10; }
static int deq(double a, double b) {                     /* prec.f90:75-88 */
  double t = DBL_EPSILON * fmax(fabs(a), fabs(b));
  return fabs(a - b) <= t;
}

/* ------------------------------------------------------------------ small tensor helpers */
static void mm(const double *A, const double *B, double *C) {          /* C = A B */
  double r[9];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    double s = 0.0; for (int k = 0; k < 3; k++) s += A[3*i+k] * B[3*k+j]; r[3*i+j] = s; }
  memcpy(C, r, sizeof r);
}
static void tr3(const double *A, double *At) { double r[9];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r[3*i+j] = A[3*j+i];
  memcpy(At, r, sizeof r); }
static double det3(const double *A) {
  return A[0]*(A[4]*A[8]-A[5]*A[7]) - A[1]*(A[3]*A[8]-A[5]*A[6]) + A[2]*(A[3]*A[7]-A[4]*A[6]);
}
/* math_invert33: Cramer; returns 1 on error */
static int invert33(const double *A, double *inv) {
  double c00 =  A[4]*A[8] - A[5]*A[7];
  double c10 = -A[3]*A[8] + A[5]*A[6];
  double c20 =  A[3]*A[7] - A[4]*A[6];
  double det = A[0]*c00 + A[1]*c10 + A[2]*c20;
  if (deq0(det)) { memset(inv, 0, 9*sizeof(double)); return 1; }
  double r[9];
  r[0] = c00; r[3] = c10; r[6] = c20;
  r[1] = -A[1]*A[8] + A[2]*A[7];
  r[4] =  A[0]*A[8] - A[2]*A[6];
  r[7] = -A[0]*A[7] + A[1]*A[6];
  r[2] =  A[1]*A[5] - A[2]*A[4];
  r[5] = -A[0]*A[5] + A[2]*A[3];
  r[8] =  A[0]*A[4] - A[1]*A[3];
  for (int i = 0; i < 9; i++) inv[i] = r[i] / det;
  return 0;
}
static void inv33(const double *A, double *inv) { if (invert33(A, inv)) memset(inv, 0, 9*sizeof(double)); }
static double ddot(const double *A, const double *B) { double s = 0; for (int i = 0; i < 9; i++) s += A[i]*B[i]; return s; }
static double norm2_9(const double *A) { return sqrt(ddot(A, A)); }
static int any_nan(const double *a, int n) { for (int i = 0; i < n; i++) if (isnan(a[i])) return 1; return 0; }
static const double I3[9] = {1,0,0, 0,1,0, 0,0,1};

/* 9x9 (row-major) product C = A B  == math_mul3333xx3333 */
static void mul99(const double *A, const double *B, double *C) {
  double r[81];
  for (int i = 0; i < 9; i++) for (int j = 0; j < 9; j++) {
    double s = 0; for (int k = 0; k < 9; k++) s += A[9*i+k]*B[9*k+j]; r[9*i+j] = s; }
  memcpy(C, r, sizeof r);
}
/* LU with partial pivoting (LAPACK dgetrf semantics); returns nonzero on exact singularity */
static int lu_factor(double *A, int n, int *piv) {
  for (int k = 0; k < n; k++) {
    int p = k; double mx = fabs(A[n*k+k]);
    for (int i = k+1; i < n; i++) if (fabs(A[n*i+k]) > mx) { mx = fabs(A[n*i+k]); p = i; }
    piv[k] = p;
    if (A[n*p+k] == 0.0) return k+1;
    if (p != k) for (int j = 0; j < n; j++) { double t = A[n*k+j]; A[n*k+j] = A[n*p+j]; A[n*p+j] = t; }
    for (int i = k+1; i < n; i++) {
      A[n*i+k] /= A[n*k+k];
      double l = A[n*i+k];
      for (int j = k+1; j < n; j++) A[n*i+j] -= l * A[n*k+j];
    }
  }
  return 0;
}
static void lu_solve(const double *LU, int n, const int *piv, double *b) {
  for (int k = 0; k < n; k++) { int p = piv[k]; if (p != k) { double t = b[k]; b[k] = b[p]; b[p] = t; } }
  for (int i = 0; i < n; i++) { double s = b[i]; for (int j = 0; j < i; j++) s -= LU[n*i+j]*b[j]; b[i] = s; }
  for (int i = n-1; i >= 0; i--) { double s = b[i]; for (int j = i+1; j < n; j++) s -= LU[n*i+j]*b[j]; b[i] = s / LU[n*i+i]; }
}
/* dgesv for one right-hand side */
static int solve_nn(double *A, int n, double *b) { int piv[16]; int e = lu_factor(A, n, piv); if (e) return e; lu_solve(A, n, piv, b); return 0; }
/* math_invert (dgetrf+dgetri) */
static int invert_nn(const double *A, int n, double *inv) {
  double LU[81]; int piv[16];
  memcpy(LU, A, sizeof(double)*n*n);
  int e = lu_factor(LU, n, piv);
  if (e) { memset(inv, 0, sizeof(double)*n*n); return e; }
  for (int c = 0; c < n; c++) {
    double col[16]; for (int i = 0; i < n; i++) col[i] = (i == c) ? 1.0 : 0.0;
    lu_solve(LU, n, piv, col);
    for (int i = 0; i < n; i++) inv[n*i+c] = col[i];
  }
  return 0;
}

/* ------------------------------------------------------------------ per point view */
typedef struct {
  const orc_phase_t *ph; const orc_num_t *num;
  double *F, *F0, *Fp, *Fp0, *Fi, *Fi0, *Fe, *Lp, *Lp0, *Li, *Li0, *S, *S0, *P;
  double *state, *state0, *dep;
} pt_t;

static pt_t point_view(const orc_phase_t *phases, const orc_num_t *num, const orc_points_t *st, int ce) {
  pt_t p; p.ph = &phases[st->phase_of[ce]]; p.num = num;
  p.F = st->F + 9*ce; p.F0 = st->F0 + 9*ce; p.Fp = st->Fp + 9*ce; p.Fp0 = st->Fp0 + 9*ce;
  p.Fi = st->Fi + 9*ce; p.Fi0 = st->Fi0 + 9*ce; p.Fe = st->Fe + 9*ce; p.Lp = st->Lp + 9*ce; p.Lp0 = st->Lp0 + 9*ce;
  p.Li = st->Li + 9*ce; p.Li0 = st->Li0 + 9*ce; p.S = st->S + 9*ce; p.S0 = st->S0 + 9*ce; p.P = st->P + 9*ce;
  p.state = st->state + (size_t)st->max_state*ce; p.state0 = st->state0 + (size_t)st->max_state*ce;
  p.dep = st->dep + (size_t)st->max_dep*ce;
  return p;
}
/* dependent-state slots (dislotwin): Lambda_sl[n_sl] | Lambda_tw[n_tw] | tau_pass[n_sl] */
#define DEP_LAMBDA_SL(p) ((p)->dep)
#define DEP_LAMBDA_TW(p) ((p)->dep + (p)->ph->n_sl)
#define DEP_TAU_PASS(p)  ((p)->dep + (p)->ph->n_sl + (p)->ph->n_tw)

/* ------------------------------------------------------------------ elasticity */
static const int VOI[6][2] = {{0,0},{1,1},{2,2},{1,2},{0,2},{0,1}};

static void homogenized_C66(const pt_t *p, double *C66) {
  const orc_phase_t *ph = p->ph;
  if (ph->law != 3) { memcpy(C66, ph->C66, 36*sizeof(double)); return; }
  /* dislotwin.f90:458-500 ; state layout rho_mob|rho_dip|gamma_sl|f_tw|f_tr */
  const double *f_tw = p->state + 3*ph->n_sl;
  double f_matrix = 1.0; double s = 0; for (int i = 0; i < ph->n_tw; i++) s += f_tw[i];
  f_matrix -= s;
  for (int k = 0; k < 36; k++) C66[k] = f_matrix * ph->C66[k];
  for (int i = 0; i < ph->n_tw; i++) for (int k = 0; k < 36; k++) C66[k] += f_tw[i] * ph->dt_C66_tw[i][k];
}

static void hooke(const pt_t *p, const double *Fe, const double *Fi, double *S, double *dS_dFe, double *dS_dFi) {
  double C66[36], C[81];
  homogenized_C66(p, C66);
  for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) {
    int i = VOI[a][0], j = VOI[a][1], k = VOI[b][0], l = VOI[b][1]; double v = C66[6*a+b];
    T4(C,i,j,k,l) = v; T4(C,j,i,k,l) = v; T4(C,i,j,l,k) = v; T4(C,j,i,l,k) = v; }
  double FeT[9], E[9], FiT[9], t[9], EE[9];
  tr3(Fe, FeT); mm(FeT, Fe, E); for (int i = 0; i < 9; i++) E[i] = 0.5*(E[i] - I3[i]);
  tr3(Fi, FiT); mm(FiT, E, t); mm(t, Fi, EE);
  double ev[6] = {EE[0], EE[4], EE[8], 2.0*EE[7], 2.0*EE[6], 2.0*EE[1]};   /* math_33toVoigt6_strain */
  double sv[6];
  for (int a = 0; a < 6; a++) { double s = 0; for (int b = 0; b < 6; b++) s += C66[6*a+b]*ev[b]; sv[a] = s; }
  S[0] = sv[0]; S[1] = sv[5]; S[2] = sv[4];
  S[3] = sv[5]; S[4] = sv[1]; S[5] = sv[3];
  S[6] = sv[4]; S[7] = sv[3]; S[8] = sv[2];
  double EFi[9]; mm(E, Fi, EFi);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    double Cij[9], a[9], b[9];
    for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) Cij[3*k+l] = T4(C,i,j,k,l);
    mm(Fi, Cij, a); mm(a, FiT, b); mm(Fe, b, a);                          /* Fe (Fi C_ij Fi^T) */
    for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) T4(dS_dFe,i,j,k,l) = a[3*k+l];
    mm(EFi, Cij, b);
    for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) T4(dS_dFi,i,j,k,l) = 2.0*b[3*k+l];
  }
}

/* ------------------------------------------------------------------ phenopowerlaw */
static void pheno_kinetics_sl(const pt_t *p, const double *Mp, double *dg, double *ddg) {
  const orc_phase_t *ph = p->ph; const double *xi_sl = p->state;
  for (int i = 0; i < ph->n_sl; i++) {
    double tp = ddot(Mp, ph->P_nS_pos[i]), tn = ddot(Mp, ph->P_nS_neg[i]);
    double tmax = fmax(tp, tn);
    dg[i] = ((tp > tn) ? 1.0 : -1.0) * ph->dot_gamma_0_sl[i] * pow(tmax / xi_sl[i], ph->n_sl_exp[i]);
    if (ddg) ddg[i] = deq0(dg[i]) ? 0.0 : dg[i] * ph->n_sl_exp[i] / tmax;
  }
}
static void pheno_kinetics_tw(const pt_t *p, const double *Mp, double *dg, double *ddg) {
  const orc_phase_t *ph = p->ph; int ns = ph->n_sl, nt = ph->n_tw;
  const double *xi_tw = p->state + ns, *gamma_tw = p->state + 2*ns + nt;
  double sumF = 0; for (int i = 0; i < nt; i++) sumF += gamma_tw[i] / ph->gamma_char_tw[i];
  for (int i = 0; i < nt; i++) {
    double tau = ddot(Mp, ph->P_tw[i]);
    dg[i] = (tau > 0.0) ? (1.0 - sumF) * ph->dot_gamma_0_tw[i] * pow(tau / xi_tw[i], ph->n_tw_exp[i]) : 0.0;
    if (ddg) ddg[i] = deq0(dg[i]) ? 0.0 : dg[i] * ph->n_tw_exp[i] / tau;
  }
}
static void pheno_Lp(const pt_t *p, const double *Mp, double *Lp, double *dLp_dMp) {
  const orc_phase_t *ph = p->ph;
  double dg[ORC_MAX_SL], ddg[ORC_MAX_SL], dgt[ORC_MAX_TW], ddgt[ORC_MAX_TW];
  memset(Lp, 0, 9*sizeof(double)); memset(dLp_dMp, 0, 81*sizeof(double));
  pheno_kinetics_sl(p, Mp, dg, ddg);
  for (int i = 0; i < ph->n_sl; i++) {
    const double *PnS = (dg[i] > 0.0) ? ph->P_nS_pos[i] : ph->P_nS_neg[i];
    for (int a = 0; a < 9; a++) Lp[a] += dg[i] * ph->P_sl[i][a];
    for (int a = 0; a < 9; a++) for (int b = 0; b < 9; b++) dLp_dMp[9*a+b] += ddg[i] * ph->P_sl[i][a] * PnS[b];
  }
  pheno_kinetics_tw(p, Mp, dgt, ddgt);
  for (int i = 0; i < ph->n_tw; i++) {
    for (int a = 0; a < 9; a++) Lp[a] += dgt[i] * ph->P_tw[i][a];
    for (int a = 0; a < 9; a++) for (int b = 0; b < 9; b++) dLp_dMp[9*a+b] += ddgt[i] * ph->P_tw[i][a] * ph->P_tw[i][b];
  }
}
static double fsign(double a, double b) { return (b >= 0.0 && !signbit(b)) ? fabs(a) : -fabs(a); }

static void pheno_dotState(const pt_t *p, const double *Mp, double *dot) {
  const orc_phase_t *ph = p->ph; int ns = ph->n_sl, nt = ph->n_tw;
  const double *xi_sl = p->state, *gamma_sl = p->state + ns + nt, *gamma_tw = p->state + 2*ns + nt;
  double *dot_xi_sl = dot, *dot_xi_tw = dot + ns, *dot_g_sl = dot + ns + nt, *dot_g_tw = dot + 2*ns + nt;
  pheno_kinetics_sl(p, Mp, dot_g_sl, NULL);
  pheno_kinetics_tw(p, Mp, dot_g_tw, NULL);
  for (int i = 0; i < ns; i++) dot_g_sl[i] = fabs(dot_g_sl[i]);
  double sumF = 0; for (int i = 0; i < nt; i++) sumF += gamma_tw[i] / ph->gamma_char_tw[i];
  double sum_gamma_sl = 0; for (int i = 0; i < ns; i++) sum_gamma_sl += gamma_sl[i];
  for (int i = 0; i < ns; i++) {
    double off = ph->f_sat_sl_tw[i] * sqrt(sumF);
    double x = 1.0 - xi_sl[i] / (ph->xi_inf_sl[i] + off);
    double left = fsign(pow(fabs(x), ph->a_sl[i]), x);
    double hs = 0; for (int j = 0; j < ns; j++) hs += ph->h_sl_sl[i][j] * dot_g_sl[j];
    double ht = 0; for (int j = 0; j < nt; j++) ht += ph->h_sl_tw[i][j] * dot_g_tw[j];
    dot_xi_sl[i] = ph->h_0_sl_sl[i] * (1.0 + ph->c_1[i] * pow(sumF, ph->c_2[i])) * left * hs + ht;
  }
  for (int i = 0; i < nt; i++) {
    double hs = 0; for (int j = 0; j < ns; j++) hs += ph->h_tw_sl[i][j] * dot_g_sl[j];
    double ht = 0; for (int j = 0; j < nt; j++) ht += ph->h_tw_tw[i][j] * dot_g_tw[j];
    dot_xi_tw[i] = ph->h_0_tw_sl[i] * pow(sum_gamma_sl, ph->c_3[i]) * hs + ph->h_0_tw_tw[i] * pow(sumF, ph->c_4[i]) * ht;
  }
}

/* ------------------------------------------------------------------ kinehardening
 * phase_mechanical_plastic_kinehardening.f90: kinetics :433-472, LpAndItsTangent :258-295, dotState :301-341,
 * deltaState :347-379.  State: xi, chi, gamma (follow dotState) | sgn_gamma, chi_flip, gamma_flip (move by deltaState only) */
static void kh_kinetics(const pt_t *p, const double *Mp, double *dg, double *ddg) {
  const orc_phase_t *ph = p->ph; int n = ph->n_sl;
  const double *xi = p->state, *chi = p->state + n;
  for (int i = 0; i < n; i++) {
    double tp = ddot(Mp, ph->P_nS_pos[i]) - chi[i], tn = ddot(Mp, ph->P_nS_neg[i]) + chi[i];
    double tmax = fmax(tp, tn);
    dg[i] = ((tp > tn) ? 1.0 : -1.0) * ph->dot_gamma_0_sl[i] * pow(tmax / xi[i], ph->n_sl_exp[i]);
    if (ddg) ddg[i] = deq0(dg[i]) ? 0.0 : dg[i] * ph->n_sl_exp[i] / tmax;
  }
}
static void kh_Lp(const pt_t *p, const double *Mp, double *Lp, double *dLp_dMp) {
  const orc_phase_t *ph = p->ph;
  double dg[ORC_MAX_SL], ddg[ORC_MAX_SL];
  memset(Lp, 0, 9*sizeof(double)); memset(dLp_dMp, 0, 81*sizeof(double));
  kh_kinetics(p, Mp, dg, ddg);
  for (int i = 0; i < ph->n_sl; i++) {
    const double *PnS = (dg[i] > 0.0) ? ph->P_nS_pos[i] : ph->P_nS_neg[i];
    for (int a = 0; a < 9; a++) Lp[a] += dg[i] * ph->P_sl[i][a];
    for (int a = 0; a < 9; a++) for (int b = 0; b < 9; b++) dLp_dMp[9*a+b] += ddg[i] * ph->P_sl[i][a] * PnS[b];
  }
}
static void kh_dotState(const pt_t *p, const double *Mp, double *dot) {
  const orc_phase_t *ph = p->ph; int n = ph->n_sl;
  const double *gamma = p->state + 2*n, *sgn = p->state + 3*n, *chi_flip = p->state + 4*n, *gamma_flip = p->state + 5*n;
  double *dot_xi = dot, *dot_chi = dot + n, *dot_gamma = dot + 2*n;
  kh_kinetics(p, Mp, dot_gamma, NULL);
  double sumGamma = 0; for (int i = 0; i < n; i++) sumGamma += gamma[i];
  for (int i = 0; i < n; i++) dot_gamma[i] = fabs(dot_gamma[i]);
  for (int i = 0; i < n; i++) {
    double hs = 0; for (int j = 0; j < n; j++) hs += ph->h_sl_sl[i][j] * dot_gamma[j];
    double x = sumGamma * ph->kh_h_0_xi[i] / ph->xi_inf_sl[i];
    dot_xi[i] = hs * (ph->kh_h_inf_xi[i] + (ph->kh_h_0_xi[i] - ph->kh_h_inf_xi[i] * (1.0 - x)) * exp(-x));
    double y = (gamma[i] - gamma_flip[i]) * ph->kh_h_0_chi[i] / (ph->kh_chi_inf[i] + chi_flip[i]);
    dot_chi[i] = sgn[i] * dot_gamma[i] * (ph->kh_h_inf_chi[i] + (ph->kh_h_0_chi[i] - ph->kh_h_inf_chi[i] * (1.0 - y)) * exp(-y));
  }
}
/* plastic_deltaState (phase_mechanical_plastic.f90:372-413) for kinehardening; returns broken */
static int kh_deltaState(pt_t *p, const double *Mp) {
  const orc_phase_t *ph = p->ph; int n = ph->n_sl;
  double dg[ORC_MAX_SL], dlt[3 * ORC_MAX_SL];
  double *chi = p->state + n, *gamma = p->state + 2*n, *sgn = p->state + 3*n, *chi_flip = p->state + 4*n, *gamma_flip = p->state + 5*n;
  kh_kinetics(p, Mp, dg, NULL);
  for (int i = 0; i < n; i++) {
    double s = (fabs(dg[i]) <= 1.0e-10) ? sgn[i] : ((dg[i] >= 0.0 && !signbit(dg[i])) ? 1.0 : -1.0);
    if (!(fabs(s - sgn[i]) <= 0.1)) { dlt[i] = s - sgn[i]; dlt[n + i] = fabs(chi[i]) - chi_flip[i]; dlt[2*n + i] = gamma[i] - gamma_flip[i]; }
    else { dlt[i] = 0.0; dlt[n + i] = 0.0; dlt[2*n + i] = 0.0; }
  }
  if (any_nan(dlt, 3 * n)) return 1;
  for (int i = 0; i < n; i++) { sgn[i] += dlt[i]; chi_flip[i] += dlt[n + i]; gamma_flip[i] += dlt[2*n + i]; }
  return 0;
}

/* ------------------------------------------------------------------ isotropic */
static void deviatoric(const double *M, double *D) { double tr = (M[0]+M[4]+M[8])/3.0; memcpy(D, M, 9*sizeof(double)); D[0]-=tr; D[4]-=tr; D[8]-=tr; }
static void iso_Lp(const pt_t *p, const double *Mp, double *Lp, double *dLp_dMp) {
  const orc_phase_t *ph = p->ph; double xi = p->state[0];
  double dev[9]; deviatoric(Mp, dev);
  double sq = ddot(dev, dev), nrm = sqrt(sq);
  if (nrm > 0.0) {
    double dgam = ph->iso_dot_gamma_0 * pow(sqrt(1.5) * nrm / (ph->iso_M * xi), ph->iso_n);
    for (int a = 0; a < 9; a++) Lp[a] = dgam * dev[a] / nrm;
    for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) for (int m = 0; m < 3; m++) for (int n = 0; n < 3; n++)
      T4(dLp_dMp,k,l,m,n) = (ph->iso_n - 1.0) * dev[3*k+l] * dev[3*m+n] / sq;
    for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) T4(dLp_dMp,k,l,k,l) += 1.0;
    for (int k = 0; k < 3; k++) for (int m = 0; m < 3; m++) T4(dLp_dMp,k,k,m,m) -= 1.0/3.0;
    for (int a = 0; a < 81; a++) dLp_dMp[a] = dgam * dLp_dMp[a] / nrm;
  } else { memset(Lp, 0, 9*sizeof(double)); memset(dLp_dMp, 0, 81*sizeof(double)); }
}
static void iso_Li(const pt_t *p, const double *Mi, double *Li, double *dLi_dMi) {
  const orc_phase_t *ph = p->ph; double xi = p->state[0];
  double tr = Mi[0] + Mi[4] + Mi[8];      /* trace of the spherical part == trace */
  if (ph->iso_dilatation && fabs(tr) > 0.0) {
    double f = ph->iso_dot_gamma_0 * pow(3.0 * ph->iso_M * xi, -ph->iso_n) * tr * pow(fabs(tr), ph->iso_n - 1.0);
    for (int a = 0; a < 9; a++) Li[a] = I3[a] * f;
    for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) for (int m = 0; m < 3; m++) for (int n = 0; n < 3; n++)
      T4(dLi_dMi,k,l,m,n) = ph->iso_n / tr * Li[3*k+l] * I3[3*m+n];
  } else { memset(Li, 0, 9*sizeof(double)); memset(dLi_dMi, 0, 81*sizeof(double)); }
}
static void iso_dotState(const pt_t *p, const double *Mp, double *dot) {
  const orc_phase_t *ph = p->ph; double xi = p->state[0];
  double nrm;
  if (ph->iso_dilatation) nrm = sqrt(ddot(Mp, Mp));
  else { double dev[9]; deviatoric(Mp, dev); nrm = sqrt(ddot(dev, dev)); }
  double dgam = ph->iso_dot_gamma_0 * pow(sqrt(1.5) * nrm / (ph->iso_M * xi), ph->iso_n);
  if (dgam > 1e-12) {
    double xi_inf_star;
    if (deq0(ph->iso_c_1)) xi_inf_star = ph->iso_xi_inf;
    else xi_inf_star = ph->iso_xi_inf + pow(asinh(pow(dgam / ph->iso_c_1, 1.0/ph->iso_c_2)), 1.0/ph->iso_c_3)
                       / ph->iso_c_4 * pow(dgam / ph->iso_dot_gamma_0, 1.0/ph->iso_n);
    double x = 1.0 - xi / xi_inf_star;
    dot[0] = dgam * (ph->iso_h_0 + ph->iso_h_ln * log(dgam)) * fsign(pow(fabs(x), ph->iso_a) * ph->iso_h, x);
  } else dot[0] = 0.0;
}

/* ------------------------------------------------------------------ dislotwin */
static void dt_kinetics_sl(const pt_t *p, const double *Mp, double *dg, double *ddg) {
  const orc_phase_t *ph = p->ph; const double *rho_mob = p->state; const double *tau_pass = DEP_TAU_PASS(p);
  for (int i = 0; i < ph->n_sl; i++) {
    double tau = ddot(Mp, ph->P_sl[i]);
    double tau_eff = fabs(tau) - tau_pass[i];
    if (tau_eff > TOL_MATH_CHECK) {
      double sr = tau_eff / ph->dt_tau_0[i], srp = pow(sr, ph->dt_p[i]);
      double QkT = ph->dt_Q_sl[i] / ph->dt_kB_T;
      double vw = exp(QkT * pow(1.0 - srp, ph->dt_q[i])) / ph->dt_v_0[i];
      double vr = ph->dt_B[i] / (tau_eff * ph->dt_b_sl[i]);
      dg[i] = fsign(rho_mob[i] * ph->dt_b_sl[i] / (vw + vr), tau);
      if (ddg) {
        double dvw = -1.0 * vw * ph->dt_p[i] * ph->dt_q[i] * QkT * pow(sr, ph->dt_p[i] - 1.0)
                     * pow(1.0 - srp, ph->dt_q[i] - 1.0) / ph->dt_tau_0[i];
        double dvr = -1.0 * vr / tau_eff;
        double dV = -1.0 * (dvw + dvr) / ((vw + vr) * (vw + vr));
        ddg[i] = dV * rho_mob[i] * ph->dt_b_sl[i];
      }
    } else { dg[i] = 0.0; if (ddg) ddg[i] = 0.0; }
  }
}
static void dt_kinetics_tw(const pt_t *p, const double *Mp, const double *adg_sl, double *dg, double *ddg) {
  const orc_phase_t *ph = p->ph; int ns = ph->n_sl;
  const double *rho_mob = p->state, *rho_dip = p->state + ns; const double *Lambda_tw = DEP_LAMBDA_TW(p);
  double mu = ph->dt_mu, nu = ph->dt_nu, G = ph->dt_Gamma_sf;
  double tau_hat = 3.0 * ph->dt_b_tw[0] * mu / ph->dt_L_tw + G / (3.0 * ph->dt_b_tw[0]);
  double x0 = mu * ph->dt_b_sl[0] * ph->dt_b_sl[0] * (2.0 + nu) / (G * 8.0 * PI_ * (1.0 - nu));
  double tau_r = mu * ph->dt_b_sl[0] / (2.0 * PI_) * (1.0 / (x0 + ph->dt_x_c) + cos(PI_/3.0) / x0);
  for (int i = 0; i < ph->n_tw; i++) {
    double tau = ddot(Mp, ph->P_tw[i]);
    if (tau > TOL_MATH_CHECK && tau < tau_r) {
      double Pn = exp(-pow(tau_hat / tau, ph->dt_r[i]));
      double dP = ph->dt_r[i] * pow(tau_hat / tau, ph->dt_r[i]) / tau * Pn;
      int s1 = ph->dt_pair[i][0], s2 = ph->dt_pair[i][1];
      double dotN0 = (adg_sl[s2] * (rho_mob[s1] + rho_dip[s1]) + adg_sl[s1] * (rho_mob[s2] + rho_dip[s2])) / (ph->dt_L_tw * 3.0);
      double Pncs = 1.0 - exp(-ph->dt_V_cs / ph->dt_kB_T * (tau_r - tau));
      double dPncs = ph->dt_V_cs / ph->dt_kB_T * (Pncs - 1.0);
      double V = PI_ / 4.0 * Lambda_tw[i] * Lambda_tw[i] * ph->dt_t_tw[i];
      dg[i] = V * dotN0 * Pncs * Pn * ph->gamma_char_tw[i];
      if (ddg) ddg[i] = V * dotN0 * (Pn * dPncs + Pncs * dP) * ph->gamma_char_tw[i];
    } else { dg[i] = 0.0; if (ddg) ddg[i] = 0.0; }
  }
}
/* symmetric 3x3 eigen-decomposition (math_eigh33 stand-in): cyclic Jacobi, eigenvalues ascending;
 * the reference uses an analytic solver with LAPACK fallback -- eigenvector sign/ordering is
 * "parity unpinned" for the shear-band term (no fixture anywhere in the reference). */
static void eigh33(const double *M, double *w, double *V) {
  double A[9]; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) A[3*i+j] = 0.5*(M[3*i+j] + M[3*j+i]);
  memcpy(V, I3, sizeof I3);
  for (int sweep = 0; sweep < 50; sweep++) {
    double off = A[1]*A[1] + A[2]*A[2] + A[5]*A[5];
    if (off < 1e-300) break;
    for (int pq = 0; pq < 3; pq++) {
      int pp = (pq == 2) ? 1 : 0, q = (pq == 0) ? 1 : 2;
      if (A[3*pp+q] == 0.0) continue;
      double th = (A[3*q+q] - A[3*pp+pp]) / (2.0 * A[3*pp+q]);
      double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th*th + 1.0));
      double c = 1.0 / sqrt(t*t + 1.0), s = t * c;
      for (int k = 0; k < 3; k++) { double akp = A[3*k+pp], akq = A[3*k+q]; A[3*k+pp] = c*akp - s*akq; A[3*k+q] = s*akp + c*akq; }
      for (int k = 0; k < 3; k++) { double apk = A[3*pp+k], aqk = A[3*q+k]; A[3*pp+k] = c*apk - s*aqk; A[3*q+k] = s*apk + c*aqk; }
      for (int k = 0; k < 3; k++) { double vkp = V[3*k+pp], vkq = V[3*k+q]; V[3*k+pp] = c*vkp - s*vkq; V[3*k+q] = s*vkp + c*vkq; }
    }
  }
  w[0] = A[0]; w[1] = A[4]; w[2] = A[8];
  for (int a = 0; a < 2; a++) for (int b = a+1; b < 3; b++) if (w[b] < w[a]) {
    double t = w[a]; w[a] = w[b]; w[b] = t;
    for (int k = 0; k < 3; k++) { double u = V[3*k+a]; V[3*k+a] = V[3*k+b]; V[3*k+b] = u; } }
}
static void dt_Lp(const pt_t *p, const double *Mp, double *Lp, double *dLp_dMp) {
  const orc_phase_t *ph = p->ph; int ns = ph->n_sl, nt = ph->n_tw;
  const double *f_tw = p->state + 3*ns;
  double dg[ORC_MAX_SL], ddg[ORC_MAX_SL], dgt[ORC_MAX_TW], ddgt[ORC_MAX_TW];
  memset(Lp, 0, 9*sizeof(double)); memset(dLp_dMp, 0, 81*sizeof(double));
  double f_matrix = 1.0; { double s = 0; for (int i = 0; i < nt; i++) s += f_tw[i]; f_matrix -= s; }
  dt_kinetics_sl(p, Mp, dg, ddg);
  for (int i = 0; i < ns; i++) {
    for (int a = 0; a < 9; a++) Lp[a] += dg[i] * ph->P_sl[i][a];
    for (int a = 0; a < 9; a++) for (int b = 0; b < 9; b++) dLp_dMp[9*a+b] += ddg[i] * ph->P_sl[i][a] * ph->P_sl[i][b];
  }
  if (nt > 0) dt_kinetics_tw(p, Mp, dg, dgt, ddgt);       /* NB: reference passes the SIGNED slip rates here (:563) */
  for (int i = 0; i < nt; i++) {
    for (int a = 0; a < 9; a++) Lp[a] += dgt[i] * ph->P_tw[i][a];
    for (int a = 0; a < 9; a++) for (int b = 0; b < 9; b++) dLp_dMp[9*a+b] += ddgt[i] * ph->P_tw[i][a] * ph->P_tw[i][b];
  }
  for (int a = 0; a < 9; a++) Lp[a] *= f_matrix;
  for (int a = 0; a < 81; a++) dLp_dMp[a] *= f_matrix;
  if (!deq0(ph->dt_gamma_0_sb)) {
    static const double sC[6][3] = {{1,0,1},{1,0,-1},{1,1,0},{1,-1,0},{0,1,1},{0,1,-1}};
    static const double mC[6][3] = {{1,0,-1},{1,0,1},{1,-1,0},{1,1,0},{0,1,-1},{0,1,1}};
    double EkT = ph->dt_E_sb / ph->dt_kB_T, w[3], V[9];
    eigh33(Mp, w, V);
    for (int i = 0; i < 6; i++) {
      double s[3], m[3], Psb[9];
      for (int r = 0; r < 3; r++) { s[r] = 0; m[r] = 0; for (int c = 0; c < 3; c++) { s[r] += V[3*r+c]*sC[i][c]; m[r] += V[3*r+c]*mC[i][c]; } }
      for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) Psb[3*r+c] = 0.5 * s[r] * m[c];
      double tau = ddot(Mp, Psb);
      if (fabs(tau) > TOL_MATH_CHECK) {
        double srp = pow(fabs(tau) / ph->dt_tau_sb, ph->dt_p_sb);
        double dgsb = fsign(ph->dt_gamma_0_sb * exp(-EkT * pow(1 - srp, ph->dt_q_sb)), tau);
        double dd = fabs(dgsb) * EkT * ph->dt_p_sb * ph->dt_q_sb / ph->dt_tau_sb
                    * pow(fabs(tau) / ph->dt_tau_sb, ph->dt_p_sb - 1.0) * pow(1.0 - srp, ph->dt_q_sb - 1.0);
        for (int a = 0; a < 9; a++) Lp[a] += dgsb * Psb[a];
        for (int a = 0; a < 9; a++) for (int b = 0; b < 9; b++) dLp_dMp[9*a+b] += dd * Psb[a] * Psb[b];
      }
    }
  }
}
static double clipf(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }
static void dt_dotState(const pt_t *p, const double *Mp, double *dot) {
  const orc_phase_t *ph = p->ph; int ns = ph->n_sl, nt = ph->n_tw;
  const double *rho_mob = p->state, *rho_dip = p->state + ns, *f_tw = p->state + 3*ns;
  const double *Lambda_sl = DEP_LAMBDA_SL(p);
  double *dot_mob = dot, *dot_dip = dot + ns, *adg = dot + 2*ns, *dot_f_tw = dot + 3*ns;
  double mu = ph->dt_mu, nu = ph->dt_nu, kT = ph->dt_kB_T;
  double f_matrix = 1.0; { double s = 0; for (int i = 0; i < nt; i++) s += f_tw[i]; f_matrix -= s; }
  double dg[ORC_MAX_SL], form[ORC_MAX_SL], climb[ORC_MAX_SL];
  dt_kinetics_sl(p, Mp, dg, NULL);
  for (int i = 0; i < ns; i++) adg[i] = fabs(dg[i]);
  for (int i = 0; i < ns; i++) {
    double tau = ddot(Mp, ph->P_sl[i]);
    if (deq0(tau) || ph->dt_omit_dipoles) { form[i] = 0.0; climb[i] = 0.0; }
    else {
      double d_hat = 3.0 * mu * ph->dt_b_sl[i] / (16.0 * PI_ * fabs(tau));
      d_hat = fmin(d_hat, Lambda_sl[i]);                 /* math_clip(right=) */
      d_hat = fmax(d_hat, ph->dt_d_caron[i]);            /* math_clip(left=)  */
      form[i] = 2.0 * (d_hat - ph->dt_d_caron[i]) / ph->dt_b_sl[i] * rho_mob[i] * adg[i];
      if (deq(d_hat, ph->dt_d_caron[i])) climb[i] = 0.0;
      else {
        const double *n = ph->dt_n0_sl[i]; double Mn[3];
        for (int r = 0; r < 3; r++) Mn[r] = Mp[3*r]*n[0] + Mp[3*r+1]*n[1] + Mp[3*r+2]*n[2];
        double sigma_cl = n[0]*Mn[0] + n[1]*Mn[1] + n[2]*Mn[2];
        double b_d = ph->dt_extended ? 24.0 * PI_ * (1.0 - nu) / (2.0 + nu) * ph->dt_Gamma_sf / (mu * ph->dt_b_sl[i]) : 1.0;
        double b3 = ph->dt_b_sl[i] * ph->dt_b_sl[i] * ph->dt_b_sl[i];
        double v_cl = 2.0 * ph->dt_omega * b_d * b_d * exp(-ph->dt_Q_cl / kT) * (exp(fabs(sigma_cl) * b3 / kT) - 1.0);
        climb[i] = 4.0 * v_cl * rho_dip[i] / (d_hat - ph->dt_d_caron[i]);
      }
    }
  }
  for (int i = 0; i < ns; i++) {
    dot_mob[i] = adg[i] / (ph->dt_b_sl[i] * Lambda_sl[i]) - form[i] - 2.0 * ph->dt_d_caron[i] / ph->dt_b_sl[i] * rho_mob[i] * adg[i];
    dot_dip[i] = form[i] - 2.0 * ph->dt_d_caron[i] / ph->dt_b_sl[i] * rho_dip[i] * adg[i] - climb[i];
  }
  if (nt > 0) {
    double dgt[ORC_MAX_TW];
    dt_kinetics_tw(p, Mp, adg, dgt, NULL);
    for (int i = 0; i < nt; i++) dot_f_tw[i] = f_matrix * dgt[i] / ph->gamma_char_tw[i];
  }
  (void)clipf;
}
static void dt_dependentState(pt_t *p) {
  const orc_phase_t *ph = p->ph; int ns = ph->n_sl, nt = ph->n_tw;
  const double *rho_mob = p->state, *rho_dip = p->state + ns, *f_tw = p->state + 3*ns;
  double *Lambda_sl = DEP_LAMBDA_SL(p), *Lambda_tw = DEP_LAMBDA_TW(p), *tau_pass = DEP_TAU_PASS(p);
  double sumf_tw = 0; for (int i = 0; i < nt; i++) sumf_tw += f_tw[i];
  double fot[ORC_MAX_TW]; for (int i = 0; i < nt; i++) fot[i] = f_tw[i] / ph->dt_t_tw[i];
  for (int i = 0; i < ns; i++) {
    double s = 0; for (int j = 0; j < ns; j++) s += ph->dt_forest[i][j] * (rho_mob[j] + rho_dip[j]);
    double inv = sqrt(s) / ph->dt_i_sl[i];
    if (nt > 0) { double h = 0; for (int j = 0; j < nt; j++) h += ph->h_sl_tw[i][j] * fot[j]; inv += h / (1.0 - sumf_tw); }
    Lambda_sl[i] = ph->dt_D / (1.0 + ph->dt_D * inv);
    double hp = 0; for (int j = 0; j < ns; j++) hp += ph->h_sl_sl[i][j] * (rho_mob[j] + rho_dip[j]);
    tau_pass[i] = ph->dt_mu * ph->dt_b_sl[i] * sqrt(hp);
  }
  for (int i = 0; i < nt; i++) {
    double h = 0; for (int j = 0; j < nt; j++) h += ph->h_tw_tw[i][j] * fot[j];
    double inv = h / (1.0 - sumf_tw);
    Lambda_tw[i] = ph->dt_i_tw * ph->dt_D / (1.0 + ph->dt_D * inv);
  }
}

/* ------------------------------------------------------------------ dislotungsten
 * phase_mechanical_plastic_dislotungsten.f90: kinetics :444-528, LpAndItsTangent :270-310, dotState :316-378, dependentState :384-405.
 * State: rho_mob | rho_dip | gamma_sl; dependent: Lambda_sl | tau_pass (the dislotwin slots with n_tw = 0).  b_sl, i_sl, p, q, B,
 * d_caron, the forest projection, mu, nu, kB T, Q_cl and D live in the dislotwin fields (dt_*), Q_s in dt_Q_sl, tau_Peierls in
 * dt_tau_0, the dipole switch in dt_omit_dipoles; h, w, omega, f_at and D_0 are the law's own (tu_*). */
static void tu_kinetics(const pt_t *p, const double *Mp, double *dg, double *ddg, double *tau_out) {
  const orc_phase_t *ph = p->ph; const double *rho_mob = p->state;
  const double *Lambda = DEP_LAMBDA_SL(p), *tau_pass = DEP_TAU_PASS(p);
  for (int i = 0; i < ph->n_sl; i++) {
    double tp = ddot(Mp, ph->P_nS_pos[i]), tn = ddot(Mp, ph->P_nS_neg[i]);
    double tau_eff = fmax(fmax(tp, tn) - tau_pass[i], 0.0);                       /* math_clip(left = 0) */
    if (tau_out) tau_out[i] = tau_eff;
    if (tau_eff > TOL_MATH_CHECK) {
      double sr = tau_eff / ph->dt_tau_0[i], srp = pow(sr, ph->dt_p[i]), srpm1 = pow(sr, ph->dt_p[i] - 1.0);
      double BR = ph->dt_Q_sl[i] / ph->dt_kB_T, effL = Lambda[i] - ph->tu_w[i];
      double t_n = ph->dt_b_sl[i] * exp(BR * pow(1.0 - srp, ph->dt_q[i])) / (ph->tu_omega[i] * effL);
      double t_k = effL * ph->dt_B[i] / (2.0 * ph->dt_b_sl[i] * tau_eff);
      dg[i] = rho_mob[i] * ph->dt_b_sl[i] * ph->tu_h[i] / (t_n + t_k) * ((tp > tn) ? 1.0 : -1.0);
      if (ddg) {
        double dtn = -1.0 * t_n * BR * ph->dt_p[i] * ph->dt_q[i] * pow(1.0 - srp, ph->dt_q[i] - 1.0) * srpm1 / ph->dt_tau_0[i];
        double dtk = -1.0 * t_k / tau_eff;
        ddg[i] = -1.0 * dg[i] * (dtn + dtk) / (t_n + t_k);
      }
    } else { dg[i] = 0.0; if (ddg) ddg[i] = 0.0; }
  }
}
static void tu_Lp(const pt_t *p, const double *Mp, double *Lp, double *dLp_dMp) {
  const orc_phase_t *ph = p->ph;
  double dg[ORC_MAX_SL], ddg[ORC_MAX_SL];
  memset(Lp, 0, 9*sizeof(double)); memset(dLp_dMp, 0, 81*sizeof(double));
  tu_kinetics(p, Mp, dg, ddg, NULL);
  for (int i = 0; i < ph->n_sl; i++) {
    const double *PnS = (dg[i] > 0.0) ? ph->P_nS_pos[i] : ph->P_nS_neg[i];
    for (int a = 0; a < 9; a++) Lp[a] += dg[i] * ph->P_sl[i][a];
    for (int a = 0; a < 9; a++) for (int b = 0; b < 9; b++) dLp_dMp[9*a+b] += ddg[i] * ph->P_sl[i][a] * PnS[b];
  }
}
static void tu_dotState(const pt_t *p, const double *Mp, double *dot) {
  const orc_phase_t *ph = p->ph; int ns = ph->n_sl;
  const double *rho_mob = p->state, *rho_dip = p->state + ns, *Lambda = DEP_LAMBDA_SL(p);
  double *dot_mob = dot, *dot_dip = dot + ns, *dot_gamma = dot + 2*ns;
  double mu = ph->dt_mu, nu = ph->dt_nu, kT = ph->dt_kB_T, tau_eff[ORC_MAX_SL];
  tu_kinetics(p, Mp, dot_gamma, NULL, tau_eff);
  for (int i = 0; i < ns; i++) dot_gamma[i] = fabs(dot_gamma[i]);
  for (int i = 0; i < ns; i++) {
    double form, climb;
    if (deq0(dot_gamma[i])) { form = 0.0; climb = 0.0; }
    else {
      double d_hat = mu * ph->dt_b_sl[i] / (8.0 * PI_ * (1.0 - nu) * tau_eff[i]);
      d_hat = fmax(d_hat, ph->dt_d_caron[i]);            /* math_clip: left = d_caron, then right = Lambda_sl (math.f90:1316-1329) */
      d_hat = fmin(d_hat, Lambda[i]);
      form = ph->dt_omit_dipoles ? 0.0 : dot_gamma[i] * 2.0 * (d_hat - ph->dt_d_caron[i]) / ph->dt_b_sl[i] * rho_mob[i];
      double v_cl = (3.0 * mu * ph->tu_D_0 * exp(-ph->dt_Q_cl / kT) * ph->tu_f_at[i] / (2.0 * PI_ * kT)) * (1.0 / (d_hat + ph->dt_d_caron[i]));
      climb = (4.0 * v_cl * rho_dip[i]) / (d_hat - ph->dt_d_caron[i]);
    }
    dot_mob[i] = dot_gamma[i] / (ph->dt_b_sl[i] * Lambda[i]) - form - dot_gamma[i] * 2.0 * ph->dt_d_caron[i] / ph->dt_b_sl[i] * rho_mob[i];
    dot_dip[i] = form - climb - dot_gamma[i] * 2.0 * ph->dt_d_caron[i] / ph->dt_b_sl[i] * rho_dip[i];
  }
}
static void tu_dependentState(pt_t *p) {
  const orc_phase_t *ph = p->ph; int ns = ph->n_sl;
  const double *rho_mob = p->state, *rho_dip = p->state + ns;
  double *Lambda = DEP_LAMBDA_SL(p), *tau_pass = DEP_TAU_PASS(p);
  for (int i = 0; i < ns; i++) {
    double hp = 0, s = 0;
    for (int j = 0; j < ns; j++) { hp += ph->h_sl_sl[i][j] * (rho_mob[j] + rho_dip[j]); s += ph->dt_forest[i][j] * (rho_mob[j] + rho_dip[j]); }
    tau_pass[i] = ph->dt_mu * ph->dt_b_sl[i] * sqrt(hp);
    double inv = 1.0 / ph->dt_D + sqrt(s) / ph->dt_i_sl[i];
    Lambda[i] = pow(inv, -1.0);
  }
}

/* ------------------------------------------------------------------ dispatchers */
static void dependentState(pt_t *p) { if (p->ph->law == 3) dt_dependentState(p); else if (p->ph->law == 5) tu_dependentState(p); }

static void Mandel(const double *Fi, const double *S, double *Mp, double *FiTFi) {
  double FiT[9]; tr3(Fi, FiT); mm(FiT, Fi, FiTFi); mm(FiTFi, S, Mp);
}
static void plastic_Lp_tangents(const pt_t *p, const double *S, const double *Fi, double *Lp, double *dLp_dS, double *dLp_dFi) {
  if (p->ph->law == 0) { memset(Lp, 0, 9*sizeof(double)); memset(dLp_dS, 0, 81*sizeof(double)); memset(dLp_dFi, 0, 81*sizeof(double)); return; }
  double Mp[9], FiTFi[9], dM[81];
  Mandel(Fi, S, Mp, FiTFi);
  if (p->ph->law == 1) iso_Lp(p, Mp, Lp, dM); else if (p->ph->law == 2) pheno_Lp(p, Mp, Lp, dM);
  else if (p->ph->law == 4) kh_Lp(p, Mp, Lp, dM); else if (p->ph->law == 5) tu_Lp(p, Mp, Lp, dM); else dt_Lp(p, Mp, Lp, dM);
  double FiS[9]; mm(Fi, S, FiS);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    double d[9], dT[9], a[9], b[9];
    for (int k = 0; k < 9; k++) d[k] = dM[9*(3*i+j)+k];
    tr3(d, dT); mm(FiS, dT, a); mm(Fi, d, b); mm(b, S, b);
    for (int k = 0; k < 9; k++) dLp_dFi[9*(3*i+j)+k] = a[k] + b[k];
    mm(FiTFi, d, a);
    for (int k = 0; k < 9; k++) dLp_dS[9*(3*i+j)+k] = a[k];
  }
}
static void Li_tangents(const pt_t *p, const double *S, const double *Fi, double *Li, double *dLi_dS, double *dLi_dFi) {
  memset(Li, 0, 9*sizeof(double)); memset(dLi_dS, 0, 81*sizeof(double)); memset(dLi_dFi, 0, 81*sizeof(double));
  if (p->ph->law != 1) return;                                   /* only isotropic contributes (active) */
  double myLi[9], my_d[81];
  iso_Li(p, S, myLi, my_d);
  for (int a = 0; a < 9; a++) Li[a] += myLi[a];
  for (int a = 0; a < 81; a++) dLi_dS[a] += my_d[a];
  double FiInv[9], t[9], tmp[9]; inv33(Fi, FiInv); double detFi = det3(Fi);
  mm(Fi, Li, t); mm(t, FiInv, Li); for (int a = 0; a < 9; a++) Li[a] *= detFi;
  mm(FiInv, Li, tmp);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    double d[9], a[9];
    for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) d[3*k+l] = T4(dLi_dS,k,l,i,j);
    mm(Fi, d, a); mm(a, FiInv, a);
    for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) T4(dLi_dS,k,l,i,j) = a[3*k+l] * detFi;
    for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) T4(dLi_dFi,k,l,i,j) += Li[3*k+l] * FiInv[3*j+i];
    for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) T4(dLi_dFi,k,i,l,j) += I3[3*k+l] * tmp[3*j+i] + Li[3*k+l] * FiInv[3*j+i];
  }
}
static void dotState(const pt_t *p, double *dot) {
  if (p->ph->law == 0) return;
  double Mp[9], FiTFi[9]; Mandel(p->Fi, p->S, Mp, FiTFi);
  if (p->ph->law == 1) iso_dotState(p, Mp, dot); else if (p->ph->law == 2) pheno_dotState(p, Mp, dot);
  else if (p->ph->law == 4) kh_dotState(p, Mp, dot); else if (p->ph->law == 5) tu_dotState(p, Mp, dot); else dt_dotState(p, Mp, dot);
}
/* plastic_deltaState (phase_mechanical_plastic.f90:372-413): instantaneous state change after a converged state integration;
 * .false. (0) for every law but kinehardening (nonlocal is not restated); returns broken */
static int deltaState(pt_t *p) {
  if (p->ph->law != 4) return 0;
  double Mp[9], FiTFi[9]; Mandel(p->Fi, p->S, Mp, FiTFi);
  return kh_deltaState(p, Mp);
}
static int n_dot_of(const orc_phase_t *ph) { return ph->n_dot > 0 ? ph->n_dot : ph->n_state; }

/* ------------------------------------------------------------------ integrateStress; returns broken */
typedef struct { long n_Lp_iter, n_Li_iter, n_stress_calls, n_state_iter, n_substeps, n_cutbacks; } orc_counters_t;

static int integrateStress(pt_t *p, const double *F, const double *Fp0, const double *Fi0, double dt, orc_counters_t *cnt) {
  const orc_num_t *num = p->num;
  double Lpguess[9], Liguess[9], invFp_cur[9], invFi_cur[9], A[9], B[9], Fe[9], S[9];
  double invFi_new[9], Fi_new[9], Lpguess_old[9], Liguess_old[9], resLp[9], resLp_old[9], resLi[9], resLi_old[9];
  double deltaLp[9] = {0}, deltaLi[9] = {0}, Lp_c[9], Li_c[9], t1[9], t2[9];
  double dS_dFe[81], dS_dFi[81], dLp_dS[81], dLp_dFi[81], dLi_dS[81], dLi_dFi[81], dFe_dLp[81], J[81], tmp[81];
  cnt->n_stress_calls++;
  dependentState(p);
  memcpy(Lpguess, p->Lp, sizeof Lpguess); memcpy(Liguess, p->Li, sizeof Liguess);
  if (invert33(Fp0, invFp_cur)) return 1;
  if (invert33(Fi0, invFi_cur)) return 1;
  mm(F, invFp_cur, A);
  int jacoLi = 0; double stepLi = 1.0; memset(resLi_old, 0, sizeof resLi_old); memcpy(Liguess_old, Liguess, sizeof Liguess);
  int nLi = 0;
  for (;;) {                                                         /* LiLoop */
    nLi++; cnt->n_Li_iter++;
    if (nLi > num->N_iter_Li_max) return 1;
    for (int a = 0; a < 9; a++) t1[a] = I3[a] - dt * Liguess[a];
    mm(invFi_cur, t1, invFi_new);
    inv33(invFi_new, Fi_new);
    int jacoLp = 0; double stepLp = 1.0; memset(resLp_old, 0, sizeof resLp_old); memcpy(Lpguess_old, Lpguess, sizeof Lpguess);
    int nLp = 0;
    for (;;) {                                                       /* LpLoop */
      nLp++; cnt->n_Lp_iter++;
      if (nLp > num->N_iter_Lp_max) return 1;
      for (int a = 0; a < 9; a++) B[a] = I3[a] - dt * Lpguess[a];
      mm(A, B, t1); mm(t1, invFi_new, Fe);
      hooke(p, Fe, Fi_new, S, dS_dFe, dS_dFi);
      plastic_Lp_tangents(p, S, Fi_new, Lp_c, dLp_dS, dLp_dFi);
      double atol = fmax(num->eps_rel_Lp * fmax(norm2_9(Lpguess), norm2_9(Lp_c)), num->eps_abs_Lp);
      for (int a = 0; a < 9; a++) resLp[a] = Lpguess[a] - Lp_c[a];
      if (any_nan(resLp, 9)) return 1;
      else if (norm2_9(resLp) < atol) break;
      else if (nLp == 1 || norm2_9(resLp) < norm2_9(resLp_old)) {
        memcpy(resLp_old, resLp, sizeof resLp); memcpy(Lpguess_old, Lpguess, sizeof Lpguess); stepLp = 1.0;
      } else {
        stepLp = num->r_linesearch_Lp * stepLp;
        for (int a = 0; a < 9; a++) Lpguess[a] = Lpguess_old[a] + deltaLp[a] * stepLp;
        continue;
      }
      if (jacoLp % num->f_update_jacobi_Lp == 0) {
        jacoLp++;
        double invFiT[9]; tr3(invFi_new, invFiT);
        for (int o = 0; o < 3; o++) for (int q = 0; q < 3; q++) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
          T4(dFe_dLp,o,i,q,j) = -dt * A[3*o+q] * invFiT[3*i+j];
        mul99(dLp_dS, dS_dFe, tmp); mul99(tmp, dFe_dLp, J);
        for (int a = 0; a < 81; a++) J[a] = -J[a];
        for (int a = 0; a < 9; a++) J[10*a] += 1.0;
        double rhs[9]; memcpy(rhs, resLp, sizeof rhs);
        if (solve_nn(J, 9, rhs)) return 1;
        for (int a = 0; a < 9; a++) deltaLp[a] = -rhs[a];
      }
      for (int a = 0; a < 9; a++) Lpguess[a] += deltaLp[a] * stepLp;
    }
    Li_tangents(p, S, Fi_new, Li_c, dLi_dS, dLi_dFi);
    double atolLi = fmax(num->eps_rel_Li * fmax(norm2_9(Liguess), norm2_9(Li_c)), num->eps_abs_Li);
    for (int a = 0; a < 9; a++) resLi[a] = Liguess[a] - Li_c[a];
    if (any_nan(resLi, 9)) return 1;
    else if (norm2_9(resLi) < atolLi) break;
    else if (nLi == 1 || norm2_9(resLi) < norm2_9(resLi_old)) {
      memcpy(resLi_old, resLi, sizeof resLi); memcpy(Liguess_old, Liguess, sizeof Liguess); stepLi = 1.0;
    } else {
      stepLi = num->r_linesearch_Li * stepLi;
      for (int a = 0; a < 9; a++) Liguess[a] = Liguess_old[a] + deltaLi[a] * stepLi;
      continue;
    }
    if (jacoLi % num->f_update_jacobi_Li == 0) {
      jacoLi++;
      double dFe_dLi[81], dFi_dLi[81], u[81], v[81];
      mm(A, B, t1); mm(t1, invFi_cur, t2);
      memset(dFe_dLi, 0, sizeof dFe_dLi); memset(dFi_dLi, 0, sizeof dFi_dLi);
      for (int o = 0; o < 3; o++) for (int q = 0; q < 3; q++) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        T4(dFe_dLi,i,o,j,q) = -dt * I3[3*o+q] * t2[3*i+j];
        T4(dFi_dLi,i,o,j,q) = -dt * I3[3*o+q] * invFi_cur[3*i+j];
      }
      for (int o = 0; o < 3; o++) for (int q = 0; q < 3; q++) {
        double d[9], a[9];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) d[3*i+j] = T4(dFi_dLi,i,j,o,q);
        mm(Fi_new, d, a); mm(a, Fi_new, a);
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T4(dFi_dLi,i,j,o,q) = a[3*i+j];
      }
      mul99(dS_dFe, dFe_dLi, u); mul99(dS_dFi, dFi_dLi, v);
      for (int a = 0; a < 81; a++) u[a] += v[a];
      mul99(dLi_dS, u, v); mul99(dLi_dFi, dFi_dLi, u);
      for (int a = 0; a < 81; a++) J[a] = -v[a] - u[a];
      for (int a = 0; a < 9; a++) J[10*a] += 1.0;
      double rhs[9]; memcpy(rhs, resLi, sizeof rhs);
      if (solve_nn(J, 9, rhs)) return 1;
      for (int a = 0; a < 9; a++) deltaLi[a] = -rhs[a];
    }
    for (int a = 0; a < 9; a++) Liguess[a] += deltaLi[a] * stepLi;
  }
  double invFp_new[9], Fp_new[9];
  mm(invFp_cur, B, invFp_new);
  if (invert33(invFp_new, Fp_new)) return 1;
  double invFpT[9]; tr3(invFp_new, invFpT);
  mm(F, invFp_new, t1); mm(S, invFpT, t2); mm(t1, t2, p->P);
  memcpy(p->S, S, sizeof S); memcpy(p->Lp, Lpguess, sizeof Lpguess); memcpy(p->Li, Liguess, sizeof Liguess);
  double sc = pow(det3(Fp_new), 1.0/3.0);
  for (int a = 0; a < 9; a++) p->Fp[a] = Fp_new[a] / sc;
  memcpy(p->Fi, Fi_new, sizeof Fi_new);
  mm(t1, invFi_new, p->Fe);
  return 0;
}

/* ------------------------------------------------------------------ integrateStateFPI; returns broken */
static double damper(const double *o0, const double *o1, const double *o2, int n) {
  double d12 = 0, d22 = 0, d01 = 0;
  for (int i = 0; i < n; i++) { d12 += (o0[i]-o1[i])*(o1[i]-o2[i]); d22 += (o1[i]-o2[i])*(o1[i]-o2[i]); d01 += o0[i]*o1[i]; }
  if (fmin(d01, d12) < 0.0 && d22 > 0.0) return 0.75 + 0.25 * tanh(2.0 + 4.0 * d12 / d22);
  return 1.0;
}
static int integrateStateFPI(pt_t *p, const double *F, const double *Fp0, const double *Fi0, const double *state0, double dt, orc_counters_t *cnt) {
  const orc_phase_t *ph = p->ph; int n = n_dot_of(ph);             /* 1:sizeDotState */
  double dot[ORC_MAX_STATE], last1[ORC_MAX_STATE], last2[ORC_MAX_STATE], r[ORC_MAX_STATE];
  int broken = 1;
  memset(dot, 0, sizeof dot); memset(last1, 0, sizeof last1); memset(last2, 0, sizeof last2);
  dotState(p, dot);
  if (any_nan(dot, n)) return 1;
  for (int i = 0; i < n; i++) p->state[i] = state0[i] + dot[i] * dt;
  for (int it = 1; it <= p->num->N_iter_state_max; it++) {
    cnt->n_state_iter++;
    for (int i = 0; i < n; i++) { last2[i] = (it > 1) ? last1[i] : 0.0; last1[i] = dot[i]; }
    broken = integrateStress(p, F, Fp0, Fi0, dt, cnt);
    if (broken) break;
    dotState(p, dot);
    if (any_nan(dot, n)) break;                      /* NB broken stays .false. here, as in the reference (:660) */
    double zeta = damper(dot, last1, last2, n);
    int conv = 1;
    for (int i = 0; i < n; i++) {
      dot[i] = dot[i] * zeta + last1[i] * (1.0 - zeta);
      r[i] = p->state[i] - state0[i] - dot[i] * dt;
      p->state[i] -= r[i];
    }
    for (int i = 0; i < n; i++) if (!(fabs(r[i]) <= fmax(ph->atol[i], p->num->eps_rel_state * fabs(p->state[i])))) conv = 0;
    if (conv) { broken = deltaState(p); break; }     /* :661-664 */
  }
  return broken;
}

/* ------------------------------------------------------------------ the other state integrators; return broken */
static int state_converged(const pt_t *p, const double *r, int n) {                 /* phase.f90:626-638 */
  for (int i = 0; i < n; i++) if (!(fabs(r[i]) <= fmax(p->ph->atol[i], p->num->eps_rel_state * fabs(p->state[i])))) return 0;
  return 1;
}
/* integrateStateEuler, phase_mechanical.f90:699-735 */
static int integrateStateEuler(pt_t *p, const double *F, const double *Fp0, const double *Fi0, const double *state0, double dt, orc_counters_t *cnt) {
  int n = n_dot_of(p->ph); double dot[ORC_MAX_STATE]; memset(dot, 0, sizeof dot);
  dotState(p, dot);
  if (any_nan(dot, n)) return 1;
  for (int i = 0; i < n; i++) p->state[i] = state0[i] + dot[i] * dt;
  if (deltaState(p)) return 1;
  return integrateStress(p, F, Fp0, Fi0, dt, cnt);
}
/* integrateStateAdaptiveEuler, :738-787 */
static int integrateStateAdaptiveEuler(pt_t *p, const double *F, const double *Fp0, const double *Fi0, const double *state0, double dt, orc_counters_t *cnt) {
  int n = n_dot_of(p->ph); double dot[ORC_MAX_STATE], r[ORC_MAX_STATE]; memset(dot, 0, sizeof dot);
  dotState(p, dot);
  if (any_nan(dot, n)) return 1;
  for (int i = 0; i < n; i++) { r[i] = -dot[i] * 0.5 * dt; p->state[i] = state0[i] + dot[i] * dt; }
  if (deltaState(p)) return 1;
  if (integrateStress(p, F, Fp0, Fi0, dt, cnt)) return 1;
  dotState(p, dot);
  if (any_nan(dot, n)) return 1;
  for (int i = 0; i < n; i++) r[i] = r[i] + 0.5 * dot[i] * dt;
  return !state_converged(p, r, n);
}
/* integrateStateRK, :862-932: explicit Runge-Kutta with Butcher tableau A (ns x ns, A[stage][n]), weights B (ns + 1),
   nodes C (ns), optional embedded error weights DB (ns + 1).  NB the stage deformation gradient is F_0 + (F - F_0) dt C(stage)
   -- times dt, as the reference has it (:902). */
static int integrateStateRK(pt_t *p, const double *F_0, const double *F, const double *Fp0, const double *Fi0, const double *state0,
                            double dt, orc_counters_t *cnt, int ns, const double *A, const double *B, const double *C, const double *DB) {
  int n = n_dot_of(p->ph); double dot[ORC_MAX_STATE], rk[6][ORC_MAX_STATE], Fs[9];
  int broken = 1;
  memset(dot, 0, sizeof dot);
  dotState(p, dot);
  if (any_nan(dot, n)) return 1;
  for (int stage = 0; stage < ns; stage++) {
    memcpy(rk[stage], dot, sizeof(double) * n);
    for (int i = 0; i < n; i++) dot[i] = A[stage * ns + 0] * rk[0][i];
    for (int m = 1; m <= stage; m++) for (int i = 0; i < n; i++) dot[i] += A[stage * ns + m] * rk[m][i];
    for (int i = 0; i < n; i++) p->state[i] = state0[i] + dot[i] * dt;
    for (int a = 0; a < 9; a++) Fs[a] = F_0[a] + (F[a] - F_0[a]) * dt * C[stage];
    broken = integrateStress(p, Fs, Fp0, Fi0, dt * C[stage], cnt);
    if (broken) break;
    dotState(p, dot);
    if (any_nan(dot, n)) break;
  }
  if (broken) return 1;
  memcpy(rk[ns], dot, sizeof(double) * n);
  for (int i = 0; i < n; i++) { double s = 0.0; for (int m = 0; m <= ns; m++) s += rk[m][i] * B[m]; dot[i] = s; }
  for (int i = 0; i < n; i++) p->state[i] = state0[i] + dot[i] * dt;
  if (DB) {
    double r[ORC_MAX_STATE];
    for (int i = 0; i < n; i++) { double s = 0.0; for (int m = 0; m <= ns; m++) s += rk[m][i] * DB[m]; r[i] = s * dt; }
    if (!state_converged(p, r, n)) return 1;
  }
  if (deltaState(p)) return 1;
  return integrateStress(p, F, Fp0, Fi0, dt, cnt);
}
static const double RK4_A[9] = {0.5, 0, 0,  0, 0.5, 0,  0, 0, 1.0};
static const double RK4_C[3] = {0.5, 0.5, 1.0};
static const double RK4_B[4] = {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0};
static const double CK_A[25] = {1. / 5., 0, 0, 0, 0,   3. / 40., 9. / 40., 0, 0, 0,   3. / 10., -9. / 10., 6. / 5., 0, 0,
                                -11. / 54., 5. / 2., -70. / 27., 35. / 27., 0,
                                1631. / 55296., 175. / 512., 575. / 13824., 44275. / 110592., 253. / 4096.};
static const double CK_C[5] = {0.2, 0.3, 0.6, 1.0, 0.875};
static const double CK_B[6] = {37. / 378., 0, 250. / 621., 125. / 594., 0, 512. / 1771.};
static const double CK_DB[6] = {37. / 378. - 2825. / 27648., 0, 250. / 621. - 18575. / 48384., 125. / 594. - 13525. / 55296.,
                                0 - 277. / 14336., 512. / 1771. - 1. / 4.};
static int integrateState(pt_t *p, const double *F_0, const double *F, const double *Fp0, const double *Fi0, const double *state0,
                          double dt, orc_counters_t *cnt) {
  switch (p->num->integrator_state) {
    case 1: return integrateStateEuler(p, F, Fp0, Fi0, state0, dt, cnt);
    case 2: return integrateStateAdaptiveEuler(p, F, Fp0, Fi0, state0, dt, cnt);
    case 3: return integrateStateRK(p, F_0, F, Fp0, Fi0, state0, dt, cnt, 3, RK4_A, RK4_B, RK4_C, NULL);
    case 4: return integrateStateRK(p, F_0, F, Fp0, Fi0, state0, dt, cnt, 5, CK_A, CK_B, CK_C, CK_DB);
    default: return integrateStateFPI(p, F, Fp0, Fi0, state0, dt, cnt);
  }
}

/* ------------------------------------------------------------------ phase_mechanical_constitutive; returns converged */
static int constitutive(pt_t *p, double Dt, orc_counters_t *cnt) {
  const orc_phase_t *ph = p->ph; const orc_num_t *num = p->num; int ns = ph->n_state;
  double state0[ORC_MAX_STATE], Li0[9], Lp0[9], Fp0[9], Fi0[9], F0[9], F[9];
  memcpy(state0, p->state0, sizeof(double)*(ns > 0 ? ns : 1));
  memcpy(Li0, p->Li0, sizeof Li0); memcpy(Lp0, p->Lp0, sizeof Lp0); memcpy(Fp0, p->Fp0, sizeof Fp0);
  memcpy(Fi0, p->Fi0, sizeof Fi0); memcpy(F0, p->F0, sizeof F0);
  double stepFrac = 0.0, step = 1.0 / num->r_cutback; int converged = 0, todo = 1;
  while (todo) {
    if (converged) {
      stepFrac += step;
      step = fmin(1.0 - stepFrac, num->r_increase * step);
      todo = step > 0.0;
      if (todo) {
        memcpy(F0, F, sizeof F0); memcpy(Lp0, p->Lp, sizeof Lp0); memcpy(Li0, p->Li, sizeof Li0);
        memcpy(Fp0, p->Fp, sizeof Fp0); memcpy(Fi0, p->Fi, sizeof Fi0); memcpy(state0, p->state, sizeof(double)*ns);
      }
    } else {
      step = num->r_cutback * step;
      if (step < 1.0) cnt->n_cutbacks++;
      memcpy(p->Fp, Fp0, sizeof Fp0); memcpy(p->Fi, Fi0, sizeof Fi0); memcpy(p->S, p->S0, 9*sizeof(double));
      if (step < 1.0) { memcpy(p->Lp, Lp0, sizeof Lp0); memcpy(p->Li, Li0, sizeof Li0); }
      memcpy(p->state, state0, sizeof(double)*ns);
      todo = step > num->r_cutback_min;
    }
    if (todo) {
      cnt->n_substeps++;
      for (int a = 0; a < 9; a++) F[a] = F0[a] + step * (p->F[a] - p->F0[a]);
      converged = !integrateState(p, F0, F, Fp0, Fi0, state0, step * Dt, cnt);
    }
  }
  return converged;
}

/* ------------------------------------------------------------------ phase_mechanical_dPdF */
static void dPdF_point(pt_t *p, double Dt, double *dPdF) {
  double dev[9], dSdFe[81], dSdFi[81], dLidS[81], dLidFi[81], dLpdS[81], dLpdFi[81], dFidS[81];
  double lhs[81], rhs[81], tmp[81], t99[81], dSdF[81], dFpinvdF[81];
  double invFp[9], invFi[9], invSubFp0[9], invSubFi0[9], t1[9], t2[9], t3[9];
  hooke(p, p->Fe, p->Fi, dev, dSdFe, dSdFi);
  Li_tangents(p, p->S, p->Fi, dev, dLidS, dLidFi);
  inv33(p->Fp, invFp); inv33(p->Fi, invFi); inv33(p->Fp0, invSubFp0); inv33(p->Fi0, invSubFi0);
  double sabs = 0; for (int a = 0; a < 81; a++) sabs += fabs(dLidS[a]);
  if (sabs < TOL_MATH_CHECK) memset(dFidS, 0, sizeof dFidS);
  else {
    memset(lhs, 0, sizeof lhs); memset(rhs, 0, sizeof rhs);
    for (int o = 0; o < 3; o++) for (int q = 0; q < 3; q++) {
      double d[9], a[9];
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) d[3*i+j] = T4(dLidFi,i,j,o,q);
      mm(invSubFi0, d, a);
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T4(lhs,i,j,o,q) += a[3*i+j] * Dt;
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T4(lhs,i,o,j,q) += invFi[3*i+j] * invFi[3*q+o];
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) d[3*i+j] = T4(dLidS,i,j,o,q);
      mm(invSubFi0, d, a);
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T4(rhs,i,j,o,q) -= a[3*i+j] * Dt;
    }
    if (invert_nn(lhs, 9, t99)) memset(dFidS, 0, sizeof dFidS);
    else mul99(t99, rhs, dFidS);
    mul99(dLidFi, dFidS, tmp);
    for (int a = 0; a < 81; a++) dLidS[a] += tmp[a];
  }
  plastic_Lp_tangents(p, p->S, p->Fi, dev, dLpdS, dLpdFi);
  mul99(dLpdFi, dFidS, tmp);
  for (int a = 0; a < 81; a++) dLpdS[a] += tmp[a];
  /* dSdF */
  mm(invFp, invFi, t1); tr3(t1, t1);
  mm(p->F, invSubFp0, t2);
  mm(p->F, invFp, t3); mm(t3, invSubFi0, t3);
  for (int o = 0; o < 3; o++) for (int q = 0; q < 3; q++) {
    double d[9], a[9], b[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) d[3*i+j] = T4(dSdFe,q,o,i,j);
    mm(d, t1, a);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T4(rhs,q,o,i,j) = a[3*i+j];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) d[3*i+j] = T4(dLpdS,i,j,q,o);
    mm(t2, d, a); mm(a, invFi, a);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) d[3*i+j] = T4(dLidS,i,j,q,o);
    mm(t3, d, b);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T4(tmp,i,j,q,o) = a[3*i+j] + b[3*i+j];
  }
  mul99(dSdFe, tmp, lhs);
  mul99(dSdFi, dFidS, t99);
  for (int a = 0; a < 81; a++) lhs[a] = lhs[a] * Dt + t99[a];
  for (int a = 0; a < 9; a++) lhs[10*a] += 1.0;
  if (invert_nn(lhs, 9, t99)) memcpy(dSdF, rhs, sizeof rhs);
  else mul99(t99, rhs, dSdF);
  /* dFpinvdF */
  mul99(dLpdS, dSdF, tmp);
  for (int o = 0; o < 3; o++) for (int q = 0; q < 3; q++) {
    double d[9], a[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) d[3*i+j] = T4(tmp,i,j,q,o);
    mm(d, invFi, a); mm(invSubFp0, a, a);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T4(dFpinvdF,i,j,q,o) = -a[3*i+j] * Dt;
  }
  /* assemble */
  double invFpT[9]; tr3(invFp, invFpT);
  mm(p->S, invFpT, t1);
  mm(p->F, invFp, t2);
  mm(t2, p->S, t3);
  memset(dPdF, 0, 81*sizeof(double));
  { double a[9]; mm(invFp, t1, a); tr3(a, a);
    for (int q = 0; q < 3; q++) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T4(dPdF,q,i,q,j) = a[3*i+j]; }
  for (int o = 0; o < 3; o++) for (int q = 0; q < 3; q++) {
    double d[9], dT[9], a[9], b[9], c[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) d[3*i+j] = T4(dFpinvdF,i,j,q,o);
    mm(p->F, d, a); mm(a, t1, a);
    tr3(d, dT); mm(t3, dT, c);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) d[3*i+j] = T4(dSdF,i,j,q,o);
    mm(t2, d, b); mm(b, invFpT, b);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T4(dPdF,i,j,q,o) += a[3*i+j] + b[3*i+j] + c[3*i+j];
  }
}

/* ================================================================== exported entry points */
/* F_in: [n][9] row-major per point (the ctypes wrapper converts from the Fortran [3][3][N] boundary layout) */
/* optional per-cell log of sub-step cut-backs (test diagnostics: cells whose Newton path is rounding-chaotic) */
static int *g_point_cutbacks = NULL;
void orc_set_point_log(int *buf) { g_point_cutbacks = buf; }
/* optional per-cell log of Lp Newton iterations (test diagnostics: cells where a stopping / line-search decision flipped) */
static int *g_point_lp_iter = NULL;
void orc_set_point_lp_log(int *buf) { g_point_lp_iter = buf; }

/* threads of the cell loops below (the reference's own loop is serial per MPI rank, homogenization.f90:233; upstream DAMASK's is
 * an OpenMP loop): 0 or less = the OpenMP default.  Used by bench.py's CPU arm to report a 1-thread figure next to the all-core one */
#ifdef _OPENMP
#include <omp.h>
static int g_omp_default = 0;
void orc_set_threads(int n) { if (!g_omp_default) g_omp_default = omp_get_max_threads(); omp_set_num_threads(n > 0 ? n : g_omp_default); }
int orc_get_threads(void) { return omp_get_max_threads(); }
#else
void orc_set_threads(int n) { (void)n; }
int orc_get_threads(void) { return 1; }
#endif

int orc_mechanical_response(const orc_phase_t *phases, const orc_num_t *num, orc_points_t *st, const double *F_in,
                            double Dt, int cell_start, int cell_end, long *counters) {
  int ill = 0;
  long c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(|:ill) reduction(+:c0,c1,c2,c3,c4,c5)
  for (int ce = cell_start; ce < cell_end; ce++) {
    pt_t p = point_view(phases, num, st, ce);
    orc_counters_t cnt; memset(&cnt, 0, sizeof cnt);
    int ns = p.ph->n_state;
    /* phase_restore(ce, includeL=.false.) */
    memcpy(p.Fp, p.Fp0, 9*sizeof(double)); memcpy(p.Fi, p.Fi0, 9*sizeof(double)); memcpy(p.S, p.S0, 9*sizeof(double));
    memcpy(p.state, p.state0, sizeof(double)*ns);
    memcpy(p.F, F_in + 9*(size_t)ce, 9*sizeof(double));             /* phase_set_F */
    if (!constitutive(&p, Dt, &cnt)) ill |= 1;
    if (g_point_cutbacks) g_point_cutbacks[ce] = (int)cnt.n_cutbacks;
    if (g_point_lp_iter) g_point_lp_iter[ce] = (int)cnt.n_Lp_iter;
    c0 += cnt.n_Lp_iter; c1 += cnt.n_Li_iter; c2 += cnt.n_stress_calls; c3 += cnt.n_state_iter; c4 += cnt.n_substeps; c5 += cnt.n_cutbacks;
  }
  if (counters) { counters[0] = c0; counters[1] = c1; counters[2] = c2; counters[3] = c3; counters[4] = c4; counters[5] = c5; }
  return ill;
}

/* P_out [n][9], dPdF_out [n][81] (row-major [i][j][k][l]); dPdF_out may be NULL */
void orc_mechanical_response2(const orc_phase_t *phases, const orc_num_t *num, orc_points_t *st, double Dt,
                              int cell_start, int cell_end, double *P_out, double *dPdF_out) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int ce = cell_start; ce < cell_end; ce++) {
    pt_t p = point_view(phases, num, st, ce);
    memcpy(P_out + 9*(size_t)ce, p.P, 9*sizeof(double));
    if (dPdF_out) dPdF_point(&p, Dt, dPdF_out + 81*(size_t)ce);
  }
}

void orc_forward(const orc_phase_t *phases, orc_points_t *st) {           /* mechanical_forward */
  for (int ce = 0; ce < st->n; ce++) {
    pt_t p = point_view(phases, NULL, st, ce);
    memcpy(p.Fi0, p.Fi, 72); memcpy(p.Fp0, p.Fp, 72); memcpy(p.F0, p.F, 72);
    memcpy(p.Li0, p.Li, 72); memcpy(p.Lp0, p.Lp, 72); memcpy(p.S0, p.S, 72);
    memcpy(p.state0, p.state, sizeof(double)*p.ph->n_state);
  }
}

void orc_init_dependent(const orc_phase_t *phases, orc_points_t *st) {    /* crystallite_init: phase.f90:564 */
  for (int ce = 0; ce < st->n; ce++) { pt_t p = point_view(phases, NULL, st, ce); dependentState(&p); }
}

int orc_sizeof_phase(void) { return (int)sizeof(orc_phase_t); }
int orc_sizeof_num(void) { return (int)sizeof(orc_num_t); }
