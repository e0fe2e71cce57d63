// TEST INFRASTRUCTURE: compiles the product's per-point device functions for the HOST so that a failing GPU parity case
// can be replayed point by point on the CPU build box (which has no GPU) and diffed against the oracle.
// Never linked into libdamask_grid_b200.so.
#include <cstring>
#include <vector>
#include "../../solver_cpp_b200/csrc/constitutive_dev.cuh"
#include "../../solver_cpp_b200/csrc/machine_dev.cuh"
#include "../../solver_cpp_b200/csrc/tangent_dev.cuh"
#include "../../solver_cpp_b200/csrc/pheno_fast_dev.cuh"
#include "../../solver_cpp_b200/csrc/pheno_lanes_dev.cuh"

void dmk_set_error(const char*, ...) {}
static int g_use_nested = 0;     // 1: the straight nested-loop formulation (cross-check of the state machine)
extern "C" void twin_use_nested(int v) { g_use_nested = v; }

template <int LAW, int NS, int ND>
static int run_point(const DmkDevPhase* ph, const DmkNumDev& num, const double* F0, const double* Ft, const double* Fp0,
                     const double* Lp0, double* Lp, const double* S0, const double* state0, double* dep, double Dt,
                     double* P, double* Fp, double* Fe, double* S, double* state, unsigned* cnt_out, double* dPdF) {
  static DmkMachine<LAW, NS, ND> mach;                // the product path: the resumable state machine
  static DmkMachineStore<LAW, NS, ND> mstore;
  machine_bind(mach, mstore);
  DmkPt<LAW, NS, ND>& p = mach.p;
  p.ph = ph; p.C66 = ph->p.C66;
  double f0[9], ft[9], fp0[9], lp0[9], s0[6];
  for (int a = 0; a < 9; a++) { f0[a] = F0[a]; ft[a] = Ft[a]; fp0[a] = Fp0[a]; lp0[a] = Lp0[a]; p.Lp[a] = Lp[a]; p.Fp[a] = Fp0[a]; p.Fe[a] = Fe[a]; p.P[a] = 0; }
  for (int a = 0; a < 6; a++) { s0[a] = S0[a]; p.S[a] = S0[a]; }
  for (int k = 0; k < ph->p.n_state; k++) p.st[k] = state0[k];
  for (int k = 0; k < ph->n_dep; k++) p.dep[k] = dep[k];
  DmkCnt cnt = {0, 0, 0, 0, 0, 0};
  bool ok;
  if (g_use_nested || num.integrator != DMK_INTEG_FPI) ok = constitutive_point(p, num, f0, ft, fp0, lp0, s0, state0, Dt, cnt);
  else {
    for (int a = 0; a < 9; a++) { mach.Ft[a] = Ft[a]; mach.F0in[a] = F0[a]; mach.Fp0[a] = Fp0[a]; mach.Lp0[a] = Lp0[a]; }
    for (int a = 0; a < 6; a++) mach.S0in[a] = S0[a];
    for (int k = 0; k < ph->p.n_state; k++) mach.state0[k] = state0[k];
    ok = machine_run_point(mach, num, Dt, cnt);
  }
  for (int a = 0; a < 9; a++) { P[a] = p.P[a]; Fp[a] = p.Fp[a]; Fe[a] = p.Fe[a]; Lp[a] = p.Lp[a]; }
  for (int a = 0; a < 6; a++) S[a] = p.S[a];
  for (int k = 0; k < ph->p.n_state; k++) state[k] = p.st[k];
  for (int k = 0; k < ph->n_dep; k++) dep[k] = p.dep[k];
  cnt_out[0] = cnt.Lp_iter; cnt_out[1] = cnt.Li_iter; cnt_out[2] = cnt.stress_calls; cnt_out[3] = cnt.state_iter; cnt_out[4] = cnt.substeps; cnt_out[5] = cnt.cutbacks;
  if (dPdF) tangent_point(p, ft, fp0, Dt, dPdF);
  return ok ? 1 : 0;
}

static int g_use_fast = 0;       // 1: route eligible phases through pheno_fast_dev.cuh; 2: through its lane schedule (pheno_lanes_dev.cuh)
extern "C" void twin_use_fast(int v) { g_use_fast = v; }
static int g_defer = 0;          // pf_stress deferral threshold (0: never defer), as DMK_PHENO_DEFER in the product
static int g_last_deferred = 0;
extern "C" void twin_set_defer(int v) { g_defer = v; }
extern "C" int twin_last_deferred() { return g_last_deferred; }
// the register-resident slip-only phenopowerlaw formulation (pheno_fast_dev.cuh) on a one-point DmkPointArrays
template <int NSL, int NEXP>
static int run_point_fast(const DmkDevPhase* ph, const DmkNumDev& num, const double* F0, const double* Ft, const double* Fp0,
                          const double* Lp0, double* Lp, const double* S0, const double* state0, double Dt,
                          double* P, double* Fp, double* Fe, double* S, double* state, unsigned* cnt_out) {
  static DmkPhenoK<NSL> K;
  dmk_pheno_fast_tables<NSL>(*ph, K);
  DmkPointArrays pa;
  memset(&pa, 0, sizeof pa);
  pa.n = 1;
  double F[9], f0[9], fp0[9], lp0[9], s0[6], st0[2 * NSL], Frec[9], Prec[9];
  int cut = 0;
  for (int a = 0; a < 9; a++) { f0[a] = F0[a]; fp0[a] = Fp0[a]; lp0[a] = Lp0[a]; F[a] = 0; }
  for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) Frec[a + 3 * b] = Ft[3 * a + b];
  for (int a = 0; a < 6; a++) s0[a] = S0[a];
  for (int k = 0; k < 2 * NSL; k++) st0[k] = state0[k];
  pa.F0 = f0; pa.F = F; pa.Fp0 = fp0; pa.Fp = Fp; pa.Fe = Fe; pa.Lp0 = lp0; pa.Lp = Lp; pa.S0 = s0; pa.S = S; pa.state0 = st0; pa.state = state;
  pa.cutbacks = &cut;
  int lpi = 0;
  pa.lp_iter = &lpi;
  double scratch[PfOff<NSL>::TOTAL], cold[PfOff<NSL>::COLD_TOTAL];
  DmkScr scr = {scratch, 1, cold, 1, &K};
  DmkCnt cnt = {0, 0, 0, 0, 0, 0};
  int rc;
  if (g_use_fast == 2 && dmk_pheno_lanes_eligible(num, g_defer == 0))
    rc = (g_defer == 0) ? pl_run_point<NSL, NEXP, true>(K, num, pa, 0, Frec, Prec, Dt, scr, cnt)
                        : pl_run_point<NSL, NEXP, false>(K, num, pa, 0, Frec, Prec, Dt, scr, cnt);
  else rc = pf_point<NSL, NEXP, 1>(K, num, g_defer, pa, 0, Frec, Prec, Dt, scr, cnt);
  if (rc == PF_DEFER) return 2;                       // nothing written: the caller replays the point on the reference-order path
  for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) P[3 * a + b] = Prec[a + 3 * b];
  cnt_out[0] = cnt.Lp_iter; cnt_out[1] = cnt.Li_iter; cnt_out[2] = cnt.stress_calls; cnt_out[3] = cnt.state_iter; cnt_out[4] = cnt.substeps; cnt_out[5] = cnt.cutbacks;
  return rc == PF_OK ? 1 : 0;
}

static void voigt_strain_of(const double* P, double* pv) { pv[0] = P[0]; pv[1] = P[4]; pv[2] = P[8]; pv[3] = P[5] + P[7]; pv[4] = P[2] + P[6]; pv[5] = P[1] + P[3]; }
static void q_host(const double* C66, const double* pv, double* Q) { for (int b = 0; b < 6; b++) { double s = 0; for (int a = 0; a < 6; a++) s += pv[a] * C66[6 * a + b]; Q[b] = s; } }

extern "C" int twin_point(const dmk_phase* phase, const dmk_numerics* n, const double* F0, const double* Ft, const double* Fp0,
                          const double* Lp0, double* Lp, const double* S0, const double* state0, double* dep, double Dt,
                          double* P, double* Fp, double* Fe, double* S, double* state, unsigned* cnt, double* dPdF) {
  static DmkDevPhase d;
  memset(&d, 0, sizeof d);
  d.p = *phase;
  const dmk_phase& p = *phase;
  bool nons = false;
  for (int a = 0; a < p.n_sl; a++) {
    const double* pos = (p.law == DMK_LAW_DISLOTWIN) ? p.P_sl[a] : p.P_nS_pos[a];
    voigt_strain_of(pos, d.pv_pos[a]); q_host(p.C66, d.pv_pos[a], d.Q_pos[a]);
    voigt_strain_of(p.P_nS_neg[a], d.pv_neg[a]); q_host(p.C66, d.pv_neg[a], d.Q_neg[a]);
    for (int k = 0; k < 9; k++) if (p.P_nS_neg[a][k] != -p.P_nS_pos[a][k]) nons = true;
  }
  for (int a = 0; a < p.n_tw; a++) { voigt_strain_of(p.P_tw[a], d.pv_tw[a]); q_host(p.C66, d.pv_tw[a], d.Q_tw[a]); }
  auto small_int = [](double v) { return (v >= 1.0 && v <= 64.0 && v == floor(v)) ? (int)v : 0; };
  for (int a = 0; a < p.n_sl; a++) d.n_sl_int[a] = small_int(p.n_sl_exp[a]);
  for (int a = 0; a < p.n_tw; a++) d.n_tw_int[a] = small_int(p.n_tw_exp[a]);
  d.iso_n_int = small_int(p.iso_n);
  d.nonschmid = nons; d.n_dep = (p.law == DMK_LAW_DISLOTWIN) ? 2 * p.n_sl + p.n_tw : (p.law == DMK_LAW_DISLOTUNGSTEN) ? 2 * p.n_sl : 0;
  DmkNumDev nd;
  nd.r_cutback_min = n->r_cutback_min; nd.r_cutback = n->r_cutback; nd.r_increase = n->r_increase; nd.eps_rel_state = n->eps_rel_state;
  nd.N_iter_state_max = n->N_iter_state_max; nd.N_iter_Lp_max = n->N_iter_Lp_max; nd.r_linesearch_Lp = n->r_linesearch_Lp;
  nd.eps_rel_Lp = n->eps_rel_Lp; nd.eps_abs_Lp = n->eps_abs_Lp; nd.f_update_jacobi_Lp = n->f_update_jacobi_Lp > 0 ? n->f_update_jacobi_Lp : 1;
  nd.integrator = n->integrator_state;
  int nexp = 0;
  g_last_deferred = 0;
  if (g_use_fast && nd.integrator == DMK_INTEG_FPI && dmk_pheno_fast_eligible(d, 12, &nexp)) {
    int ok = (nexp == 20) ? run_point_fast<12, 20>(&d, nd, F0, Ft, Fp0, Lp0, Lp, S0, state0, Dt, P, Fp, Fe, S, state, cnt)
                          : run_point_fast<12, 0>(&d, nd, F0, Ft, Fp0, Lp0, Lp, S0, state0, Dt, P, Fp, Fe, S, state, cnt);
    if (ok == 2) { g_last_deferred = 1; goto generic; }
    if (dPdF) {
      static DmkPt<DMK_LAW_PHENOPOWERLAW, 2 * DMK_MAX_SYS, 0> p;
      static DmkPtStore<DMK_LAW_PHENOPOWERLAW, 2 * DMK_MAX_SYS, 0> ps;
      dmk_pt_bind(p, ps);
      p.ph = &d; p.C66 = d.p.C66;
      double ft[9], fp0[9];
      for (int a = 0; a < 9; a++) { ft[a] = Ft[a]; fp0[a] = Fp0[a]; p.Fp[a] = Fp[a]; p.Fe[a] = Fe[a]; p.Lp[a] = Lp[a]; p.P[a] = P[a]; }
      for (int a = 0; a < 6; a++) p.S[a] = S[a];
      for (int k = 0; k < d.p.n_state; k++) p.st[k] = state[k];
      tangent_point(p, ft, fp0, Dt, dPdF);
    }
    return ok;
  }
generic:
#define ARGS &d, nd, F0, Ft, Fp0, Lp0, Lp, S0, state0, dep, Dt, P, Fp, Fe, S, state, cnt, dPdF
  switch (p.law) {
    case DMK_LAW_NONE: return run_point<DMK_LAW_NONE, 0, 0>(ARGS);
    case DMK_LAW_ISOTROPIC: return run_point<DMK_LAW_ISOTROPIC, 1, 0>(ARGS);
    case DMK_LAW_PHENOPOWERLAW: return run_point<DMK_LAW_PHENOPOWERLAW, 2 * DMK_MAX_SYS, 0>(ARGS);
    case DMK_LAW_KINEHARDENING: return run_point<DMK_LAW_KINEHARDENING, DMK_MAX_STATE, 0>(ARGS);
    case DMK_LAW_DISLOTUNGSTEN: return run_point<DMK_LAW_DISLOTUNGSTEN, 3 * DMK_MAX_SL, 2 * DMK_MAX_SL>(ARGS);
    default: return run_point<DMK_LAW_DISLOTWIN, DMK_MAX_STATE, DMK_MAX_DEP>(ARGS);
  }
}
