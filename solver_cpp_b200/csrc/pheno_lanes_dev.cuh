// Lane-persistent schedule of the register-resident slip-only phenopowerlaw update (pheno_fast_dev.cuh).
//
// Same arithmetic as pf_point / pf_stress, i.e. the reference's
//   phase_mechanical_constitutive   /root/reference/src/phase_mechanical.f90:1013-1100   (first, full-size sub-step only)
//   integrateStateFPI               phase_mechanical.f90:611-693
//   integrateStress                 phase_mechanical.f90:380-604
// cut into three resumable pieces so that a warp never waits for its slowest point: with one thread walking one point
// through the nested state / Newton loops, ncu showed 9.8 of 32 lanes active per executed instruction on a running
// solve's iterate (profiles/r01_summary.md) -- the points of a warp need 2..8 Newton iterations per stress integration
// and 3..10 stress integrations, and the warp pays the maximum at every level.  Here every lane owns one point at a
// time and a trip of the kernel's loop is
//     refill      lanes without a point take the next cell of the warp's chunk           (pl_begin)
//     state block lanes that just got a point / whose Newton loop just converged        (pl_state_block)
//     Newton      ONE Lp Newton iteration for every lane that is inside a Newton loop    (pl_lp_iteration)
// so the cost of a warp is the sum, not the nested maximum, of its lanes' work.
//
// A sub-step that FAILS (the Newton loop runs out of iterations, a non-positive pivot, a singular matrix, a NaN) -- what
// the reference answers with a cut-back -- hands the point, untouched, to the reference-order kernel (constitutive.cu),
// as pf_point does with deferral enabled; that regime is rounding-chaotic (tests/test_oracle_golden.py) and rare.  The
// line search of the Newton loop (every third point of a running solve's iterate needs it) runs here.  The lane schedule
// exists for a full first sub-step (r_cutback * (1 / r_cutback) >= 1) and a Jacobian update every iteration.
//
// Live across trips: A = F Fp0^-1, the Lp iterate and the last Newton step (registers), the last accepted Lp iterate
// (scratch slot LOLD), the state-integrator arrays (scratch, as in pf_point).  B, Fe, S are recomputed from (A, Lp) and
// Fp0^-1 from Fp0 in the state block instead of being carried.
#pragma once
#include "pheno_fast_dev.cuh"

#define PL_NEED 0        // no point
#define PL_INIT 1        // point loaded: state predictor pending
#define PL_LP 2          // inside an Lp Newton loop
#define PL_CONV 3        // Newton loop converged: Fp, P, state corrector pending
#define PL_NEWTON 5      // residual accepted: Newton step pending (three-way vote only)
#define PL_EXIT 4        // no work left

#define PL_CONTINUE 0
#define PL_DONE 1
#define PL_DEFER 2

// re-convergence point of the lanes of `group` (device); nothing on the host twin, which runs one point at a time
DMK_HD __forceinline__ void pl_converge(unsigned group) {
#ifdef __CUDA_ARCH__
  __syncwarp(group);
#else
  (void)group;
#endif
}

template <int NSL>
struct PlLane {
  double A[9], Lp[9];
  double delta[9];       // last Newton step (the line search shortens it)
  double res_old;        // residual norm of the last accepted Newton iterate
  double step;           // line-search factor of the running step
  long long i;           // cell
  int mode;
  int nLp;               // Newton iterations of the running stress integration
  int it;                // state iteration (= stress integrations started for this point)
  unsigned nLp_pt;       // Newton iterations of the completed stress integrations of this point
};

// Loads point i and sets up the (only) sub-step [F0 -> F0 + step (F - F0)].  S <- S0.  False: singular Fp0 (defer).
template <int NSL>
DMK_HD __forceinline__ bool pl_begin(const DmkPointArrays& pa, long long i, const double* __restrict__ Fin, double step,
                                     const DmkScr& scr, PlLane<NSL>& L, double (&S)[6]) {
  typedef PfOff<NSL> O;
  const long long n = pa.n;
  double F[9], Fp0[9], invFp0[9];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int cc = 0; cc < 3; cc++) {
      const int a = 3 * r + cc;
      const double f0 = pa.F0[a * n + i];
      F[a] = f0 + step * (Fin[i * 9 + r + 3 * cc] - f0);                  // boundary record: c = i + 3 j
    }
#pragma unroll
  for (int a = 0; a < 9; a++) { Fp0[a] = pa.Fp0[a * n + i]; L.Lp[a] = pa.Lp[a * n + i]; }
  if (invert33r(Fp0, invFp0)) return false;
  mm3(F, invFp0, L.A);
#pragma unroll
  for (int a = 0; a < 6; a++) S[a] = pa.S0[a * n + i];
  {                                                   // unrolled, pointer bumped: these two copies run at ~4 of 32 lanes
    const double* __restrict__ src = pa.state0 + i;
#pragma unroll
    for (int k = 0; k < 2 * NSL; k++) { scr(O::ST + k) = *src; src += n; }
  }
  pf_rxi<NSL>(scr);                                   // 1 / xi of state0, for the predictor's dotState
  L.i = i; L.it = 0; L.nLp = 0; L.nLp_pt = 0; L.res_old = 0.0; L.step = 1.0;
  return true;
}

// integrateStress' Lp loop (pf_stress), one trip = pl_residual [+ pl_newton].
// pl_residual: residual at L.Lp and the loop's decisions.  PL_DONE: converged at the current L.Lp; PL_CONTINUE: the last
// step was shortened (line search), evaluate the residual again; PL_STEP: iterate accepted, L.delta holds the residual
// and a Newton step is due (pl_newton).  The state block sets nLp = 0, step = 1 before a loop.
#define PL_STEP 3
template <int NSL, int NEXP, int UNR>
DMK_HD __forceinline__ int pl_residual(const DmkPhenoK<NSL>& K, const DmkNumDev& num, double dt, const DmkScr& scr, PlLane<NSL>& L) {
  typedef PfOff<NSL> O;
  L.nLp++;
  if (L.nLp > num.N_iter_Lp_max) return PL_DEFER;
  double B[9], Fe[9], S[6];
#pragma unroll
  for (int a = 0; a < 9; a++) B[a] = ((a % 4 == 0) ? 1.0 : 0.0) - dt * L.Lp[a];
  mm3(L.A, B, Fe);
  hooke_S(K.C66, Fe, S);
  double res[9];
#pragma unroll
  for (int a = 0; a < 9; a++) res[a] = 0.0;
#pragma unroll (unr_rate(UNR))
  for (int s = 0; s < NSL; s++) {
    double dg, c;
    pf_rate<NSL, NEXP>(K, s, S, scr(O::RXI + s), dg, c);
    scr(O::C + s) = dt * c;
#pragma unroll
    for (int a = 0; a < 9; a++) res[a] += dg * K.P[s][a];
  }
  const double atol = fmax(num.eps_rel_Lp * fmax(norm9(L.Lp), norm9(res)), num.eps_abs_Lp);
#pragma unroll
  for (int a = 0; a < 9; a++) res[a] = L.Lp[a] - res[a];
  const double rn = norm9(res);
  if (any_nan9(res)) return PL_DEFER;
  if (rn < atol) return PL_DONE;
  if (L.nLp == 1 || rn < L.res_old) {
    L.res_old = rn;
#pragma unroll
    for (int a = 0; a < 9; a++) { scr(O::LOLD + a) = L.Lp[a]; L.delta[a] = res[a]; }   // the old step is dead: delta carries the residual
    L.step = 1.0;
    return PL_STEP;
  }
  L.step = num.r_linesearch_Lp * L.step;
#pragma unroll
  for (int a = 0; a < 9; a++) L.Lp[a] = scr(O::LOLD + a) + L.delta[a] * L.step;
  return PL_CONTINUE;
}

// Newton step from the accepted iterate L.Lp with residual L.delta: delta <- -(dR/dLp)^-1 R, Lp += delta.
template <int NSL, int UNR>
DMK_HD __forceinline__ int pl_newton(const DmkPhenoK<NSL>& K, double dt, const DmkScr& scr, PlLane<NSL>& L) {
  double B[9], Fe[9], res[9];
#pragma unroll
  for (int a = 0; a < 9; a++) { B[a] = ((a % 4 == 0) ? 1.0 : 0.0) - dt * L.Lp[a]; res[a] = L.delta[a]; }
  mm3(L.A, B, Fe);
  if (pf_newton_step<NSL, UNR>(K, L.A, Fe, scr, res)) return PL_DEFER;
#pragma unroll
  for (int a = 0; a < 9; a++) { L.delta[a] = -res[a]; L.Lp[a] += L.delta[a] * L.step; }
  return PL_CONTINUE;
}

template <int NSL, int NEXP, int UNR>
DMK_HD __forceinline__ int pl_lp_iteration(const DmkPhenoK<NSL>& K, const DmkNumDev& num, double dt, const DmkScr& scr,
                                           PlLane<NSL>& L) {
  const int rc = pl_residual<NSL, NEXP, UNR>(K, num, dt, scr, L);
  return rc == PL_STEP ? pl_newton<NSL, UNR>(K, dt, scr, L) : rc;
}

// State block of integrateStateFPI.  init: predictor (state = state0 + dotState(S0) dt).  Else: S of the stress
// integration that just converged, dotState, damper, corrector, convergence.  PL_CONTINUE: a Newton loop follows (1/xi
// refreshed); PL_DONE: the point is finished and stored; PL_DEFER: nothing stored.
// Predictor and corrector lanes run ONE copy of every loop (the predictor is the corrector with last(1) = last(2) = 0,
// zeta = 1 and state == state0, bit for bit); `group` = the warp's lanes inside this call, re-converged before each
// shared stretch -- with separate branches the compiler duplicated dotState and the loops and ran them once per branch,
// at 3.8 and 17 of 32 lanes (ncu source view, profiles/r01_summary.md).
template <int NSL, int NEXP>
DMK_HD __forceinline__ int pl_state_block(const DmkPhenoK<NSL>& K, const DmkNumDev& num, const DmkPointArrays& pa,
                                          const double* __restrict__ Fin, double* __restrict__ Pout, double dt, const DmkScr& scr,
                                          PlLane<NSL>& L, double (&S)[6], bool init, unsigned group) {
  typedef PfOff<NSL> O;
  constexpr int NST = 2 * NSL;
  const long long n = pa.n, i = L.i;
  double B[9], Fe[9];
  bool defer = false;
  if (!init) {
    L.nLp_pt += (unsigned)L.nLp;
#pragma unroll
    for (int a = 0; a < 9; a++) B[a] = ((a % 4 == 0) ? 1.0 : 0.0) - dt * L.Lp[a];
    // The reference forms Fp and P after every stress integration (phase_mechanical.f90:586-604) and fails the sub-step if
    // Fp_new^-1 = Fp0^-1 B is singular; only the last Fp and P survive, so they are formed once, when the point is stored.
    // Fp0^-1 is regular (pl_begin), so the singularity test is made on B here.
    defer = deq0(det3(B));
    mm3(L.A, B, Fe);
    hooke_S(K.C66, Fe, S);
  }
  pl_converge(group);
  const bool nan = pf_dot_state<NSL, NEXP>(K, S, scr);
  // reference: a NaN rate fails the predictor's sub-step (cut back: defer), but ENDS the corrector loop with
  // broken = .false. (:660): the point is stored with the state of the previous iterate
  defer = defer || (init && nan);
  bool finished = !init && nan;
  double zeta = 1.0;
  if (!init && !nan) {
    double d12 = 0.0, d22 = 0.0, d01 = 0.0;           // damper (:677-690); D = last(1) - last(2)
#pragma unroll 1
    for (int k = 0; k < NST; k++) {
      const double l1 = scr(O::L1 + k), dd = scr(O::D + k), dk = pf_dot<NSL>(scr, k);
      d12 += (dk - l1) * dd;
      d22 += dd * dd;
      d01 += dk * l1;
    }
    if (fmin(d01, d12) < 0.0 && d22 > 0.0) zeta = 0.75 + 0.25 * tanh(2.0 + 4.0 * d12 / d22);
  }
  pl_converge(group);
  if (!nan) {
    bool conv = true;
#pragma unroll 1
    for (int k = 0; k < NST; k++) {
      const double l1 = init ? 0.0 : scr(O::L1 + k);
      const double d = pf_dot<NSL>(scr, k) * zeta + l1 * (1.0 - zeta);
      scr(O::D + k) = d - l1;
      scr(O::L1 + k) = d;
      double s = scr(O::ST + k);
      const double r = s - pa.state0[k * n + i] - d * dt;       // predictor: s == state0, r = -d dt
      s -= r;
      scr(O::ST + k) = s;
      if (!(fabs(r) <= fmax(K.atol[k], num.eps_rel_state * fabs(s)))) conv = false;
    }
    finished = !init && (conv || L.it >= num.N_iter_state_max);   // a loop that runs out leaves broken = .false. (reference)
  }
  pl_converge(group);
  if (!finished && !defer) {
    L.it++; L.nLp = 0; L.step = 1.0;
    pf_rxi<NSL>(scr);
  }
  if (defer) return PL_DEFER;
  if (!finished) return PL_CONTINUE;
  // ---- Fp, P (phase_mechanical.f90:586-604; F invFp_new == A B == Fe) and store the point
  double P[9], Fp[9];
  {
    double Fp0[9], invFp0[9], invFp_new[9], Fp_new[9], Sm[9], t2[9];
#pragma unroll
    for (int a = 0; a < 9; a++) Fp0[a] = pa.Fp0[a * n + i];
    invert33r(Fp0, invFp0);
    mm3(invFp0, B, invFp_new);
    if (invert33r(invFp_new, Fp_new)) return PL_DEFER;
    Sm[0] = S[0]; Sm[1] = S[5]; Sm[2] = S[4]; Sm[3] = S[5]; Sm[4] = S[1]; Sm[5] = S[3]; Sm[6] = S[4]; Sm[7] = S[3]; Sm[8] = S[2];
    mm3_nt(Sm, invFp_new, t2);
    mm3(Fe, t2, P);
    const double rsc = 1.0 / cbrt(det3(Fp_new));
#pragma unroll
    for (int a = 0; a < 9; a++) Fp[a] = Fp_new[a] * rsc;
  }
  pa.cutbacks[i] = 0;
  pa.lp_iter[i] = (int)L.nLp_pt;
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int cc = 0; cc < 3; cc++) {
      pa.F[(3 * r + cc) * n + i] = Fin[i * 9 + r + 3 * cc];
      Pout[i * 9 + r + 3 * cc] = P[3 * r + cc];
    }
#pragma unroll
  for (int a = 0; a < 9; a++) { pa.Fp[a * n + i] = Fp[a]; pa.Lp[a * n + i] = L.Lp[a]; pa.Fe[a * n + i] = Fe[a]; }
#pragma unroll
  for (int a = 0; a < 6; a++) pa.S[a * n + i] = S[a];
  {
    double* __restrict__ dst = pa.state + i;
#pragma unroll
    for (int k = 0; k < NST; k++) { *dst = scr(O::ST + k); dst += n; }
  }
  return PL_DONE;
}

// the lane schedule exists for these numerics only (see the header); `step` = size of the first sub-step
inline bool dmk_pheno_lanes_eligible(const DmkNumDev& num, double* step) {
  *step = num.r_cutback * (1.0 / num.r_cutback);
  return *step >= 1.0 && num.f_update_jacobi_Lp == 1 && num.N_iter_state_max >= 1;
}

// Host twin / reference driver of the three pieces for one point (what one lane does).  PF_OK, PF_DEFER.
template <int NSL, int NEXP>
DMK_HD inline int pl_run_point(const DmkPhenoK<NSL>& K, const DmkNumDev& num, const DmkPointArrays& pa, long long i,
                               const double* Fin, double* Pout, double step, double Dt, const DmkScr& scr, DmkCnt& cnt) {
  PlLane<NSL> L;
  double S[6];
  const double dt = step * Dt;
  if (!pl_begin<NSL>(pa, i, Fin, step, scr, L, S)) return PF_DEFER;
  int rc = pl_state_block<NSL, NEXP>(K, num, pa, Fin, Pout, dt, scr, L, S, true, 0xffffffffu);
  while (rc == PL_CONTINUE) {
    do { rc = pl_lp_iteration<NSL, NEXP, 1>(K, num, dt, scr, L); } while (rc == PL_CONTINUE);
    if (rc == PL_DEFER) return PF_DEFER;
    rc = pl_state_block<NSL, NEXP>(K, num, pa, Fin, Pout, dt, scr, L, S, false, 0xffffffffu);
  }
  if (rc == PL_DEFER) return PF_DEFER;
  cnt.Lp_iter += L.nLp_pt; cnt.Li_iter += L.it; cnt.stress_calls += L.it; cnt.state_iter += L.it; cnt.substeps += 1;
  return PF_OK;
}
