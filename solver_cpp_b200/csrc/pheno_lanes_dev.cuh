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
// phase_mechanical_constitutive answers with a cut-back (:1040-1085) -- is handled in one of two ways (template CUTBACK):
//   false (default): the point is handed, untouched, to the reference-order kernel (constitutive.cu).  The cut-back regime
//          is rounding-chaotic (tests/test_oracle_golden.py); the reference-order kernel follows the oracle's path on most
//          such cells (> 70 % within 1e-9 on the rough Cu test), this formulation on 22 % -- so by default they go there.
//   true  (DMK_PHENO_CUTBACK=lanes): cut back and repeated in-lane -- the lane re-enters the refill block with the same
//          cell (pl_substep_begin); sub-step start values are the *0 arrays until a partial sub-step has converged, then
//          they are parked in the output arrays F, Fp, Lp, state, as pf_point does (bit-identical to it on the host twin).
//          5.4x faster on an evaluation in which every cell cuts back (virgin Cu, 128^3: 214 ms against 1159 ms).
// The line search of the Newton loop runs in-lane in both.  The lane schedule exists for a Jacobian update every
// iteration (the default) and, with CUTBACK = false, a full first sub-step (r_cutback * (1 / r_cutback) >= 1).
//
// Live across trips: A = F Fp0^-1, the Lp iterate and the last Newton step / pending residual (registers), the last
// accepted Lp iterate (scratch slot LOLD), the state-integrator arrays (scratch, as in pf_point), the sub-step
// bookkeeping.  B, Fe, S are recomputed from (A, Lp) and Fp0^-1 from Fp0 in the state block instead of being carried.
#pragma once
#include "pheno_fast_dev.cuh"

#define PL_NEED 0        // no point
#define PL_RESTART 6     // has a point, a (first / next / cut-back) sub-step is to be set up
#define PL_INIT 1        // point loaded: state predictor pending
#define PL_LP 2          // inside an Lp Newton loop
#define PL_CONV 3        // Newton loop converged: Fp, P, state corrector pending
#define PL_NEWTON 5      // residual accepted: Newton step pending (three-way vote only)
#define PL_EXIT 4        // no work left

#define PL_CONTINUE 0
#define PL_DONE 1
#define PL_FAIL 2         // the running sub-step failed: cut back
#define PL_ILL 4          // sub-steps exhausted: point stored as terminally ill
#define PL_GIVEUP 7       // CUTBACK = false: the first sub-step failed; the point goes, untouched, to the reference-order kernel

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
  double delta[9];       // last Newton step (the line search shortens it); between pl_residual and pl_newton: the residual
  double res_old;        // residual norm of the last accepted Newton iterate
  double step;           // line-search factor of the running step
  double sstep, frac;    // sub-stepping (phase_mechanical.f90:1040-1085): size of the running sub-step, fraction done
  double dt;             // time step of the running sub-step
  long long i;           // cell
  int mode;
  int nLp;               // Newton iterations of the running stress integration
  int it;                // state iteration of the running sub-step
  unsigned nLp_pt, calls_pt, nsub, ncut;    // per-point counters: Newton iterations, stress integrations, sub-steps, cut-backs
  bool sub;              // sub-step start values are parked in pa.F / Fp / Lp / state
  bool ok;               // the previous sub-step converged (false: fresh point or failed sub-step)
};

// a lane takes cell i
template <int NSL>
DMK_HD __forceinline__ void pl_fetch(const DmkPointArrays& pa, const DmkNumDev& num, long long i, PlLane<NSL>& L) {
  const long long n = pa.n;
#pragma unroll
  for (int a = 0; a < 9; a++) L.Lp[a] = pa.Lp[a * n + i];          // phase_restore(includeL = .false.): Lp keeps the last guess
  L.i = i; L.sub = false; L.ok = false; L.frac = 0.0; L.sstep = 1.0 / num.r_cutback;
  L.nLp_pt = 0; L.calls_pt = 0; L.nsub = 0; L.ncut = 0; L.nLp = 0; L.it = 0;
}

// Sets up the next sub-step of the lane's point: after a failure (or for a fresh point) the cut-back branch of the
// reference's loop, then F, Fp0^-1, A, 1/xi.  PL_CONTINUE: the predictor follows; PL_FAIL: singular Fp0 (cut back again);
// PL_ILL: sub-step size exhausted, point stored.
template <int NSL, bool CUTBACK>
DMK_HD __forceinline__ int pl_substep_begin(const DmkNumDev& num, const DmkPointArrays& pa, const double* __restrict__ Fin,
                                            double* __restrict__ Pout, double Dt, const DmkScr& scr, PlLane<NSL>& L, double (&S)[6]) {
  typedef PfOff<NSL> O;
  const long long n = pa.n, i = L.i;
  if (!CUTBACK && L.nsub > 0) return PL_GIVEUP;       // the full first sub-step failed: not ours
  const bool sub = CUTBACK && L.sub;
  const double* F0s = sub ? pa.F : pa.F0;
  const double* Fp0s = sub ? pa.Fp : pa.Fp0;
  if (!L.ok) {
    L.sstep = num.r_cutback * L.sstep;
    if (CUTBACK && L.sstep < 1.0) {                   // (CUTBACK = false: the only sub-step set up here is the full first one)
      L.ncut++;
      const double* Lp0s = sub ? pa.Lp : pa.Lp0;
#pragma unroll
      for (int a = 0; a < 9; a++) L.Lp[a] = Lp0s[a * n + i];
    }
#pragma unroll
    for (int a = 0; a < 6; a++) S[a] = pa.S0[a * n + i];
    {                                                 // unrolled, pointer bumped: this copy runs at ~4 of 32 lanes
      const double* __restrict__ src = (sub ? pa.state : pa.state0) + i;
      constexpr int UC = PL_U_COPY;
#pragma unroll (UC)
      for (int k = 0; k < 2 * NSL; k++) { scr(O::ST + k) = *src; src += n; }
    }
    if (CUTBACK && !(L.sstep > num.r_cutback_min)) {  // terminally ill: what pf_point leaves behind (P of no converged sub-step: 0)
      pa.cutbacks[i] = (int)L.ncut;
      pa.lp_iter[i] = (int)L.nLp_pt;
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int cc = 0; cc < 3; cc++) { pa.F[(3 * r + cc) * n + i] = Fin[i * 9 + r + 3 * cc]; Pout[i * 9 + r + 3 * cc] = 0.0; }
#pragma unroll
      for (int a = 0; a < 9; a++) { const double f = Fp0s[a * n + i]; pa.Fp[a * n + i] = f; pa.Lp[a * n + i] = L.Lp[a]; }
#pragma unroll
      for (int a = 0; a < 6; a++) pa.S[a * n + i] = S[a];
#pragma unroll 1
      for (int k = 0; k < 2 * NSL; k++) pa.state[k * n + i] = scr(O::ST + k);
      return PL_ILL;
    }
  }
  L.nsub++;
  L.dt = L.sstep * Dt;
  double F[9], Fp0[9], invFp0[9];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int cc = 0; cc < 3; cc++) {
      const int a = 3 * r + cc;
      F[a] = F0s[a * n + i] + L.sstep * (Fin[i * 9 + r + 3 * cc] - pa.F0[a * n + i]);     // boundary record: c = i + 3 j
    }
#pragma unroll
  for (int a = 0; a < 9; a++) Fp0[a] = Fp0s[a * n + i];
  if (invert33r(Fp0, invFp0)) { L.ok = false; return PL_FAIL; }
  mm3(F, invFp0, L.A);
  pf_rxi<NSL>(scr);                                   // 1 / xi of the sub-step's start state, for the predictor's dotState
  L.it = 0; L.nLp = 0; L.res_old = 0.0; L.step = 1.0;
  return PL_CONTINUE;
}

// integrateStress' Lp loop (pf_stress), one trip = pl_residual [+ pl_newton].
// pl_residual: residual at L.Lp and the loop's decisions.  PL_DONE: converged at the current L.Lp; PL_CONTINUE: the last
// step was shortened (line search), evaluate the residual again; PL_STEP: iterate accepted, L.delta holds the residual
// and a Newton step is due (pl_newton).  The state block sets nLp = 0, step = 1 before a loop.
#define PL_STEP 3
template <int NSL, int NEXP, int UNR>
DMK_HD __forceinline__ int pl_residual(const DmkPhenoK<NSL>& K, const DmkNumDev& num, const DmkScr& scr, PlLane<NSL>& L) {
  typedef PfOff<NSL> O;
  const double dt = L.dt;
  L.nLp++;
  if (L.nLp > num.N_iter_Lp_max) return PL_FAIL;
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
    pf_rate<NSL, NEXP>(pf_tab<NSL>(scr), s, S, scr(O::RXI + s), dg, c);
    scr(O::C + s) = dt * c;
#pragma unroll
    for (int a = 0; a < 9; a++) res[a] += dg * pf_tab<NSL>(scr).P[s][a];
  }
  // the three Frobenius norms of the stopping test, their square roots scheduled into each other (pf_sqrt_fast)
  double n2[3] = {0.0, 0.0, 0.0}, nr[3];
#pragma unroll
  for (int a = 0; a < 9; a++) { n2[0] += L.Lp[a] * L.Lp[a]; n2[1] += res[a] * res[a]; }
#pragma unroll
  for (int a = 0; a < 9; a++) { res[a] = L.Lp[a] - res[a]; n2[2] += res[a] * res[a]; }
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 3; k++) { ok = ok & pf_fast_ok(n2[k]); nr[k] = pf_sqrt_fast(n2[k]); }
  if (!ok) {
#pragma unroll
    for (int k = 0; k < 3; k++) nr[k] = sqrt(n2[k]);
  }
  const double atol = fmax(num.eps_rel_Lp * fmax(nr[0], nr[1]), num.eps_abs_Lp);
  const double rn = nr[2];
  if (isnan(n2[2])) return PL_FAIL;                   // any_nan9(res)
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
DMK_HD __forceinline__ int pl_newton(const DmkPhenoK<NSL>& K, const DmkScr& scr, PlLane<NSL>& L) {
  const double dt = L.dt;
  double B[9], Fe[9], res[9];
#pragma unroll
  for (int a = 0; a < 9; a++) { B[a] = ((a % 4 == 0) ? 1.0 : 0.0) - dt * L.Lp[a]; res[a] = L.delta[a]; }
  mm3(L.A, B, Fe);
  if (pf_newton_step<NSL, UNR>(K, L.A, Fe, scr, res)) return PL_FAIL;
#pragma unroll
  for (int a = 0; a < 9; a++) { L.delta[a] = -res[a]; L.Lp[a] += L.delta[a] * L.step; }
  return PL_CONTINUE;
}

template <int NSL, int NEXP, int UNR>
DMK_HD __forceinline__ int pl_lp_iteration(const DmkPhenoK<NSL>& K, const DmkNumDev& num, const DmkScr& scr, PlLane<NSL>& L) {
  const int rc = pl_residual<NSL, NEXP, UNR>(K, num, scr, L);
  return rc == PL_STEP ? pl_newton<NSL, UNR>(K, scr, L) : rc;
}

// State block of integrateStateFPI.  init: predictor (state = state0 + dotState(S) dt).  Else: S of the stress integration
// that just converged, dotState, damper, corrector, convergence.  PL_CONTINUE: a Newton loop follows (1/xi refreshed);
// PL_RESTART: the sub-step converged and another one follows (its start values are parked); PL_DONE: the point is finished
// and stored; PL_FAIL: the sub-step failed (cut back).
// Predictor and corrector lanes run ONE copy of every loop (the predictor is the corrector with last(1) = last(2) = 0,
// zeta = 1 and state == state0, bit for bit); `group` = the warp's lanes inside this call, re-converged before each
// shared stretch -- with separate branches the compiler duplicated dotState and the loops and ran them once per branch,
// at 3.8 and 17 of 32 lanes (ncu source view, profiles/r01_summary.md).
template <int NSL, int NEXP, bool CUTBACK>
DMK_HD __forceinline__ int pl_state_block(const DmkPhenoK<NSL>& K, const DmkNumDev& num, const DmkPointArrays& pa,
                                          const double* __restrict__ Fin, double* __restrict__ Pout, const DmkScr& scr,
                                          PlLane<NSL>& L, double (&S)[6], bool init, unsigned group) {
  typedef PfOff<NSL> O;
  constexpr int NST = 2 * NSL;
  const long long n = pa.n, i = L.i;
  const double dt = L.dt;
  const bool sub = CUTBACK && L.sub;
  double B[9], Fe[9];
  bool fail = false;
  if (!init) {
    L.nLp_pt += (unsigned)L.nLp;
#pragma unroll
    for (int a = 0; a < 9; a++) B[a] = ((a % 4 == 0) ? 1.0 : 0.0) - dt * L.Lp[a];
    // The reference forms Fp and P after every stress integration (phase_mechanical.f90:586-604) and fails the sub-step if
    // Fp_new^-1 = Fp0^-1 B is singular; only the last Fp and P survive, so they are formed once per sub-step, below.
    // Fp0^-1 is regular (pl_substep_begin), so the singularity test is made on B here.
    fail = deq0(det3(B));
    mm3(L.A, B, Fe);
    hooke_S(K.C66, Fe, S);
  }
  pl_converge(group);
  const bool nan = pf_dot_state<NSL, NEXP>(K, S, scr);
  // reference: a NaN rate fails the predictor's sub-step (cut back), but ENDS the corrector loop with broken = .false.
  // (:660): the sub-step counts as converged with the state of the previous iterate
  fail = fail || (init && nan);
  bool finished = !init && nan;
  double zeta = 1.0;
  if (!init && !nan) {
    double d12 = 0.0, d22 = 0.0, d01 = 0.0;           // damper (:677-690); D = last(1) - last(2)
    // L1 and D (cold part of the scratch) are loaded PL_U_STATE entries at a time so that the loads of a batch are in flight
    // together; the sums run in the order k = 0, 1, 2, ... as before
    constexpr int UB = (NST % PL_U_STATE == 0) ? PL_U_STATE : 4;
    static_assert(NST % 4 == 0, "state count: batches of four");
#pragma unroll 1
    for (int k0 = 0; k0 < NST; k0 += UB) {
      double l1[UB], dd[UB], dk[UB];
#pragma unroll
      for (int j = 0; j < UB; j++) { l1[j] = scr.cold(O::L1 + k0 + j); dd[j] = scr.cold(O::D + k0 + j); dk[j] = pf_dot<NSL>(scr, k0 + j); }
#pragma unroll
      for (int j = 0; j < UB; j++) {
        d12 += (dk[j] - l1[j]) * dd[j];
        d22 += dd[j] * dd[j];
        d01 += dk[j] * l1[j];
      }
    }
    if (fmin(d01, d12) < 0.0 && d22 > 0.0) zeta = 0.75 + 0.25 * tanh(2.0 + 4.0 * d12 / d22);
  }
  pl_converge(group);
  if (!nan) {
    bool conv = true;
    const double* __restrict__ st0s = (sub ? pa.state : pa.state0) + i;
    constexpr int UB = (NST % PL_U_STATE == 0) ? PL_U_STATE : 4;
#pragma unroll 1
    for (int k0 = 0; k0 < NST; k0 += UB) {
      double l1[UB], s0[UB], dk[UB], sk[UB];
#pragma unroll
      for (int j = 0; j < UB; j++) s0[j] = st0s[(k0 + j) * n];                       // global: issued first, all in flight together
#pragma unroll
      for (int j = 0; j < UB; j++) { l1[j] = init ? 0.0 : scr.cold(O::L1 + k0 + j); dk[j] = pf_dot<NSL>(scr, k0 + j); sk[j] = scr(O::ST + k0 + j); }
#pragma unroll
      for (int j = 0; j < UB; j++) {
        const int k = k0 + j;
        const double d = dk[j] * zeta + l1[j] * (1.0 - zeta);
        scr.cold(O::D + k) = d - l1[j];
        scr.cold(O::L1 + k) = d;
        double s = sk[j];
        const double r = s - s0[j] - d * dt;          // predictor: s == state0, r = -d dt
        s -= r;
        scr(O::ST + k) = s;
        if (!(fabs(r) <= fmax(pf_tab<NSL>(scr).atol[k], num.eps_rel_state * fabs(s)))) conv = false;
      }
    }
    finished = !init && (conv || L.it >= num.N_iter_state_max);   // a loop that runs out leaves broken = .false. (reference)
  }
  pl_converge(group);
  if (!finished && !fail) {
    L.it++; L.calls_pt++; L.nLp = 0; L.step = 1.0;
    pf_rxi<NSL>(scr);
  }
  if (fail) return PL_FAIL;
  if (!finished) return PL_CONTINUE;
  // ---- the sub-step converged: Fp, P (phase_mechanical.f90:586-604; F invFp_new == A B == Fe)
  double P[9], Fp[9];
  {
    const double* Fp0s = sub ? pa.Fp : pa.Fp0;
    double Fp0[9], invFp0[9], invFp_new[9], Fp_new[9], Sm[9], t2[9];
#pragma unroll
    for (int a = 0; a < 9; a++) Fp0[a] = Fp0s[a * n + i];
    invert33r(Fp0, invFp0);
    mm3(invFp0, B, invFp_new);
    if (invert33r(invFp_new, Fp_new)) return PL_FAIL;
    Sm[0] = S[0]; Sm[1] = S[5]; Sm[2] = S[4]; Sm[3] = S[5]; Sm[4] = S[1]; Sm[5] = S[3]; Sm[6] = S[4]; Sm[7] = S[3]; Sm[8] = S[2];
    mm3_nt(Sm, invFp_new, t2);
    mm3(Fe, t2, P);
    const double rsc = 1.0 / cbrt(det3(Fp_new));
#pragma unroll
    for (int a = 0; a < 9; a++) Fp[a] = Fp_new[a] * rsc;
  }
  // ---- sub-step bookkeeping (:1043-1058): next sub-step from the values just reached, or done
  const double step_done = L.sstep;
  L.frac += L.sstep;
  L.sstep = fmin(1.0 - L.frac, num.r_increase * L.sstep);
  if (CUTBACK && L.sstep > 0.0) {                     // park the new sub-step start in the output arrays
    const double* F0s = sub ? pa.F : pa.F0;
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int cc = 0; cc < 3; cc++) {
        const int a = 3 * r + cc;
        pa.F[a * n + i] = F0s[a * n + i] + step_done * (Fin[i * 9 + r + 3 * cc] - pa.F0[a * n + i]);
      }
#pragma unroll
    for (int a = 0; a < 9; a++) { pa.Lp[a * n + i] = L.Lp[a]; pa.Fp[a * n + i] = Fp[a]; }
#pragma unroll 1
    for (int k = 0; k < NST; k++) pa.state[k * n + i] = scr(O::ST + k);
    L.sub = true; L.ok = true;
    return PL_RESTART;
  }
  pa.cutbacks[i] = (int)L.ncut;
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
    constexpr int UC = PL_U_COPY;
#pragma unroll (UC)
    for (int k = 0; k < NST; k++) { *dst = scr(O::ST + k); dst += n; }
  }
  return PL_DONE;
}

// the lane schedule exists for these numerics only (see the header)
inline bool dmk_pheno_lanes_eligible(const DmkNumDev& num, bool cutback) {
  return num.f_update_jacobi_Lp == 1 && num.N_iter_state_max >= 1 && (cutback || num.r_cutback * (1.0 / num.r_cutback) >= 1.0);
}

// Host twin / reference driver of the pieces for one point (what one lane does).  PF_OK, PF_BROKEN (terminally ill) or,
// with CUTBACK = false, PF_DEFER (nothing written).
template <int NSL, int NEXP, bool CUTBACK>
DMK_HD inline int pl_run_point(const DmkPhenoK<NSL>& K, const DmkNumDev& num, const DmkPointArrays& pa, long long i,
                               const double* Fin, double* Pout, double Dt, const DmkScr& scr, DmkCnt& cnt) {
  PlLane<NSL> L;
  double S[6];
  pl_fetch<NSL>(pa, num, i, L);
  int result = PF_OK;
  for (;;) {
    int rc = pl_substep_begin<NSL, CUTBACK>(num, pa, Fin, Pout, Dt, scr, L, S);
    if (rc == PL_GIVEUP) return PF_DEFER;
    if (rc == PL_ILL) { result = PF_BROKEN; break; }
    if (rc == PL_FAIL) continue;
    rc = pl_state_block<NSL, NEXP, CUTBACK>(K, num, pa, Fin, Pout, scr, L, S, true, 0xffffffffu);
    while (rc == PL_CONTINUE) {
      do { rc = pl_lp_iteration<NSL, NEXP, 1>(K, num, scr, L); } while (rc == PL_CONTINUE);
      if (rc == PL_FAIL) { L.nLp_pt += (unsigned)L.nLp; break; }
      rc = pl_state_block<NSL, NEXP, CUTBACK>(K, num, pa, Fin, Pout, scr, L, S, false, 0xffffffffu);
    }
    if (rc == PL_DONE) break;
    if (rc == PL_FAIL) L.ok = false;                  // cut back (or, CUTBACK = false, give up on the next visit)
  }
  cnt.Lp_iter += L.nLp_pt; cnt.Li_iter += L.calls_pt; cnt.stress_calls += L.calls_pt; cnt.state_iter += L.calls_pt;
  cnt.substeps += L.nsub; cnt.cutbacks += L.ncut;
  return result;
}
