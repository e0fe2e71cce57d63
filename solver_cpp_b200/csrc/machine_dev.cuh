// The per-point integrator of constitutive_dev.cuh unrolled into a resumable state machine, so that the 32 lanes of a
// warp stay busy although every material point needs a different number of Lp Newton iterations, state iterations and
// sub-steps: the nested loops of phase_mechanical_constitutive -> integrateStateFPI -> integrateStress
// (/root/reference/src/phase_mechanical.f90:1013-1100, 611-693, 380-604) reconverge only at their ends, which left 12
// of 32 lanes active on a polycrystal (ncu, profiles/).  Here one trip of the kernel's main loop is exactly ONE Lp
// Newton iteration for every lane that is inside an Lp loop; everything between two Lp loops (stress epilogue, state
// update, sub-step bookkeeping, storing a finished point and fetching the next) is a short transition executed by the
// lanes that need it.  A lane that finishes its point fetches another one at once (lane-persistent).
// The arithmetic and its order are those of the nested functions -- tests/host_twin checks bit equality on the host.
#pragma once
#include "constitutive_dev.cuh"

enum { M_FETCH = 0, M_SUBSTEP = 1, M_STRESS_PRO = 2, M_LP = 3, M_STRESS_EPI = 4, M_EXIT = 5 };

template <int LAW, int NS, int ND>
struct DmkMachine {
  DmkPt<LAW, NS, ND> p;
  int state;
  // the point's increment: start and target
  double Ft[9], F0in[9], S0in[6];
  // sub-step
  double F0[9], F[9], Fp0[9], Lp0[9];
  double stepFrac, step, dtsub;
  bool converged, todo;
  // state integration
  int it;
  // stress integration
  double Lpg[9], Lpg_old[9], delta[9], A[9], invFp[9], B[9], S[6];
  double res_old_norm, stepLp;
  int jaco, nLp;
  // dynamically indexed arrays, kept outside the aggregate (see DmkPt): pointers into a DmkMachineStore
  double *state0, *dot, *last1, *last2, *ddg;
};
template <int LAW, int NS, int ND>
struct DmkMachineStore {
  DmkPtStore<LAW, NS, ND> pt;
  double state0[NS > 0 ? NS : 1], dot[NS > 0 ? NS : 1], last1[NS > 0 ? NS : 1], last2[NS > 0 ? NS : 1];
  double ddg[2 * DMK_MAX_SYS];
};
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void machine_bind(DmkMachine<LAW, NS, ND>& m, DmkMachineStore<LAW, NS, ND>& s) {
  dmk_pt_bind(m.p, s.pt);
  m.state0 = s.state0; m.dot = s.dot; m.last1 = s.last1; m.last2 = s.last2; m.ddg = s.ddg;
}

// phase_mechanical_constitutive, set-up (phase_mechanical.f90:1030-1045); the caller has filled p.{ph,C66,Lp,Fe,dep}
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void machine_begin(DmkMachine<LAW, NS, ND>& m, const DmkNumDev& num) {
  const int ns = m.p.ph->p.n_state;
  for (int i = 0; i < ns; i++) m.p.st[i] = m.state0[i];
#pragma unroll
  for (int a = 0; a < 9; a++) { m.F0[a] = m.F0in[a]; m.F[a] = m.F0in[a]; m.p.Fp[a] = m.Fp0[a]; m.p.P[a] = 0.0; }
#pragma unroll
  for (int a = 0; a < 6; a++) m.p.S[a] = m.S0in[a];
  m.stepFrac = 0.0; m.step = 1.0 / num.r_cutback; m.converged = false; m.todo = true;
  m.state = M_SUBSTEP;
}

// one trip of the sub-step loop up to the start of the state integration (phase_mechanical.f90:1047-1095, 628-640)
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void machine_substep(DmkMachine<LAW, NS, ND>& m, const DmkNumDev& num, double Dt, DmkCnt& cnt) {
  DmkPt<LAW, NS, ND>& p = m.p;
  const int ns = p.ph->p.n_state;
  if (m.converged) {
    m.stepFrac += m.step;
    m.step = fmin(1.0 - m.stepFrac, num.r_increase * m.step);
    m.todo = m.step > 0.0;
    if (m.todo) {
#pragma unroll
      for (int a = 0; a < 9; a++) { m.F0[a] = m.F[a]; m.Lp0[a] = p.Lp[a]; m.Fp0[a] = p.Fp[a]; }
      for (int i = 0; i < ns; i++) m.state0[i] = p.st[i];
    }
  } else {
    m.step = num.r_cutback * m.step;
    if (m.step < 1.0) cnt.cutbacks++;
#pragma unroll
    for (int a = 0; a < 9; a++) p.Fp[a] = m.Fp0[a];
#pragma unroll
    for (int a = 0; a < 6; a++) p.S[a] = m.S0in[a];
    if (m.step < 1.0) {
#pragma unroll
      for (int a = 0; a < 9; a++) p.Lp[a] = m.Lp0[a];
    }
    for (int i = 0; i < ns; i++) p.st[i] = m.state0[i];
    m.todo = m.step > num.r_cutback_min;
  }
  if (!m.todo) { m.state = M_FETCH; return; }                       // point finished: m.converged is its verdict
  cnt.substeps++;
#pragma unroll
  for (int a = 0; a < 9; a++) m.F[a] = m.F0[a] + m.step * (m.Ft[a] - m.F0in[a]);
  m.dtsub = m.step * Dt;
  // integrateStateFPI, head (over the integrated part of the state, 1:sizeDotState)
  const int nd = n_dot_state(p);
  for (int i = 0; i < nd; i++) { m.dot[i] = 0.0; m.last1[i] = 0.0; m.last2[i] = 0.0; }
  dot_state(p, m.dot);
  bool nan = false;
  for (int i = 0; i < nd; i++) nan = nan || isnan(m.dot[i]);
  if (nan) { m.converged = false; m.state = M_SUBSTEP; return; }     // broken
  for (int i = 0; i < nd; i++) p.st[i] = m.state0[i] + m.dot[i] * m.dtsub;
  m.it = 1;
  m.state = M_STRESS_PRO;
}

// top of one state iteration + head of integrateStress (phase_mechanical.f90:642-652, 420-450)
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void machine_stress_prologue(DmkMachine<LAW, NS, ND>& m, const DmkNumDev& num, DmkCnt& cnt) {
  DmkPt<LAW, NS, ND>& p = m.p;
  const int ns = n_dot_state(p);
  if (m.it > num.N_iter_state_max) { m.converged = true; m.state = M_SUBSTEP; return; }   // loop ran out with broken = .false. (as in the reference)
  cnt.state_iter++;
  for (int i = 0; i < ns; i++) { m.last2[i] = (m.it > 1) ? m.last1[i] : 0.0; m.last1[i] = m.dot[i]; }
  cnt.stress_calls++;
  dependent_state(p);
#pragma unroll
  for (int a = 0; a < 9; a++) m.Lpg[a] = p.Lp[a];
  if (invert33(m.Fp0, m.invFp)) { m.converged = false; m.state = M_SUBSTEP; return; }
  mm3(m.F, m.invFp, m.A);
  cnt.Li_iter++;
#pragma unroll
  for (int a = 0; a < 9; a++) { m.Lpg_old[a] = m.Lpg[a]; m.delta[a] = 0.0; }
  m.res_old_norm = 0.0; m.stepLp = 1.0; m.jaco = 0; m.nLp = 0;
  m.state = M_LP;
}

// ONE trip of the Lp loop (phase_mechanical.f90:470-545)
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void machine_lp_iteration(DmkMachine<LAW, NS, ND>& m, const DmkNumDev& num, DmkCnt& cnt) {
  DmkPt<LAW, NS, ND>& p = m.p;
  m.nLp++; cnt.Lp_iter++;
  if (m.nLp > num.N_iter_Lp_max) { m.converged = false; m.state = M_SUBSTEP; return; }
  const double dt = m.dtsub;
  double Fe[9];
#pragma unroll
  for (int a = 0; a < 9; a++) m.B[a] = ((a % 4 == 0) ? 1.0 : 0.0) - dt * m.Lpg[a];
  mm3(m.A, m.B, Fe);
  hooke_S(p.C66, Fe, m.S);
  double Lp_c[9], res[9];
  double iso_c = 0.0, iso_sq = 1.0, iso_dev[9];
  if constexpr (LAW == DMK_LAW_NONE) {
#pragma unroll
    for (int a = 0; a < 9; a++) Lp_c[a] = 0.0;
  } else if constexpr (LAW == DMK_LAW_ISOTROPIC) iso_Lp(p, m.S, Lp_c, iso_c, iso_dev, iso_sq);
  else if constexpr (LAW == DMK_LAW_PHENOPOWERLAW) pheno_Lp(p, m.S, Lp_c, m.ddg);
  else if constexpr (LAW == DMK_LAW_KINEHARDENING) kh_Lp(p, m.S, Lp_c, m.ddg);
  else if constexpr (LAW == DMK_LAW_DISLOTUNGSTEN) tu_Lp(p, m.S, Lp_c, m.ddg);
  else dt_Lp(p, m.S, Lp_c, m.ddg);
  const double atol = fmax(num.eps_rel_Lp * fmax(norm9(m.Lpg), norm9(Lp_c)), num.eps_abs_Lp);
#pragma unroll
  for (int a = 0; a < 9; a++) res[a] = m.Lpg[a] - Lp_c[a];
  const double rn = norm9(res);
  if (any_nan9(res)) { m.converged = false; m.state = M_SUBSTEP; return; }
  if (rn < atol) { m.state = M_STRESS_EPI; return; }
  if (m.nLp == 1 || rn < m.res_old_norm) {
    m.res_old_norm = rn;
#pragma unroll
    for (int a = 0; a < 9; a++) m.Lpg_old[a] = m.Lpg[a];
    m.stepLp = 1.0;
  } else {
    m.stepLp = num.r_linesearch_Lp * m.stepLp;
#pragma unroll
    for (int a = 0; a < 9; a++) m.Lpg[a] = m.Lpg_old[a] + m.delta[a] * m.stepLp;
    return;
  }
  if (m.jaco % num.f_update_jacobi_Lp == 0) {
    m.jaco++;
    double G[9], J[81];
    mm3_tn(m.A, Fe, G);
#pragma unroll
    for (int a = 0; a < 81; a++) J[a] = (a % 10 == 0) ? 1.0 : 0.0;
    if constexpr (LAW == DMK_LAW_ISOTROPIC) iso_jac(p, dt, G, iso_c, iso_dev, iso_sq, J);
    else if constexpr (LAW == DMK_LAW_PHENOPOWERLAW) pheno_jac(p, dt, G, m.ddg, J);
    else if constexpr (LAW == DMK_LAW_DISLOTWIN) dt_jac(p, dt, G, m.ddg, J);
    else if constexpr (LAW == DMK_LAW_KINEHARDENING || LAW == DMK_LAW_DISLOTUNGSTEN) kh_jac(p, dt, G, m.ddg, J);
    if (solve9(J, res)) { m.converged = false; m.state = M_SUBSTEP; return; }
#pragma unroll
    for (int a = 0; a < 9; a++) m.delta[a] = -res[a];
  }
#pragma unroll
  for (int a = 0; a < 9; a++) m.Lpg[a] += m.delta[a] * m.stepLp;
}

// tail of integrateStress + rest of the state iteration (phase_mechanical.f90:586-604, 654-693)
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ void machine_stress_epilogue(DmkMachine<LAW, NS, ND>& m, const DmkNumDev& num) {
  DmkPt<LAW, NS, ND>& p = m.p;
  const int ns = n_dot_state(p);
  double invFp_new[9], Fp_new[9], t1[9], Sm[9], t2[9];
  mm3(m.invFp, m.B, invFp_new);
  if (invert33(invFp_new, Fp_new)) { m.converged = false; m.state = M_SUBSTEP; return; }
  mm3(m.F, invFp_new, t1);
  Sm[0] = m.S[0]; Sm[1] = m.S[5]; Sm[2] = m.S[4]; Sm[3] = m.S[5]; Sm[4] = m.S[1]; Sm[5] = m.S[3]; Sm[6] = m.S[4]; Sm[7] = m.S[3]; Sm[8] = m.S[2];
  mm3_nt(Sm, invFp_new, t2);
  mm3(t1, t2, p.P);
  const double sc = pow(det3(Fp_new), 1.0 / 3.0);
#pragma unroll
  for (int a = 0; a < 9; a++) { p.Fp[a] = Fp_new[a] / sc; p.Lp[a] = m.Lpg[a]; p.Fe[a] = t1[a]; }
#pragma unroll
  for (int a = 0; a < 6; a++) p.S[a] = m.S[a];
  // state iteration, tail
  dot_state(p, m.dot);
  bool nan = false;
  for (int i = 0; i < ns; i++) nan = nan || isnan(m.dot[i]);
  if (nan) { m.converged = true; m.state = M_SUBSTEP; return; }      // leaves the loop with broken = .false. (reference :660)
  double d12 = 0.0, d22 = 0.0, d01 = 0.0;
  for (int i = 0; i < ns; i++) {
    d12 += (m.dot[i] - m.last1[i]) * (m.last1[i] - m.last2[i]);
    d22 += (m.last1[i] - m.last2[i]) * (m.last1[i] - m.last2[i]);
    d01 += m.dot[i] * m.last1[i];
  }
  double zeta = 1.0;
  if (fmin(d01, d12) < 0.0 && d22 > 0.0) zeta = 0.75 + 0.25 * tanh(2.0 + 4.0 * d12 / d22);
  bool conv = true;
  for (int i = 0; i < ns; i++) {
    m.dot[i] = m.dot[i] * zeta + m.last1[i] * (1.0 - zeta);
    const double r = p.st[i] - m.state0[i] - m.dot[i] * m.dtsub;
    p.st[i] -= r;
    if (!(fabs(r) <= fmax(p.ph->p.atol[i], num.eps_rel_state * fabs(p.st[i])))) conv = false;
  }
  if (conv) { m.converged = !delta_state(p); m.state = M_SUBSTEP; return; }     // :661-664
  m.it++;
  m.state = M_STRESS_PRO;
}

// runs one point to completion on a single lane (host twin, single-cell launches); returns converged
template <int LAW, int NS, int ND>
DMK_HD __forceinline__ bool machine_run_point(DmkMachine<LAW, NS, ND>& m, const DmkNumDev& num, double Dt, DmkCnt& cnt) {
  machine_begin(m, num);
  while (m.state != M_FETCH) {
    if (m.state == M_SUBSTEP) machine_substep(m, num, Dt, cnt);
    else if (m.state == M_STRESS_PRO) machine_stress_prologue(m, num, cnt);
    else if (m.state == M_LP) machine_lp_iteration(m, num, cnt);
    else machine_stress_epilogue(m, num);
  }
  return m.converged;
}
