// Per-material-point constitutive state held in HBM and the device-side phase parameter tables.
//
// Replaces the Fortran module arrays of the reference's constitutive library (phase_mechanical_F/Fp/Fe/Lp/S and
// their *0 twins, plasticState(ph)%state/state0: /root/reference/src/phase_mechanical.f90:183-260, phase.f90:26-46).
// Layout: structure of arrays, component-major -- X[c * n + cell] -- so that the 32 lanes of a warp (32 neighbouring
// cells along x) read and write 256 contiguous bytes per component.  Tensors are row-major (c = 3 i + j) internally;
// the AoS record fields that cross the boundary (F in, P out) keep the reference's Fortran order (c = i + 3 j).
// The intermediate configuration is fixed to Fi = I, Li = 0 on the device (no dilatational law is offered), which
// removes Fi, Fi0, Li, Li0 from the per-point state without changing any result bit.
#pragma once
#include "ctx.h"
#include "dmk_common.cuh"

#define DMK_MAX_DEP (2 * DMK_MAX_SL + DMK_MAX_TW)      // dislotwin dependent state: Lambda_sl | Lambda_tw | tau_pass
#define DMK_MAX_SYS (DMK_MAX_SL + DMK_MAX_TW)

struct DmkDevPhase {
  dmk_phase p;
  // Q = C : P_nonSchmid as symmetric Voigt-6 (11,22,33,23,13,12), for the analytic Newton Jacobian of the Lp loop
  double Q_pos[DMK_MAX_SL][6], Q_neg[DMK_MAX_SL][6], Q_tw[DMK_MAX_TW][6];
  // strain-like Voigt projections of the (non-)Schmid tensors: pv[alpha] = P(ij) + P(ji) (shear), P(ii) (normal)
  double pv_pos[DMK_MAX_SL][6], pv_neg[DMK_MAX_SL][6], pv_tw[DMK_MAX_TW][6];
  int nonschmid;      // P_nS_neg != -P_nS_pos (cI with non-Schmid coefficients)
  // stress exponents that are small positive integers (n = 20 in every shipped parameter set) are evaluated by repeated
  // squaring instead of pow(): 0 = use pow
  int n_sl_int[DMK_MAX_SL], n_tw_int[DMK_MAX_TW], iso_n_int;
  int n_dep;
};

struct DmkPointArrays {
  long long n;                         // local cells
  int n_phases;
  int max_state, max_dep;
  int* phase;                          // [n] phase index of each cell
  double *F0, *F, *Fp0, *Fp, *Fe, *Lp0, *Lp;   // [9][n]
  double *S0, *S;                      // [6][n] Voigt (11,22,33,23,13,12)
  double *state0, *state;              // [max_state][n]
  double* dep;                         // [max_dep][n]  (dislotwin only, else nullptr)
  double* nrm;                         // [n] ||dPdF||^2 scratch of the tangent pass
  int* cutbacks;                       // [n] sub-step cut-backs of each cell in the last response call (diagnostic)
  int* lp_iter;                        // [n] Lp Newton iterations of each cell in the last response call (diagnostic)
  int* defer_list;                     // [n] cells the register-resident kernel handed to the reference-order kernel
  unsigned* d_defer_count;             // entries of defer_list written by the last register-resident launch
  double* d_defer_total;               // length of the given-up lists of all ranks (all-reduced), as a double
  unsigned long long* d_work;          // next unclaimed cell of the lane-persistent kernel's chunk queue (pheno_fast.cu)
  DmkDevPhase* d_phases;               // device array
  dmk_phase* h_phases;                 // host copy
  DmkDevPhase* h_dev_phases;           // host copy of the device tables
  int law_of_phase[DMK_MAX_PHASES], nstate_of_phase[DMK_MAX_PHASES], nsys_of_phase[DMK_MAX_PHASES];
  unsigned long long* d_counters;      // 16: Lp iters, Li iters, stress calls, state iters, substeps, cutbacks, ill cells, deferred cells | 8 lane-schedule statistics
  int* d_ill;                          // device flag (OR of broken cells since the last reset)
  long long n_cells_of_phase[DMK_MAX_PHASES];
  // isostrain homogenisation (homogenization_mechanical_isostrain.f90; N_constituents > 1): n = n_const * n_cells material points,
  // point k * n_cells + cell is constituent k of the cell.  All per-point arrays above are sized by n; the cell fields crossing
  // the boundary are partitioned into / homogenised from the point fields below.  n_const = 1: n == n_cells, nothing allocated
  int n_const;
  long long n_cells;
  double* v;                           // [n] volume fraction of each point within its cell
  double *Fpt, *Ppt;                   // [n][9] records: partitioned deformation gradient, per-point first Piola-Kirchhoff stress
};

struct DmkNumDev {
  double r_cutback_min, r_cutback, r_increase, eps_rel_state;
  int N_iter_state_max, N_iter_Lp_max;
  double r_linesearch_Lp, eps_rel_Lp, eps_abs_Lp;
  int f_update_jacobi_Lp;
  int integrator;                      // DMK_INTEG_* (constitutive_dev.cuh)
};

int dmk_points_constitutive(dmk_ctx* c, const double* d_F, double* d_P, double dt);
int dmk_wait_staged_input(dmk_ctx* c, const double* d_F);   // kernels without chunked launches: wait for the whole staged field
int dmk_points_pheno_fast(dmk_ctx* c, int ph, const DmkNumDev& nd, const double* d_F, double* d_P, double dt, bool* handled,
                          bool* deferred /*out: cells may be waiting in pa.defer_list*/,
                          double* defer_max /*out: >= 0: the reference-order kernel takes the list only if *pa.d_defer_total (its
                                              length over all ranks) is at most this; < 0: always*/);
int dmk_mass_deferral_setting(const dmk_ctx* c);     // 0 nested, 1 machine, 2 lanes (default): who takes a long deferred list (constitutive.cu)
int dmk_points_tangent(dmk_ctx* c, double dt, double* d_Csum81 /*device, row-major sum over local cells*/,
                       double* d_dPdF /*optional [n][81] Fortran order*/);
int dmk_cells_tangent(dmk_ctx* c, double dt, double* d_Csum81 /*may be null under isostrain*/, double* d_cell_dPdF /*[cell][81] Fortran order, homogenised over the constituents*/);
int dmk_points_tangent_cell(dmk_ctx* c, double dt, long long cell, double* d_out81 /*row-major*/);
