/* dmk_grid.h -- C ABI of libdamask_grid_b200.so: the B200-native hot path of the DAMASK grid solver.
 *
 * Drop-in boundary for the C++ driver of dmentock/solver_cpp.  Two groups of entry points:
 *
 *  (1) the Fortran-named symbols the reference's `Spectral` class binds today
 *      (/root/reference/src/grid/spectral.h:16-23, damask_grid.h:15-18), with identical signatures and meaning;
 *      they act on the context made current with dmk_bind_global() and exchange data through HOST mirrors,
 *      exactly like the Fortran module arrays they replace;
 *
 *  (2) dmk_* functions: the same operations with explicit context and DEVICE-resident fields, plus the spectral
 *      half (`spectral_utilities` / `mechanical_spectral_basic` surface: FFT, Gamma convolution, divergence RMS,
 *      masked compliance, forward/rate, residual, convergence test, NGMRES-driven solve).
 *
 * Conventions: all field pointers use the reference's Fortran memory order -- F(3,3,nx,ny,nz_local), first index
 * fastest, i.e. one 72-byte record per cell with component c = i + 3*j.  3x3 / 3x3x3x3 host arguments are also
 * Fortran order (T(i,j) at [i + 3*j], C(i,j,k,l) at [i + 3*j + 9*k + 27*l]).  Every function returns 0 on
 * success; dmk_last_error() describes the last failure.  Nothing here falls back to a CPU implementation.
 */
#ifndef DMK_GRID_H
#define DMK_GRID_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dmk_ctx dmk_ctx;

/* ---- return codes of every int-returning dmk_* entry point; dmk_last_error() holds the message of the last failure */
#ifndef DMK_OK
#define DMK_OK 0
#define DMK_ERR_CUDA 1           /* a CUDA call failed, or no CUDA device (there is no CPU path) */
#define DMK_ERR_ARG 2            /* bad argument, unreadable or malformed input file */
#define DMK_ERR_UNSUPPORTED 3    /* valid input the device path does not offer (other laws / lattices, RGC, ...) */
#define DMK_ERR_STATE 4          /* call out of order, failed collective, or no solution found for an increment */
#endif

/* ---- limits of the device constitutive kernels */
#define DMK_MAX_SL 24
#define DMK_MAX_TW 12
#define DMK_MAX_STATE (3 * DMK_MAX_SL + DMK_MAX_TW)
#define DMK_MAX_PHASES 8

enum { DMK_LAW_NONE = 0, DMK_LAW_ISOTROPIC = 1, DMK_LAW_PHENOPOWERLAW = 2, DMK_LAW_DISLOTWIN = 3, DMK_LAW_KINEHARDENING = 4, DMK_LAW_DISLOTUNGSTEN = 5 };
enum { DMK_DERIVATIVE_CONTINUOUS = 0, DMK_DERIVATIVE_CENTRAL_DIFF = 1, DMK_DERIVATIVE_FWBW_DIFF = 2 };
enum { DMK_DIVCORR_NONE = 0, DMK_DIVCORR_SIZE = 1, DMK_DIVCORR_SIZE_GRID = 2 };

/* grid + decomposition (reference: DiscretizationGrid, /root/reference/src/grid/discretization_grid.h:24-60;
 * Config::NumFFT, config.h:60-72) */
typedef struct {
  int cells[3];               /* nx, ny, nz (global)                                   */
  double size[3];             /* geom_size / m                                         */
  int world_size, rank;       /* z-slab decomposition: nz % world_size == 0, ny % world_size == 0 */
  int derivative;             /* DMK_DERIVATIVE_*                                      */
  int divergence_correction;  /* DMK_DIVCORR_*                                         */
  int kt_y, kt_z;             /* kx records per tile in the y / z pass (0 = automatic) */
} dmk_grid_desc;

/* numerics of the per-point integrator (phase_mechanical.f90:277-294) and of the grid solver (config.h:50-58) */
typedef struct {
  double r_cutback_min, r_cutback, r_increase, eps_rel_state;
  int N_iter_state_max, N_iter_Lp_max;
  double r_linesearch_Lp, eps_rel_Lp, eps_abs_Lp;
  int f_update_jacobi_Lp;
  int integrator_state;       /* 0 FPI (default), 1 Euler, 2 AdaptiveEuler, 3 RK4, 4 RKCK45 (phase_mechanical.f90:310-323) */
  int itmin, itmax, update_gamma;
  double eps_abs_div, eps_rel_div, eps_abs_P, eps_rel_P;
} dmk_numerics;

/* one phase: Hooke elasticity + plastic law, expanded per system (reference tParameters of each law).
 * Matrices are row-major: P_sl[a][3*i+j]; h_x_y[reacting][acting]. */
typedef struct {
  int law, n_sl, n_tw, n_state;
  double C66[36];                          /* lattice-symmetrised Voigt stiffness at T */
  double atol[DMK_MAX_STATE];
  double state_init[DMK_MAX_STATE];
  double P_sl[DMK_MAX_SL][9], P_nS_pos[DMK_MAX_SL][9], P_nS_neg[DMK_MAX_SL][9], P_tw[DMK_MAX_TW][9];
  double h_sl_sl[DMK_MAX_SL][DMK_MAX_SL], h_sl_tw[DMK_MAX_SL][DMK_MAX_TW];
  double h_tw_sl[DMK_MAX_TW][DMK_MAX_SL], h_tw_tw[DMK_MAX_TW][DMK_MAX_TW];
  /* isotropic (phase_mechanical_plastic_isotropic.f90:11-40); dilatation must be 0 on the device path */
  double iso_xi_inf, iso_dot_gamma_0, iso_n, iso_h_0, iso_h, iso_M, iso_h_ln, iso_c_1, iso_c_2, iso_c_3, iso_c_4, iso_a;
  int iso_dilatation;
  /* phenopowerlaw (phase_mechanical_plastic_phenopowerlaw.f90:11-58) */
  double dot_gamma_0_sl[DMK_MAX_SL], n_sl_exp[DMK_MAX_SL], a_sl[DMK_MAX_SL], h_0_sl_sl[DMK_MAX_SL], c_1[DMK_MAX_SL],
      c_2[DMK_MAX_SL], xi_inf_sl[DMK_MAX_SL], f_sat_sl_tw[DMK_MAX_SL];
  double dot_gamma_0_tw[DMK_MAX_TW], n_tw_exp[DMK_MAX_TW], c_3[DMK_MAX_TW], c_4[DMK_MAX_TW], h_0_tw_sl[DMK_MAX_TW],
      h_0_tw_tw[DMK_MAX_TW], gamma_char_tw[DMK_MAX_TW];
  /* dislotwin (phase_mechanical_plastic_dislotwin.f90:12-92), twinning only (no transformation systems) */
  double dt_mu, dt_nu, dt_kB_T, dt_Gamma_sf, dt_Q_cl, dt_omega, dt_D, dt_p_sb, dt_q_sb, dt_i_tw, dt_L_tw, dt_x_c,
      dt_V_cs, dt_tau_sb, dt_gamma_0_sb, dt_E_sb;
  int dt_extended, dt_omit_dipoles;
  double dt_b_sl[DMK_MAX_SL], dt_Q_sl[DMK_MAX_SL], dt_v_0[DMK_MAX_SL], dt_i_sl[DMK_MAX_SL], dt_p[DMK_MAX_SL],
      dt_q[DMK_MAX_SL], dt_tau_0[DMK_MAX_SL], dt_B[DMK_MAX_SL], dt_d_caron[DMK_MAX_SL];
  double dt_n0_sl[DMK_MAX_SL][3], dt_forest[DMK_MAX_SL][DMK_MAX_SL];
  double dt_b_tw[DMK_MAX_TW], dt_t_tw[DMK_MAX_TW], dt_r[DMK_MAX_TW];
  int dt_pair[DMK_MAX_TW][2];              /* 0-based slip pair nucleating twin i (crystal.f90:82-95) */
  double dt_C66_tw[DMK_MAX_TW][36];
  /* kinehardening (phase_mechanical_plastic_kinehardening.f90:11-40); shares dot_gamma_0_sl, n_sl_exp, xi_inf_sl, h_sl_sl and the
   * (non-)Schmid tensors with phenopowerlaw.  State: xi | chi | gamma (n_sl each, rates from dotState) followed by
   * sgn_gamma | chi_flip | gamma_flip (n_sl each, moved only by the deltaState jump): n_state = 6 n_sl, n_sl <= DMK_MAX_STATE / 6 */
  double kh_chi_inf[DMK_MAX_SL], kh_h_0_xi[DMK_MAX_SL], kh_h_0_chi[DMK_MAX_SL], kh_h_inf_xi[DMK_MAX_SL], kh_h_inf_chi[DMK_MAX_SL];
  /* dislotungsten (phase_mechanical_plastic_dislotungsten.f90:10-47).  State rho_mob | rho_dip | gamma_sl (3 n_sl), dependent
   * state Lambda_sl | tau_pass.  b_sl, i_sl, p, q, B, d_caron, the forest projection, mu, nu, kB T, Q_cl, D and the dipole switch are
   * the dislotwin members above (dt_*), Q_s is dt_Q_sl, tau_Peierls is dt_tau_0; the law's own: kink-pair height and width,
   * attempt frequency, atomic volume f_at b^3, self-diffusion prefactor */
  double tu_h[DMK_MAX_SL], tu_w[DMK_MAX_SL], tu_omega[DMK_MAX_SL], tu_f_at[DMK_MAX_SL], tu_D_0;
} dmk_phase;

/* collective callbacks for world_size > 1 (all stream-ordered on the context's stream; device pointers).
 * The Python host implements them with torch.distributed / NCCL, a C++ host with ncclSend/ncclRecv. */
typedef struct {
  void* user;
  int (*alltoall)(void* user, const void* d_send, void* d_recv, size_t bytes_per_peer);
  int (*allreduce_sum)(void* user, double* d_buf, int n);
  int (*allreduce_max)(void* user, double* d_buf, int n);
} dmk_comm;

/* solution status, mirrors Config::SolutionState (config.h:35-40) */
typedef struct {
  int required_iterations, converged, stag_converged, terminally_ill;
} dmk_solution_state;

/* ------------------------------------------------------------------ lifecycle */
const char* dmk_last_error(void);
int dmk_create(dmk_ctx** ctx, const dmk_grid_desc* grid);
void dmk_destroy(dmk_ctx* ctx);
int dmk_set_stream(dmk_ctx* ctx, void* cuda_stream);
int dmk_set_comm(dmk_ctx* ctx, const dmk_comm* comm);
/* native NCCL binding (csrc/nccl_comm.cpp): the library issues the collectives itself on its stream -- ncclAllReduce (sum, max)
 * and the all-to-all of the distributed-FFT transpose as grouped ncclSend / ncclRecv -- in place of the callback table.  One rank
 * makes the 128-byte ncclUniqueId, the host distributes it by its own means (MPI_Bcast in the reference's driver;
 * torch.distributed in this repo's Python host), every rank calls dmk_nccl_init (collective).  replaces the exchange of FFTW-MPI's
 * transposed plans and the MPI_Allreduce calls of spectral.cpp:160-165 / mech_base.cpp:395-402.  libnccl is resolved at run time. */
int dmk_nccl_unique_id(unsigned char id[128]);
int dmk_nccl_init(dmk_ctx* ctx, const unsigned char id[128]);
int dmk_nccl_version(int* version);
int dmk_synchronize(dmk_ctx* ctx);
int dmk_local_cells(const dmk_ctx* ctx, int* nz_local, int* z_offset, int* ny_local, int* y_offset);
int dmk_kernel_launches(const dmk_ctx* ctx, long long* n);        /* kernels launched so far by this context */
/* CUDA-event section timing of residual evaluations: out[0..4] = accumulated ms of {constitutive kernel, P sum + aim,
 * forward FFT x+y, fused z / Gamma / divergence pass, inverse FFT y+x}, out[5] = evaluations accumulated */
/* run-time switches for comparison runs and tests: "pheno_mass_deferral" = 0 nested | 1 machine | 2 lanes | -1 default */
int dmk_set_option(dmk_ctx* ctx, const char* name, int value);
int dmk_set_profiling(dmk_ctx* ctx, int enable);
int dmk_get_profile(dmk_ctx* ctx, double out[6], int reset);
/* measured FP64 FMA throughput of the device (register-resident DFMA chains), TFLOP/s: the constitutive roofline */
int dmk_fp64_peak(double* tflops);
/* self-test of the per-point kernel's branch-free reciprocal / square root (csrc/pheno_fast_dev.cuh) against the IEEE
 * operators on n arguments of random sign-less mantissa and exponent in [2^-lim, 2^lim]: number of results that differ */
int dmk_selftest_special(long long n, unsigned long long seed, int lim, long long* bad_rcp, long long* bad_sqrt);
int dmk_sizeof_phase(void);                                       /* sizeof(dmk_phase), for binding checks */
int dmk_sizeof_numerics(void);

/* ------------------------------------------------------------------ crystal tables (host, pure; crystal.f90)
 * fill the geometry-derived members of a dmk_phase for lattice "cF" or "cI" */
int dmk_crystal_schmid_slip(const char* lattice, const int* N_sl, int n_fam, double* P /* [sum][9] */);
/* with the cI non-Schmid coefficients of family 1 and the sense (+1 / -1) of the slip direction (crystal.f90:1353-1431) */
int dmk_crystal_schmid_slip_ns(const char* lattice, const int* N_sl, int n_fam, const double* non_schmid, int n_coeff,
                               int sense, double* P);
int dmk_crystal_schmid_twin(const char* lattice, const int* N_tw, int n_fam, double* P);
int dmk_crystal_slip_normal(const char* lattice, const int* N_sl, int n_fam, double* n /* [sum][3] */);
int dmk_crystal_interaction(const char* kind /* "sl-sl","tw-tw","sl-tw","tw-sl" */, const char* lattice,
                            const int* N_react, int nf_react, const int* N_act, int nf_act, const double* values,
                            int n_values, double* h /* [sum_react][sum_act] */);
int dmk_crystal_forest_projection(const char* lattice, const int* N_sl, int n_fam, const double* f_edge, double* proj);
int dmk_crystal_C66_twin(const char* lattice, const int* N_tw, int n_fam, const double* C66, double* C66_tw);
int dmk_crystal_twin_nucleation_pairs(const int* N_tw, int n_fam, int* pairs /* [sum][2], 0-based */);
void dmk_qu2om(const double q[4], double om[9] /* row-major */);

/* ------------------------------------------------------------------ constitutive half (device-resident)
 * replaces f_materialpoint_initDamask + f_discretization_init for the hot path:
 * phase_of_cell / O_of_cell describe the LOCAL z-slab, cell order x fastest (VTI order). */
int dmk_material_init(dmk_ctx* ctx, const dmk_phase* phases, int n_phases, const int* phase_of_cell,
                      const double* O_of_cell /* [n_local][9] row-major rotation matrices (Fp0) */,
                      const dmk_numerics* num);
/* isostrain homogenisation (homogenization_mechanical_isostrain.f90:36-45 + mechanical_homogenize, homogenization_mechanical.f90:
 * 125-145): n_constituents material points per cell, every one of them sees the cell's deformation gradient, the cell's P and dPdF
 * are the volume-fraction weighted sums.  Point k * n_local_cells + cell is constituent k of the cell: phase_of_point [n_const *
 * cells], O_of_point [..][9] row-major, v_of_point [..] (positive, summing to 1 per cell).  Per-point arrays (dmk_get_state /
 * dmk_set_state) then hold n_const * cells rows; every field crossing the boundary stays per cell.  n_constituents = 1 is
 * dmk_material_init. */
int dmk_material_init_constituents(dmk_ctx* ctx, const dmk_phase* phases, int n_phases, int n_constituents, const int* phase_of_point,
                                   const double* O_of_point, const double* v_of_point, const dmk_numerics* num);
int dmk_n_constituents(dmk_ctx* ctx);
/* f_homogenization_mechanical_response + _response2 (P only): integrate every local cell from its *0 state to the
 * deformation gradient field d_F; writes d_P, the ill flag, and the slab's sum of P. */
int dmk_constitutive_response(dmk_ctx* ctx, const double* d_F, double* d_P, double dt);
/* phase_mechanical_dPdF reduced on the fly (spectral.cpp:189-218): C_volAvg, C_minMaxAvg from the state left by the
 * last dmk_constitutive_response; host outputs, Fortran order; d_dPdF (81 doubles/cell) is optional. */
int dmk_tangent_response(dmk_ctx* ctx, double dt, double* C_volAvg, double* C_minMaxAvg, double* d_dPdF,
                         long long minmax_cells[2]);
/* the same reductions on GIVEN device fields, as Spectral::constitutive_response (spectral.cpp:160-218) runs them over
 * homogenization_P / homogenization_dPdF: d_P records [cell][9], d_dPdF [cell][81] in the reference's Fortran order;
 * n_cells = cells of this rank's slab; host outputs in Fortran order; minmax_cells = global indices of the two cells */
int dmk_homogenise(dmk_ctx* ctx, const double* d_P, const double* d_dPdF, long long n_cells, double* P_av, double* C_volAvg,
                   double* C_minMaxAvg, long long minmax_cells[2]);
int dmk_materialpoint_forward(dmk_ctx* ctx);                       /* f_materialpoint_forward */
/* host copy, cell-major [cell][component]: "F","F0","Fp","Fp0","Fe","Lp","Lp0","S","S0" (row-major 3x3), "state","state0",
 * "dep", and "cutbacks" (sub-step cut-backs of each cell in the last response call) */
int dmk_get_state(dmk_ctx* ctx, const char* name, double* host_out, size_t n_doubles);
/* inverse of dmk_get_state for "F","F0","Fp","Fp0","Fe","Lp","Lp0","S","S0","state","state0","dep": restart of the per-point state
 * (what phase_mechanical.f90:1264-1298 writes to / reads from the restart file) */
int dmk_set_state(dmk_ctx* ctx, const char* name, const double* host_in, size_t n_doubles);
int dmk_max_state(dmk_ctx* ctx);                                   /* components of "state" / "state0" */
int dmk_max_dep(dmk_ctx* ctx);                                     /* components of "dep" (0: none) */
int dmk_get_counters(dmk_ctx* ctx, long long counters[7]);         /* Lp iters, stress calls, state iters, substeps, cutbacks, ill,
                                                                     * cells deferred to the reference-order kernel */
/* cumulative lane-schedule statistics of the lane-persistent constitutive kernel: {Newton-block warp trips, lanes served,
 * state-block trips, lanes, refill trips, lanes, Newton steps solved, 0} */
int dmk_get_schedule_stats(dmk_ctx* ctx, long long stats[8]);
int dmk_terminally_ill(dmk_ctx* ctx, int* flag, int reset);

/* ------------------------------------------------------------------ spectral half (device-resident fields)
 * FFT<5>::forward / backward (fft.h:37-60): half spectrum layout [kz][ky_local][kx][9] complex, unnormalised */
int dmk_fft_forward(dmk_ctx* ctx, const double* d_real, double* d_fourier);
int dmk_fft_backward(dmk_ctx* ctx, const double* d_fourier, double* d_real, double wgt);
size_t dmk_fourier_doubles(const dmk_ctx* ctx);                    /* size of a half-spectrum buffer in doubles */
int dmk_update_gamma(dmk_ctx* ctx, const double* C /* host, 81 */);/* MechBase::update_gamma (mech_base.cpp:156) */
/* calculate_divergence_rms + gamma_convolution fused (mech_base.cpp:302-477): d_field in: stress, out: Gamma*field;
 * field_aim host 3x3 (already rotated to lab); err_div returned on the host */
int dmk_divergence_and_gamma(dmk_ctx* ctx, double* d_field, const double* field_aim, double* err_div);
/* Fused exchange of the distributed FFT over peer memory (NVLink / NVSwitch): each rank exports the CUDA IPC handles of its
 * two spectral buffers (128 bytes), the host side all-gathers them, dmk_p2p_import opens the peers' buffers.  From then on
 * the y pass stores its tiles straight into the owning peer's z-pass buffer and the (Gamma-fused) inverse z pass stores
 * straight into the owning peer's y-pass buffer; the all-to-all callback is no longer used, two 1-double all-reduces order
 * the stores.  *enabled = 0: the grid does not qualify (line lengths), the all-to-all path stays. */
int dmk_p2p_export(dmk_ctx* ctx, unsigned char* handles128);
int dmk_p2p_import(dmk_ctx* ctx, const unsigned char* all_handles, int* enabled);
int dmk_p2p_disable(dmk_ctx* ctx);                                 /* back to the all-to-all callback (collective decision) */
/* MechBase::base_update_coords (mech_base.cpp:45-154): deformed node and cell-centre coordinates of the local slab from F
 * (device field).  Host outputs x_n[nz_loc+1][ny+1][nx+1][3] and x_p[nz_loc][ny][nx][3] (= the reference's x_n(3,i,j,k),
 * x_p(3,i,j,k), first index fastest). */
int dmk_update_coords(dmk_ctx* ctx, const double* d_F, double* h_x_n, double* h_x_p);
int dmk_masked_compliance(const double* C /*81*/, const double* R /*9 row-major*/, const int* stress_mask /*9, 1='x'*/,
                          double* S /*81*/);                        /* MechBase::calculate_masked_compliance */
int dmk_forward_field(dmk_ctx* ctx, double dt, const double* d_field_last, const double* d_rate, const double* aim,
                      double* d_out);                               /* MechBase::forward_field  (aim may be NULL) */
int dmk_calculate_rate(dmk_ctx* ctx, int heterogeneous, const double* d_field0, const double* d_field, double dt,
                       const double* av_rate, double* d_rate);      /* MechBase::calculate_rate */

/* ------------------------------------------------------------------ mechanical_spectral_basic surface
 * MechSolverBasic::{init, forward, calculate_solution, form_residual, converged} (mech_solver_basic.cpp) with the
 * solution vector, its history (NGMRES, m = 30) and all work fields resident in HBM. */
int dmk_mech_init(dmk_ctx* ctx, const long long* gamma_pair_cells /* NULL or 2 global cell indices */);
/* the polarisation variant of the scheme (`solver: mechanical: spectral_polarisation`; upstream DAMASK's
 * grid_mechanical_spectral_polarisation -- the reference tree parses its numerics only: alpha, beta, eps_abs_curl(F),
 * eps_rel_curl(F), config.cpp:450-457, config.h:68-71).  Solution vector (F | F_tau): dmk_mech_forward /
 * dmk_mech_calculate_solution / dmk_mech_get work as for the basic scheme; dmk_mech_form_residual takes and returns
 * 2 x 9 doubles per cell; fields "F_tau", "F_tau_lastInc" and the scalar "err_curl" are available in addition. */
int dmk_mech_init_polarisation(dmk_ctx* ctx, double alpha, double beta, double eps_abs_curl, double eps_rel_curl);
/* BC values / masks / R: row-major 3x3 (index 3*i+j); mask 1 = 'x' (component not controlled) */
int dmk_mech_forward(dmk_ctx* ctx, int cut_back, int guess, double dt, double dt_old, double t_remaining,
                     const char* deformation_type, const double* deformation_values, const int* deformation_mask,
                     const char* stress_type, const double* stress_values, const int* stress_mask,
                     const double* R /* 9 row-major */);
int dmk_mech_calculate_solution(dmk_ctx* ctx, dmk_solution_state* out);
int dmk_mech_prepare(dmk_ctx* ctx);   /* the pre-SNES part of calculate_solution: tangent reductions, masked compliance */
int dmk_mech_set(dmk_ctx* ctx, const char* what, const double* values9);   /* "F_aim","F_aim_lastInc","P_aim" (Fortran order) */
int dmk_mech_set_field_device(dmk_ctx* ctx, const char* what, const double* d_src); /* "F","F_lastInc","F_dot" */
int dmk_mech_form_residual(dmk_ctx* ctx, const double* d_F, double* d_r);  /* one evaluation, device pointers */
int dmk_mech_form_residual_host(dmk_ctx* ctx, const double* F, double* r); /* same with host buffers (H2D + D2H) */
int dmk_mech_get(dmk_ctx* ctx, const char* what, double* out, size_t n);   /* "P_av","F_aim","err_div","err_BC","C_volAvg",... */
int dmk_mech_get_field(dmk_ctx* ctx, const char* what, double* host_out);  /* "F","P","F_lastInc": local slab, 9*n doubles */
int dmk_mech_set_field(dmk_ctx* ctx, const char* what, const double* host_in);
int dmk_mech_device_field(dmk_ctx* ctx, const char* what, double** d_ptr);  /* device pointer of "F","F_lastInc","F_dot","r" */
int dmk_mech_copy_field(dmk_ctx* ctx, const char* what, double* d_dst);      /* device-to-device copy of a solver field */
/* ------------------------------------------------------------------ load case: Config::parse_load + DamaskGrid::loop_over_loadcase
 * One load step (Config::LoadStep, config.h:74-96).  Values / masks / R row-major 3x3; mask 1 = 'x' (component not controlled
 * by this boundary condition); R = the rotation of the load frame as a matrix (the file gives a quaternion w, x, y, z). */
typedef struct {
  char deformation_type[8];                       /* "F", "dot_F" or "L" */
  double deformation_values[9]; int deformation_mask[9];
  char stress_type[8];                            /* "P" or "dot_P" */
  double stress_values[9]; int stress_mask[9];
  double R[9];
  double t, r; int N;                             /* discretization: time, geometric ratio of the time steps (1: equal), increments */
  int estimate_rate, f_out, f_restart;
} dmk_load_step;
typedef struct { int inc, step, inc_in_step, f_out; double t; int converged, iterations; } dmk_increment_record;
/* what the increment loop asks of the solver underneath; all return 0 or an error code that ends the loop */
typedef struct {
  void* user;
  int (*forward)(void* user, int cut_back, int guess, double dt, double dt_old, double t_remaining, const dmk_load_step* step);
  int (*materialpoint_forward)(void* user);
  int (*calculate_solution)(void* user, dmk_solution_state* out);
  int (*result)(void* user, const dmk_increment_record* rec);      /* may be NULL; called for the increments f_out selects */
} dmk_loop_ops;
/* Config::parse_load (config.cpp:247-358): load file -> load steps; mechanical_solver (may be NULL) receives solver/mechanical */
int dmk_load_from_yaml(const char* load_yaml_path, dmk_load_step* steps, int max_steps, int* n_steps, char* mechanical_solver,
                       int solver_name_capacity);
/* DamaskGrid::loop_over_loadcase (damask_grid.cpp:82-196) with upstream DAMASK's cut-back bookkeeping.  ops = NULL: the context's
 * own solver (dmk_mech_forward / dmk_materialpoint_forward / dmk_mech_calculate_solution; dmk_mech_init must have run); otherwise
 * ctx may be NULL.  max_increments <= 0: the whole load case.  Up to max_records records are stored (records may be NULL);
 * DMK_ERR_STATE when an increment finds no solution at the last cut-back level. */
int dmk_loop_over_loadcase(dmk_ctx* ctx, const dmk_load_step* steps, int n_steps, int max_cut_back, int max_increments,
                           const dmk_loop_ops* ops, dmk_increment_record* records, int max_records, int* n_records);
/* test aid: keep host copies of the F handed to and the r returned by every residual evaluation of the solver */
int dmk_mech_capture(dmk_ctx* ctx, int enable);
int dmk_mech_get_capture(dmk_ctx* ctx, double* F_out, double* r_out, size_t max_evals, size_t* n_evals);
int dmk_mech_eval_log(dmk_ctx* ctx, double* out, size_t max_rows, size_t* n_rows); /* per evaluation: 9 P_av (row-major), err_div, err_BC, total_iter */
int dmk_mech_eval_stats(dmk_ctx* ctx, double* out, size_t max_rows, size_t* n_rows); /* per evaluation: cells of this slab whose first sub-step failed (cut back), reserved */
int dmk_mech_reports(dmk_ctx* ctx, double* out, size_t max_rows, size_t* n_rows);  /* per convergence check: total_iter, err_div, divTol, err_BC, BCTol, reason */

/* ------------------------------------------------------------------ (1) Fortran-named drop-in symbols */
int dmk_bind_global(dmk_ctx* ctx);
/* init symbols of the reference's main.cpp:9-14,80-86 and discretization_grid.h:10-19 (csrc/init_dropin.cpp): initBase keeps
 * the file names (lengths by value, as Fortran character arguments), f_discretization_init the 1-based material index and
 * the cell-centre coordinates of this rank's cells (x fastest), initDamask reads material.yaml / numerics.yaml, creates
 * a context for the local slab, initialises the per-point state on the device and binds it (dmk_bound_context) */
void f_materialpoint_initBase(const char* material_path, int material_path_len, const char* load_path, int load_path_len,
                              const char* grid_path, int grid_path_len, const char* numerics_path, int numerics_path_len);
void f_materialpoint_initDamask(void);
void f_discretization_init(int* material_at, int* n_materialpoints, double* IPcoords0, int* n_ips, double* NodeCoords0,
                           int* n_nodes, int* shared_nodes_begin);
void f_discretization_setIPcoords(double* IPcoords0);
void f_discretization_setNodeCoords(double* NodeCoords);
void f_discretization_fetch_ip_node_coord_pointers(double** ip_coords_ptr, double** node_coords_ptr);
void f_CLI_get_geomFilename(char* filename);
/* geometry file readers of discretization_grid.h:20-22 (VTI.f90:146-266, 473-497): cells / size / origin of a VTK ImageData
 * file's text, and a cell-data array as 32-bit integers PLUS ONE (material IDs become 1-based); on error cells = -1 /
 * *dataset_size = -1 and dmk_last_error() has the message */
void f_VTI_readCellsSizeOrigin(const char* fileContent, int string_size, int* cells, double* size, double* origin);
void f_VTI_readDataset_int(const char* fileContent, int string_size, const char* label, int label_length, int* dataset_ptr,
                           int* dataset_size);
dmk_ctx* dmk_bound_context(void);       /* the context f_materialpoint_initDamask made (NULL if it failed: dmk_last_error) */
/* material.yaml -> phase tables and, per `material:` entry, phase index and orientation matrix (row-major); numerics.yaml ->
 * dmk_numerics (a NULL / empty path gives the defaults) -- the C++ twin of solver_cpp_b200/material.py */
int dmk_material_from_yaml(const char* material_yaml_path, double T, dmk_phase* phases, int max_phases, int* n_phases,
                           int* phase_of_material, double* O_of_material, int max_materials, int* n_materials);
/* ... with several constituents per material entry (material.f90:135-161; homogenisation types pass and isostrain): entry (m, k)
 * of the per-material tables sits at m * max_constituents + k; v = volume fraction (the fractions of an entry sum to 1) */
int dmk_material_constituents_from_yaml(const char* material_yaml_path, double T, dmk_phase* phases, int max_phases, int* n_phases,
                                        int max_constituents, int* n_constituents_of_material, int* phase_of_material,
                                        double* O_of_material, double* v_of_material, int max_materials, int* n_materials);
int dmk_numerics_from_yaml(const char* numerics_yaml_path, dmk_numerics* num, int* derivative, int* divergence_correction);
/* the grid-level keys of Config::parse_numerics (config.cpp:379-384, 449-457): limits of the increment loop (N_cutback_max is
 * dmk_loop_over_loadcase's max_cut_back) and the polarisation scheme's parameters (dmk_mech_init_polarisation) */
typedef struct { int N_staggered_iter_max, N_cutback_max; double alpha, beta, eps_abs_curl, eps_rel_curl; } dmk_grid_numerics;
int dmk_grid_numerics_from_yaml(const char* numerics_yaml_path, dmk_grid_numerics* out);
void f_homogenization_fetch_tensor_pointers(double** homogenization_F0, double** homogenization_F,
                                            double** homogenization_P, double** homogenization_dPdF,
                                            void** terminally_ill);
void f_homogenization_mechanical_response(double* delta_t, int* cell_start, int* cell_end);
void f_homogenization_thermal_response(double* delta_t, int* cell_start, int* cell_end);
void f_homogenization_mechanical_response2(double* delta_t, int* FEsolving_execIP, int* FEsolving_execElem);
void f_materialpoint_forward(void);
void f_materialpoint_result(int* inc, double* t);
/* helpers of utilities_fortran.h:8-15 */
void f_math_invert(double* InvA, int* err, double* A, int* n);
void f_math_3333to99(double* m3333, double* m99);
void f_math_99to3333(double* m99, double* m3333);
void f_math_mul3333xx33(double* A, double* B, double* res);
void f_rotate_tensor2(double* qu, double* T, double* rotated, int* active);
void f_rotate_tensor4(double* qu, double* T, double* rotated, int* active);

#ifdef __cplusplus
}
#endif
#endif /* DMK_GRID_H */
