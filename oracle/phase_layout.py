"""ORACLE (test infrastructure only). Single source of truth for the C structs shared between
oracle/constitutive.c and its ctypes wrapper: the field tables below generate both the C header
(oracle/_gen_layout.h, written by oracle/build.py) and the ctypes.Structure classes.

Field meaning follows the reference's tParameters containers:
  phenopowerlaw  phase_mechanical_plastic_phenopowerlaw.f90:11-58
  isotropic      phase_mechanical_plastic_isotropic.f90:11-40
  kinehardening  phase_mechanical_plastic_kinehardening.f90:11-40
  dislotwin      phase_mechanical_plastic_dislotwin.f90:12-92
  numerics       phase_mechanical.f90:277-294
"""
import ctypes as C

MAX_SL, MAX_TW, MAX_TR = 48, 12, 12
MAX_STATE = 3 * MAX_SL + MAX_TW + MAX_TR

LAW_NONE, LAW_ISOTROPIC, LAW_PHENOPOWERLAW, LAW_DISLOTWIN, LAW_KINEHARDENING, LAW_DISLOTUNGSTEN = 0, 1, 2, 3, 4, 5

# (name, 'i'|'d', shape tuple or ())
PHASE_FIELDS = [
    ('law', 'i', ()), ('n_sl', 'i', ()), ('n_tw', 'i', ()), ('n_tr', 'i', ()), ('n_state', 'i', ()),
    ('n_dot', 'i', ()),        # sizeDotState: the leading part of the state that follows dotState; the rest moves by deltaState only
    ('C66', 'd', (36,)),
    ('atol', 'd', (MAX_STATE,)),
    ('state_init', 'd', (MAX_STATE,)),
    # Schmid matrices, row-major 3x3: P[a][3*i+j]
    ('P_sl', 'd', (MAX_SL, 9)), ('P_nS_pos', 'd', (MAX_SL, 9)), ('P_nS_neg', 'd', (MAX_SL, 9)),
    ('P_tw', 'd', (MAX_TW, 9)),
    # interaction matrices [reacting][acting]
    ('h_sl_sl', 'd', (MAX_SL, MAX_SL)), ('h_sl_tw', 'd', (MAX_SL, MAX_TW)),
    ('h_tw_sl', 'd', (MAX_TW, MAX_SL)), ('h_tw_tw', 'd', (MAX_TW, MAX_TW)),
    # ---- isotropic
    ('iso_xi_inf', 'd', ()), ('iso_dot_gamma_0', 'd', ()), ('iso_n', 'd', ()), ('iso_h_0', 'd', ()),
    ('iso_h', 'd', ()), ('iso_M', 'd', ()), ('iso_h_ln', 'd', ()), ('iso_c_1', 'd', ()), ('iso_c_2', 'd', ()),
    ('iso_c_3', 'd', ()), ('iso_c_4', 'd', ()), ('iso_a', 'd', ()), ('iso_dilatation', 'i', ()),
    # ---- phenopowerlaw
    ('dot_gamma_0_sl', 'd', (MAX_SL,)), ('n_sl_exp', 'd', (MAX_SL,)), ('a_sl', 'd', (MAX_SL,)),
    ('h_0_sl_sl', 'd', (MAX_SL,)), ('c_1', 'd', (MAX_SL,)), ('c_2', 'd', (MAX_SL,)),
    ('xi_inf_sl', 'd', (MAX_SL,)), ('f_sat_sl_tw', 'd', (MAX_SL,)),
    ('dot_gamma_0_tw', 'd', (MAX_TW,)), ('n_tw_exp', 'd', (MAX_TW,)), ('c_3', 'd', (MAX_TW,)),
    ('c_4', 'd', (MAX_TW,)), ('h_0_tw_sl', 'd', (MAX_TW,)), ('h_0_tw_tw', 'd', (MAX_TW,)),
    ('gamma_char_tw', 'd', (MAX_TW,)),
    # ---- dislotwin
    ('dt_mu', 'd', ()), ('dt_nu', 'd', ()), ('dt_kB_T', 'd', ()), ('dt_Gamma_sf', 'd', ()),
    ('dt_Q_cl', 'd', ()), ('dt_omega', 'd', ()), ('dt_D', 'd', ()), ('dt_p_sb', 'd', ()), ('dt_q_sb', 'd', ()),
    ('dt_i_tw', 'd', ()), ('dt_L_tw', 'd', ()), ('dt_x_c', 'd', ()), ('dt_V_cs', 'd', ()),
    ('dt_tau_sb', 'd', ()), ('dt_gamma_0_sb', 'd', ()), ('dt_E_sb', 'd', ()),
    ('dt_extended', 'i', ()), ('dt_omit_dipoles', 'i', ()), ('dt_fcc_nucleation', 'i', ()),
    ('dt_b_sl', 'd', (MAX_SL,)), ('dt_Q_sl', 'd', (MAX_SL,)), ('dt_v_0', 'd', (MAX_SL,)),
    ('dt_i_sl', 'd', (MAX_SL,)), ('dt_p', 'd', (MAX_SL,)), ('dt_q', 'd', (MAX_SL,)),
    ('dt_tau_0', 'd', (MAX_SL,)), ('dt_B', 'd', (MAX_SL,)), ('dt_d_caron', 'd', (MAX_SL,)),
    ('dt_n0_sl', 'd', (MAX_SL, 3)), ('dt_forest', 'd', (MAX_SL, MAX_SL)),
    ('dt_b_tw', 'd', (MAX_TW,)), ('dt_t_tw', 'd', (MAX_TW,)), ('dt_r', 'd', (MAX_TW,)),
    ('dt_pair', 'i', (MAX_TW, 2)),                      # 0-based slip indices
    ('dt_C66_tw', 'd', (MAX_TW, 36)),
    # ---- kinehardening (phase_mechanical_plastic_kinehardening.f90:11-40; dot_gamma_0_sl, n_sl_exp, xi_inf_sl are shared)
    ('kh_chi_inf', 'd', (MAX_SL,)), ('kh_h_0_xi', 'd', (MAX_SL,)), ('kh_h_0_chi', 'd', (MAX_SL,)),
    ('kh_h_inf_xi', 'd', (MAX_SL,)), ('kh_h_inf_chi', 'd', (MAX_SL,)),
    # ---- dislotungsten (phase_mechanical_plastic_dislotungsten.f90:10-47; b_sl, i_sl, p, q, B, d_caron, Q_s -> dt_Q_sl,
    #      tau_Peierls -> dt_tau_0, forest projection, mu, nu, kB T, Q_cl, D and the dipole switch are the dislotwin fields)
    ('tu_h', 'd', (MAX_SL,)), ('tu_w', 'd', (MAX_SL,)), ('tu_omega', 'd', (MAX_SL,)), ('tu_f_at', 'd', (MAX_SL,)), ('tu_D_0', 'd', ()),
]

NUM_FIELDS = [
    ('r_cutback_min', 'd', ()), ('r_cutback', 'd', ()), ('r_increase', 'd', ()), ('eps_rel_state', 'd', ()),
    ('N_iter_state_max', 'i', ()), ('N_iter_Lp_max', 'i', ()), ('r_linesearch_Lp', 'd', ()),
    ('eps_rel_Lp', 'd', ()), ('eps_abs_Lp', 'd', ()), ('f_update_jacobi_Lp', 'i', ()),
    ('N_iter_Li_max', 'i', ()), ('r_linesearch_Li', 'd', ()), ('eps_rel_Li', 'd', ()), ('eps_abs_Li', 'd', ()),
    ('f_update_jacobi_Li', 'i', ()),
    ('integrator_state', 'i', ()),      # 0 FPI, 1 Euler, 2 AdaptiveEuler, 3 RK4, 4 RKCK45 (phase_mechanical.f90:310-323)
]

# per-point arrays (struct of pointers); tensors are row-major 3x3 ([3*i+j]) per point
STATE_FIELDS = ['F', 'F0', 'Fp', 'Fp0', 'Fi', 'Fi0', 'Fe', 'Lp', 'Lp0', 'Li', 'Li0', 'S', 'S0', 'P',
                'state', 'state0', 'dep']


def _ctype(kind, shape):
    t = C.c_int if kind == 'i' else C.c_double
    for n in reversed(shape):
        t = t * n
    return t


def _cdecl(name, kind, shape):
    return '  %s %s%s;' % ('int' if kind == 'i' else 'double', name, ''.join('[%d]' % n for n in shape))


class Phase(C.Structure):
    _fields_ = [(n, _ctype(k, s)) for n, k, s in PHASE_FIELDS]


class Numerics(C.Structure):
    _fields_ = [(n, _ctype(k, s)) for n, k, s in NUM_FIELDS]


class PointState(C.Structure):
    _fields_ = ([('n', C.c_int), ('max_state', C.c_int), ('max_dep', C.c_int)]
                + [('phase_of', C.POINTER(C.c_int))]
                + [(n, C.POINTER(C.c_double)) for n in STATE_FIELDS])


def header_text():
    lines = ['/* GENERATED by oracle/phase_layout.py -- do not edit */', '#pragma once',
             '#define ORC_MAX_SL %d' % MAX_SL, '#define ORC_MAX_TW %d' % MAX_TW, '#define ORC_MAX_TR %d' % MAX_TR,
             '#define ORC_MAX_STATE %d' % MAX_STATE, 'typedef struct {']
    lines += [_cdecl(*f) for f in PHASE_FIELDS]
    lines += ['} orc_phase_t;', 'typedef struct {']
    lines += [_cdecl(*f) for f in NUM_FIELDS]
    lines += ['} orc_num_t;', 'typedef struct {', '  int n; int max_state; int max_dep;', '  int *phase_of;']
    lines += ['  double *%s;' % n for n in STATE_FIELDS]
    lines += ['} orc_points_t;', '']
    return '\n'.join(lines)
