"""ctypes binding of libdamask_grid_b200.so (the C ABI declared in include/dmk_grid.h).

This is the only way Python reaches the CUDA path; there is no Python or CPU implementation behind it: if the
library or a CUDA device is missing, calls raise."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# DMK_LIB_PATH: another build of the SAME library (tools/build_variant.py: tuning builds with other compile-time factors)
LIB_PATH = os.environ.get('DMK_LIB_PATH') or os.path.join(HERE, 'libdamask_grid_b200.so')

MAX_SL, MAX_TW, MAX_PHASES = 24, 12, 8
MAX_STATE = 3 * MAX_SL + MAX_TW


class GridDesc(C.Structure):
    _fields_ = [('cells', C.c_int * 3), ('size', C.c_double * 3), ('world_size', C.c_int), ('rank', C.c_int),
                ('derivative', C.c_int), ('divergence_correction', C.c_int), ('kt_y', C.c_int), ('kt_z', C.c_int)]


class Numerics(C.Structure):
    _fields_ = [('r_cutback_min', C.c_double), ('r_cutback', C.c_double), ('r_increase', C.c_double),
                ('eps_rel_state', C.c_double), ('N_iter_state_max', C.c_int), ('N_iter_Lp_max', C.c_int),
                ('r_linesearch_Lp', C.c_double), ('eps_rel_Lp', C.c_double), ('eps_abs_Lp', C.c_double),
                ('f_update_jacobi_Lp', C.c_int), ('integrator_state', C.c_int), ('itmin', C.c_int), ('itmax', C.c_int), ('update_gamma', C.c_int),
                ('eps_abs_div', C.c_double), ('eps_rel_div', C.c_double), ('eps_abs_P', C.c_double),
                ('eps_rel_P', C.c_double)]


def _d(*shape):
    t = C.c_double
    for n in reversed(shape):
        t = t * n
    return t


class Phase(C.Structure):
    _fields_ = [
        ('law', C.c_int), ('n_sl', C.c_int), ('n_tw', C.c_int), ('n_state', C.c_int),
        ('C66', _d(36)), ('atol', _d(MAX_STATE)), ('state_init', _d(MAX_STATE)),
        ('P_sl', _d(MAX_SL, 9)), ('P_nS_pos', _d(MAX_SL, 9)), ('P_nS_neg', _d(MAX_SL, 9)), ('P_tw', _d(MAX_TW, 9)),
        ('h_sl_sl', _d(MAX_SL, MAX_SL)), ('h_sl_tw', _d(MAX_SL, MAX_TW)), ('h_tw_sl', _d(MAX_TW, MAX_SL)),
        ('h_tw_tw', _d(MAX_TW, MAX_TW)),
        ('iso_xi_inf', C.c_double), ('iso_dot_gamma_0', C.c_double), ('iso_n', C.c_double), ('iso_h_0', C.c_double),
        ('iso_h', C.c_double), ('iso_M', C.c_double), ('iso_h_ln', C.c_double), ('iso_c_1', C.c_double),
        ('iso_c_2', C.c_double), ('iso_c_3', C.c_double), ('iso_c_4', C.c_double), ('iso_a', C.c_double),
        ('iso_dilatation', C.c_int),
        ('dot_gamma_0_sl', _d(MAX_SL)), ('n_sl_exp', _d(MAX_SL)), ('a_sl', _d(MAX_SL)), ('h_0_sl_sl', _d(MAX_SL)),
        ('c_1', _d(MAX_SL)), ('c_2', _d(MAX_SL)), ('xi_inf_sl', _d(MAX_SL)), ('f_sat_sl_tw', _d(MAX_SL)),
        ('dot_gamma_0_tw', _d(MAX_TW)), ('n_tw_exp', _d(MAX_TW)), ('c_3', _d(MAX_TW)), ('c_4', _d(MAX_TW)),
        ('h_0_tw_sl', _d(MAX_TW)), ('h_0_tw_tw', _d(MAX_TW)), ('gamma_char_tw', _d(MAX_TW)),
        ('dt_mu', C.c_double), ('dt_nu', C.c_double), ('dt_kB_T', C.c_double), ('dt_Gamma_sf', C.c_double),
        ('dt_Q_cl', C.c_double), ('dt_omega', C.c_double), ('dt_D', C.c_double), ('dt_p_sb', C.c_double),
        ('dt_q_sb', C.c_double), ('dt_i_tw', C.c_double), ('dt_L_tw', C.c_double), ('dt_x_c', C.c_double),
        ('dt_V_cs', C.c_double), ('dt_tau_sb', C.c_double), ('dt_gamma_0_sb', C.c_double), ('dt_E_sb', C.c_double),
        ('dt_extended', C.c_int), ('dt_omit_dipoles', C.c_int),
        ('dt_b_sl', _d(MAX_SL)), ('dt_Q_sl', _d(MAX_SL)), ('dt_v_0', _d(MAX_SL)), ('dt_i_sl', _d(MAX_SL)),
        ('dt_p', _d(MAX_SL)), ('dt_q', _d(MAX_SL)), ('dt_tau_0', _d(MAX_SL)), ('dt_B', _d(MAX_SL)),
        ('dt_d_caron', _d(MAX_SL)),
        ('dt_n0_sl', _d(MAX_SL, 3)), ('dt_forest', _d(MAX_SL, MAX_SL)),
        ('dt_b_tw', _d(MAX_TW)), ('dt_t_tw', _d(MAX_TW)), ('dt_r', _d(MAX_TW)),
        ('dt_pair', (C.c_int * 2) * MAX_TW), ('dt_C66_tw', _d(MAX_TW, 36)),
        ('kh_chi_inf', _d(MAX_SL)), ('kh_h_0_xi', _d(MAX_SL)), ('kh_h_0_chi', _d(MAX_SL)), ('kh_h_inf_xi', _d(MAX_SL)),
        ('kh_h_inf_chi', _d(MAX_SL)),
        ('tu_h', _d(MAX_SL)), ('tu_w', _d(MAX_SL)), ('tu_omega', _d(MAX_SL)), ('tu_f_at', _d(MAX_SL)), ('tu_D_0', C.c_double),
    ]


ALLTOALL_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int)


class Comm(C.Structure):
    _fields_ = [('user', C.c_void_p), ('alltoall', ALLTOALL_FN), ('allreduce_sum', ALLREDUCE_FN),
                ('allreduce_max', ALLREDUCE_FN)]


class SolutionState(C.Structure):
    _fields_ = [('required_iterations', C.c_int), ('converged', C.c_int), ('stag_converged', C.c_int),
                ('terminally_ill', C.c_int)]


class GridNumerics(C.Structure):
    """dmk_grid_numerics: the grid-level keys of Config::parse_numerics (config.cpp:379-384, 449-457)"""
    _fields_ = [('N_staggered_iter_max', C.c_int), ('N_cutback_max', C.c_int), ('alpha', C.c_double), ('beta', C.c_double),
                ('eps_abs_curl', C.c_double), ('eps_rel_curl', C.c_double)]


class LoadStep(C.Structure):
    """dmk_load_step (Config::LoadStep, config.h:74-96): values / masks / R row-major 3x3, mask 1 = 'x'"""
    _fields_ = [('deformation_type', C.c_char * 8), ('deformation_values', C.c_double * 9), ('deformation_mask', C.c_int * 9),
                ('stress_type', C.c_char * 8), ('stress_values', C.c_double * 9), ('stress_mask', C.c_int * 9),
                ('R', C.c_double * 9), ('t', C.c_double), ('r', C.c_double), ('N', C.c_int),
                ('estimate_rate', C.c_int), ('f_out', C.c_int), ('f_restart', C.c_int)]


class IncrementRecord(C.Structure):
    _fields_ = [('inc', C.c_int), ('step', C.c_int), ('inc_in_step', C.c_int), ('f_out', C.c_int), ('t', C.c_double),
                ('converged', C.c_int), ('iterations', C.c_int)]


LOOP_FORWARD = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.POINTER(LoadStep))
LOOP_POINT_FORWARD = C.CFUNCTYPE(C.c_int, C.c_void_p)
LOOP_SOLVE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(SolutionState))
LOOP_RESULT = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(IncrementRecord))


class LoopOps(C.Structure):
    """dmk_loop_ops: what dmk_loop_over_loadcase asks of the solver underneath"""
    _fields_ = [('user', C.c_void_p), ('forward', LOOP_FORWARD), ('materialpoint_forward', LOOP_POINT_FORWARD),
                ('calculate_solution', LOOP_SOLVE), ('result', LOOP_RESULT)]


class DmkError(RuntimeError):
    pass


_lib = None


def load():
    """Loads the shared library (no fallback: raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DmkError('%s not built: run `python -c "import __graft_entry__ as g; g.build()"`' % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.dmk_last_error.restype = C.c_char_p
        _lib.dmk_fourier_doubles.restype = C.c_size_t
    return _lib


def check(rc):
    if rc != 0:
        raise DmkError('libdamask_grid_b200: rc=%d: %s' % (rc, load().dmk_last_error().decode()))
