"""material.yaml -> device phase tables (dmk_phase) and per-cell phase / orientation arrays.

Host-side set-up the reference performs once in f_materialpoint_initDamask:
  elastic Hooke C66(T)                /root/reference/src/phase_mechanical_elastic.f90:24-104
  phenopowerlaw parameters            /root/reference/src/phase_mechanical_plastic_phenopowerlaw.f90:72-290
  isotropic parameters                /root/reference/src/phase_mechanical_plastic_isotropic.f90:48-148
  dislotwin parameters                /root/reference/src/phase_mechanical_plastic_dislotwin.f90:129-452
  Fp0 = O, numerics                   /root/reference/src/phase_mechanical.f90:183-339
  polynomials in (T - T_ref)          /root/reference/src/polynomials.f90:65-110
Crystallography comes from the library's own C functions (csrc/crystal.cpp); nothing here touches oracle/."""
import ctypes as C

import numpy as np

from . import lib as L

T_ROOM = 293.15                # constants.f90:12
K_B = 1.380649e-23             # constants.f90:13
LAW_NONE, LAW_ISOTROPIC, LAW_PHENOPOWERLAW, LAW_DISLOTWIN, LAW_KINEHARDENING, LAW_DISLOTUNGSTEN = 0, 1, 2, 3, 4, 5


def _ia(v):
    a = np.ascontiguousarray(v, dtype=np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int))


def _da(v):
    a = np.ascontiguousarray(v, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def _check(rc):
    if rc != 0:
        raise L.DmkError(L.load().dmk_last_error().decode())


def schmid_slip(lattice, N_sl, non_schmid=None, sense=1):
    n = int(np.sum(N_sl))
    out = np.zeros((n, 9))
    Ni, Np = _ia(N_sl)
    ns, nsp = _da(non_schmid if non_schmid is not None and len(non_schmid) else [0.0])
    ncoef = 0 if non_schmid is None else len(non_schmid)
    _check(L.load().dmk_crystal_schmid_slip_ns(lattice.encode(), Np, len(Ni), nsp, ncoef, int(sense),
                                                out.ctypes.data_as(C.POINTER(C.c_double))))
    return out


def schmid_twin(lattice, N_tw):
    out = np.zeros((int(np.sum(N_tw)), 9))
    Ni, Np = _ia(N_tw)
    _check(L.load().dmk_crystal_schmid_twin(lattice.encode(), Np, len(Ni), out.ctypes.data_as(C.POINTER(C.c_double))))
    return out


def slip_normal(lattice, N_sl):
    out = np.zeros((int(np.sum(N_sl)), 3))
    Ni, Np = _ia(N_sl)
    _check(L.load().dmk_crystal_slip_normal(lattice.encode(), Np, len(Ni), out.ctypes.data_as(C.POINTER(C.c_double))))
    return out


def interaction(kind, lattice, N_react, N_act, values):
    out = np.zeros((int(np.sum(N_react)), int(np.sum(N_act))))
    Nr, Nrp = _ia(N_react)
    Na, Nap = _ia(N_act)
    v, vp = _da(values)
    _check(L.load().dmk_crystal_interaction(kind.encode(), lattice.encode(), Nrp, len(Nr), Nap, len(Na), vp, len(v),
                                            out.ctypes.data_as(C.POINTER(C.c_double))))
    return out


def forest_projection(lattice, N_sl, f_edge):
    n = int(np.sum(N_sl))
    out = np.zeros((n, n))
    Ni, Np = _ia(N_sl)
    f, fp = _da(f_edge)
    _check(L.load().dmk_crystal_forest_projection(lattice.encode(), Np, len(Ni), fp, out.ctypes.data_as(C.POINTER(C.c_double))))
    return out


def C66_twin(lattice, N_tw, C66):
    out = np.zeros((int(np.sum(N_tw)), 36))
    Ni, Np = _ia(N_tw)
    c, cp = _da(np.asarray(C66).reshape(36))
    _check(L.load().dmk_crystal_C66_twin(lattice.encode(), Np, len(Ni), cp, out.ctypes.data_as(C.POINTER(C.c_double))))
    return out


def twin_nucleation_pairs(N_tw):
    out = np.zeros((int(np.sum(N_tw)), 2), dtype=np.int32)
    Ni, Np = _ia(N_tw)
    _check(L.load().dmk_crystal_twin_nucleation_pairs(Np, len(Ni), out.ctypes.data_as(C.POINTER(C.c_int))))
    return out


def qu2om(q):
    qa, qp = _da(q)
    out = np.zeros(9)
    L.load().dmk_qu2om(qp, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out.reshape(3, 3)


# ------------------------------------------------------------------------------------------------
def poly_at(dct, key, T, var='T'):
    """value of a (possibly temperature dependent) parameter: key, key,T, key,T^2 ... around T_ref"""
    coef = [float(dct[key])]
    if f'{key},{var}' in dct:
        coef.append(float(dct[f'{key},{var}']))
    for o in (2, 3, 4):
        k = f'{key},{var}^{o}'
        if k in dct:
            coef += [0.0] * (o - len(coef)) + [float(dct[k])]
    if len(coef) == 1:
        return coef[0]
    x = T - float(dct[f'{var}_ref'])
    y = coef[-1]
    for c in coef[-2::-1]:
        y = y * x + c
    return y


def cubic_C66(C11, C12, C44):
    """crystal_symmetrize_C66 for cubic lattices (crystal.f90:1680-1720)"""
    Cm = np.zeros((6, 6))
    Cm[:3, :3] = C12
    Cm[np.arange(3), np.arange(3)] = C11
    Cm[np.arange(3, 6), np.arange(3, 6)] = C44
    return Cm


def isotropic_mu(C66, bound='isostrain'):
    b = bound.lower()
    if b == 'isostrain':
        return (C66[0, 0] - C66[0, 1] + C66[3, 3] * 3.0) / 5.0
    if b == 'isostress':
        return 5.0 / (4.0 / (C66[0, 0] - C66[0, 1]) + 3.0 / C66[3, 3])
    raise ValueError(bound)


def isotropic_nu(C66, bound='isostrain'):
    b = bound.lower()
    if b == 'isostrain':
        K = np.sum(C66[:3, :3]) / 9.0
    elif b == 'isostress':
        K = 1.0 / np.sum(np.linalg.inv(C66)[:3, :3])
    else:
        raise ValueError(bound)
    mu = isotropic_mu(C66, bound)
    return (1.5 * K - mu) / (3.0 * K + mu)


def _expand(vals, N):
    return np.repeat(np.asarray(vals, float), N)


def _set(field, values):
    a = np.ctypeslib.as_array(field)
    v = np.asarray(values)
    if any(s > t for s, t in zip(v.shape, a.shape)):
        raise L.DmkError('phase table overflow: %s into %s' % (v.shape, a.shape))
    a[tuple(slice(0, s) for s in v.shape)] = v


def make_phase(cfg, T=T_ROOM):
    """one `phase:` entry of material.yaml -> lib.Phase"""
    ph = L.Phase()
    if not isinstance(cfg, dict):
        raise L.DmkError('phase entry is empty')
    for key in ('lattice', 'mechanical'):                       # the same refusals, in the same words, as csrc/init_dropin.cpp
        if key not in cfg:
            raise L.DmkError("missing key '%s'" % key)
    lattice = cfg['lattice']
    if lattice not in ('cF', 'cI'):
        raise L.DmkError('lattice %s is not offered on the device path (cF, cI)' % lattice)
    mech = cfg['mechanical']
    if not isinstance(mech, dict) or 'elastic' not in mech:
        raise L.DmkError("missing key 'elastic'")
    el = mech['elastic']
    if el.get('type', 'Hooke') != 'Hooke':
        raise L.DmkError('elastic type %s' % el.get('type'))
    C66 = cubic_C66(poly_at(el, 'C_11', T), poly_at(el, 'C_12', T), poly_at(el, 'C_44', T))
    _set(ph.C66, C66.reshape(36))
    pl = mech.get('plastic')
    kind = 'none' if pl is None else pl['type']
    state_init, atol = np.zeros(0), np.zeros(0)
    if kind == 'none':
        ph.law = LAW_NONE
    elif kind == 'isotropic':
        ph.law = LAW_ISOTROPIC
        ph.iso_xi_inf, ph.iso_dot_gamma_0, ph.iso_n = pl['xi_inf'], pl['dot_gamma_0'], pl['n']
        ph.iso_h_0, ph.iso_h, ph.iso_M, ph.iso_h_ln = pl['h_0'], pl.get('h', 3.0), pl['M'], pl.get('h_ln', 0.0)
        ph.iso_c_1, ph.iso_c_2, ph.iso_c_3, ph.iso_c_4 = (pl.get('c_1', 0.0), pl.get('c_2', 0.0), pl.get('c_3', 0.0),
                                                          pl.get('c_4', 0.0))
        ph.iso_a = pl['a']
        ph.iso_dilatation = int(bool(pl.get('dilatation', False)))
        state_init = np.array([float(pl['xi_0'])])
        atol = np.array([float(pl.get('atol_xi', 1.0))])
    elif kind == 'phenopowerlaw':
        ph.law = LAW_PHENOPOWERLAW
        N_sl, N_tw = list(pl.get('N_sl', [])), list(pl.get('N_tw', []))
        ns, nt = int(np.sum(N_sl)), int(np.sum(N_tw))
        ph.n_sl, ph.n_tw = ns, nt
        xi0_sl, xi0_tw = np.zeros(0), np.zeros(0)
        if ns:
            one = np.ones(len(N_sl))
            _set(ph.dot_gamma_0_sl, _expand(pl.get('dot_gamma_0_sl', one), N_sl))
            _set(ph.n_sl_exp, _expand(pl.get('n_sl', one), N_sl))
            _set(ph.a_sl, _expand(pl.get('a_sl', one), N_sl))
            _set(ph.h_0_sl_sl, _expand(pl.get('h_0_sl-sl', 0 * one), N_sl))
            _set(ph.c_1, _expand(pl.get('c_1', 0 * one), N_sl))
            _set(ph.c_2, _expand(pl.get('c_2', one), N_sl))
            xi0_sl = _expand(pl['xi_0_sl'], N_sl)
            _set(ph.xi_inf_sl, _expand(pl['xi_inf_sl'], N_sl))
            _set(ph.f_sat_sl_tw, _expand(pl.get('f_sat_sl-tw', 0 * one), N_sl))
            _set(ph.h_sl_sl, interaction('sl-sl', lattice, N_sl, N_sl, pl['h_sl-sl']))
            P = schmid_slip(lattice, N_sl)
            _set(ph.P_sl, P)
            if lattice == 'cI':
                a_nS = pl.get('a_nonSchmid_110', [])
                _set(ph.P_nS_pos, schmid_slip(lattice, N_sl, a_nS, +1))
                _set(ph.P_nS_neg, schmid_slip(lattice, N_sl, a_nS, -1))
            else:
                _set(ph.P_nS_pos, P)
                _set(ph.P_nS_neg, -P)
        if nt:
            one = np.ones(len(N_tw))
            _set(ph.dot_gamma_0_tw, _expand(pl.get('dot_gamma_0_tw', one), N_tw))
            _set(ph.n_tw_exp, _expand(pl.get('n_tw', one), N_tw))
            _set(ph.c_3, _expand(pl.get('c_3', 0 * one), N_tw))
            _set(ph.c_4, _expand(pl.get('c_4', 0 * one), N_tw))
            _set(ph.h_0_tw_tw, _expand(pl.get('h_0_tw-tw', 0 * one), N_tw))
            xi0_tw = _expand(pl['xi_0_tw'], N_tw)
            _set(ph.gamma_char_tw, np.full(nt, 0.5 * np.sqrt(2.0)))          # crystal.f90:466-473 (cF, cI)
            _set(ph.h_tw_tw, interaction('tw-tw', lattice, N_tw, N_tw, pl['h_tw-tw']))
            _set(ph.P_tw, schmid_twin(lattice, N_tw))
        if ns and nt:
            _set(ph.h_0_tw_sl, _expand(pl.get('h_0_tw-sl', np.zeros(len(N_tw))), N_tw))
            _set(ph.h_sl_tw, interaction('sl-tw', lattice, N_sl, N_tw, pl['h_sl-tw']))
            _set(ph.h_tw_sl, interaction('tw-sl', lattice, N_tw, N_sl, pl['h_tw-sl']))
        state_init = np.concatenate([xi0_sl, xi0_tw, np.zeros(ns), np.zeros(nt)])
        atol = np.concatenate([np.full(ns + nt, float(pl.get('atol_xi', 1.0))),
                               np.full(ns + nt, float(pl.get('atol_gamma', 1.0e-6)))])
    elif kind == 'kinehardening':                    # plastic_kinehardening_init, phase_mechanical_plastic_kinehardening.f90:72-252
        ph.law = LAW_KINEHARDENING
        N_sl = list(pl.get('N_sl', []))
        ns = int(np.sum(N_sl))
        if ns == 0:
            raise L.DmkError('kinehardening without slip systems')
        if 6 * ns > L.MAX_STATE:
            raise L.DmkError('kinehardening: at most %d slip systems on the device path' % (L.MAX_STATE // 6))
        ph.n_sl, ph.n_tw = ns, 0
        P = schmid_slip(lattice, N_sl)
        _set(ph.P_sl, P)
        if lattice == 'cI':
            a_nS = pl.get('a_nonSchmid_110', [])
            _set(ph.P_nS_pos, schmid_slip(lattice, N_sl, a_nS, +1))
            _set(ph.P_nS_neg, schmid_slip(lattice, N_sl, a_nS, -1))
        else:
            _set(ph.P_nS_pos, P)
            _set(ph.P_nS_neg, -P)
        _set(ph.h_sl_sl, interaction('sl-sl', lattice, N_sl, N_sl, pl['h_sl-sl']))
        _set(ph.dot_gamma_0_sl, _expand(pl['dot_gamma_0'], N_sl))
        _set(ph.n_sl_exp, _expand(pl['n'], N_sl))
        _set(ph.xi_inf_sl, _expand(pl['xi_inf'], N_sl))
        _set(ph.kh_chi_inf, _expand(pl['chi_inf'], N_sl))
        _set(ph.kh_h_0_xi, _expand(pl['h_0_xi'], N_sl))
        _set(ph.kh_h_0_chi, _expand(pl['h_0_chi'], N_sl))
        _set(ph.kh_h_inf_xi, _expand(pl['h_inf_xi'], N_sl))
        _set(ph.kh_h_inf_chi, _expand(pl['h_inf_chi'], N_sl))
        for key in ('dot_gamma_0', 'n', 'xi_0', 'xi_inf', 'chi_inf'):           # the reference's sanity checks (:164-169)
            if np.any(np.asarray(pl[key], float) <= 0.0):
                raise L.DmkError('kinehardening: %s must be positive' % key)
        state_init = np.concatenate([_expand(pl['xi_0'], N_sl), np.zeros(5 * ns)])
        atol = np.concatenate([np.full(2 * ns, float(pl.get('atol_xi', 1.0))), np.full(ns, float(pl.get('atol_gamma', 1.0e-6))),
                               np.zeros(3 * ns)])
    elif kind == 'dislotungsten':                    # plastic_dislotungsten_init, phase_mechanical_plastic_dislotungsten.f90:79-268
        ph.law = LAW_DISLOTUNGSTEN
        N_sl = list(pl.get('N_sl', []))
        ns = int(np.sum(N_sl))
        if ns == 0:
            raise L.DmkError('dislotungsten without slip systems')
        ph.n_sl, ph.n_tw = ns, 0
        bound = pl.get('isotropic_bound', 'isostrain')
        ph.dt_mu, ph.dt_nu = isotropic_mu(C66, bound), isotropic_nu(C66, bound)
        ph.dt_kB_T = K_B * T
        P = schmid_slip(lattice, N_sl)
        _set(ph.P_sl, P)
        if lattice == 'cI':
            a_nS = pl.get('a_nonSchmid_110', [])
            _set(ph.P_nS_pos, schmid_slip(lattice, N_sl, a_nS, +1))
            _set(ph.P_nS_neg, schmid_slip(lattice, N_sl, a_nS, -1))
        else:
            _set(ph.P_nS_pos, P)
            _set(ph.P_nS_neg, -P)
        ph.dt_omit_dipoles = int(bool(pl.get('no_dipole_formation', False)))
        ph.dt_D, ph.tu_D_0, ph.dt_Q_cl = pl['D'], pl['D_0'], pl['Q_cl']
        f_edge = _expand(pl.get('f_edge', [0.5] * len(N_sl)), N_sl)
        rho_mob_0, rho_dip_0 = _expand(pl['rho_mob_0'], N_sl), _expand(pl['rho_dip_0'], N_sl)
        b_sl = _expand(pl['b_sl'], N_sl)
        _set(ph.dt_b_sl, b_sl); _set(ph.dt_Q_sl, _expand(pl['Q_s'], N_sl)); _set(ph.dt_i_sl, _expand(pl['i_sl'], N_sl))
        _set(ph.dt_tau_0, _expand(pl['tau_Peierls'], N_sl)); _set(ph.dt_p, _expand(pl['p_sl'], N_sl))
        _set(ph.dt_q, _expand(pl['q_sl'], N_sl)); _set(ph.dt_B, _expand(pl['B'], N_sl))
        _set(ph.tu_h, _expand(pl['h'], N_sl)); _set(ph.tu_w, _expand(pl['w'], N_sl)); _set(ph.tu_omega, _expand(pl['omega'], N_sl))
        _set(ph.dt_d_caron, b_sl * float(pl['D_a'])); _set(ph.tu_f_at, b_sl * b_sl * b_sl * float(pl['f_at']))    # b_sl**3 as gfortran evaluates it: two products, not pow()
        _set(ph.h_sl_sl, interaction('sl-sl', lattice, N_sl, N_sl, pl['h_sl-sl']))
        _set(ph.dt_forest, forest_projection(lattice, N_sl, f_edge))
        if ph.tu_D_0 < 0.0 or ph.dt_Q_cl <= 0.0 or np.any(rho_mob_0 < 0.0) or np.any(rho_dip_0 < 0.0) or np.any(b_sl <= 0.0) \
                or np.any(np.asarray(pl['Q_s'], float) <= 0.0) or np.any(np.asarray(pl['tau_Peierls'], float) < 0.0) \
                or np.any(np.asarray(pl['B'], float) < 0.0) or float(pl['D_a']) < 0.0 or float(pl['f_at']) <= 0.0:
            raise L.DmkError('dislotungsten: parameter out of range')                   # the reference's sanity checks (:205-215)
        state_init = np.concatenate([rho_mob_0, rho_dip_0, np.zeros(ns)])
        atol = np.concatenate([np.full(2 * ns, float(pl.get('atol_rho', 1.0))), np.full(ns, float(pl.get('atol_gamma', 1.0e-6)))])
    elif kind == 'dislotwin':
        ph.law = LAW_DISLOTWIN
        N_sl, N_tw, N_tr = list(pl.get('N_sl', [])), list(pl.get('N_tw', [])), list(pl.get('N_tr', []))
        ns, nt = int(np.sum(N_sl)), int(np.sum(N_tw))
        if int(np.sum(N_tr)) > 0:
            raise L.DmkError('dislotwin transformation systems (N_tr) are not offered on the device path')
        if ns == 0:
            raise L.DmkError('dislotwin without slip systems')
        ph.n_sl, ph.n_tw = ns, nt
        bound = pl.get('isotropic_bound', 'isostrain')
        ph.dt_mu, ph.dt_nu = isotropic_mu(C66, bound), isotropic_nu(C66, bound)
        ph.dt_kB_T = K_B * T
        _set(ph.P_sl, schmid_slip(lattice, N_sl))
        _set(ph.dt_n0_sl, slip_normal(lattice, N_sl))
        ph.dt_extended = int(bool(pl.get('extend_dislocations', False)))
        ph.dt_omit_dipoles = int(bool(pl.get('omit_dipoles', False)))
        ph.dt_Q_cl = pl['Q_cl']
        nf = len(N_sl)
        f_edge = _expand(pl.get('f_edge', [0.5] * nf), N_sl)
        rho_mob_0, rho_dip_0 = _expand(pl['rho_mob_0'], N_sl), _expand(pl['rho_dip_0'], N_sl)
        b_sl = _expand(pl['b_sl'], N_sl)
        _set(ph.dt_v_0, _expand(pl['v_0'], N_sl)); _set(ph.dt_b_sl, b_sl); _set(ph.dt_Q_sl, _expand(pl['Q_sl'], N_sl))
        _set(ph.dt_i_sl, _expand(pl['i_sl'], N_sl)); _set(ph.dt_p, _expand(pl['p_sl'], N_sl))
        _set(ph.dt_q, _expand(pl['q_sl'], N_sl)); _set(ph.dt_tau_0, _expand(pl['tau_0'], N_sl))
        _set(ph.dt_B, _expand(pl.get('B', [0.0] * nf), N_sl)); _set(ph.dt_d_caron, b_sl * float(pl['D_a']))
        _set(ph.h_sl_sl, interaction('sl-sl', lattice, N_sl, N_sl, pl['h_sl-sl']))
        _set(ph.dt_forest, forest_projection(lattice, N_sl, f_edge))
        fcc_nucleation = lattice == 'cF' and N_sl[0] == 12
        ph.dt_omega = float(pl.get('omega', 1000.0)) * (12.0 if lattice in ('cF', 'hP') else 8.0)
        if nt:
            if not fcc_nucleation or len(N_tw) != 1:
                raise L.DmkError('twinning in dislotwin needs cF with the 12 {111} slip systems')
            _set(ph.P_tw, schmid_twin(lattice, N_tw))
            _set(ph.gamma_char_tw, np.full(nt, 0.5 * np.sqrt(2.0)))
            ph.dt_L_tw, ph.dt_i_tw = pl['L_tw'], pl['i_tw']
            _set(ph.dt_b_tw, _expand(pl['b_tw'], N_tw)); _set(ph.dt_t_tw, _expand(pl['t_tw'], N_tw))
            _set(ph.dt_r, _expand(pl['p_tw'], N_tw))
            _set(ph.h_tw_tw, interaction('tw-tw', lattice, N_tw, N_tw, pl['h_tw-tw']))
            _set(ph.h_sl_tw, interaction('sl-tw', lattice, N_sl, N_tw, pl['h_sl-tw']))
            _set(ph.dt_pair, twin_nucleation_pairs(N_tw))
            _set(ph.dt_C66_tw, C66_twin(lattice, N_tw, C66))
            ph.dt_x_c, ph.dt_V_cs = pl['x_c'], pl['V_cs']
        if nt or ph.dt_extended:
            ph.dt_Gamma_sf = poly_at(pl, 'Gamma_sf', T)
        ph.dt_gamma_0_sb = float(pl.get('gamma_0_sb', 0.0))
        ph.dt_tau_sb = ph.dt_E_sb = ph.dt_p_sb = ph.dt_q_sb = 1.0
        if ph.dt_gamma_0_sb > 0.0:
            ph.dt_tau_sb, ph.dt_E_sb, ph.dt_p_sb, ph.dt_q_sb = pl['tau_sb'], pl['Q_sb'], pl['p_sb'], pl['q_sb']
        ph.dt_D = pl['D']
        state_init = np.concatenate([rho_mob_0, rho_dip_0, np.zeros(ns), np.zeros(nt)])
        atol = np.concatenate([np.full(2 * ns, float(pl.get('atol_rho', 1.0))), np.full(ns, float(pl.get('atol_gamma', 1.0e-6))),
                               np.full(nt, float(pl.get('atol_f_tw', 1.0e-6)))])
    else:
        raise L.DmkError('plastic law %s is not offered on the device path' % kind)
    ph.n_state = len(state_init)
    _set(ph.state_init, state_init)
    _set(ph.atol, atol)
    return ph


NUMERICS_DEFAULTS = dict(r_cutback_min=1.0e-3, r_cutback=0.25, r_increase=1.5, eps_rel_state=1.0e-6, N_iter_state_max=20,
                         N_iter_Lp_max=40, r_linesearch_Lp=0.5, eps_rel_Lp=1.0e-6, eps_abs_Lp=1.0e-8, f_update_jacobi_Lp=1,
                         integrator_state=0,
                         itmin=1, itmax=100, update_gamma=0, eps_abs_div=1.0e-4, eps_rel_div=5.0e-4, eps_abs_P=1.0e3,
                         eps_rel_P=1.0e-3)


INTEGRATORS = {'FPI': 0, 'Euler': 1, 'AdaptiveEuler': 2, 'RK4': 3, 'RKCK45': 4}


def make_numerics(phase_mech=None, grid_mech=None):
    """numerics.yaml `phase: mechanical:` and `solver: grid: mechanical:` dicts -> lib.Numerics (defaults:
    phase_mechanical.f90:277-294, /root/reference/src/grid/config.h:50-58)"""
    d = dict(NUMERICS_DEFAULTS)
    src = dict(phase_mech or {})
    pl = src.pop('plastic', {}) or {}
    known_top = ('r_cutback_min', 'r_cutback', 'r_increase', 'eps_rel_state', 'N_iter_state_max')
    known_pl = ('N_iter_Lp_max', 'r_linesearch_Lp', 'eps_rel_Lp', 'eps_abs_Lp', 'f_update_jacobi_Lp')
    # Li keys (phase_mechanical.f90:285-294) are accepted and unused: no law with Li != 0 is offered on the device path
    known_ei = ('N_iter_Li_max', 'r_linesearch_Li', 'eps_rel_Li', 'eps_abs_Li', 'f_update_jacobi_Li')
    ei = src.pop('eigen', {}) or {}
    for name, dct, known in (('phase/mechanical', src, known_top), ('phase/mechanical/plastic', pl, known_pl + ('integrator_state',)),
                             ('phase/mechanical/eigen', ei, known_ei)):
        bad = [k for k in dct if k not in known]
        if bad:
            raise L.DmkError('numerics: unknown key(s) %s under %s' % (bad, name))
    for k in known_top:
        if k in src:
            d[k] = src[k]
    for k in known_pl:
        if k in pl:
            d[k] = pl[k]
    integ = pl.get('integrator_state', 'FPI')
    if integ not in INTEGRATORS:
        raise L.DmkError('unknown state integrator %s' % integ)
    d['integrator_state'] = INTEGRATORS[integ]
    # `solver: grid: mechanical:` keys exactly as Config::parse_numerics reads them (config.cpp:386-402)
    g = dict(grid_mech or {})
    ren = {'N_iter_min': 'itmin', 'N_iter_max': 'itmax', 'eps_abs_div(P)': 'eps_abs_div', 'eps_rel_div(P)': 'eps_rel_div',
           'eps_abs_P': 'eps_abs_P', 'eps_rel_P': 'eps_rel_P', 'update_gamma': 'update_gamma',
           'itmin': 'itmin', 'itmax': 'itmax', 'eps_abs_div': 'eps_abs_div', 'eps_rel_div': 'eps_rel_div'}   # + the struct's own names
    for k, v in g.items():
        if k not in ren:
            raise L.DmkError('numerics: unknown key %r under solver/grid/mechanical (known: %s)' % (k, ', '.join(sorted(ren))))
        k = ren[k]
        d[k] = int(v) if k in ('itmin', 'itmax', 'update_gamma') else float(v)
    # the range checks of upstream DAMASK (grid_mechanical_spectral_basic_init, IO_error 301; phase_mechanical.f90:296-309):
    # Config::parse_numerics does not check, the values would be used as they come
    wrong = [k for k in ('eps_abs_div', 'eps_rel_div', 'eps_abs_P', 'eps_rel_P', 'r_cutback_min', 'r_cutback', 'r_increase',
                         'eps_rel_state', 'r_linesearch_Lp', 'eps_rel_Lp', 'eps_abs_Lp') if not float(d[k]) > 0.0]
    wrong += [k for k in ('N_iter_state_max', 'N_iter_Lp_max', 'f_update_jacobi_Lp') if int(d[k]) < 1]
    if int(d['itmax']) <= 1:
        wrong.append('N_iter_max')
    if int(d['itmin']) > int(d['itmax']) or int(d['itmin']) < 1:
        wrong.append('N_iter_min')
    if wrong:
        raise L.DmkError('numerics: value out of range: ' + ', '.join(wrong))
    n = L.Numerics()
    for k, v in d.items():
        setattr(n, k, v)
    return n


class Material:
    """phases + per-point assignment for one z-slab.

    One constituent per cell ("pass" homogenisation) or several ("isostrain", homogenization_mechanical_isostrain.f90:36-45):
    material point k * n_cells + cell is constituent k of the cell.  All materials in use must have the same number of
    constituents; their volume fractions sum to 1 (material.f90:145,161)."""

    def __init__(self, material, material_at, T=T_ROOM):
        names = list(material['phase'].keys())
        self.phase_names = names
        built = []
        for k in names:
            try:
                built.append(make_phase(material['phase'][k], T))
            except KeyError as e:                               # a parameter the law needs is not in the deck
                raise L.DmkError('phase %s: missing key %s' % (k, e)) from None
            except L.DmkError as e:
                raise L.DmkError('phase %s: %s' % (k, e)) from None
        self.phases = (L.Phase * len(names))(*built)
        mats = material['material']
        material_at = np.asarray(material_at).reshape(-1)
        used = np.unique(material_at)
        if used.size and (used[0] < 0 or used[-1] >= len(mats)):
            raise L.DmkError('more materials requested than found in the material configuration')      # material.f90:116
        counts = {len(mats[m]['constituents']) for m in used}
        if len(counts) != 1:
            raise L.DmkError('materials with different numbers of constituents in one grid are not offered')
        self.n_constituents = nc = counts.pop()
        ph_of_mat = np.zeros((len(mats), nc), dtype=np.int32)
        O_of_mat = np.tile(np.eye(3).reshape(9), (len(mats), nc, 1))
        v_of_mat = np.zeros((len(mats), nc))
        for m in used:
            entry = mats[m]
            cons = entry['constituents']
            homog = (material.get('homogenization') or {}).get(entry.get('homogenization'))
            if homog is not None:
                if int(homog.get('N_constituents', len(cons))) != len(cons):
                    raise L.DmkError('material %d: %d constituents, its homogenization asks for %s'
                                     % (m, len(cons), homog.get('N_constituents')))
                htype = (homog.get('mechanical') or {}).get('type', 'pass')
                if htype not in ('pass', 'isostrain') or (htype == 'pass' and len(cons) != 1):
                    raise L.DmkError('mechanical homogenisation %r with %d constituents is not offered (pass, isostrain)'
                                     % (htype, len(cons)))
            elif len(cons) != 1:
                raise L.DmkError('material %d: %d constituents without a homogenization entry' % (m, len(cons)))
            for k, con in enumerate(cons):
                ph_of_mat[m, k] = names.index(con['phase'])
                O_of_mat[m, k] = qu2om(con['O']).reshape(9)
                v_of_mat[m, k] = float(con.get('v', 1.0))
            if abs(v_of_mat[m].sum() - 1.0) > 1.0e-9:
                raise L.DmkError('material %d: volume fractions of the constituents sum to %r' % (m, v_of_mat[m].sum()))
        n = material_at.size
        self.n_cells = n
        self.phase_of_point = np.ascontiguousarray(ph_of_mat[material_at].T.reshape(nc * n), dtype=np.int32)
        self.O_of_point = np.ascontiguousarray(O_of_mat[material_at].transpose(1, 0, 2).reshape(nc * n, 9))
        self.v_of_point = np.ascontiguousarray(v_of_mat[material_at].T.reshape(nc * n))
        # one constituent per cell: the per-cell names of dmk_material_init
        self.phase_of_cell = self.phase_of_point
        self.O_of_cell = self.O_of_point
