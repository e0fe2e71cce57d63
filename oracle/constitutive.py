"""ORACLE (test infrastructure only -- never imported by the product path).

ctypes front-end of oracle/constitutive.c plus the parameter set-up the reference does at init
(material.yaml -> per-phase parameter containers):
  elastic_init / elastic_C66          /root/reference/src/phase_mechanical_elastic.f90:24-104
  plastic_phenopowerlaw_init          .../phase_mechanical_plastic_phenopowerlaw.f90:72-290
  plastic_isotropic_init              .../phase_mechanical_plastic_isotropic.f90:48-148
  plastic_dislotwin_init              .../phase_mechanical_plastic_dislotwin.f90:129-452
  mechanical_init (Fp0 = O, numerics) .../phase_mechanical.f90:183-339
  polynomial evaluation               .../polynomials.f90:65-110 (Horner in (T - T_ref))
All tensors here are "natural" numpy arrays [..., i, j]; the reference's Fortran memory order is the
transpose of the last two axes (see oracle/layout notes in DESIGN.md).
"""
import ctypes as C
import numpy as np

from . import crystal
from .phase_layout import (Phase, Numerics, PointState, STATE_FIELDS, MAX_SL, MAX_TW,
                           LAW_NONE, LAW_ISOTROPIC, LAW_PHENOPOWERLAW, LAW_DISLOTWIN, LAW_KINEHARDENING, LAW_DISLOTUNGSTEN)
from .build import build

T_ROOM = 293.15            # constants.f90:12
K_B = 1.380649e-23         # constants.f90:13

NUMERICS_DEFAULTS = dict(r_cutback_min=1.0e-3, r_cutback=0.25, r_increase=1.5, eps_rel_state=1.0e-6,
                         N_iter_state_max=20, N_iter_Lp_max=40, r_linesearch_Lp=0.5, eps_rel_Lp=1.0e-6,
                         eps_abs_Lp=1.0e-8, f_update_jacobi_Lp=1, N_iter_Li_max=40, r_linesearch_Li=0.5,
                         eps_rel_Li=None, eps_abs_Li=None, f_update_jacobi_Li=None)


INTEGRATORS = {'FPI': 0, 'Euler': 1, 'AdaptiveEuler': 2, 'RK4': 3, 'RKCK45': 4}


def make_numerics(num_phase_mech=None):
    """num_phase_mech: the `phase: mechanical:` dict of numerics.yaml (may be None)."""
    d = dict(NUMERICS_DEFAULTS)
    src = dict(num_phase_mech or {})
    pl = src.pop('plastic', {}) or {}
    ei = src.pop('eigen', {}) or {}
    for k in ('r_cutback_min', 'r_cutback', 'r_increase', 'eps_rel_state', 'N_iter_state_max'):
        if k in src:
            d[k] = src[k]
    for k in ('N_iter_Lp_max', 'r_linesearch_Lp', 'eps_rel_Lp', 'eps_abs_Lp', 'f_update_jacobi_Lp'):
        if k in pl:
            d[k] = pl[k]
    for k in ('N_iter_Li_max', 'r_linesearch_Li', 'eps_rel_Li', 'eps_abs_Li', 'f_update_jacobi_Li'):
        if k in ei:
            d[k] = ei[k]
    if d['eps_rel_Li'] is None:
        d['eps_rel_Li'] = d['eps_rel_Lp']
    if d['eps_abs_Li'] is None:
        d['eps_abs_Li'] = d['eps_abs_Lp']
    if d['f_update_jacobi_Li'] is None:
        d['f_update_jacobi_Li'] = d['f_update_jacobi_Lp']
    d['integrator_state'] = INTEGRATORS[pl.get('integrator_state', 'FPI')]
    n = Numerics()
    for k, v in d.items():
        setattr(n, k, v)
    return n


def poly_at(dct, key, T, var='T'):
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


def _expand(vals, N):
    return np.repeat(np.asarray(vals, float), N)


def _set(arr_field, values):
    a = np.ctypeslib.as_array(arr_field)
    v = np.asarray(values)
    a[tuple(slice(0, s) for s in v.shape)] = v


def make_phase(name, phase_cfg, T=T_ROOM):
    """material.yaml phase entry -> Phase struct."""
    ph = Phase()
    lattice = phase_cfg['lattice']
    if lattice not in ('cF', 'cI'):
        raise NotImplementedError('lattice ' + lattice)
    mech = phase_cfg['mechanical']
    el = mech['elastic']
    if el.get('type', 'Hooke') != 'Hooke':
        raise NotImplementedError(el.get('type'))
    C66 = crystal.symmetrize_C66(poly_at(el, 'C_11', T), poly_at(el, 'C_12', T), poly_at(el, 'C_44', T))
    _set(ph.C66, C66.reshape(36))
    pl = mech.get('plastic', None)
    ptype = 'none' if pl is None else pl['type']
    state_init = np.zeros(0)
    atol = np.zeros(0)
    if ptype == 'none':
        ph.law = LAW_NONE
    elif ptype == 'isotropic':
        ph.law = LAW_ISOTROPIC
        ph.iso_xi_inf = pl['xi_inf']; ph.iso_dot_gamma_0 = pl['dot_gamma_0']; ph.iso_n = pl['n']
        ph.iso_h_0 = pl['h_0']; ph.iso_h = pl.get('h', 3.0); ph.iso_M = pl['M']; ph.iso_h_ln = pl.get('h_ln', 0.0)
        ph.iso_c_1 = pl.get('c_1', 0.0); ph.iso_c_2 = pl.get('c_2', 0.0); ph.iso_c_3 = pl.get('c_3', 0.0)
        ph.iso_c_4 = pl.get('c_4', 0.0); ph.iso_a = pl['a']; ph.iso_dilatation = int(bool(pl.get('dilatation', False)))
        state_init = np.array([float(pl['xi_0'])])
        atol = np.array([float(pl.get('atol_xi', 1.0))])
    elif ptype == 'phenopowerlaw':
        ph.law = LAW_PHENOPOWERLAW
        N_sl = list(pl.get('N_sl', [])); N_tw = list(pl.get('N_tw', []))
        ns, nt = int(np.sum(N_sl)), int(np.sum(N_tw))
        ph.n_sl, ph.n_tw = ns, nt
        xi0_sl = np.zeros(0); xi0_tw = np.zeros(0)
        if ns > 0:
            nf = len(N_sl); ones = np.ones(nf)
            _set(ph.dot_gamma_0_sl, _expand(pl.get('dot_gamma_0_sl', ones), N_sl))
            _set(ph.n_sl_exp, _expand(pl.get('n_sl', ones), N_sl))
            _set(ph.a_sl, _expand(pl.get('a_sl', ones), N_sl))
            _set(ph.h_0_sl_sl, _expand(pl.get('h_0_sl-sl', 0 * ones), N_sl))
            _set(ph.c_1, _expand(pl.get('c_1', 0 * ones), N_sl))
            _set(ph.c_2, _expand(pl.get('c_2', ones), N_sl))
            xi0_sl = _expand(pl['xi_0_sl'], N_sl)
            _set(ph.xi_inf_sl, _expand(pl['xi_inf_sl'], N_sl))
            _set(ph.f_sat_sl_tw, _expand(pl.get('f_sat_sl-tw', 0 * ones), N_sl))
            _set(ph.h_sl_sl, crystal.interaction_slip_slip(N_sl, pl['h_sl-sl'], lattice))
            P = crystal.schmid_slip(N_sl, lattice)
            _set(ph.P_sl, P.reshape(ns, 9))
            if lattice == 'cI':
                a_nS = pl.get('a_nonSchmid_110', [])
                _set(ph.P_nS_pos, crystal.schmid_slip(N_sl, lattice, a_nS, +1).reshape(ns, 9))
                _set(ph.P_nS_neg, crystal.schmid_slip(N_sl, lattice, a_nS, -1).reshape(ns, 9))
            else:
                _set(ph.P_nS_pos, P.reshape(ns, 9)); _set(ph.P_nS_neg, -P.reshape(ns, 9))
        if nt > 0:
            nf = len(N_tw); ones = np.ones(nf)
            _set(ph.dot_gamma_0_tw, _expand(pl.get('dot_gamma_0_tw', ones), N_tw))
            _set(ph.n_tw_exp, _expand(pl.get('n_tw', ones), N_tw))
            _set(ph.c_3, _expand(pl.get('c_3', 0 * ones), N_tw))
            _set(ph.c_4, _expand(pl.get('c_4', 0 * ones), N_tw))
            _set(ph.h_0_tw_tw, _expand(pl.get('h_0_tw-tw', 0 * ones), N_tw))
            xi0_tw = _expand(pl['xi_0_tw'], N_tw)
            _set(ph.gamma_char_tw, crystal.characteristic_shear_twin(N_tw, lattice))
            _set(ph.h_tw_tw, crystal.interaction_twin_twin(N_tw, pl['h_tw-tw'], lattice))
            _set(ph.P_tw, crystal.schmid_twin(N_tw, lattice).reshape(nt, 9))
        if ns > 0 and nt > 0:
            _set(ph.h_0_tw_sl, _expand(pl.get('h_0_tw-sl', 0 * np.ones(len(N_tw))), N_tw))
            _set(ph.h_sl_tw, crystal.interaction_slip_twin(N_sl, N_tw, pl['h_sl-tw'], lattice))
            _set(ph.h_tw_sl, crystal.interaction_twin_slip(N_tw, N_sl, pl['h_tw-sl'], lattice))
        state_init = np.concatenate([xi0_sl, xi0_tw, np.zeros(ns), np.zeros(nt)])
        atol = np.concatenate([np.full(ns + nt, float(pl.get('atol_xi', 1.0))),
                               np.full(ns + nt, float(pl.get('atol_gamma', 1.0e-6)))])
    elif ptype == 'kinehardening':                   # plastic_kinehardening_init, phase_mechanical_plastic_kinehardening.f90:72-252
        ph.law = LAW_KINEHARDENING
        N_sl = list(pl.get('N_sl', []))
        ns = int(np.sum(N_sl))
        if ns == 0:
            raise NotImplementedError('kinehardening without slip')
        ph.n_sl, ph.n_tw = ns, 0
        P = crystal.schmid_slip(N_sl, lattice)
        _set(ph.P_sl, P.reshape(ns, 9))
        if lattice == 'cI':
            a_nS = pl.get('a_nonSchmid_110', [])
            _set(ph.P_nS_pos, crystal.schmid_slip(N_sl, lattice, a_nS, +1).reshape(ns, 9))
            _set(ph.P_nS_neg, crystal.schmid_slip(N_sl, lattice, a_nS, -1).reshape(ns, 9))
        else:
            _set(ph.P_nS_pos, P.reshape(ns, 9)); _set(ph.P_nS_neg, -P.reshape(ns, 9))
        _set(ph.h_sl_sl, crystal.interaction_slip_slip(N_sl, pl['h_sl-sl'], lattice))
        _set(ph.dot_gamma_0_sl, _expand(pl['dot_gamma_0'], N_sl)); _set(ph.n_sl_exp, _expand(pl['n'], N_sl))
        _set(ph.xi_inf_sl, _expand(pl['xi_inf'], N_sl)); _set(ph.kh_chi_inf, _expand(pl['chi_inf'], N_sl))
        _set(ph.kh_h_0_xi, _expand(pl['h_0_xi'], N_sl)); _set(ph.kh_h_0_chi, _expand(pl['h_0_chi'], N_sl))
        _set(ph.kh_h_inf_xi, _expand(pl['h_inf_xi'], N_sl)); _set(ph.kh_h_inf_chi, _expand(pl['h_inf_chi'], N_sl))
        ph.n_dot = 3 * ns
        state_init = np.concatenate([_expand(pl['xi_0'], N_sl), np.zeros(5 * ns)])
        atol = np.concatenate([np.full(2 * ns, float(pl.get('atol_xi', 1.0))), np.full(ns, float(pl.get('atol_gamma', 1.0e-6))),
                               np.zeros(3 * ns)])
    elif ptype == 'dislotungsten':                   # plastic_dislotungsten_init, phase_mechanical_plastic_dislotungsten.f90:79-268
        ph.law = LAW_DISLOTUNGSTEN
        N_sl = list(pl.get('N_sl', []))
        ns = int(np.sum(N_sl))
        if ns == 0:
            raise NotImplementedError('dislotungsten without slip')
        ph.n_sl, ph.n_tw = ns, 0
        bound = pl.get('isotropic_bound', 'isostrain')
        ph.dt_mu = crystal.isotropic_mu(C66, bound); ph.dt_nu = crystal.isotropic_nu(C66, bound)
        ph.dt_kB_T = K_B * T
        P = crystal.schmid_slip(N_sl, lattice)
        _set(ph.P_sl, P.reshape(ns, 9))
        if lattice == 'cI':
            a_nS = pl.get('a_nonSchmid_110', [])
            _set(ph.P_nS_pos, crystal.schmid_slip(N_sl, lattice, a_nS, +1).reshape(ns, 9))
            _set(ph.P_nS_neg, crystal.schmid_slip(N_sl, lattice, a_nS, -1).reshape(ns, 9))
        else:
            _set(ph.P_nS_pos, P.reshape(ns, 9)); _set(ph.P_nS_neg, -P.reshape(ns, 9))
        ph.dt_omit_dipoles = int(bool(pl.get('no_dipole_formation', False)))
        ph.dt_D = pl['D']; ph.tu_D_0 = pl['D_0']; ph.dt_Q_cl = pl['Q_cl']
        nf = len(N_sl)
        f_edge = _expand(pl.get('f_edge', [0.5] * nf), N_sl)
        b_sl = _expand(pl['b_sl'], N_sl)
        _set(ph.dt_b_sl, b_sl); _set(ph.dt_Q_sl, _expand(pl['Q_s'], N_sl)); _set(ph.dt_i_sl, _expand(pl['i_sl'], N_sl))
        _set(ph.dt_tau_0, _expand(pl['tau_Peierls'], N_sl)); _set(ph.dt_p, _expand(pl['p_sl'], N_sl)); _set(ph.dt_q, _expand(pl['q_sl'], N_sl))
        _set(ph.tu_h, _expand(pl['h'], N_sl)); _set(ph.tu_w, _expand(pl['w'], N_sl)); _set(ph.tu_omega, _expand(pl['omega'], N_sl))
        _set(ph.dt_B, _expand(pl['B'], N_sl)); _set(ph.dt_d_caron, b_sl * float(pl['D_a'])); _set(ph.tu_f_at, b_sl * b_sl * b_sl * float(pl['f_at']))    # b_sl**3 as gfortran evaluates it: two products, not pow()
        _set(ph.h_sl_sl, crystal.interaction_slip_slip(N_sl, pl['h_sl-sl'], lattice))
        _set(ph.dt_forest, crystal.forest_projection(N_sl, lattice, f_edge))
        state_init = np.concatenate([_expand(pl['rho_mob_0'], N_sl), _expand(pl['rho_dip_0'], N_sl), np.zeros(ns)])
        atol = np.concatenate([np.full(2 * ns, float(pl.get('atol_rho', 1.0))), np.full(ns, float(pl.get('atol_gamma', 1.0e-6)))])
    elif ptype == 'dislotwin':
        ph.law = LAW_DISLOTWIN
        N_sl = list(pl.get('N_sl', [])); N_tw = list(pl.get('N_tw', [])); N_tr = list(pl.get('N_tr', []))
        ns, nt, ntr = int(np.sum(N_sl)), int(np.sum(N_tw)), int(np.sum(N_tr))
        if ntr > 0:
            raise NotImplementedError('dislotwin transformation systems are not restated (N_tr must be absent)')
        if ns == 0:
            raise NotImplementedError('dislotwin without slip')
        ph.n_sl, ph.n_tw, ph.n_tr = ns, nt, 0
        bound = pl.get('isotropic_bound', 'isostrain')
        ph.dt_mu = crystal.isotropic_mu(C66, bound); ph.dt_nu = crystal.isotropic_nu(C66, bound)
        ph.dt_kB_T = K_B * T
        _set(ph.P_sl, crystal.schmid_slip(N_sl, lattice).reshape(ns, 9))
        _set(ph.dt_n0_sl, crystal.slip_normal(N_sl, lattice))
        ph.dt_extended = int(bool(pl.get('extend_dislocations', False)))
        ph.dt_omit_dipoles = int(bool(pl.get('omit_dipoles', False)))
        ph.dt_Q_cl = pl['Q_cl']
        nf = len(N_sl)
        f_edge = _expand(pl.get('f_edge', [0.5] * nf), N_sl)
        rho_mob_0 = _expand(pl['rho_mob_0'], N_sl); rho_dip_0 = _expand(pl['rho_dip_0'], N_sl)
        b_sl = _expand(pl['b_sl'], N_sl)
        _set(ph.dt_v_0, _expand(pl['v_0'], N_sl)); _set(ph.dt_b_sl, b_sl); _set(ph.dt_Q_sl, _expand(pl['Q_sl'], N_sl))
        _set(ph.dt_i_sl, _expand(pl['i_sl'], N_sl)); _set(ph.dt_p, _expand(pl['p_sl'], N_sl))
        _set(ph.dt_q, _expand(pl['q_sl'], N_sl)); _set(ph.dt_tau_0, _expand(pl['tau_0'], N_sl))
        _set(ph.dt_B, _expand(pl.get('B', [0.0] * nf), N_sl)); _set(ph.dt_d_caron, b_sl * float(pl['D_a']))
        _set(ph.h_sl_sl, crystal.interaction_slip_slip(N_sl, pl['h_sl-sl'], lattice))
        _set(ph.dt_forest, crystal.forest_projection(N_sl, lattice, f_edge))
        ph.dt_fcc_nucleation = int(lattice == 'cF' and N_sl[0] == 12)
        ph.dt_omega = float(pl.get('omega', 1000.0)) * (12.0 if lattice in ('cF', 'hP') else 8.0)
        if nt > 0:
            if not ph.dt_fcc_nucleation or len(N_tw) != 1:
                raise ValueError('TWIP needs fcc with 12 {111} slip systems')
            _set(ph.P_tw, crystal.schmid_twin(N_tw, lattice).reshape(nt, 9))
            _set(ph.gamma_char_tw, crystal.characteristic_shear_twin(N_tw, lattice))
            ph.dt_L_tw = pl['L_tw']; ph.dt_i_tw = pl['i_tw']
            _set(ph.dt_b_tw, _expand(pl['b_tw'], N_tw)); _set(ph.dt_t_tw, _expand(pl['t_tw'], N_tw))
            _set(ph.dt_r, _expand(pl['p_tw'], N_tw))
            _set(ph.h_tw_tw, crystal.interaction_twin_twin(N_tw, pl['h_tw-tw'], lattice))
            _set(ph.h_sl_tw, crystal.interaction_slip_twin(N_sl, N_tw, pl['h_sl-tw'], lattice))
            pairs = crystal.CF_TWIN_NUCLEATION_SLIP_PAIR[crystal._active_rows(N_tw, crystal.CF_NTWIN)] - 1
            _set(ph.dt_pair, pairs.astype(np.int32))
            _set(ph.dt_C66_tw, crystal.C66_twin(N_tw, C66, lattice).reshape(nt, 36))
            ph.dt_x_c = pl['x_c']; ph.dt_V_cs = pl['V_cs']
        if nt > 0 or ph.dt_extended:
            ph.dt_Gamma_sf = poly_at(pl, 'Gamma_sf', T)
        ph.dt_gamma_0_sb = float(pl.get('gamma_0_sb', 0.0))
        ph.dt_tau_sb = 1.0; ph.dt_E_sb = 1.0; ph.dt_p_sb = 1.0; ph.dt_q_sb = 1.0
        if ph.dt_gamma_0_sb > 0.0:
            ph.dt_tau_sb = pl['tau_sb']; ph.dt_E_sb = pl['Q_sb']; ph.dt_p_sb = pl['p_sb']; ph.dt_q_sb = pl['q_sb']
        ph.dt_D = pl['D']
        state_init = np.concatenate([rho_mob_0, rho_dip_0, np.zeros(ns), np.zeros(nt)])
        atol = np.concatenate([np.full(2 * ns, float(pl.get('atol_rho', 1.0))),
                               np.full(ns, float(pl.get('atol_gamma', 1.0e-6))),
                               np.full(nt, float(pl.get('atol_f_tw', 1.0e-6)))])
    else:
        raise NotImplementedError('plastic law ' + ptype)
    ph.n_state = len(state_init)
    _set(ph.state_init, state_init)
    _set(ph.atol, atol)
    return ph


class ConstitutiveOracle:
    """Holds the per-point state of all material points and exposes the six reference entry points
    (homogenization.f90:219-324,491-510; materialpoint.f90:208).

    Homogenisation: "pass" (one constituent per cell) or "isostrain" (homogenization_mechanical_isostrain.f90:36-45: every
    constituent of a cell sees the cell's F; mechanical_homogenize, homogenization_mechanical.f90:125-145: P and dPdF of the cell
    are the sums over the constituents weighted by their volume fractions, constituent 1 first).  Material point
    k * n_cells + cell is constituent k of the cell; all materials in use must have the same number of constituents."""

    def __init__(self, material, material_at, num_phase_mech=None, T=T_ROOM, threads=None):
        self.lib = C.CDLL(build())
        assert self.lib.orc_sizeof_phase() == C.sizeof(Phase), 'struct layout mismatch'
        assert self.lib.orc_sizeof_num() == C.sizeof(Numerics)
        material_at = np.asarray(material_at).reshape(-1)
        self.n_cells = material_at.size
        phase_names = list(material['phase'].keys())
        self.phase_names = phase_names
        self.phases = (Phase * len(phase_names))(*[make_phase(k, material['phase'][k], T) for k in phase_names])
        self.num = make_numerics(num_phase_mech)
        mats = material['material']
        used = np.unique(material_at)
        counts = {len(mats[m]['constituents']) for m in used}
        if len(counts) != 1:
            raise NotImplementedError('materials with different numbers of constituents in one grid')
        self.n_const = nc = counts.pop()
        ph_of_mat = np.zeros((len(mats), nc), dtype=np.int32)
        O_of_mat = np.tile(np.eye(3), (len(mats), nc, 1, 1))
        v_of_mat = np.zeros((len(mats), nc))
        for m in used:
            entry = mats[m]
            cons = entry['constituents']
            homog = material.get('homogenization', {}).get(entry.get('homogenization'), None)
            if homog is not None:
                # material.f90:145 (IO_error 148) and the mechanical types this oracle restates
                if int(homog.get('N_constituents', len(cons))) != len(cons):
                    raise ValueError('material %d: %d constituents, homogenization asks for %s' % (m, len(cons), homog['N_constituents']))
                htype = homog.get('mechanical', {}).get('type', 'pass')
                if htype not in ('pass', 'isostrain') or (htype == 'pass' and len(cons) != 1):
                    raise NotImplementedError('mechanical homogenization %s with %d constituents' % (htype, len(cons)))
            for k, con in enumerate(cons):
                ph_of_mat[m, k] = phase_names.index(con['phase'])
                O_of_mat[m, k] = crystal.qu2om(con['O'])
                v_of_mat[m, k] = float(con.get('v', 1.0))
            if abs(v_of_mat[m].sum() - 1.0) > 1.0e-9:                # material.f90:161 (IO_error 153)
                raise ValueError('material %d: volume fractions sum to %r' % (m, v_of_mat[m].sum()))
        self.n = n = self.n_cells * nc                                # material points
        # point k * n_cells + cell
        self.v_of = np.ascontiguousarray(v_of_mat[material_at].T.reshape(n))
        O_of_pt = np.ascontiguousarray(O_of_mat[material_at].transpose(1, 0, 2, 3).reshape(n, 3, 3))
        ph_of_pt = ph_of_mat[material_at].T.reshape(n)
        self.phase_of = np.ascontiguousarray(ph_of_pt)
        self.max_state = max(1, max(p.n_state for p in self.phases))
        self.max_dep = max(1, max(2 * p.n_sl + p.n_tw for p in self.phases))
        eye = np.broadcast_to(np.eye(3).reshape(9), (n, 9))
        self.a = {}
        for k in STATE_FIELDS:
            self.a[k] = np.zeros((n, 9))
        self.a['state'] = np.zeros((n, self.max_state)); self.a['state0'] = np.zeros((n, self.max_state))
        self.a['dep'] = np.zeros((n, self.max_dep))
        # mechanical_init: F = I, Fp = O_0, Fe = V_e0 Fp^T (V_e0 = I), Fi = I; *0 = current
        Fp = O_of_pt.reshape(n, 9)
        self.a['F'][:] = eye; self.a['F0'][:] = eye
        self.a['Fp'][:] = Fp; self.a['Fp0'][:] = Fp
        self.a['Fe'][:] = O_of_pt.transpose(0, 2, 1).reshape(n, 9)
        self.a['Fi'][:] = eye; self.a['Fi0'][:] = eye
        for i, p in enumerate(self.phases):
            sel = self.phase_of == i
            init = np.ctypeslib.as_array(p.state_init)[:self.max_state]
            self.a['state'][sel] = init; self.a['state0'][sel] = init
        self.st = PointState()
        self.st.n = n; self.st.max_state = self.max_state; self.st.max_dep = self.max_dep
        self.st.phase_of = self.phase_of.ctypes.data_as(C.POINTER(C.c_int))
        for k in STATE_FIELDS:
            setattr(self.st, k, self.a[k].ctypes.data_as(C.POINTER(C.c_double)))
        self.lib.orc_init_dependent(self.phases, C.byref(self.st))
        self.lib.orc_mechanical_response.restype = C.c_int
        self.counters = np.zeros(6, dtype=np.int64)
        self.point_cutbacks = np.zeros(n, dtype=np.int32)       # per-cell cut-backs of the last response call
        self.point_lp_iter = np.zeros(n, dtype=np.int32)        # per-cell Lp Newton iterations of the last response call
        self.terminally_ill = False
        self.threads = threads

    # f_homogenization_mechanical_response (mechanical_partition: homogenization_mechanical.f90:95-122)
    def mechanical_response(self, F, dt, cell_start=0, cell_end=None):
        F = np.ascontiguousarray(np.asarray(F, float).reshape(self.n_cells, 9))
        if self.n_const > 1:
            if cell_start != 0 or cell_end not in (None, self.n_cells):
                raise NotImplementedError('cell ranges with several constituents per cell')
            F = np.ascontiguousarray(np.tile(F, (self.n_const, 1)))  # isostrain_partitionDeformation: F = spread(avgF)
            cell_end = self.n
        cell_end = self.n if cell_end is None else cell_end
        cnt = np.zeros(6, dtype=np.int64)
        self.point_cutbacks[:] = 0
        self.lib.orc_set_point_log(self.point_cutbacks.ctypes.data_as(C.POINTER(C.c_int)))
        self.point_lp_iter[:] = 0
        self.lib.orc_set_point_lp_log(self.point_lp_iter.ctypes.data_as(C.POINTER(C.c_int)))
        ill = self.lib.orc_mechanical_response(self.phases, C.byref(self.num), C.byref(self.st),
                                               F.ctypes.data_as(C.POINTER(C.c_double)), C.c_double(dt),
                                               C.c_int(cell_start), C.c_int(cell_end),
                                               cnt.ctypes.data_as(C.POINTER(C.c_long)))
        self.counters = cnt
        if ill:
            self.terminally_ill = True
        return bool(ill)

    # f_homogenization_mechanical_response2 (mechanical_homogenize: homogenization_mechanical.f90:125-145)
    def mechanical_response2(self, dt, tangent=True):
        P = np.zeros((self.n, 9))
        dPdF = np.zeros((self.n, 81)) if tangent else None
        self.lib.orc_mechanical_response2(self.phases, C.byref(self.num), C.byref(self.st), C.c_double(dt),
                                          C.c_int(0), C.c_int(self.n), P.ctypes.data_as(C.POINTER(C.c_double)),
                                          dPdF.ctypes.data_as(C.POINTER(C.c_double)) if tangent else None)
        if self.n_const > 1:
            nc, m = self.n_const, self.n_cells
            v = self.v_of.reshape(nc, m, 1)
            Pp = P.reshape(nc, m, 9)
            P = Pp[0] * v[0]
            for k in range(1, nc):
                P = P + Pp[k] * v[k]
            if tangent:
                Tp = dPdF.reshape(nc, m, 81)
                dPdF = Tp[0] * v[0]
                for k in range(1, nc):
                    dPdF = dPdF + Tp[k] * v[k]
        m = self.n_cells
        return P.reshape(m, 3, 3), (dPdF.reshape(m, 3, 3, 3, 3) if tangent else None)

    def set_threads(self, n):
        """OpenMP threads of the cell loops (0: the default); -> the number now in effect"""
        self.lib.orc_set_threads(C.c_int(int(n)))
        return int(self.lib.orc_get_threads())

    # f_materialpoint_forward
    def forward(self):
        self.lib.orc_forward(self.phases, C.byref(self.st))

    def snapshot(self):
        return {k: v.copy() for k, v in self.a.items()}

    def restore(self, snap):
        for k, v in snap.items():
            self.a[k][:] = v
