"""ORACLE (test infrastructure only -- never imported by the product path).

Restates the reference's mechanical spectral "basic" solver and load-case loop on top of oracle/spectral.py,
oracle/constitutive.py and oracle/ngmres.py, with upstream semantics (D3, D9, D10 applied):
  MechSolverBasic::init / calculate_solution / forward / converged / form_residual
                                 /root/reference/src/grid/mech_solver_basic.cpp:19-108,110-161,163-220,242-284,286-339
  Spectral::constitutive_response  /root/reference/src/grid/spectral.cpp:132-219
  DamaskGrid::loop_over_loadcase / forward_time / solve_fields   /root/reference/src/grid/damask_grid.cpp:36-196
  numerics defaults                /root/reference/src/grid/config.h:50-78
"""
import numpy as np

from . import spectral as sp
from .constitutive import ConstitutiveOracle
from .ngmres import NGMRES

GRID_MECH_DEFAULTS = dict(itmin=1, itmax=100, update_gamma=False, eps_abs_div=1.0e-4, eps_rel_div=5.0e-4,
                          eps_abs_P=1.0e3, eps_rel_P=1.0e-3)
SNES_BREAKOUT_INNER_ITER, SNES_DIVERGED_MAX_IT = 6, -5      # any reason > 0 counts as converged


class BoundaryCondition:
    def __init__(self, values=None, mask=None, kind=''):
        self.values = np.zeros((3, 3)) if values is None else np.asarray(values, float)
        self.mask = np.zeros((3, 3), bool) if mask is None else np.asarray(mask, bool)   # True = 'x' (free)
        self.type = kind


def parse_bc(entry):
    vals = np.zeros((3, 3)); mask = np.zeros((3, 3), bool)
    for i in range(3):
        for j in range(3):
            v = entry[i][j]
            if isinstance(v, str) and v.strip().lower() == 'x':
                mask[i, j] = True
            else:
                vals[i, j] = float(v)
    return vals, mask


def parse_load(load_dict):
    steps = []
    for st in load_dict['loadstep']:
        mech = st['boundary_conditions']['mechanical']
        step = dict(r=1, estimate_rate=True, f_out=1, R=np.eye(3))
        for k in ('F', 'dot_F', 'L'):
            if k in mech:
                step['deformation'] = BoundaryCondition(*parse_bc(mech[k]), k)
        for k in ('P', 'dot_P'):
            if k in mech:
                step['stress'] = BoundaryCondition(*parse_bc(mech[k]), k)
        if 'stress' not in step:
            # upstream DAMASK_grid.f90: no stress entry = no component is stress controlled (mask all 'x'); the deformation must
            # then prescribe all nine.  (Config::parse_load, config.cpp:309-311, throws here -- on the reference's own example
            # decks examples/grid/shearXY.yaml and shearZX.yaml.)
            step['stress'] = BoundaryCondition(np.zeros((3, 3)), np.ones((3, 3), bool), 'P')
        if 'R' in mech:
            from .crystal import qu2om
            step['R'] = qu2om(mech['R'])
        step['t'] = float(st['discretization']['t']); step['N'] = int(st['discretization']['N'])
        step['r'] = float(st['discretization'].get('r', 1))
        step['estimate_rate'] = bool(st.get('estimate_rate', True))
        step['f_out'] = int(st.get('f_out', 1))
        steps.append(step)
    return steps


class MechSolverBasicOracle:
    def __init__(self, cells, size, material, material_at, grid_mech=None, num_phase_mech=None,
                 derivative=sp.DERIVATIVE_CONTINUOUS, divergence_correction=sp.DIVCORR_SIZE_GRID,
                 tangent_every_eval=True):
        self.grid = sp.Grid(cells, size)
        self.num = dict(GRID_MECH_DEFAULTS); self.num.update(grid_mech or {})
        self.con = ConstitutiveOracle(material, material_at, num_phase_mech)
        self.xi1st, self.xi2nd = sp.xi_tables(self.grid, derivative)
        self.scaled_size = sp.scaled_geom_size(self.grid, divergence_correction)
        nx, ny, nz = self.grid.cells
        self.shape = (nz, ny, nx, 3, 3)
        I = np.eye(3)
        self.F = np.broadcast_to(I, self.shape).copy()
        self.F_last_inc = self.F.copy()
        self.F_dot = np.zeros(self.shape)
        self.F_aim = I.copy(); self.F_aim_last_inc = I.copy(); self.F_aim_dot = np.zeros((3, 3))
        self.P_av = np.zeros((3, 3)); self.P_aim = np.zeros((3, 3))
        self.S = np.zeros((3, 3, 3, 3))
        self.C_volAvg = self.C_minMaxAvg = None
        self.C_volAvgLastInc = self.C_minMaxAvgLastInc = None
        self.params = dict(stress_mask=np.ones((3, 3), bool), R=I.copy(), dt=0.0)
        self.err_div = self.err_BC = 0.0
        self.total_iter = 0
        self.tangent_every_eval = tangent_every_eval
        self._tangent_dt = None
        self.log_fields = True        # keep F and r of every evaluation in the log (tests); bench.py's CPU arm switches it off
        self.log = []                 # one dict per residual evaluation
        self.reports = []             # one dict per convergence check
        self.minmax_cells = None
        self.P = None

    # ------------------------------------------------------------------ Spectral::constitutive_response
    def constitutive_response(self, F, dt):
        ill = self.con.mechanical_response(F, dt)
        if not self.tangent_every_eval:
            # what the device path does (and what bench.py's CPU arm times): dPdF and its reductions are consumed once per
            # increment (mech_solver_basic.cpp:116, update_gamma = false), so they are formed when asked for, from the state
            # the LAST evaluation left -- same values, 1/n_evaluations of the work
            P, _ = self.con.mechanical_response2(dt, tangent=False)
            self.P_av = sp.rotate2(self.params['R'], P.sum(axis=0) * self.grid.wgt)
            self._tangent_dt = dt
            self.P = P.reshape(self.shape)
            return self.P
        P, dPdF = self.con.mechanical_response2(dt, tangent=True)
        self.P_av, self.C_volAvg, self.C_minMaxAvg, self.minmax_cells = sp.homogenise(P, dPdF, self.grid)
        self.P_av = sp.rotate2(self.params['R'], self.P_av)
        self.P = P.reshape(self.shape)
        return self.P

    def ensure_tangent(self):
        if self.tangent_every_eval or self._tangent_dt is None:
            return
        P, dPdF = self.con.mechanical_response2(self._tangent_dt, tangent=True)
        _, self.C_volAvg, self.C_minMaxAvg, self.minmax_cells = sp.homogenise(P, dPdF, self.grid)
        self._tangent_dt = None

    def init(self, gamma_pair_cells=None):
        self.constitutive_response(self.F, 0.0)
        self.ensure_tangent()
        if gamma_pair_cells is not None:            # replay aid: force the two cells that define C_ref (V7)
            _, dPdF = self.con.mechanical_response2(0.0, tangent=True)
            a, b = gamma_pair_cells
            self.C_minMaxAvg = 0.5 * (dPdF[a] + dPdF[b])
        self.update_gamma(self.C_minMaxAvg)

    def update_gamma(self, C):
        self.C_ref = C / self.grid.wgt

    # ------------------------------------------------------------------ form_residual
    def form_residual(self, X, nfuncs, petsc_iter):
        if nfuncs == 0 and petsc_iter == 0:
            self.total_iter = -1
        if self.total_iter <= petsc_iter:
            self.total_iter += 1
        F = X.reshape(self.shape)
        R = self.params['R']
        P = self.constitutive_response(F, self.params['dt'])
        hat = sp.fft_forward(P)
        self.err_div, raw = sp.divergence_rms(hat, self.xi1st, self.grid, self.scaled_size)
        delta_F_aim = np.einsum('ijkl,kl->ij', self.S, self.P_av - self.P_aim)      # upstream (D3)
        self.F_aim = self.F_aim - delta_F_aim
        mask = self.params['stress_mask']
        self.err_BC = np.abs(np.where(mask, 0.0, self.P_av - self.P_aim)).max()
        out = sp.gamma_hat_apply(hat, self.xi1st, self.C_ref)
        out[0, 0, 0] = sp.rotate2(R, delta_F_aim, active=True)
        r = sp.fft_backward(out, self.grid.cells, 1.0)
        keep = self.log_fields
        self.log.append(dict(F=F.copy() if keep else None, r=r.copy() if keep else None, P_av=self.P_av.copy(), err_div=self.err_div, err_BC=self.err_BC,
                             div_raw=raw, total_iter=self.total_iter, counters=self.con.counters.copy(),
                             ill=self.con.terminally_ill))
        return r.reshape(-1)

    # ------------------------------------------------------------------ converged
    def converged(self, petsc_iter):
        n = self.num
        maxP = self.P_av.max()
        divTol = max(maxP * n['eps_rel_div'], n['eps_abs_div'])
        BCTol = max(maxP * n['eps_rel_P'], n['eps_abs_P'])
        if (self.total_iter >= n['itmin'] and max(self.err_div / divTol, self.err_BC / BCTol) < 1.0) \
                or self.con.terminally_ill:
            reason = SNES_BREAKOUT_INNER_ITER
        elif self.total_iter >= n['itmax']:
            reason = SNES_DIVERGED_MAX_IT
        else:
            reason = 0
        self.reports.append(dict(total_iter=self.total_iter, err_div=self.err_div, divTol=divTol,
                                 err_BC=self.err_BC, BCTol=BCTol, reason=reason))
        return reason

    # ------------------------------------------------------------------ calculate_solution
    def calculate_solution(self):
        self.ensure_tangent()
        self.S = sp.masked_compliance(self.C_volAvg, self.params['R'], self.params['stress_mask'])
        if self.num['update_gamma']:
            self.update_gamma(self.C_minMaxAvg)
        snes = NGMRES()
        X, reason, its = snes.solve(self.F.reshape(-1).copy(), self.form_residual, self.converged)
        self.F = X.reshape(self.shape).copy()
        self.snes_trace = snes.trace
        sol = dict(converged=reason > 0, required_iterations=self.total_iter, terminally_ill=self.con.terminally_ill,
                   stag_converged=True)
        self.con.terminally_ill = False
        self.P_aim = np.where(self.params['stress_mask'], self.P_av, self.P_aim)
        return sol

    # ------------------------------------------------------------------ forward
    def forward(self, cut_back, guess, dt, dt_old, t_remaining, deformation_bc, stress_bc, R):
        self.ensure_tangent()
        if cut_back:
            self.C_volAvg = self.C_volAvgLastInc; self.C_minMaxAvg = self.C_minMaxAvgLastInc
        else:
            self.C_volAvgLastInc = self.C_volAvg; self.C_minMaxAvgLastInc = self.C_minMaxAvg
            if guess:
                self.F_aim_dot = np.where(stress_bc.mask, 0.0, (self.F_aim - self.F_aim_last_inc) / dt_old)
            else:
                self.F_aim_dot = np.zeros((3, 3))
            self.F_aim_last_inc = self.F_aim.copy()
            if deformation_bc.type == 'L':                 # upstream: L . F_aim_lastInc
                step = deformation_bc.values @ self.F_aim_last_inc
            elif deformation_bc.type == 'dot_F':
                step = deformation_bc.values
            elif deformation_bc.type == 'F':
                step = (deformation_bc.values - self.F_aim_last_inc) / t_remaining
            else:
                raise ValueError(deformation_bc.type)
            self.F_aim_dot = self.F_aim_dot + np.where(deformation_bc.mask, 0.0, step)
            self.F_dot = sp.calculate_rate(guess, self.F_last_inc, self.F, dt_old,
                                           sp.rotate2(R, self.F_aim_dot, active=True))
            self.F_last_inc = self.F.copy()
        self.F_aim = self.F_aim_last_inc + self.F_aim_dot * dt
        if stress_bc.type == 'P':
            self.P_aim = self.P_aim + np.where(stress_bc.mask, 0.0, (stress_bc.values - self.P_aim) / t_remaining) * dt
        else:
            self.P_aim = self.P_aim + np.where(stress_bc.mask, 0.0, stress_bc.values) * dt
        self.F = sp.forward_field(dt, self.F_last_inc, self.F_dot, self.grid, sp.rotate2(R, self.F_aim, active=True))
        self.params = dict(stress_mask=stress_bc.mask.copy(), R=np.asarray(R, float).copy(), dt=dt)


def run_loadcase(solver, steps, max_cut_back=3, callback=None):
    """DamaskGrid::loop_over_loadcase (damask_grid.cpp:82-196). Returns a list of per-increment records."""
    records = []
    cut_back = False
    sub_step_factor, cut_back_level, total_inc = 2, 0, 0
    t = 0.0; dt = 1.0; dt_prev = 0.0
    for l, st in enumerate(steps):
        t0 = t
        guess = st['estimate_rate'] and l > 0
        for inc in range(st['N']):
            total_inc += 1
            dt_prev = dt
            if abs(st['r'] - 1.0) < 1e-9:
                dt = st['t'] / st['N']
            else:
                dt = st['t'] * (st['r'] ** inc - st['r'] ** (inc + 1)) / (1.0 - st['r'] ** st['N'])
            dt = dt * float(sub_step_factor) ** (-cut_back_level)
            step_fraction = 0
            while step_fraction < sub_step_factor ** cut_back_level:
                t_remaining = st['t'] + t0 - t
                t += dt
                step_fraction += 1
                solver.forward(cut_back, guess, dt, dt_prev, t_remaining, st['deformation'], st['stress'], st['R'])
                if not cut_back:
                    solver.con.forward()                                     # f_materialpoint_forward
                sol = solver.calculate_solution()
                found = sol['converged'] and sol['stag_converged'] and not sol['terminally_ill']
                if found:
                    dt_prev = dt; cut_back = False; guess = True
                elif cut_back_level < max_cut_back:
                    cut_back = True
                    step_fraction = (step_fraction - 1) * sub_step_factor
                    cut_back_level += 1
                    t -= dt
                    dt = dt / sub_step_factor
                else:
                    raise RuntimeError('no solution found for increment %d' % total_inc)
            cut_back_level = max(0, cut_back_level - 1)                       # upstream (D10)
            rec = dict(inc=total_inc, t=t, converged=sol['converged'], iterations=sol['required_iterations'],
                       P_av=solver.P_av.copy(), F_av=solver.F.mean(axis=(0, 1, 2)), F_aim=solver.F_aim.copy())
            records.append(rec)
            if callback:
                callback(rec, solver)
    return records
