"""ORACLE (test infrastructure only -- never imported by the product path).  PARITY UNPINNED.

The polarisation variant of the mechanical spectral solver (Eisenlohr et al. 2013, Shanthraj et al. 2015), which
BASELINE.json's north star names next to the basic scheme.  It does NOT exist in /root/reference: damask_grid.cpp:8-12
accepts only `spectral_basic`; what the tree has are its numerics keys -- alpha, beta, eps_abs_curl(F), eps_rel_curl(F)
(src/grid/config.cpp:450-457, defaults config.h:68-71: alpha = beta = 1, 1e-10, 5e-4).  This file therefore restates UPSTREAM
DAMASK's grid_mechanical_spectral_polarisation (v3.0.0-alpha7: init, solution, forward, converged, formResidual) on top of the
pinned pieces of the basic oracle (constitutive update, FFT, Gamma convolution, divergence RMS, masked compliance, NGMRES).
There is no fixture for it anywhere in the reference: device-vs-oracle agreement is self-consistency, not pinned parity.

Solution vector: (F, F_tau), both [z,y,x,3,3]; residual (r_F, r_F_tau).  One evaluation:
    T        = beta C_scale:(F - I) - alpha F.(C_scale:(F_tau - F - I))
    T       <- Gamma * T, DC term rot(beta F_aim)
    r_F_tau  = beta F - T
    P        = constitutive response at F - r_F_tau / beta ;  F_aim -= S:(P_av - P_aim)
    err_BC   = max | mask ? C_scale:(F_aim - rot(F_av)) : P_av - P_aim |      (mask: component not stress controlled)
    err_div  = divergence RMS of P ;  err_curl = curl RMS of F
    r_F      = (dPdF + C_scale)^-1 : (P - F.(C_scale:(F_tau - F - I))) + r_F_tau
"""
import numpy as np

from . import spectral as sp
from .ngmres import NGMRES
from .solver import MechSolverBasicOracle, SNES_BREAKOUT_INNER_ITER, SNES_DIVERGED_MAX_IT

I3 = np.eye(3)


class MechSolverPolarisationOracle(MechSolverBasicOracle):
    def __init__(self, *a, alpha=1.0, beta=1.0, eps_abs_curl=1.0e-10, eps_rel_curl=5.0e-4, **k):
        k['tangent_every_eval'] = True            # the residual needs dPdF of every cell at every evaluation
        super().__init__(*a, **k)
        self.alpha, self.beta, self.eps_abs_curl, self.eps_rel_curl = alpha, beta, eps_abs_curl, eps_rel_curl
        self.F_tau = 2.0 * self.F
        self.F_tau_last_inc = 2.0 * self.F_last_inc
        self.F_tau_dot = np.zeros(self.shape)
        self.F_av = I3.copy()
        self.err_curl = 0.0
        self.C_scale = self.S_scale = None
        self.dPdF = None

    def constitutive_response(self, F, dt):
        self.con.mechanical_response(F, dt)
        P, dPdF = self.con.mechanical_response2(dt, tangent=True)
        self.P_av, self.C_volAvg, self.C_minMaxAvg, self.minmax_cells = sp.homogenise(P, dPdF, self.grid)
        self.P_av = sp.rotate2(self.params['R'], self.P_av)
        self.dPdF = dPdF.reshape(self.shape[:3] + (3, 3, 3, 3))
        self.P = P.reshape(self.shape)
        return self.P

    def init(self, gamma_pair_cells=None):
        super().init(gamma_pair_cells)
        self.C_scale = self.C_minMaxAvg.copy()
        self.S_scale = sp.invsym3333(self.C_minMaxAvg)

    # ------------------------------------------------------------------ formResidual
    def form_residual(self, X, nfuncs, petsc_iter):
        if nfuncs == 0 and petsc_iter == 0:
            self.total_iter = -1
        if self.total_iter <= petsc_iter:
            self.total_iter += 1
        n9 = int(np.prod(self.shape))
        F, F_tau = X[:n9].reshape(self.shape), X[n9:].reshape(self.shape)
        R, Cs, a, b = self.params['R'], self.C_scale, self.alpha, self.beta
        self.F_av = F.sum(axis=(0, 1, 2)) * self.grid.wgt
        lam = np.einsum('ijkl,...kl->...ij', Cs, F_tau - F - I3)                     # C_scale:(F_tau - F - I)
        T = b * np.einsum('ijkl,...kl->...ij', Cs, F - I3) - a * np.einsum('...ij,...jk->...ik', F, lam)
        T = sp.gamma_convolution(T, sp.rotate2(R, b * self.F_aim, active=True), self.xi1st, self.C_ref, self.grid)
        r_F_tau = b * F - T
        P = self.constitutive_response(F - r_F_tau / b, self.params['dt'])
        self.F_aim = self.F_aim - np.einsum('ijkl,kl->ij', self.S, self.P_av - self.P_aim)
        mask = self.params['stress_mask']
        self.err_BC = np.abs(np.where(mask, np.einsum('ijkl,kl->ij', Cs, self.F_aim - sp.rotate2(R, self.F_av)),
                                      self.P_av - self.P_aim)).max()
        self.err_div, _ = sp.divergence_rms(sp.fft_forward(P), self.xi1st, self.grid, self.scaled_size)
        Ainv = sp.invsym3333(self.dPdF + Cs)
        r_F = np.einsum('...ijkl,...kl->...ij', Ainv, P - np.einsum('...ij,...jk->...ik', F, lam)) + r_F_tau
        self.err_curl, _ = sp.curl_rms(sp.fft_forward(F), self.xi1st, self.grid, self.scaled_size)
        self.log.append(dict(P_av=self.P_av.copy(), err_div=self.err_div, err_BC=self.err_BC, err_curl=self.err_curl,
                             total_iter=self.total_iter, F=None, r=None))
        return np.concatenate([r_F.reshape(-1), r_F_tau.reshape(-1)])

    # ------------------------------------------------------------------ converged
    def converged(self, petsc_iter):
        n = self.num
        curlTol = max(np.abs(self.F_aim - I3).max() * self.eps_rel_curl, self.eps_abs_curl)
        divTol = max(np.abs(self.P_av).max() * n['eps_rel_div'], n['eps_abs_div'])
        BCTol = max(np.abs(self.P_av).max() * n['eps_rel_P'], n['eps_abs_P'])
        ok = max(self.err_div / divTol, self.err_curl / curlTol, self.err_BC / BCTol) < 1.0
        if (self.total_iter >= n['itmin'] and ok) or self.con.terminally_ill:
            reason = SNES_BREAKOUT_INNER_ITER
        elif self.total_iter >= n['itmax']:
            reason = SNES_DIVERGED_MAX_IT
        else:
            reason = 0
        self.reports.append(dict(total_iter=self.total_iter, err_div=self.err_div, divTol=divTol, err_curl=self.err_curl,
                                 curlTol=curlTol, err_BC=self.err_BC, BCTol=BCTol, reason=reason))
        return reason

    # ------------------------------------------------------------------ solution
    def calculate_solution(self):
        self.S = sp.masked_compliance(self.C_volAvg, self.params['R'], self.params['stress_mask'])
        if self.num['update_gamma']:
            self.update_gamma(self.C_minMaxAvg)
            self.C_scale = self.C_minMaxAvg.copy()
            self.S_scale = sp.invsym3333(self.C_minMaxAvg)
        snes = NGMRES()
        X0 = np.concatenate([self.F.reshape(-1), self.F_tau.reshape(-1)])
        X, reason, its = snes.solve(X0, self.form_residual, self.converged)
        n9 = int(np.prod(self.shape))
        self.F, self.F_tau = X[:n9].reshape(self.shape).copy(), X[n9:].reshape(self.shape).copy()
        sol = dict(converged=reason > 0, required_iterations=self.total_iter, terminally_ill=self.con.terminally_ill, stag_converged=True)
        self.con.terminally_ill = False
        self.P_aim = np.where(self.params['stress_mask'], self.P_av, self.P_aim)
        return sol

    # ------------------------------------------------------------------ forward
    def forward(self, cut_back, guess, dt, dt_old, t_remaining, deformation_bc, stress_bc, R):
        if cut_back:
            self.C_volAvg = self.C_volAvgLastInc; self.C_minMaxAvg = self.C_minMaxAvgLastInc
        else:
            self.C_volAvgLastInc = self.C_volAvg; self.C_minMaxAvgLastInc = self.C_minMaxAvg
            if guess:
                self.F_aim_dot = np.where(stress_bc.mask, 0.0, (self.F_aim - self.F_aim_last_inc) / dt_old)
            else:
                self.F_aim_dot = np.zeros((3, 3))
            self.F_aim_last_inc = self.F_aim.copy()
            if deformation_bc.type == 'L':
                step = deformation_bc.values @ self.F_aim_last_inc
            elif deformation_bc.type == 'dot_F':
                step = deformation_bc.values
            elif deformation_bc.type == 'F':
                step = (deformation_bc.values - self.F_aim_last_inc) / t_remaining
            else:
                raise ValueError(deformation_bc.type)
            self.F_aim_dot = self.F_aim_dot + np.where(deformation_bc.mask, 0.0, step)
            rot_rate = sp.rotate2(R, self.F_aim_dot, active=True)
            self.F_dot = sp.calculate_rate(guess, self.F_last_inc, self.F, dt_old, rot_rate)
            self.F_tau_dot = sp.calculate_rate(guess, self.F_tau_last_inc, self.F_tau, dt_old, rot_rate)
            self.F_last_inc = self.F.copy()
            self.F_tau_last_inc = self.F_tau.copy()
        self.F_aim = self.F_aim_last_inc + self.F_aim_dot * dt
        if stress_bc.type == 'P':
            self.P_aim = self.P_aim + np.where(stress_bc.mask, 0.0, (stress_bc.values - self.P_aim) / t_remaining) * dt
        else:
            self.P_aim = self.P_aim + np.where(stress_bc.mask, 0.0, stress_bc.values) * dt
        self.F = sp.forward_field(dt, self.F_last_inc, self.F_dot, self.grid, sp.rotate2(R, self.F_aim, active=True))
        if guess:
            self.F_tau = sp.forward_field(dt, self.F_tau_last_inc, self.F_tau_dot, self.grid)
        else:
            Fl = self.F_tau - self.F
            E2 = np.einsum('...ki,...kj->...ij', Fl, Fl) - I3                                   # F_lambda^T F_lambda - I
            Fl = I3 + np.einsum('ijkl,...kl->...ij', self.S_scale,
                                0.5 * np.einsum('...ij,...jk->...ik', Fl, np.einsum('ijkl,...kl->...ij', self.C_scale, E2)))
            self.F_tau = Fl + self.F
        self.params = dict(stress_mask=stress_bc.mask.copy(), R=np.asarray(R, float).copy(), dt=dt)
