"""ORACLE (test infrastructure only).

Behavioural restatement of PETSc 3.18 SNESNGMRES with its defaults and no nonlinear preconditioner, as driven by
/root/reference/src/grid/mech_solver_basic.cpp:29-64,110-161 (`-mechanical_snes_type ngmres`, custom
convergence test, residual callback with side effects).  PETSc is NOT in /root/reference (third-party, pinned
version 3.18.6 by output_damask_f.sh:128); this restates the published algorithm (Oosterlee & Washio 2000;
PETSc manual "SNESNGMRES") and is pinned behaviourally by the 108 logged iterates of the golden run (G1):
every evaluation point X^M / X^A must be reproduced.

Defaults: m = 30, restart_type = difference, select_type = difference, gammaA = 2, gammaC = 2, epsilonB = 0.1,
deltaB = 0.9, restart_it = 2, no additive line search, candidate = false, basic line search with lambda = 1
(X^M = X - F(X)).  Per iteration: evaluate F(X^M), small least squares over the stored history, evaluate F(X^A),
select one of them, possibly restart, append to history, convergence test.
"""
import numpy as np

CONVERGED_ITERATING = 0


class NGMRES:
    def __init__(self, msize=30, gammaA=2.0, gammaC=2.0, epsilonB=0.1, deltaB=0.9, restart_it=2, max_its=10000):
        self.msize, self.gammaA, self.gammaC = msize, gammaA, gammaC
        self.epsilonB, self.deltaB, self.restart_it, self.max_its = epsilonB, deltaB, restart_it, max_its
        self.trace = []

    def solve(self, X, residual, converged):
        """X: flat vector (modified copy returned); residual(X, n_funcs, it) -> F; converged(it) -> reason (0 = go on).
        Returns (X, reason, iterations)."""
        m = self.msize
        Xdot = [None] * m; Fdot = [None] * m
        Q = np.zeros((m, m))
        nfuncs = 0
        it = 0
        F = residual(X, nfuncs, it); nfuncs += 1
        fnorm = np.linalg.norm(F)
        fminnorm = fnorm
        reason = converged(0)
        if reason:
            return X, reason, 0
        Xdot[0] = X.copy(); Fdot[0] = F.copy()
        k_restart, l, ivec, restart_count = 1, 1, 0, 0
        for k in range(1, self.max_its + 1):
            XM = X - F                                              # basic line search, lambda = 1
            FM = residual(XM, nfuncs, it); nfuncs += 1
            fMnorm = np.linalg.norm(FM)
            # ---- combined solution
            nu = fMnorm * fMnorm
            xi = np.array([FM @ Fdot[i] for i in range(l)])
            beta = np.array([Fdot[ivec] @ Fdot[i] for i in range(l)])
            Q[:l, ivec] = beta; Q[ivec, :l] = beta
            rhs = nu - xi
            H = Q[:l, :l] - xi[:, None] - xi[None, :] + nu
            if l == 1:
                alpha = np.array([rhs[0] / H[0, 0] if H[0, 0] != 0.0 else 0.0])
            else:
                alpha = np.linalg.lstsq(H, rhs, rcond=-1)[0]      # LAPACK gelss with rcond = -1 in PETSc
            XA = XM * (1.0 - alpha.sum())
            for i in range(l):
                XA = XA + alpha[i] * Xdot[i]
            FA = residual(XA, nfuncs, it); nfuncs += 1
            if fminnorm > fMnorm:
                fminnorm = fMnorm
            # ---- norms for selection / restart
            fAnorm = np.linalg.norm(FA)
            dnorm = np.linalg.norm(XA - XM)
            dminnorm = min(np.linalg.norm(Xdot[i] - XA) for i in range(l))
            # ---- selection (difference)
            selectA = True
            if fAnorm >= self.gammaA * fminnorm:
                selectA = False
            if not (self.epsilonB * dnorm < dminnorm or np.sqrt(fnorm) < self.deltaB * np.sqrt(fminnorm)):
                selectA = False
            if selectA:
                X, F, fnorm = XA, FA, fAnorm
            else:
                X, F, fnorm = XM, FM, fMnorm
            # ---- restart test (difference)
            selectRestart = False
            if self.epsilonB * dnorm > dminnorm and np.sqrt(fAnorm) > self.deltaB * np.sqrt(fminnorm) and l > 0:
                selectRestart = True
            if np.sqrt(fAnorm) > self.gammaC * np.sqrt(fminnorm):
                selectRestart = True
            restart_count = restart_count + 1 if selectRestart else 0
            ivec = k_restart % m
            self.trace.append(dict(k=k, l=l, alpha=alpha.copy(), selectA=selectA, restart=selectRestart,
                                   fMnorm=fMnorm, fAnorm=fAnorm))
            if restart_count >= self.restart_it:
                restart_count, k_restart, l, ivec = 0, 1, 1, 0
                Xdot[0] = X.copy(); Fdot[0] = F.copy()
            else:
                if l < m:
                    l += 1
                k_restart += 1
                if fminnorm > fnorm:
                    fminnorm = fnorm
                Xdot[ivec] = X.copy(); Fdot[ivec] = F.copy()
            it = k
            reason = converged(it)
            if reason:
                return X, reason, it
        return X, -5, it                                            # SNES_DIVERGED_MAX_IT
