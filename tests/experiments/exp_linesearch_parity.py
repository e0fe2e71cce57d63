"""CPU experiment: the register-resident formulation WITH its own line search (host twin, defer off) against the oracle
on bench.py's kind of state.  Which cells deviate by more than 1e-9, and do they coincide with differing iteration counts?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'host_twin'))
import numpy as np
import bench, twin
from solver_cpp_b200 import workloads as W, material as M
from oracle import spectral as osp
from oracle.solver import MechSolverBasicOracle, parse_load

V6 = [0, 4, 8, 5, 2, 1]
n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n_relax = int(sys.argv[2]) if len(sys.argv) > 2 else 4
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cells = (n1, n1, n1)
grains = max(8, int(round(500 * np.prod(cells) / 256.0 ** 3)))
material, material_at, size = W.polycrystal(cells, grains, 'Cu_pheno', 0)
o = MechSolverBasicOracle(cells, size, material, material_at)
o.init()
nx, ny, nz = cells
DT = bench.DT
for inc in (1, 2):
    F = np.broadcast_to(bench.taylor_F(inc), (nz, ny, nx, 3, 3)).copy()
    o.con.mechanical_response(F.reshape(-1, 9), DT)
    o.con.forward()
st = parse_load(W.TENSION_X)[0]
o.F = np.broadcast_to(bench.taylor_F(2), (nz, ny, nx, 3, 3)).copy()
o.F_last_inc = np.broadcast_to(bench.taylor_F(1), (nz, ny, nx, 3, 3)).copy()
o.F_aim = bench.taylor_F(2); o.F_aim_last_inc = bench.taylor_F(1)
o.forward(False, True, DT, DT, 5.0 - 2 * DT, st['deformation'], st['stress'], st['R'])
o.S = osp.masked_compliance(o.C_volAvg, o.params['R'], o.params['stress_mask'])
X = o.F.reshape(-1).copy()
lib = twin.build()
m = M.Material(material, material_at)
num = M.make_numerics()
con = o.con
n = con.n
ns = con.phases[0].n_state
lib.twin_use_fast(mode)
if len(sys.argv) > 4:
    lib.twin_set_defer(int(sys.argv[4]))
for it in range(n_relax + 1):
    F = X.reshape(o.shape)
    snap = con.snapshot()
    con.mechanical_response(F, DT)
    P, _ = con.mechanical_response2(DT, tangent=False)
    Fc = F.reshape(n, 9)
    errP = np.zeros(n); errA = np.zeros(n); dcnt = np.zeros(n, int); deferred = np.zeros(n, bool)
    for i in range(n):
        r = twin.run_point(lib, m.phases[int(con.phase_of[i])], num, snap['F0'][i], Fc[i], snap['Fp0'][i], snap['Lp0'][i], snap['Lp'][i],
                           snap['S0'][i][V6], snap['state0'][i][:ns], np.zeros(1), DT)
        deferred[i] = lib.twin_last_deferred()
        errP[i] = np.abs(r['P'] - con.a['P'][i]).max() / np.abs(con.a['P'][i]).max()
        errA[i] = max(np.abs(r['Fp'] - con.a['Fp'][i]).max(),
                      (np.abs(r['state'][:ns] - con.a['state'][i, :ns]) / np.maximum(np.abs(con.a['state'][i, :ns]), 1.0)).max())
        dcnt[i] = int(r['cnt'][0]) - int(con.point_lp_iter[i])
    bad = errP > 1e-9
    print('iter %d: Lp/cell %.2f cutback cells %d deferred %d | P err>1e-9: %d (max %.2e)  aux>1e-8: %d (max %.2e) | count mismatch %d | bad&mismatch %d bad&match %d | max err among matching %.2e'
          % (it, con.point_lp_iter.mean(), (con.point_cutbacks > 0).sum(), deferred.sum(), bad.sum(), errP.max(), (errA > 1e-8).sum(), errA.max(),
             (dcnt != 0).sum(), (bad & (dcnt != 0)).sum(), (bad & (dcnt == 0)).sum(), errP[dcnt == 0].max()), flush=True)
    P = P.reshape(o.shape)
    P_av = P.sum(axis=(0, 1, 2)) * o.grid.wgt
    hat = osp.fft_forward(P)
    dF = np.einsum('ijkl,kl->ij', o.S, P_av - o.P_aim)
    out = osp.gamma_hat_apply(hat, o.xi1st, o.C_ref)
    out[0, 0, 0] = dF
    X = X - osp.fft_backward(out, o.grid.cells, 1.0).reshape(-1)
