"""Multi-GPU parity script (run under torchrun, one rank per GPU): the z-slab decomposed solver against the single-GPU
solver on the same problem.  Used by tests/test_gpu_multi.py and by hand:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_parity.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    world, rank, local = int(os.environ['WORLD_SIZE']), int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
    # a hang (a rank waiting in a collective the others never enter) must cost a bounded time: after DMK_PARITY_TIMEOUT seconds
    # every rank dumps the stack of all its threads to stderr and exits
    import faulthandler
    faulthandler.dump_traceback_later(float(os.environ.get('DMK_PARITY_TIMEOUT', '240')), exit=True)
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from solver_cpp_b200 import mech, workloads as W, parallel
    from solver_cpp_b200.grid import GridContext
    ok = True
    # ---- spectral half: forward FFT + Gamma + inverse on a random field, slab vs whole
    cells, size = (16, 16 * world, 8 * world if os.environ.get('DMK_PARITY_SMALL_Z') else 16 * world), (2e-5, 1e-5, 1.5e-5)
    nx, ny, nz = cells
    rng = np.random.default_rng(1)
    P = 1e8 * rng.standard_normal((nz, ny, nx, 3, 3))
    aim = 1e-4 * rng.standard_normal((3, 3))
    C = rng.standard_normal((3, 3, 3, 3)) * 1e8
    for i in range(3):
        for j in range(3):
            C[i, j, i, j] += 1e11
            C[i, j, j, i] += 2e10
    z0, z1 = parallel.slab_range(nz, world, rank)
    for deriv in (0, 2):
        g = GridContext(cells, size, derivative=deriv, world_size=world, rank=rank)
        g.set_comm(parallel.make_comm(world, rank))
        fused = parallel.fused_exchange_wanted() and g.enable_fused_exchange()
        g.update_gamma(C)
        d = g.to_device(P[z0:z1])
        rms = g.divergence_and_gamma(d, aim)
        out = g.to_natural(d)
        g1 = GridContext(cells, size, derivative=deriv)
        g1.update_gamma(C)
        d1 = g1.to_device(P)
        rms1 = g1.divergence_and_gamma(d1, aim)
        ref = g1.to_natural(d1)[z0:z1]
        e = np.abs(out - ref).max() / np.abs(ref).max()
        ok &= e < 1e-12 and abs(rms - rms1) <= 1e-12 * rms1
        if rank == 0:
            print('spectral deriv %d (fused exchange over peer memory: %s): rel err %.2e, rms %.15e vs %.15e' % (deriv, fused, e, rms, rms1))
    # ---- full solver: 2 increments of tensionX on a polycrystal
    cells = (16, 16, 8 * world)
    nz = cells[2]
    z0, z1 = parallel.slab_range(nz, world, rank)
    material, mat_slab, size = W.polycrystal(cells, 12, 'Al_pheno', 0, z_range=(z0, z1))
    _, mat_all, _ = W.polycrystal(cells, 12, 'Al_pheno', 0)
    s = mech.make_solver(cells, size, material, mat_slab, world_size=world, rank=rank, comm=parallel.make_comm(world, rank))
    s.init()
    if rank == 0:
        cm = s.g._comm_keepalive
        print('collectives: %s%s' % (type(cm).__name__, ' (library-native ncclAllReduce / ncclSend / ncclRecv)' if getattr(cm, 'native', False) else ''))
    recs = mech.loop_over_loadcase(s, mech.parse_load(W.TENSION_X), max_increments=2)
    s1 = mech.make_solver(cells, size, material, mat_all)
    s1.init()
    recs1 = mech.loop_over_loadcase(s1, mech.parse_load(W.TENSION_X), max_increments=2)
    it, it1 = [r['iterations'] for r in recs], [r['iterations'] for r in recs1]
    eP = max(np.abs(a['P_av'] - b['P_av']).max() / np.abs(b['P_av']).max() for a, b in zip(recs, recs1))
    F = s.g.to_natural(s.field('F'))
    F1 = s1.g.to_natural(s1.field('F'))[z0:z1]
    eF = np.abs(F - F1).max()
    sel = tuple(int(v) for v in s.get('minmax_cells', 2)), tuple(int(v) for v in s1.get('minmax_cells', 2))
    ok &= it == it1 and eP < 1e-9 and eF < 1e-10 and sel[0] == sel[1]
    # ---- update_coords: slab (halo layers through the all-to-all callback) vs whole grid
    x_n, x_p = s.update_coords()
    x_n1, x_p1 = s1.update_coords()
    eX = max(np.abs(x_p - x_p1[z0:z1]).max(), np.abs(x_n - x_n1[z0:z1 + 1]).max()) / max(size)
    ok &= eX < 1e-12
    # ---- result output of a slab job: rank 0 gathers the slabs over NCCL (solver_cpp_b200.dadf5.gather_slabs)
    from solver_cpp_b200 import dadf5
    whole = dadf5.gather_slabs({'F': F.reshape(-1, 9), 'u_n': x_n, 'u_p': x_p.reshape(-1, 3)}, None, s.g.device)
    if rank == 0:
        F_all = s1.g.to_natural(s1.field('F')).reshape(-1, 9)
        eG = max(np.abs(whole['F'] - F_all).max(), np.abs(whole['u_n'] - x_n1).max() / max(size), np.abs(whole['u_p'] - x_p1.reshape(-1, 3)).max() / max(size))
        ok &= whole['F'].shape == F_all.shape and whole['u_n'].shape == x_n1.shape and eG < 1e-10
        print('gather_slabs over NCCL: shapes %s %s, max err %.2e' % (whole['F'].shape, whole['u_n'].shape, eG))
    else:
        ok &= whole is None
    if rank == 0:
        print('solver: iterations %s vs %s, P_av rel %.2e, F abs %.2e, minmax cells %s vs %s' % (it, it1, eP, eF, sel[0], sel[1]))
        print('update_coords: rel err %.2e (solver ran with the fused exchange: %s)' % (eX, getattr(s.g, 'fused_exchange', False)))
    t = torch.tensor([1.0 if ok else 0.0], device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print('MGPU_PARITY_OK' if t.item() == 1.0 else 'MGPU_PARITY_FAILED')
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == '__main__':
    main()
