#!/usr/bin/env python
"""bench.py -- one JSON line for the DAMASK grid hot path on B200 (contract: see the task statement).

Workload (config.workload): BASELINE.json configs[2] (C3), the 256^3 500-grain phenopowerlaw Cu polycrystal under tensionX,
at N=1; weak scaling with 256^3 cells per GPU: 256x256x512 on 2, 256x512x512 on 4 and 512^3 (configs[3]'s grid) on 8 GPUs,
z-slab decomposed, grain density kept at 500 per 256^3.  --config C1|C2|C4|C5 selects the other BASELINE.json workloads.

Protocol `solve` (default): virgin state (F = I), then the first K load increments SOLVED -- per increment
MechSolverBasic::forward, f_materialpoint_forward and calculate_solution, i.e. the consistent tangent and its reductions, the
masked compliance, and the NGMRES solve with every residual evaluation (per-point constitutive update, homogenisation, forward
FFT, fused divergence-RMS / Gamma pass, inverse FFT) and every NGMRES vector operation.  A step = one load increment.
Warm-up = W residual evaluations on the virgin state (all kernels), after which the per-point state and the solver are
re-initialised.  metric = material-point updates / s = cells x residual evaluations / wall time of the K increments.

Protocol `evaluation` (round 1's line, kept as a kernel-level detail): a step = ONE residual evaluation of a hand-built iterate.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config C3] [--protocol solve|evaluation] [--cells nx ny nz]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DT, RATE = 0.5, 1.0e-3
WEAK_CELLS = {1: (256, 256, 256), 2: (256, 256, 512), 4: (256, 512, 512), 8: (512, 512, 512)}
N_RELAX = 2                                 # protocol `evaluation`: fixed-point iterations before the timed evaluations


def default_cells(world, config='C3'):
    """C3/C4/C5: 256^3 cells per GPU; the grid grows z, then y, then x, so that every line length stays a power of two <= 512.
    C1/C2: the configuration's own grid (z-slabs need nz % world == 0)"""
    from solver_cpp_b200 import workloads as W
    if config in ('C1', 'C2'):
        return W.CONFIGS[config]['cells']
    return WEAK_CELLS.get(world, (256, 256, 256 * world))


def grains_for(config, cells):
    from solver_cpp_b200 import workloads as W
    c = W.CONFIGS[config]
    return max(8, int(round(c['grains'] * float(np.prod(cells)) / float(np.prod(c['cells'])))))


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def kernel_constants():
    """per-kernel figures taken from this round's `ncu --set full` captures (profiles/r02_kernel_constants.json): DFMA-equivalent
    flops per Lp Newton iteration of the shipped per-point kernel, DRAM traffic of the FFT/Gamma passes"""
    p = os.path.join(ROOT, 'profiles', 'r02_kernel_constants.json')
    if os.path.exists(p):
        return json.load(open(p))
    return {'flop_per_lp_iter': 4520.0, 'flop_source': 'round-1 ncu capture of an earlier kernel version (no round-2 capture found)',
            'fft_section_traffic_256': None, 'constitutive_traffic_256': None}


def taylor_F(inc):
    """homogeneous tensionX deformation gradient after `inc` increments (lateral contraction ~ incompressible flow)"""
    e = RATE * DT * inc
    return np.diag([1.0 + e, 1.0 - 0.5 * e, 1.0 - 0.5 * e])


def load_steps(parse_load, load, K):
    """the configuration's load case with enough increments of its own step size for K of them"""
    steps = parse_load(load)
    st = steps[0]
    dt = st['t'] / st['N']
    if K > st['N']:
        st['N'] = K
        st['t'] = dt * K
    return steps


class ClockSampler(threading.Thread):
    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = set()
        for r in self.rows:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(self.rows[0][1]), 'power_w_max': max(float(r[2]) for r in self.rows),
                'reasons': sorted(reasons), 'samples': len(self.rows)}


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port)
def cpu_sample_cells(K, budget_s=75.0):
    """edge of the cubic CPU sample so that K solved increments stay near budget_s (measured on the GPU box's 16 host threads:
    14.3 us per cell and evaluation -- 539 evaluations of the 20^3 sample in 61.6 s for the first three increments of C3, most
    of them in increment 1, ~45 per later one; profiles/r02_bench_n1.json)"""
    evals = {1: 380.0, 2: 470.0}.get(K, 540.0) + 45.0 * max(K - 3, 0)      # increment 1 of C3 is the long one (two attempts end at itmax)
    n = budget_s / (evals * 14.3e-6)
    edge = int(round(n ** (1.0 / 3.0) / 4.0)) * 4
    return (max(16, min(40, edge)),) * 3


def cpu_solve(config, cells, K, threads):
    """the oracle (CPU restatement of the reference algorithm; oracle/) on a bounded sample of the same workload, same
    protocol: virgin state, the first K increments solved (oracle NGMRES, every evaluation, tangent once per increment)."""
    os.environ['OMP_NUM_THREADS'] = str(threads)
    from solver_cpp_b200 import workloads as W
    from oracle.solver import MechSolverBasicOracle, parse_load, run_loadcase
    cfg = W.CONFIGS[config]
    material, material_at, size = W.polycrystal(cells, grains_for(config, cells), cfg['phase'], 0)
    o = MechSolverBasicOracle(cells, size, material, material_at, tangent_every_eval=False)
    o.log_fields = False
    o.init()
    steps = load_steps(parse_load, cfg['load'], K)
    recs = []

    class Done(Exception):
        pass

    def cb(rec, s):
        recs.append(dict(iterations=rec['iterations'], evaluations=len(o.log), P11=float(rec['P_av'][0, 0])))
        if len(recs) >= K:
            raise Done
    o.log = []
    t0 = time.perf_counter()
    try:
        run_loadcase(o, steps, callback=cb)
    except Done:
        pass
    wall = time.perf_counter() - t0
    evals = len(o.log)
    o.log = []
    n = float(np.prod(cells))
    # the per-point update alone, on the last iterate of the solve, with all threads and with one: the reference's own cell loop is
    # serial per MPI rank (homogenization.f90:233; upstream's OpenMP directive was dropped in the fork), SURVEY 8(d)
    per_point = {}
    try:
        snap = o.con.snapshot()
        for label, th in (('all_threads', 0), ('one_thread', 1)):
            o.con.set_threads(th)
            tc = time.perf_counter()
            o.con.mechanical_response(o.F.reshape(-1, 9), o.params['dt'])
            per_point[label] = n / (time.perf_counter() - tc)
            o.con.restore(snap)
        o.con.set_threads(0)
    except Exception as e:                                  # a reported extra, never a reason to lose the line
        per_point = {'error': str(e)[:100]}
    return dict(value=n * evals / wall, wall_s=wall, evaluations=evals, iterations=[r['iterations'] for r in recs],
                evaluations_cumulative=[r['evaluations'] for r in recs], P11=[r['P11'] for r in recs],
                per_point_update_cells_per_s=per_point)


def cpu_evaluations(cells, steps, warmup, threads):
    """protocol `evaluation` on the CPU: 2 Taylor increments, N_RELAX fixed-point iterations, then `steps` residual
    evaluations.  Returns (cells/s, seconds per evaluation)."""
    os.environ['OMP_NUM_THREADS'] = str(threads)
    from solver_cpp_b200 import workloads as W
    from oracle import spectral as osp
    from oracle.solver import MechSolverBasicOracle, parse_load
    grains = max(8, int(round(500 * np.prod(cells) / 256.0 ** 3)))
    material, material_at, size = W.polycrystal(cells, grains, 'Cu_pheno', 0)
    o = MechSolverBasicOracle(cells, size, material, material_at)
    o.init()
    nx, ny, nz = cells
    for inc in (1, 2):
        F = np.broadcast_to(taylor_F(inc), (nz, ny, nx, 3, 3)).copy()
        o.con.mechanical_response(F.reshape(-1, 9), DT)
        o.con.forward()
    st = parse_load(W.TENSION_X)[0]
    o.F = np.broadcast_to(taylor_F(2), (nz, ny, nx, 3, 3)).copy()
    o.F_last_inc = np.broadcast_to(taylor_F(1), (nz, ny, nx, 3, 3)).copy()
    o.F_aim = taylor_F(2); o.F_aim_last_inc = taylor_F(1)
    o.forward(False, True, DT, DT, 5.0 - 2 * DT, st['deformation'], st['stress'], st['R'])
    o.S = osp.masked_compliance(o.C_volAvg, o.params['R'], o.params['stress_mask'])
    X = o.F.reshape(-1).copy()
    split = {'constitutive_s': 0.0, 'spectral_s': 0.0}

    def one():
        F = X.reshape(o.shape)
        tc = time.perf_counter()
        o.con.mechanical_response(F, o.params['dt'])
        P, _ = o.con.mechanical_response2(o.params['dt'], tangent=False)
        split['constitutive_s'] += time.perf_counter() - tc
        ts = time.perf_counter()
        P = P.reshape(o.shape)
        P_av = P.sum(axis=(0, 1, 2)) * o.grid.wgt
        hat = osp.fft_forward(P)
        osp.divergence_rms(hat, o.xi1st, o.grid, o.scaled_size)
        dF = np.einsum('ijkl,kl->ij', o.S, P_av - o.P_aim)
        out = osp.gamma_hat_apply(hat, o.xi1st, o.C_ref)
        out[0, 0, 0] = dF
        r = osp.fft_backward(out, o.grid.cells, 1.0)
        split['spectral_s'] += time.perf_counter() - ts
        return r
    for _ in range(N_RELAX):
        X -= one().reshape(-1)
    for _ in range(warmup):
        one()
    split['constitutive_s'] = split['spectral_s'] = 0.0
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / steps
    cpu_evaluations.split = {k: v / steps for k, v in split.items()}
    return float(np.prod(cells)) / dt, dt


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    from solver_cpp_b200 import workloads as W
    cells = tuple(args.cells) if args.cells else default_cells(args.gpus, args.config)
    note = ('oracle/ = CPU restatement of the reference algorithm (the reference itself cannot be built here: no gfortran/'
            'MPI/FFTW/PETSc/HDF5); constitutive loop OpenMP over cells, FFT scipy pocketfft with all cores')
    if args.protocol == 'solve':
        K = max(1, args.steps)
        sample = cpu_sample_cells(K) if args.config not in ('C1', 'C2') else W.CONFIGS[args.config]['cells']
        r = cpu_solve(args.config, sample, K, threads)
        desc = '%dx%dx%d sub-problem of the same workload (same material, load and protocol), first %d increments solved' % (sample + (K,))
        line = {'impl': 'reference', 'metric': 'material-point updates/s', 'value': r['value'], 'unit': 'cells*evaluations/s',
                'n_gpus': args.gpus, 'steps': K, 'warmup': 0, 'ms_per_step': r['wall_s'] / K * 1e3, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                'config': {'workload': '%s: %s; CPU arm on a bounded sample' % (args.config, W.CONFIGS[args.config]['name']),
                           'cells': list(cells), 'sample_cells': list(sample), 'protocol': 'solve', 'step': 'one load increment solved',
                           'warmup_note': 'none: a solve from the virgin state is not repeatable without re-initialisation; the CPU arm has no clocks or caches to warm'},
                'ngmres_iterations': r['iterations'], 'evaluations': r['evaluations'],
                'newton_iters_per_s': sum(r['iterations']) / r['wall_s'],
                'cpu_baseline': {'value': r['value'], 'unit': 'cells*evaluations/s', 'cores': threads, 'kind': 'port', 'sample': desc,
                                 'per_point_update_cells_per_s': r['per_point_update_cells_per_s']},
                'e2e': {'value': r['value'], 'unit': 'cells*evaluations/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                'note': note}
    else:
        steps, warmup = min(max(1, args.steps), 5), max(1, min(args.warmup, 2))
        sample = (40, 40, 40)
        v, sec = cpu_evaluations(sample, steps, warmup, threads)
        line = {'impl': 'reference', 'metric': 'material-point updates/s', 'value': v, 'unit': 'cells*evaluations/s', 'n_gpus': args.gpus,
                'steps': steps, 'warmup': warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                'config': {'workload': 'C3: 256^3-per-GPU 500-grain phenopowerlaw Cu polycrystal, tensionX; CPU arm on a bounded sample',
                           'cells': list(cells), 'sample_cells': list(sample), 'protocol': 'evaluation'},
                'cpu_baseline': {'value': v, 'unit': 'cells*evaluations/s', 'cores': threads, 'kind': 'port',
                                 'sample': '%dx%dx%d sub-problem of the same workload (same material, load and protocol)' % sample,
                                 'seconds_per_evaluation': dict(cpu_evaluations.split)},
                'e2e': {'value': v, 'unit': 'cells*evaluations/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                'note': note}
    emit(line)


# ------------------------------------------------------------------------------------------------ GPU arm
def make_problem(config, cells, world, rank):
    from solver_cpp_b200 import workloads as W, mech, parallel
    cfg = W.CONFIGS[config]
    grains = grains_for(config, cells)
    z0, z1 = parallel.slab_range(cells[2], world, rank)
    material, material_at, size = W.polycrystal(cells, grains, cfg['phase'], 0, z_range=(z0, z1))
    comm = parallel.make_comm(world, rank) if world > 1 else None
    s = mech.make_solver(cells, size, material, material_at, world_size=world, rank=rank, comm=comm)
    return s, material, material_at, grains, cfg


def parity_check(world, rank):
    """N > 1 only: two residual evaluations of a small polycrystal on the z-slab decomposed solver against the same on this
    rank's GPU alone -- P_av, err_div, err_BC and this rank's slab of r must agree, so that every multi-GPU bench line carries
    a correctness check of the exchange it times.  Raises on disagreement.  The material is the Al phenopowerlaw of C1, not
    the workload's Cu: on virgin Cu every cell cuts back, and that regime amplifies the 1e-16 by which the homogenised stress of
    a slab sum and a whole-grid sum differ to 1e-5 in the next evaluation (the per-point kernel itself is reproducible bit
    for bit and independent of the decomposition: tests/experiments/exp_determinism.py)."""
    import torch
    from solver_cpp_b200 import workloads as W, mech, parallel
    cells = (32, 32, 8 * world)
    z0, z1 = parallel.slab_range(cells[2], world, rank)
    material, mat_slab, size = W.polycrystal(cells, 12, 'Al_pheno', 0, z_range=(z0, z1))
    _, mat_all, _ = W.polycrystal(cells, 12, 'Al_pheno', 0)
    st = mech.parse_load(W.TENSION_X)[0]
    out = []
    for slab in (True, False):
        s = mech.make_solver(cells, size, material, mat_slab if slab else mat_all, world_size=world if slab else 1, rank=rank if slab else 0,
                             comm=parallel.make_comm(world, rank) if slab else None)
        s.init()
        s.forward(False, False, DT, 1.0, 5.0, st['deformation'], st['stress'], st['R'])
        s.g.materialpoint_forward()
        s.prepare_evaluation()
        X = s.field('F')
        vals = []
        for _ in range(2):
            R = s.form_residual(X)
            vals.append((s.get33('P_av').copy(), float(s.get('err_div')[0]), float(s.get('err_BC')[0]), R.clone()))
            X = X - R
        out.append(vals)
        fused = bool(getattr(s.g, 'fused_exchange', False)) if slab else None
        if slab:
            fused_on = fused
        s.g.close()
    worst = dict(P_av_rel=0.0, err_div_rel=0.0, err_BC_rel=0.0, r_abs=0.0)
    for a, b in zip(*out):
        worst['P_av_rel'] = max(worst['P_av_rel'], float(np.abs(a[0] - b[0]).max() / np.abs(b[0]).max()))
        worst['err_div_rel'] = max(worst['err_div_rel'], abs(a[1] - b[1]) / abs(b[1]))
        worst['err_BC_rel'] = max(worst['err_BC_rel'], abs(a[2] - b[2]) / max(abs(b[2]), 1e-300))
        worst['r_abs'] = max(worst['r_abs'], float((a[3] - b[3][z0:z1]).abs().max().item()))
    ok = worst['P_av_rel'] < 1e-10 and worst['err_div_rel'] < 1e-10 and worst['err_BC_rel'] < 1e-8 and worst['r_abs'] < 1e-12
    worst.update(grid=list(cells), evaluations=2, fused_exchange=fused_on, ok=bool(ok))
    t = torch.tensor([1.0 if ok else 0.0], device='cuda')
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MIN)
    if t.item() != 1.0:
        raise RuntimeError('multi-GPU parity check failed on rank %d: %s' % (rank, worst))
    return worst


def run_solve(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device: the product path has no CPU fallback')
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from solver_cpp_b200 import mech, material as M, workloads as W
    cells = tuple(args.cells) if args.cells else default_cells(world, args.config)
    nx, ny, nz = cells
    pc = parity_check(world, rank) if world > 1 else None
    s, material, material_at, grains, cfg = make_problem(args.config, cells, world, rank)
    g = s.g
    n_local, n_global = g.n_local, nx * ny * nz
    K, W_ = max(1, args.steps), max(3, args.warmup)
    steps = load_steps(mech.parse_load, cfg['load'], K)
    st = steps[0]
    dt_inc = st['t'] / st['N']

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up: W residual evaluations on the virgin state (first iterate of increment 1 and its fixed-point successors), every
    # kernel of the path incl. the tangent; then the per-point state and the solver are re-initialised
    s.init()
    s.forward(False, False, dt_inc, 1.0, st['t'], st['deformation'], st['stress'], st['R'])
    g.materialpoint_forward()
    s.prepare_evaluation()
    X = s.field('F')
    R = g.new_field()
    for _ in range(W_):
        s.form_residual(X, R)
        X.sub_(R)
    g.tangent_response(dt_inc)
    del X, R
    g.material_init(M.Material(material, material_at), g.numerics)
    s.init()
    torch.cuda.empty_cache()

    # ---- timed: the first K increments solved
    g.set_profiling(True)
    g.get_profile(reset=True)
    c0, l0 = g.counters(), g.kernel_launches()
    n_log0 = len(s.eval_log())
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    incs = []
    t_last = [0.0]

    def cb(rec, solver):
        now = time.perf_counter()
        incs.append(dict(inc=rec['inc'], iterations=int(rec['iterations']), seconds=now - t_last[0], P11=float(rec['P_av'][0, 0]),
                         F11=float(rec['F_av'][0, 0])))
        t_last[0] = now
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); t_last[0] = t0
    e0.record()
    ts0 = s.get('solver_seconds', 12)
    mech.loop_over_loadcase(s, steps, max_increments=K, callback=cb)
    ts1 = s.get('solver_seconds', 12)
    e1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    ms = max(e0.elapsed_time(e1), 0.0)
    launches = g.kernel_launches() - l0
    prof = g.get_profile(reset=True)
    c1 = g.counters()
    g.set_profiling(False)
    log = s.eval_log()[n_log0:]
    stats = s.eval_stats()[-len(log):] if len(log) else np.zeros((0, 2))
    n_eval = len(log)
    t = torch.tensor([ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())

    # ---- end to end through the reference-facing call: the first iterates of increment K+1 evaluated through
    # dmk_mech_form_residual_host (pinned host F in, host r out; what a host-side SNES calling the library does per evaluation)
    t_rem = st['t'] - dt_inc * K
    s.forward(False, True, dt_inc, dt_inc, t_rem if t_rem > 0 else dt_inc, st['deformation'], st['stress'], st['R'])
    g.materialpoint_forward()
    s.prepare_evaluation()
    Xd = s.field('F')
    Rd = s.form_residual(Xd)
    hosts = []
    for k in range(2):                      # two consecutive iterates, alternated: the Lp guess always comes from the OTHER one
        h = torch.empty(Xd.shape, dtype=torch.float64).pin_memory()
        h.copy_((Xd if k == 0 else Xd - Rd).cpu())
        hosts.append(h.numpy())
    Ke = 6
    dev = [Xd, Xd - Rd]
    s.form_residual(dev[1], Rd); s.form_residual(dev[0], Rd)
    barrier()
    e0.record()
    for k in range(Ke):                     # the same evaluations with the fields resident in HBM (what the copies cost)
        s.form_residual(dev[(k + 1) % 2], Rd)
    e1.record()
    barrier()
    ms_same = e0.elapsed_time(e1)
    del Xd, Rd, dev
    rh = torch.empty(g.field_shape, dtype=torch.float64).pin_memory()
    rn = rh.numpy()
    s.form_residual_host(hosts[1], rn)
    s.form_residual_host(hosts[0], rn)
    barrier()
    te = time.perf_counter()
    e0.record()
    for k in range(Ke):
        s.form_residual_host(hosts[(k + 1) % 2], rn)
    e1.record()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), (time.perf_counter() - te) * 1e3)
    t = torch.tensor([ms_e2e], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t.item())
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    fp64_peak = g.fp64_peak_tflops()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm_peak, peak_src = peaks()
    kc = kernel_constants()
    ne = max(prof['evaluations'], 1)
    fft_ms = (prof['fft_forward_ms'] + prof['gamma_zpass_ms'] + prof['fft_inverse_ms']) / ne
    con_ms = prof['constitutive_ms'] / ne
    achieved_hbm = 720.0 * n_local / (fft_ms * 1e-3) / 1e9          # 5 passes x 144 B per cell, this GPU's slab
    lp_total = c1['Lp_iter'] - c0['Lp_iter']
    fp64_achieved = lp_total * kc['flop_per_lp_iter'] / (prof['constitutive_ms'] * 1e-3) / 1e12 if prof['constitutive_ms'] > 0 else None
    cut_share = float((stats[:, 0] > 0.01 * n_local).mean()) if len(stats) else 0.0
    # evaluations per increment from the log's iteration counter (total_iter restarts at 0 with every increment)
    starts = [i for i in range(n_eval) if i == 0 or log[i, 11] < log[i - 1, 11]]
    evals_per_inc = [(starts[j + 1] if j + 1 < len(starts) else n_eval) - starts[j] for j in range(len(starts))]
    iters = [r['iterations'] for r in incs]
    sec = ms * 1e-3
    phen = cfg['phase'] in ('Cu_pheno', 'Al_pheno')
    line = {
        'metric': 'material-point updates/s', 'value': n_global * n_eval / sec, 'unit': 'cells*evaluations/s', 'n_gpus': world,
        'steps': K, 'warmup': W_, 'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': '%s: %s (BASELINE.json); %d^3-per-GPU weak-scaling family for N > 1' % (args.config, cfg['name'], 256)
                   if args.config not in ('C1', 'C2') else '%s: %s (BASELINE.json)' % (args.config, cfg['name']),
                   'cells': list(cells), 'cells_per_gpu': n_local, 'grains': grains, 'protocol': 'solve',
                   'step': 'one load increment solved from the virgin state on: forward + materialpoint_forward + calculate_solution '
                           '(tangent, masked compliance, NGMRES with every residual evaluation and vector operation)',
                   'warmup_step': 'one residual evaluation on the virgin state; per-point state and solver re-initialised afterwards',
                   'l2': 'inputs larger than L2 (field %.2f GB per GPU, per-point state %.1f GB)' % (72e-9 * n_local, 1.0e-6 * n_local),
                   'decomposition': 'z-slabs, 1 per GPU'},
        'evaluations': n_eval, 'evaluations_per_increment': evals_per_inc, 'ngmres_iterations': iters,
        'newton_iters_per_s': sum(iters) / sec, 'seconds_per_increment': [round(r['seconds'], 4) for r in incs],
        'P11_MPa': [round(r['P11'] / 1e6, 4) for r in incs],
        'share_of_evaluations_with_over_1pct_cells_cut_back': cut_share,
        'cut_back_cell_fraction_per_evaluation': {'max': float(stats[:, 0].max() / n_local) if len(stats) else 0.0,
                                                  'mean': float(stats[:, 0].mean() / n_local) if len(stats) else 0.0},
        'inner_Lp_iterations_per_cell_and_evaluation': lp_total / n_local / max(n_eval, 1),
        'per_cell_and_evaluation': {k: (c1[k] - c0[k]) / max(n_eval, 1) / n_local for k in ('stress_calls', 'state_iter', 'substeps', 'cutbacks')},
        'gpu_launches': int(launches),
        'wall_ms': wall_ms,
        'kernel_ms': {'per': 'residual evaluation, mean over the %d evaluations of the timed solve (CUDA events on the solver stream)' % int(prof['evaluations']),
                      'constitutive': con_ms, 'reduce_aim': prof['reduce_aim_ms'] / ne, 'fft_forward_xy': prof['fft_forward_ms'] / ne,
                      'gamma_div_zpass': prof['gamma_zpass_ms'] / ne, 'fft_inverse_yx': prof['fft_inverse_ms'] / ne,
                      'ngmres_wall_ms_per_iteration': dict(zip(('evaluations', 'dot_products', 'least_squares_host', 'combination_and_distances',
                                                                'norm_of_F_A', 'XM', 'whole_solve', 'iterations', 'selection_and_allocation',
                                                                'history_copies', 'convergence_test', 'first_evaluation_and_setup'),
                                                               [float(1e3 * (ts1[k] - ts0[k]) / max(ts1[7] - ts0[7], 1.0)) for k in range(12)])),
                      'outside_evaluations_ms_per_step': (ms - sum(prof[k] for k in ('constitutive_ms', 'reduce_aim_ms', 'fft_forward_ms', 'gamma_zpass_ms', 'fft_inverse_ms'))) / K},
        # the dominant kernel by time: the per-point constitutive update (FP64 vector pipe; tensor cores do not apply)
        'roofline': {'bound': 'fp64', 'kernel': 'dmk_pheno_lanes_kernel (per-point constitutive update), %.1f %% of the timed solve' % (100.0 * prof['constitutive_ms'] / ms) if phen
                     else 'per-point constitutive kernel, %.1f %% of the timed solve' % (100.0 * prof['constitutive_ms'] / ms),
                     'achieved': fp64_achieved, 'peak': fp64_peak, 'unit': 'TFLOP/s',
                     'frac': (fp64_achieved / fp64_peak) if (fp64_peak and fp64_achieved is not None and phen) else None,
                     'traffic': kc.get('constitutive_traffic_256') if (world == 1 and cells == (256, 256, 256)) else None,
                     'flops': 'Lp Newton iterations (device counters, timed region) x %.0f flop per iteration (%s)' % (kc['flop_per_lp_iter'], kc.get('flop_source', '')),
                     'peak_source': 'FP64 DFMA rate measured in this run (dmk_fp64_peak; MEASURED_PEAKS.json has no FP64 figure)'},
        'roofline_hbm': {'bound': 'hbm', 'kernel': 'FFT + Gamma/div section (x r2c, y, fused z, y^-1, x c2r): 5 passes x 144 B/cell, mean over the timed solve',
                         'achieved': achieved_hbm, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': achieved_hbm / hbm_peak,
                         'traffic': kc.get('fft_section_traffic_256') if (world == 1 and cells == (256, 256, 256)) else None,
                         'peak_source': peak_src},
        'e2e': {'value': n_global / (ms_e2e / Ke * 1e-3), 'unit': 'cells*evaluations/s', 'h2d_bytes_per_step': int(72 * n_local),
                'd2h_bytes_per_step': int(72 * n_local), 'ms_per_step': ms_e2e / Ke,
                'step': 'one residual evaluation of a running-solve iterate (first iterates of increment %d) through dmk_mech_form_residual_host: '
                        'pinned host F in, host r out, both copies inside the timed region' % (K + 1),
                'call': 'dmk_mech_form_residual_host',
                'device_resident_same_evaluations': {'value': n_global / (ms_same / Ke * 1e-3), 'ms_per_step': ms_same / Ke},
                'note': 'these late-solve iterates are cheaper than the mean evaluation of the timed solve (fewer inner iterations, no '
                        'cut-backs): compare e2e with device_resident_same_evaluations for the cost of the copies, with `value` only as throughputs '
                        'of different evaluations'},
        'clocks': sampler.summary(),
    }
    if pc is not None:
        line['parity_check'] = pc
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        sample = (20, 20, 20) if args.config not in ('C1', 'C2') else W.CONFIGS[args.config]['cells']
        if args.config == 'C2':
            sample = (32, 32, 32)
        r = cpu_solve(args.config, sample, min(K, 3), threads)
        line['cpu_baseline'] = {'value': r['value'], 'unit': 'cells*evaluations/s', 'cores': threads, 'kind': 'port',
                                'sample': '%dx%dx%d sub-problem of the same workload, same protocol (first %d increments solved): %d evaluations, '
                                          'NGMRES iterations %s, %.1f s' % (sample + (min(K, 3), r['evaluations'], r['iterations'], r['wall_s'])),
                                'per_point_update_cells_per_s': r['per_point_update_cells_per_s']}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def setup_state(cells, world=1, rank=0):
    """protocol `evaluation`: solver `s` on a polycrystal after two homogeneous (Taylor) increments and N_RELAX fixed-point
    iterations of the third; returns (s, X, R, grains) with X the iterate to evaluate, R a residual buffer"""
    import torch
    from solver_cpp_b200 import workloads as W, mech
    s, material, material_at, grains, cfg = make_problem('C3', cells, world, rank)
    g = s.g
    s.init()
    F = g.new_field()
    P = g.new_field()
    for inc in (1, 2):
        F.copy_(torch.from_numpy(np.ascontiguousarray(taylor_F(inc).T.reshape(9))).to(F.device).expand_as(F))
        g.constitutive_response(F, DT, P)
        g.materialpoint_forward()
    st = mech.parse_load(W.TENSION_X)[0]
    Fl = g.new_field()
    Fl.copy_(torch.from_numpy(np.ascontiguousarray(taylor_F(1).T.reshape(9))).to(F.device).expand_as(F))
    s.set_history(F_lastInc=Fl, F=F, F_aim=taylor_F(2), F_aim_lastInc=taylor_F(1))
    s.forward(False, True, DT, DT, 5.0 - 2 * DT, st['deformation'], st['stress'], st['R'])
    s.prepare_evaluation()
    X = s.field('F')
    R = g.new_field()
    for _ in range(N_RELAX):
        s.form_residual(X, R)
        X.sub_(R)
    g.synchronize()
    return s, X, R, grains


def run_evaluation(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device: the product path has no CPU fallback')
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    cells = tuple(args.cells) if args.cells else default_cells(world)
    nx, ny, nz = cells
    s, X, R, grains = setup_state(cells, world, rank)
    g = s.g
    n_local, n_global = g.n_local, nx * ny * nz

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    W_, K = max(3, args.warmup), max(1, args.steps)
    for _ in range(W_):
        s.form_residual(X, R)
    g.set_profiling(True)
    g.get_profile(reset=True)
    c0 = g.counters()
    l0 = g.kernel_launches()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        s.form_residual(X, R)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = g.kernel_launches() - l0
    prof = g.get_profile(reset=True)
    c1 = g.counters()
    g.set_profiling(False)
    t = torch.tensor([ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    fp64_peak = g.fp64_peak_tflops()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm_peak, peak_src = peaks()
    kc = kernel_constants()
    per_step = ms / K
    ne = max(prof['evaluations'], 1)
    fft_ms = (prof['fft_forward_ms'] + prof['gamma_zpass_ms'] + prof['fft_inverse_ms']) / ne
    con_ms = prof['constitutive_ms'] / ne
    achieved = 720.0 * n_local / (fft_ms * 1e-3) / 1e9
    lp_iters = (c1['Lp_iter'] - c0['Lp_iter']) / K
    fp64_achieved = lp_iters * kc['flop_per_lp_iter'] / (con_ms * 1e-3) / 1e12
    line = {
        'metric': 'material-point updates/s', 'value': n_global / (per_step * 1e-3), 'unit': 'cells*evaluations/s', 'n_gpus': world,
        'steps': K, 'warmup': W_, 'ms_per_step': per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'C3: 256^3-per-GPU 500-grain phenopowerlaw Cu polycrystal, tensionX (BASELINE.json configs[2])',
                   'cells': list(cells), 'cells_per_gpu': n_local, 'grains': grains, 'protocol': 'evaluation',
                   'step': 'one residual evaluation (form_residual)',
                   'state': 'after 2 homogeneous (Taylor) increments of dot_F11 = 1e-3, dt = 0.5 s, and %d fixed-point iterations F <- F - r(F) of increment 3' % N_RELAX,
                   'l2': 'inputs larger than L2 (field 1.2 GB per GPU, per-point state 16 GB)', 'decomposition': 'z-slabs, 1 per GPU'},
        'inner_Lp_iterations_per_cell': lp_iters / n_local,
        'per_cell': {k: (c1[k] - c0[k]) / K / n_local for k in ('stress_calls', 'state_iter', 'substeps', 'cutbacks')},
        'gpu_launches': int(launches),
        'kernel_ms': {'constitutive': con_ms, 'reduce_aim': prof['reduce_aim_ms'] / ne, 'fft_forward_xy': prof['fft_forward_ms'] / ne,
                      'gamma_div_zpass': prof['gamma_zpass_ms'] / ne, 'fft_inverse_yx': prof['fft_inverse_ms'] / ne},
        'roofline': {'bound': 'fp64', 'kernel': 'dmk_pheno_lanes_kernel (per-point constitutive update)', 'achieved': fp64_achieved,
                     'peak': fp64_peak, 'unit': 'TFLOP/s', 'frac': fp64_achieved / fp64_peak if fp64_peak else None, 'traffic': None,
                     'flops': 'Lp Newton iterations x %.0f flop (%s)' % (kc['flop_per_lp_iter'], kc.get('flop_source', '')),
                     'peak_source': 'FP64 DFMA rate measured in this run (dmk_fp64_peak)'},
        'roofline_hbm': {'bound': 'hbm', 'kernel': 'FFT + Gamma/div section (x r2c, y, fused z, y^-1, x c2r): 5 passes x 144 B/cell',
                         'achieved': achieved, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': achieved / hbm_peak, 'traffic': None,
                         'peak_source': peak_src},
        'clocks': sampler.summary(),
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + '\n').encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # Libraries (NCCL's version banner, torch) write to file descriptor 1; the contract is ONE JSON line on stdout.  Everything
    # but that line is sent to stderr: fd 1 is pointed at fd 2 and the line is written to the saved descriptor (emit).
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--config', default='C3', choices=['C1', 'C2', 'C3', 'C4', 'C5'])
    ap.add_argument('--protocol', default='solve', choices=['solve', 'evaluation'])
    ap.add_argument('--cells', type=int, nargs=3, default=None)
    ap.add_argument('--no-cpu', action='store_true')
    args = ap.parse_args()
    # bounded time: a rank that waits for ever (a peer that died, a collective that never completes) ends the run with a
    # diagnostic instead of holding the GPUs until the caller's own limit; DMK_BENCH_TIMEOUT_S = 0 turns the guard off
    limit = float(os.environ.get('DMK_BENCH_TIMEOUT_S', '2700'))
    if limit > 0:
        import faulthandler
        import threading

        def expired():
            sys.stderr.write('bench.py: rank %s still running after %.0f s (DMK_BENCH_TIMEOUT_S); stacks follow\n' % (os.environ.get('RANK', '0'), limit))
            faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
            sys.stderr.flush()
            os._exit(124)
        guard = threading.Timer(limit, expired)
        guard.daemon = True
        guard.start()
    if args.impl == 'reference':
        run_reference(args)
    elif args.protocol == 'solve':
        run_solve(args)
    else:
        run_evaluation(args)


if __name__ == '__main__':
    main()
