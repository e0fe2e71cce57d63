#!/usr/bin/env python
"""bench.py -- one JSON line for the DAMASK grid hot path on B200 (contract: see the task statement).

Workload (config.workload): BASELINE.json config 3, the 256^3 500-grain phenopowerlaw Cu polycrystal under tensionX, at
N=1; weak scaling with 256^3 cells per GPU: 256x256x512 on 2, 256x512x512 on 4 and 512^3 (BASELINE.json config 4's grid)
on 8 GPUs, z-slab decomposed, grain density kept at 500 per 256^3.
A step = ONE residual evaluation of the spectral basic scheme (MechSolverBasic::form_residual): per-point constitutive
update of every cell from its increment-start state + homogenisation + forward r2c FFT + fused divergence-RMS/Gamma
pass + inverse c2r FFT.  State before timing: two homogeneous (Taylor) increments of the load case, so every cell is
flowing plastically; the timed evaluations are those of the third increment's first iterate.
metric = material-point updates / s (cells x evaluations / s, whole job).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--cells nx ny nz]
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


def default_cells(world):
    """256^3 cells per GPU; the grid grows z, then y, then x, so that every line length stays a power of two <= 512"""
    return WEAK_CELLS.get(world, (256, 256, 256 * world))

CPU_SAMPLE_CELLS = (40, 40, 40)
N_RELAX = 2                                 # fixed-point iterations before the timed evaluations (both arms)


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def section_traffic(cells, world):
    """dram__bytes_read + dram__bytes_write of the five FFT/Gamma kernels of one evaluation from the round's `ncu --set full`
    capture (profiles/r01_step_ncu_metrics.csv), bytes per evaluation; None for configurations that were not captured"""
    p = os.path.join(ROOT, 'profiles', 'r01_step_ncu_metrics.csv')
    if world != 1 or tuple(cells) != (256, 256, 256) or not os.path.exists(p):
        return None
    import csv
    rows = list(csv.reader(open(p)))
    h = rows[0]
    tot = 0.0
    for r in rows[2:]:
        if 'dmk_fast_' in r[h.index('Kernel Name')]:
            tot += (float(r[h.index('dram__bytes_read.sum')]) + float(r[h.index('dram__bytes_write.sum')])) * 1e9
    return tot or None


def taylor_F(inc, shape_prefix=None):
    """homogeneous tensionX deformation gradient after `inc` increments (lateral contraction ~ incompressible flow)"""
    e = RATE * DT * inc
    return np.diag([1.0 + e, 1.0 - 0.5 * e, 1.0 - 0.5 * e])


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
def cpu_evaluations(cells, steps, warmup, threads):
    """the oracle (CPU restatement of the reference algorithm; oracle/) on a bounded sample of the same workload, same
    protocol: 2 Taylor increments, then `steps` residual evaluations.  Returns (cells/s, seconds per evaluation)."""
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
        # residual evaluation without the per-evaluation tangent (the product computes it once per increment, too)
        F = X.reshape(o.shape)
        tc = time.perf_counter()
        o.con.mechanical_response(F, o.params['dt'])
        P, _ = o.con.mechanical_response2(o.params['dt'], tangent=False)
        split['constitutive_s'] += time.perf_counter() - tc
        ts = time.perf_counter()
        try:
            return spectral_tail(P)
        finally:
            split['spectral_s'] += time.perf_counter() - ts

    def spectral_tail(P):
        P = P.reshape(o.shape)
        P_av = P.sum(axis=(0, 1, 2)) * o.grid.wgt
        hat = osp.fft_forward(P)
        osp.divergence_rms(hat, o.xi1st, o.grid, o.scaled_size)
        dF = np.einsum('ijkl,kl->ij', o.S, P_av - o.P_aim)
        out = osp.gamma_hat_apply(hat, o.xi1st, o.C_ref)
        out[0, 0, 0] = dF
        return osp.fft_backward(out, o.grid.cells, 1.0)
    for _ in range(N_RELAX):                   # same state as the GPU arm: N_RELAX fixed-point iterations X <- X - r(X)
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
    steps, warmup = max(1, args.steps), max(1, min(args.warmup, 2))
    steps = min(steps, 5)                       # bounded: each step is ~2-4 s of CPU work on the sample
    v, sec = cpu_evaluations(CPU_SAMPLE_CELLS, steps, warmup, threads)
    cells = tuple(args.cells) if args.cells else default_cells(args.gpus)
    sample = '%dx%dx%d sub-problem of the same workload (same material, load and protocol)' % CPU_SAMPLE_CELLS
    line = {'impl': 'reference', 'metric': 'material-point updates/s', 'value': v, 'unit': 'cells*evaluations/s', 'n_gpus': args.gpus,
            'steps': steps, 'warmup': warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'C3: 256^3-per-GPU 500-grain phenopowerlaw Cu polycrystal, tensionX; CPU arm on a bounded sample',
                       'cells': list(cells), 'sample_cells': list(CPU_SAMPLE_CELLS)},
            'cpu_baseline': {'value': v, 'unit': 'cells*evaluations/s', 'cores': threads, 'kind': 'port', 'sample': sample,
                             'seconds_per_evaluation': dict(cpu_evaluations.split)},
            'e2e': {'value': v, 'unit': 'cells*evaluations/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'note': 'oracle/ = CPU restatement of the reference algorithm (the reference itself cannot be built here: no gfortran/'
                    'MPI/FFTW/PETSc/HDF5); constitutive loop OpenMP over cells, FFT scipy pocketfft with all cores'}
    emit(line)


# ------------------------------------------------------------------------------------------------ GPU arm
def setup_state(cells, world=1, rank=0):
    """the bench state on this rank's z-slab: solver `s` on a polycrystal after two homogeneous (Taylor) increments and
    N_RELAX fixed-point iterations of the third; returns (s, X, R) with X the iterate to evaluate, R a residual buffer"""
    import torch
    from solver_cpp_b200 import workloads as W, mech
    from solver_cpp_b200 import parallel
    nx, ny, nz = cells
    grains = max(8, int(round(500 * np.prod(cells) / 256.0 ** 3)))
    z0, z1 = parallel.slab_range(nz, world, rank)
    material, material_at, size = W.polycrystal(cells, grains, 'Cu_pheno', 0, z_range=(z0, z1))
    comm = parallel.make_comm(world, rank) if world > 1 else None
    s = mech.make_solver(cells, size, material, material_at, world_size=world, rank=rank, comm=comm)
    g = s.g
    s.init()
    # ---- state: two homogeneous increments
    F = g.new_field()
    P = g.new_field()
    for inc in (1, 2):
        F.copy_(torch.from_numpy(np.ascontiguousarray(taylor_F(inc).T.reshape(9))).to(F.device).expand_as(F))
        g.constitutive_response(F, DT, P)
        g.materialpoint_forward()
    st = mech.parse_load(W.TENSION_X)[0]
    # the solver's own fields follow the same history: F_lastInc = F(1), F = F(2), aims alike
    Fl = g.new_field()
    Fl.copy_(torch.from_numpy(np.ascontiguousarray(taylor_F(1).T.reshape(9))).to(F.device).expand_as(F))
    s.set_history(F_lastInc=Fl, F=F, F_aim=taylor_F(2), F_aim_lastInc=taylor_F(1))
    s.forward(False, True, DT, DT, 5.0 - 2 * DT, st['deformation'], st['stress'], st['R'])
    s.prepare_evaluation()                     # masked compliance + device scalar block, as calculate_solution does
    X = s.field('F')
    R = g.new_field()
    # a realistic iterate: N_RELAX fixed-point iterations of the basic scheme, X <- X - r(X), so that F fluctuates inside
    # the grains as it does in a running solve (after the Taylor increments alone every cell of a grain is identical and
    # the warps of the per-point kernel would be unrealistically coherent)
    for _ in range(N_RELAX):
        s.form_residual(X, R)
        X.sub_(R)
    g.synchronize()
    return s, X, R, grains


def run_ours(args):
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

    # ---- kernel-resident timing: K evaluations, inputs already in HBM
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
    # ---- end to end through the reference-facing call: host buffers in, host buffers out, copies inside the timed region
    Fh = torch.empty(X.shape, dtype=torch.float64).pin_memory()
    rh = torch.empty(X.shape, dtype=torch.float64).pin_memory()
    Fh.copy_(X.cpu())
    Fn, rn = Fh.numpy(), rh.numpy()
    s.form_residual_host(Fn, rn)
    barrier()
    Ke = max(1, min(K, 5))
    t0 = time.perf_counter()
    e0.record()
    for _ in range(Ke):
        s.form_residual_host(Fn, rn)
    e1.record()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), 0.0)
    wall_e2e = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([max(ms_e2e, wall_e2e)], dtype=torch.float64, device='cuda')
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
    per_step = ms / K
    value = n_global / (per_step * 1e-3)
    fft_ms = (prof['fft_forward_ms'] + prof['gamma_zpass_ms'] + prof['fft_inverse_ms']) / max(prof['evaluations'], 1)
    con_ms = prof['constitutive_ms'] / max(prof['evaluations'], 1)
    alg_bytes = 720.0 * n_local                              # 5 passes x 144 B per cell, this GPU's slab
    achieved = alg_bytes / (fft_ms * 1e-3) / 1e9
    lp_iters = (c1['Lp_iter'] - c0['Lp_iter']) / K
    flop_per_lp_iter = 2.0 * 2260.0                          # DFMA-equivalent flops per Lp Newton iteration, see DESIGN.md (ncu-calibrated)
    fp64_achieved = lp_iters * flop_per_lp_iter / (con_ms * 1e-3) / 1e12
    line = {
        'metric': 'material-point updates/s', 'value': value, 'unit': 'cells*evaluations/s', 'n_gpus': world, 'steps': K, 'warmup': W_,
        'ms_per_step': per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'C3: 256^3-per-GPU 500-grain phenopowerlaw Cu polycrystal, tensionX (BASELINE.json configs[2])',
                   'cells': list(cells), 'cells_per_gpu': n_local, 'grains': grains, 'step': 'one residual evaluation (form_residual)',
                   'state': 'after 2 homogeneous (Taylor) increments of dot_F11 = 1e-3, dt = 0.5 s, and %d fixed-point iterations F <- F - r(F) of increment 3' % N_RELAX,
                   'l2': 'inputs larger than L2 (field 1.2 GB per GPU, per-point state 16 GB)', 'decomposition': 'z-slabs, 1 per GPU'},
        'newton_iters_per_s': 0.5 * 1e3 / per_step,
        'inner_Lp_iterations_per_cell': lp_iters / n_local,
        'deferred_cell_fraction': (c1['deferred'] - c0['deferred']) / K / n_local,
        'per_cell': {k: (c1[k] - c0[k]) / K / n_local for k in ('stress_calls', 'state_iter', 'substeps', 'cutbacks')},
        'gpu_launches': int(launches),
        'kernel_ms': {'constitutive': con_ms, 'reduce_aim': prof['reduce_aim_ms'] / max(prof['evaluations'], 1),
                      'fft_forward_xy': prof['fft_forward_ms'] / max(prof['evaluations'], 1),
                      'gamma_div_zpass': prof['gamma_zpass_ms'] / max(prof['evaluations'], 1),
                      'fft_inverse_yx': prof['fft_inverse_ms'] / max(prof['evaluations'], 1)},
        'roofline': {'bound': 'hbm', 'kernel': 'FFT + Gamma/div section (x r2c, y, fused z, y^-1, x c2r): 5 passes x 144 B/cell',
                     'achieved': achieved, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': achieved / hbm_peak,
                     'traffic': section_traffic(cells, world), 'peak_source': peak_src},
        'roofline_fp64': {'bound': 'fp64', 'kernel': 'dmk_constitutive_kernel (dominant by time)', 'achieved': fp64_achieved,
                          'peak': fp64_peak, 'unit': 'TFLOP/s', 'frac': fp64_achieved / fp64_peak if fp64_peak else None,
                          'peak_source': 'measured in this run (DFMA chains, dmk_fp64_peak)',
                          'flops': 'Lp Newton iterations x %.0f flop (DESIGN.md)' % flop_per_lp_iter},
        'e2e': {'value': n_global / (ms_e2e / Ke * 1e-3), 'unit': 'cells*evaluations/s', 'h2d_bytes_per_step': int(72 * n_local),
                'd2h_bytes_per_step': int(72 * n_local), 'ms_per_step': ms_e2e / Ke,
                'call': 'dmk_mech_form_residual_host (pinned host F in, host r out)'},
        'clocks': sampler.summary(),
    }
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        v, sec = cpu_evaluations(CPU_SAMPLE_CELLS, 3, 1, threads)
        line['cpu_baseline'] = {'value': v, 'unit': 'cells*evaluations/s', 'cores': threads, 'kind': 'port',
                                'sample': '%dx%dx%d sub-problem of the same workload, same protocol, %.2f s per evaluation'
                                          % (CPU_SAMPLE_CELLS + (sec,)),
                                'seconds_per_evaluation': dict(cpu_evaluations.split)}
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
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--cells', type=int, nargs=3, default=None)
    ap.add_argument('--no-cpu', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
