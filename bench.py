#!/usr/bin/env python
"""bench.py — LM iterations/s of the bundle-adjustment hot path on B200 (BASELINE.json metric).

A "step" is one Levenberg-Marquardt iteration (levenberg_marquardt.m:117-206) on the
synthetic self-calibration scene of BASELINE config 4 (1000 cameras x 200k points x 2M
observations, shared Brown IO, depend datum): residual + Jacobian + normal-equation
assembly, damped Schur reduction, dense FP64 Cholesky of the 6002-order reduced system,
back-substitution, |Jp| statistics and the trial-point residual.  Accepted trial points are
taken, so successive steps walk the real LM path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N>1 (torchrun, one rank per GPU): points are sharded across ranks (weak scaling: 200k points
and 2M observations per rank over the same 1000 cameras), reduced system combined by NCCL
allreduce; value counts 2M-observation equivalents so it aggregates over ranks.
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

C4 = dict(nImg=1000, nOP=200000, rays=10)
ALG_BYTES = {   # algorithmic HBM bytes per observation (DESIGN.md "Kernels")
    'k_cam_side': 64.0,      # uv 16 + isig 16 + pt idx 4 + (img via chunk) + point gather 24 + chunk partials ~4
    'k_point_side': 64.0 + 144.0 + 41.6,   # reads as above, writes W_o 144 B/obs + point record 416 B/pt (10 rays)
    'k_resid': 64.0,
    'k_jp': 64.0 + 24.0,
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--nimg', type=int, default=C4['nImg'])
    ap.add_argument('--nop', type=int, default=C4['nOP'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md): one
    long-running `nvidia-smi -lms 50` whose lines are collected by this thread."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None
        self.t_mark = None

    def run(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                line = line.strip()
                if line:
                    self.rows.append((time.perf_counter(), [c.strip() for c in line.split(',')]))
        except Exception:
            pass

    def mark(self):
        """Only samples taken after this call count (start of the timed region)."""
        self.t_mark = time.perf_counter()

    def stop(self):
        self.stop_flag = True
        if self.proc is not None:
            try:
                self.proc.kill()
            except Exception:
                pass

    def summary(self):
        rows = [r for t, r in self.rows if self.t_mark is None or t >= self.t_mark]
        sm = [float(r[0]) for r in rows if r and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[k] for r in rows if len(r) >= 7 for k in range(4)
                          if r[3 + k].lower().startswith('active')})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(rows)}


def cpu_lm_iteration_sample(nImg=100, nOP=20000, iters=2):
    """Reference CPU path (oracle port of levenberg_marquardt.m: sparse J, J'J, sparse solve of the
    FULL damped system) on a bounded 1/10-scale sample of the workload.  Returns (it/s, desc)."""
    import scipy.sparse as sp
    from dbat_b200.synth import make_scene
    from oracle import lsa
    from oracle.cameramodel import brown_euler_cam4
    from oracle.dbatstruct import buildweightmatrix, serialize
    s, _ = make_scene(nImg, nOP, rays=10)
    x = serialize(s)
    R = np.sqrt(buildweightmatrix(s))
    nOPx = len(s.bundle.serial.OP.dest)
    lsa.set_ordering(np.concatenate([np.arange(len(x) - nOPx, len(x)), np.arange(len(x) - nOPx)[::-1]]))
    I = sp.identity(len(x), format='csc')
    t0 = time.perf_counter()
    for _ in range(iters):
        f, J = brown_euler_cam4(x, s, True)
        r = R * f
        Jw = (sp.diags(R) @ J).tocsc()
        N = (Jw.T @ Jw).tocsc()
        g = Jw.T @ r
        p, _ = lsa._solve_spd(N + 1e-10 * N.diagonal().sum() / len(x) * I, -g)
        Jp = Jw @ p
        fNew = brown_euler_cam4(x + p, s, False)[0]
        if np.sum((R * fNew) ** 2) < r @ r:
            x = x + p
    dt = time.perf_counter() - t0
    nobs = len(s.IP.img)
    return iters / dt, nobs, ('%d LM iterations of the oracle (SciPy sparse J, J\'J, SuperLU solve of the full '
                              'damped system) on a %d-camera x %d-point x %d-observation scene of the same '
                              'generator' % (iters, nImg, nOP, nobs))


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU algorithm (oracle port; no MATLAB/Octave exists
    here) on the host cores, bounded sample of the same workload."""
    if rank != 0:
        return
    steps = max(1, args.steps)
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_lm_iteration_sample(20, 2000, 1)
    nI, nP = (100, 20000) if args.nimg >= 100 else (args.nimg, args.nop)
    t0 = time.perf_counter()
    its, nobs, desc = cpu_lm_iteration_sample(nI, nP, steps)
    dt = time.perf_counter() - t0
    # 2M-observation equivalents, same unit as the GPU arm
    value = its * nobs / 2.0e6
    cores = os.cpu_count() or 1
    line = {
        'impl': 'reference', 'metric': 'lm_iterations_per_s', 'value': value,
        'unit': 'LM iterations/s (2M-observation equivalents)', 'n_gpus': 0, 'steps': steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 / its, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'synthetic self-calibration LM, 1/10-scale sample of BASELINE config 4 '
                               '(%d cameras x %d points x %d observations); value scaled by observations/2M'
                               % (nI, nP, nobs)},
        'cpu_baseline': {'value': value, 'unit': 'LM iterations/s (2M-observation equivalents)',
                         'cores': cores, 'kind': 'port', 'sample': desc},
        'e2e': {'value': value, 'unit': 'LM iterations/s (2M-observation equivalents)',
                'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0, 'seconds': dt,
    }
    print(json.dumps(line), flush=True)


def fp64_peak_tflops():
    """cuBLAS DGEMM yardstick for the FP64 tensor pipe (MEASURED_PEAKS.json has no FP64 figure)."""
    import torch
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device='cuda')
    b = torch.randn(n, n, dtype=torch.float64, device='cuda')
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def main():
    args = parse()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    import dbat_b200
    from dbat_b200.parallel import ShardedProblem
    from dbat_b200.synth import make_scene

    # scene: config 4 per rank (weak scaling over points; the same 1000 stations)
    nImg, nOP = args.nimg, args.nop * world
    s, _ = make_scene(nImg, nOP, rays=C4['rays'], cache_dir=os.environ.get('DBAT_SCENE_CACHE', '/tmp'))
    nObsGlobal = len(s.IP.img)
    x0 = dbat_b200.serialize(s)
    if world > 1:
        P = ShardedProblem(s, rank, world)
    else:
        P = dbat_b200.Problem(s)
    n = P.n

    # pinned host buffers for the end-to-end leg
    xh = torch.empty(n, dtype=torch.float64).pin_memory()
    ph = torch.empty(n, dtype=torch.float64).pin_memory()
    xh.numpy()[:] = x0
    lam = 0.0

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(k, resident):
        """k LM iterations; returns (wall seconds, device ms list, launches)."""
        dev_ms, launches = [], 0
        sync()
        t0 = time.perf_counter()
        for _ in range(k):
            if resident:
                _, st = P.normal_step(None, lam, trial=True, accept=True, want_p=False)
            else:
                _, st = P.normal_step(xh.numpy(), lam, trial=True, accept=False, p_out=ph.numpy())
            dev_ms.append(st['device_ms'])
            launches += st['launches']
        sync()
        return time.perf_counter() - t0, dev_ms, launches

    # device-resident arm
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    P.normal_step(x0, lam, trial=True, accept=False, want_p=False)       # upload x0
    timed(args.warmup, True)
    P.normal_step(x0, lam, trial=True, accept=False, want_p=False)       # restart the path at x0
    sampler.mark()
    wall, dev_ms, launches = timed(args.steps, True)
    phases = P.phase_times() if hasattr(P, 'phase_times') else {}
    # end-to-end arm (host x in, host p out every step)
    timed(min(args.warmup, 3), False)
    wall_e2e, _, _ = timed(args.steps, False)
    sampler.stop()

    t = torch.tensor([wall, wall_e2e, float(np.sum(dev_ms))], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall, wall_e2e, dev_total = [float(v) for v in t.cpu()]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    equiv = nObsGlobal / 2.0e6
    its = args.steps / (dev_total * 1e-3)
    value = its * equiv
    e2e_value = args.steps / wall_e2e * equiv
    line = {
        'metric': 'lm_iterations_per_s', 'value': value,
        'unit': 'LM iterations/s (2M-observation equivalents)', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dev_total / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'BASELINE config 4: synthetic %d cameras x %d points x %d observations, shared IO '
                               '+ Brown self-calibration (model 3), LM iteration; n=%d unknowns, reduced order %d'
                               % (nImg, nOP, nObsGlobal, n, n - 3 * nOP),
                   'l2': 'inputs (>600 MB of observation, cross-block and reduced-system arrays) exceed the 126 MB L2',
                   'timing': 'CUDA events on the library stream around each step (max over ranks); wall '
                             'clock %.3f ms/step' % (1e3 * wall / args.steps),
                   'parallelism': 'points sharded over %d rank(s), NCCL allreduce of the reduced system' % world},
        'e2e': {'value': e2e_value, 'unit': 'LM iterations/s (2M-observation equivalents)',
                'h2d_bytes_per_step': 8 * n, 'd2h_bytes_per_step': 8 * n + 64},
        'gpu_launches': int(launches),
        'clocks': sampler.summary(),
        'phases_ms_last_step': {k: v[0] for k, v in phases.items()},
    }
    # roofline of the dominant kernel family of the step
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        hbm_peak, peak_src = float(peaks['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        hbm_peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
    nRed = n - 3 * nOP
    ph_ms = {k: v[0] for k, v in phases.items() if k != 'total'}
    dom = max(ph_ms, key=ph_ms.get) if ph_ms else 'cholesky'
    line['dominant_phase'] = dom
    if dom == 'cholesky':
        fp64 = fp64_peak_tflops()
        ach = (nRed ** 3 / 3.0) / (ph_ms['cholesky'] * 1e-3) / 1e12
        # DRAM bytes of one factorisation from the ncu capture profiles/chol_factor_n6002_dram.csv
        # (dram__bytes_read.sum + dram__bytes_write.sum over its k_potrf128 / k_gemm_nt launches, cold
        # caches): known for the config-4 size only
        traffic = 3.306e9 if nRed == 6002 else None
        line['roofline'] = {'bound': 'tensor', 'achieved': ach, 'peak': fp64, 'unit': 'TFLOP/s',
                            'frac': ach / fp64, 'traffic': traffic,
                            'kernel': 'k_gemm_nt (DMMA blocked Cholesky, %d^3/3 flop per factorisation)' % nRed,
                            'peak_source': 'cuBLAS DGEMM 8192^3 measured in this run (FP64; MEASURED_PEAKS.json '
                                           'holds no FP64 figure)'}
    else:
        nobs_rank = nObsGlobal / world
        key, kern = {'eval_jac_assembly': (ALG_BYTES['k_cam_side'] + ALG_BYTES['k_point_side'], 'k_cam_side+k_point_side'),
                     'trial_residual': (ALG_BYTES['k_resid'], 'k_resid'),
                     'jp_stats': (ALG_BYTES['k_jp'], 'k_jp')}.get(dom, (144.0 + 41.6, 'k_schur'))
        ach = key * nobs_rank / (ph_ms.get(dom, 1.0) * 1e-3) / 1e9
        line['roofline'] = {'bound': 'hbm', 'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s',
                            'frac': ach / hbm_peak, 'traffic': None, 'kernel': kern, 'peak_source': peak_src}
    # HBM view of the streaming kernels (explains the step; algorithmic bytes per observation)
    nobs_rank = nObsGlobal / world
    hbm = {}
    for phn, (b, kern) in {'eval_jac_assembly': (ALG_BYTES['k_cam_side'] + ALG_BYTES['k_point_side'], 'k_cam_side+k_point_side'),
                           'trial_residual': (ALG_BYTES['k_resid'], 'k_resid'),
                           'jp_stats': (ALG_BYTES['k_jp'], 'k_jp')}.items():
        if phn in ph_ms and ph_ms[phn] > 0:
            g = b * nobs_rank / (ph_ms[phn] * 1e-3) / 1e9
            hbm[kern] = {'GB/s': g, 'frac_of_measured_hbm': g / hbm_peak, 'ms': ph_ms[phn], 'bytes_per_obs': b}
    line['hbm_streams'] = hbm
    if world == 1 and not args.no_cpu_baseline:
        nI, nPt = (100, 20000) if nImg >= 100 else (nImg, nOP)
        cits, cobs, desc = cpu_lm_iteration_sample(nI, nPt, 2)
        line['cpu_baseline'] = {'value': cits * cobs / 2.0e6, 'unit': 'LM iterations/s (2M-observation equivalents)',
                                'cores': os.cpu_count() or 1, 'kind': 'port', 'sample': desc}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
