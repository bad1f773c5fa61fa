#!/usr/bin/env python
"""bench.py -- time-to-tolerance and space-time DOF/s of the MGRIT hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg5|cfg2]

Workload (BASELINE.json configs[4], the one the metric's target is quoted on; fits one GPU):
    heat_1d backward Euler, nx = 1025 (1023 dofs), nt = 2^20 + 1 on t in [0, 2], FCF-relaxation V-cycles, 3 levels
    (coarsening 64 x 16, chosen by the builder: BASELINE leaves it open), nested iteration, tol 1e-10; strong
    scaling over the time ranks (one process per GPU).
One "step" = one complete solve: setup (incl. nested iteration) + MGRIT iterations until conv < 1e-10.
Prints ONE JSON line (rank 0).  `value` times the solve with every table already in HBM; `e2e` times the public API
from host NumPy inputs (application objects, Mgrit(), solve(), result copied back to the host).
`--impl reference` times the reference's CPU algorithm (the oracle port: per-point Python loop + SciPy SuperLU, exactly
what pymgrit does) on a bounded sample of the same workload.
"""
import argparse
import json
import logging
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

METRIC = 'MGRIT space-time DOF/s (heat_1d nx=1025, solve to 1e-10)'
UNIT = 'DOF/s'


def rhs(x, t):
    return -np.sin(np.pi * x) * (np.sin(t) - np.pi ** 2 * np.cos(t))


def init_cond(x):
    return np.sin(np.pi * x)


WORKLOADS = {
    # name: (nt, coarsening factor per level transition).  cfg2 is BASELINE.json configs[1] verbatim; for cfg5 BASELINE
    # leaves the hierarchy to the builder: (64, 16) -> 16385 and 1025 coarse points converges in 3 FCF V-cycles and is
    # (with (128, 8), which ends closer to the tolerance) the fastest of the 24 hierarchies tried on 1 and on 8 GPUs
    # (scripts/hierarchy_sweep.py, profiles/r01o_hierarchy_sweep.txt; (16, 16, 8) was the choice of the earlier kernels).
    'cfg5': (2 ** 20 + 1, (64, 16)),
    'cfg2': (16385, (4, 4)),
}
# Bounded CPU sample: the first nt_sample time points of the SAME problem (same dt = 2 / 2^20, same coarsening factors,
# same cycle), i.e. a time window of the workload, so that the work per space-time DOF is the workload's.
CPU_SAMPLE = (2049, (64, 16))          # one core: about 10 s
CPU_SAMPLE_MP = (32769, (64, 16))      # several cores (time-parallel workers): about 10 s on 16 cores
HEAT_KW = dict(x_start=0, x_end=1, nx=1025, a=1, init_cond=init_cond, rhs=rhs, t_start=0, t_stop=2)
SOLVER_KW = dict(cf_iter=1, cycle_type='V', nested_iteration=True, tol=1e-10)


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get('hbm_gbs', 6650.0), 'measured'
    return 6650.0, 'fallback'


def ncu_traffic(sweep, intervals, m0=None):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `sweep` from the committed ncu --set full capture
    (profiles/ncu_traffic.json, written by scripts/ncu_summary.py traffic), scaled to this run's number of coarse
    intervals (traffic is proportional to it); None if the sweep was not captured."""
    path = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    if not os.path.exists(path):
        return None
    with open(path) as f:
        cap = json.load(f)
    ent = cap['sweeps'].get(sweep)
    if ent is None or (m0 is not None and cap.get('coarsening') not in (None, m0)):
        return None                    # not captured, or captured for intervals of another length
    return (ent['dram_read_bytes'] + ent['dram_write_bytes']) * intervals / cap['intervals']


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled under load (B200_PROFILING.md).  The timed region of the default
    run lasts tens of milliseconds, shorter than nvidia-smi's sampling period, so the sampler runs from the warm-up to
    the end of the per-sweep timing (the GPU is busy throughout); samples that fall inside the timed region are
    preferred when there are any, and `window` says which were used."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '20', '-i', str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(',')]))

    def mark_timed(self, t0, t1):
        self.t0, self.t1 = t0, t1

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.05)
        self.proc.terminate()
        inside = [r for ts, r in self.rows if self.t0 is not None and self.t0 <= ts <= self.t1]
        window = 'timed region'
        rows = inside
        if not rows:
            rows, window = [r for _, r in self.rows], 'warm-up + timed region + per-sweep timing (all under load)'
        sm, mx, reasons = [], None, set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm), 'window': window}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on a bounded sample
# ------------------------------------------------------------------------------------------------
def hierarchy(make, nt, coarsening, t_stop=None):
    """[fine, coarse, ...]: level l+1 lives on every coarsening[l]-th point of level l (t_interval = t[::m])."""
    kw = {k: v for k, v in HEAT_KW.items() if k not in ('t_start', 't_stop')}
    fine_kw = dict(HEAT_KW)
    if t_stop is not None:
        fine_kw['t_stop'] = t_stop
    levels = [make(nt=nt, **fine_kw)]
    for m in coarsening:
        levels.append(make(t_interval=levels[-1].t[::m], **kw))
    return levels


def describe(name, nt, coarsening):
    return (f'{name}: heat_1d backward Euler nx=1025 nt={nt} on [0,2], {len(coarsening) + 1}-level, coarsening '
            f'{"x".join(str(m) for m in coarsening)}, FCF V-cycle, nested iteration, tol 1e-10')


def cpu_cores():
    from oracle import mgrit_oracle_mp as OM
    return max(1, min(OM.host_cores(), 64))


def cpu_sample(cores=None, solver='spsolve'):
    """Full MGRIT solve of the same problem at reduced nt with the reference's arithmetic (per-point Python loop +
    SciPy SuperLU per step), time-parallel over `cores` worker processes the way the reference spreads time points
    over mpi4py ranks.  Returns (DOF/s, seconds, iterations, cores, description)."""
    from oracle import mgrit_oracle as O
    from oracle import mgrit_oracle_mp as OM
    cores = cpu_cores() if cores is None else cores
    nt_sample, coarsening = CPU_SAMPLE_MP if cores >= 4 else CPU_SAMPLE
    if os.environ.get('MGRIT_BENCH_CPU_SAMPLE_NT'):          # tests/test_bench_cpu.py: a sample that runs in a second
        nt_sample = int(os.environ['MGRIT_BENCH_CPU_SAMPLE_NT'])
    nt_full = WORKLOADS['cfg5'][0]
    t_stop = HEAT_KW['t_stop'] * (nt_sample - 1) / (nt_full - 1)          # same dt as the workload
    t0 = time.time()
    prob = hierarchy(lambda **kw: O.Heat1DOracle(solver=solver, **kw), nt_sample, coarsening, t_stop=t_stop)
    if cores > 1:
        mg = OM.ParallelMgritOracle(prob, workers=cores, **SOLVER_KW)
    else:
        mg = O.MgritOracle(prob, **SOLVER_KW)
    info = mg.solve()
    sec = time.time() - t0
    if cores > 1:
        mg.close()
    its = len(info['conv'])
    what = (f'the first {nt_sample} of the workload\'s {nt_full} time points (t in [0, {t_stop:.6g}], same dt, {len(coarsening) + 1}'
            f'-level, coarsening {"x".join(str(m) for m in coarsening)}, FCF V-cycle, nested iteration, tol 1e-10; {its} '
            f'iterations): per-point Python loop + SciPy SuperLU per step as in the reference, time-parallel over {cores} '
            f'worker process(es) like its mpi4py time ranks; DOF/s is linear in nt')
    return 1023 * nt_sample / sec, sec, its, cores, what


def reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    nt, coarsening = WORKLOADS[args.workload]
    if args.coarsening:
        coarsening = tuple(int(x) for x in args.coarsening.split(','))
    vals, times = [], []
    warm = min(args.warmup, 1)                # a CPU loop has nothing to warm beyond imports; keeps the run in minutes
    cores, sample = 1, ''
    for k in range(warm + args.steps):
        dofs, sec, its, cores, sample = cpu_sample()
        if k >= warm:
            vals.append(dofs)
            times.append(sec)
    sec = float(np.mean(times))
    val = float(np.mean(vals))
    line = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': describe(args.workload, nt, coarsening), 'sample': sample},
            'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def gpu_arm(args):
    import torch
    import torch.distributed as dist
    import pymgrit_b200 as P
    from pymgrit_b200 import _lib

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    nt, coarsening = WORKLOADS[args.workload]
    if args.coarsening:
        coarsening = tuple(int(x) for x in args.coarsening.split(','))
    ndof = 1023

    def make_problem():
        return hierarchy(P.Heat1D, nt, coarsening)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- device-resident timing: tables in HBM, time setup sweeps (nested iteration) + iterations ----
    solver = P.Mgrit(problem=make_problem(), logging_lvl=logging.WARNING, **SOLVER_KW)
    info = None
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        solver.restart()
        info = solver.solve()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = solver.launches
    wall0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        solver.restart()
        info = solver.solve()
    ev1.record()
    barrier()
    sampler.mark_timed(wall0, time.perf_counter())
    ms_step = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    launches = (solver.launches - launches0)
    iters = len(info['conv'])
    value = ndof * nt / (ms_step * 1e-3)

    # ---- end to end through the public API from host inputs, result back on the host ----
    e2e_ms = []
    h2d = d2h = 0
    for k in range(max(1, min(args.steps, 3)) + 1):
        barrier()
        t0 = time.perf_counter()
        s2 = P.Mgrit(problem=make_problem(), logging_lvl=logging.WARNING, **SOLVER_KW)
        inf2 = s2.solve()
        last = s2.u[0][-1].get_values()             # solution at the final time (device -> host)
        conv_host = np.array(inf2['conv'])
        torch.cuda.synchronize()
        dt_ms = (time.perf_counter() - t0) * 1e3
        if k > 0:
            e2e_ms.append(max_over_ranks(dt_ms))
        h2d = s2.h2d_bytes
        d2h = last.nbytes + 8 * len(conv_host)
        del s2
    e2e_ms = float(np.mean(e2e_ms))

    # ---- per-kernel roofline on level 0 (CUDA events around single launches on the solved state) ----
    kernels = solver.time_level0_sweeps(repeats=5)
    clocks = sampler.stop() if rank == 0 else None
    hbm, which = peaks()
    dom = max([k for k in kernels if k['bound'] == 'hbm'], key=lambda k: k['share_ms'])
    roofline = {'bound': 'hbm', 'kernel': dom['name'], 'achieved': dom['gbs'], 'peak': hbm, 'unit': 'GB/s',
                'frac': dom['gbs'] / hbm, 'traffic': ncu_traffic(dom['name'], dom['intervals'], coarsening[0]), 'peak_source': which,
                'frac_of_nominal_7700': dom['gbs'] / 7700.0,          # HGX B200 data-sheet figure (B200_PROFILING.md)
                'frac_of_nominal_8000': dom['gbs'] / 8000.0,          # the ~8 TB/s BASELINE.json's north star quotes
                'algorithmic_bytes': dom['algorithmic_bytes'], 'ms': dom['ms']}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            # in a fresh process: the CPU port forks its time-parallel workers, which a process that has initialised CUDA
            # and started threads should not do
            r = subprocess.run([sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--steps', '1',
                                '--warmup', '0', '--workload', args.workload], stdout=subprocess.PIPE,
                               stderr=subprocess.DEVNULL, text=True, timeout=900)
            lines = [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
            if lines:
                ref = json.loads(lines[-1])
                cpu = dict(ref['cpu_baseline'])
                cpu['sample'] += f" ({ref['ms_per_step'] / 1e3:.1f} s)"
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64',
                'data': 'synthetic',
                'config': {'workload': describe(args.workload, nt, coarsening), 'iterations': iters,
                           'conv': [float(c) for c in info['conv']],
                           'l2': ('working set (level 0: %.1f GB) is far larger than the 126 MB L2' if ndof * nt * 8 > 1e9
                                  else 'working set (level 0: %.2f GB) is of the order of the 126 MB L2 and is not flushed '
                                       'between steps: parity-size workload, not the headline') % (ndof * nt * 8 / 1e9),
                           'parallelism': f'time-slab x{world}'},
                'time_to_tolerance_s': ms_step * 1e-3, 'clocks': clocks, 'gpu_launches': launches // args.steps,
                'e2e': {'value': ndof * nt / (e2e_ms * 1e-3), 'unit': UNIT, 'ms_per_step': e2e_ms,
                        'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h)},
                'roofline': roofline, 'kernels': kernels, 'cpu_baseline': cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--workload', default='cfg5', choices=sorted(WORKLOADS))
    ap.add_argument('--coarsening', default='', help='comma-separated coarsening factors per level (overrides the workload)')
    ap.add_argument('--no-cpu', action='store_true')
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: whatever libraries print to file descriptor 1 meanwhile (NCCL's version
    # banner when NCCL_DEBUG is set in the environment, torchrun notices) is sent to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, 'w', buffering=1)
    if args.impl == 'reference':
        reference_arm(args)
    else:
        gpu_arm(args)
    sys.stdout.flush()


if __name__ == '__main__':
    main()
