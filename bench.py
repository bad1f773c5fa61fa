#!/usr/bin/env python
"""bench.py -- time-to-tolerance and space-time DOF/s of the MGRIT hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg5|cfg1|cfg2|cfg3|cfg4]

Default workload (BASELINE.json configs[4], the one the metric's target is quoted on; fits one GPU):
    heat_1d backward Euler, nx = 1025 (1023 dofs), nt = 2^20 + 1 on t in [0, 2], FCF-relaxation V-cycles, 3 levels
    (coarsening 64 x 16, chosen by the builder: BASELINE leaves it open), nested iteration, tol 1e-10; strong
    scaling over the time ranks (one process per GPU).
The other BASELINE configs are selected with --workload (same JSON line; profiles/r02_bench_cfg*.json).
One "step" = one complete solve: setup (incl. nested iteration) + MGRIT iterations until conv < 1e-10.
Prints ONE JSON line (rank 0).
  e2e    the headline: the public API from host NumPy inputs -- application objects, Mgrit(), solve(), the solution at
         the final time copied back to the host -- i.e. the reference's time_setup + time_solve (mgrit.py:142, 240, 602,
         638).  `cold_ms` is the very first call of the process.
  value  the same solve with every table already in HBM (restart() + solve()): the device-side part of e2e.
  parity after the timed region: sampled level-0 points (both sides of every slab boundary, the last point, random
         ones) are checked against the CPU oracle's Phi (u[i] = Phi(u[i-1]) to 1e-10 at F-points, within the reported
         residual at C-points) and the residual history against the committed single-GPU history; failure -> exit 1.
`--impl reference` times the reference's CPU algorithm (per-point Python loop + SciPy SuperLU: the unmodified reference
when /root/reference/src is present, else the oracle port, time-parallel over the host cores) on a bounded sample.
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

UNIT = 'DOF/s'


def rhs(x, t):
    return -np.sin(np.pi * x) * (np.sin(t) - np.pi ** 2 * np.cos(t))


def init_cond(x):
    return np.sin(np.pi * x)


def rhs_2d(x, y, t):          # docs/source/usage/parallelism.rst:121-125
    return -np.sin(x) * np.sin(y) * (np.sin(t) - 2 * np.cos(t))


def adv_init(x):              # examples/example_advection.py:30-31 (the reference's default initial condition)
    return np.exp(-x ** 2)


HEAT_KW = dict(x_start=0, x_end=1, nx=1025, a=1, init_cond=init_cond, rhs=rhs, t_start=0, t_stop=2)
SOLVER_KW = dict(cf_iter=1, cycle_type='V', nested_iteration=True, tol=1e-10)

# name: application ('heat1d' ...), its constructor arguments without the time grid, (t_start, t_stop, nt), coarsening
# factor per level transition, solver arguments, spatial dofs per time point, text.  cfg1-cfg4 are BASELINE.json
# configs[0..3] verbatim; for cfg4 and cfg5 BASELINE leaves the hierarchy to the builder: cfg5 (64, 16) ->
# 16385 and 1025 coarse points converges in 3 FCF V-cycles and is the fastest of the 24 hierarchies tried on 1 and on 8
# GPUs (scripts/hierarchy_sweep.py, profiles/r01o_hierarchy_sweep.txt); cfg4 (advection, coarsening 2): 4 levels -- MGRIT
# contracts slowly on this hyperbolic problem and deep hierarchies stall (10 levels: 129 V-cycles to 1e-10, 4 levels: 10;
# scripts/cfg4_sweep.py, profiles/r02_cfg4_sweep.txt).  Since the coarsest level is solved time-parallel in Fourier space
# (csrc/fourier.cu) a long coarsest level costs a few milliseconds instead of a chain of dependent solves, and the
# two-level hierarchy -- 4 V-cycles -- is the fastest of the 11 depth / cycle combinations (profiles/r02r_cfg4_sweep.txt).
_HEAT1D = {k: v for k, v in HEAT_KW.items() if k not in ('t_start', 't_stop')}
WORKLOADS = {
    'cfg5': dict(app='heat1d', kw=_HEAT1D, t=(0, 2, 2 ** 20 + 1), coarsening=(64, 16), solver=SOLVER_KW, ndof=1023,
                 text='heat_1d backward Euler nx=1025 nt=1048577 on [0,2]'),
    'cfg2': dict(app='heat1d', kw=_HEAT1D, t=(0, 2, 16385), coarsening=(4, 4), solver=SOLVER_KW, ndof=1023,
                 text='heat_1d backward Euler nx=1025 nt=16385 on [0,2]'),
    'cfg1': dict(app='dahlquist', kw=dict(), t=(0, 5, 101), coarsening=(2,), solver=dict(tol=1e-10), ndof=1,
                 text='Dahlquist nt=101 on [0,5]'),
    'cfg3': dict(app='heat2d', kw=dict(x_start=0, x_end=1, y_start=0, y_end=1, nx=512, ny=512, a=1, rhs=rhs_2d),
                 t=(0, 5, 4097), coarsening=(8, 8, 8), solver=dict(cycle_type='F', tol=1e-10), ndof=512 * 512,
                 text='heat_2d backward Euler 512x512 nt=4097 on [0,5]'),
    'cfg4': dict(app='advection1d', kw=dict(c=1, x_start=-1, x_end=1, nx=4096), t=(0, 2, 65537), coarsening=(2,),
                 solver=dict(cf_iter=1, cycle_type='V', nested_iteration=True, tol=1e-10), ndof=4095,
                 text='advection 1D upwind nx=4096 nt=65537 on [0,2]'),
}
# Bounded CPU samples: a time window [t_start, t_start + (nt_sample - 1) dt] of the SAME problem (same dt, same coarsening
# factors, same cycle), so that the work per space-time DOF is the workload's.  (one core, several cores)
CPU_SAMPLES = {
    'cfg5': ((2049, (64, 16)), (32769, (64, 16))),
    'cfg2': ((1025, (4, 4)), (4097, (4, 4))),
    'cfg1': ((101, (2,)), (101, (2,))),
    'cfg4': ((129, (2,)), (513, (2,))),
    'cfg3': None,                     # see cpu_sample(): a 512 x 512 SuperLU solve takes seconds, a spatially reduced grid is timed
}


def metric_name(wl):
    return f'MGRIT space-time DOF/s ({WORKLOADS[wl]["text"].split(" on ")[0]}, solve to 1e-10)'


def describe(name, coarsening=None):
    w = WORKLOADS[name]
    co = w['coarsening'] if coarsening is None else coarsening
    s = w['solver']
    return (f'{name}: {w["text"]}, {len(co) + 1}-level, coarsening {"x".join(str(m) for m in co)}, '
            f'{"FCF" if s.get("cf_iter", 1) == 1 else "F(CF)^%d" % s["cf_iter"]} {s.get("cycle_type", "V")}-cycle, '
            f'{"nested iteration, " if s.get("nested_iteration", True) else ""}tol {s.get("tol", 1e-7):g}')


def workload_grid(name):
    """(nt, coarsening factors) of a workload."""
    return WORKLOADS[name]['t'][2], WORKLOADS[name]['coarsening']


def hierarchy(make, nt, coarsening, t_stop=None):
    """heat_1d levels [fine, coarse, ...] (kept for the tests and scripts): level l+1 lives on every coarsening[l]-th
    point of level l."""
    return build_levels(make, _HEAT1D, (HEAT_KW['t_start'], HEAT_KW['t_stop'] if t_stop is None else t_stop, nt), coarsening)


def build_levels(make, kw, t, coarsening):
    """[fine, coarse, ...]: level l+1 lives on every coarsening[l]-th point of level l (t_interval = t[::m])."""
    levels = [make(t_start=t[0], t_stop=t[1], nt=t[2], **kw)]
    for m in coarsening:
        levels.append(make(t_interval=levels[-1].t[::m], **kw))
    return levels


def app_class(ns, app):
    """The application class of namespace `ns` (pymgrit_b200, pymgrit, or the oracle adapter)."""
    return getattr(ns, {'heat1d': 'Heat1D', 'heat2d': 'Heat2D', 'advection1d': 'Advection1D', 'dahlquist': 'Dahlquist'}[app])


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
                                          '-lms', os.environ.get('MGRIT_BENCH_SMI_MS', '20'), '-i', str(self.index)],
                                         stdout=subprocess.PIPE, text=True)
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
# reference arm / cpu baseline: the reference's CPU algorithm on a bounded sample
# ------------------------------------------------------------------------------------------------
class _OracleNS:
    """The oracle's problem classes under the reference's names (oracle/ is the CPU restatement of the reference: test
    and baseline infrastructure, imported only by the CPU legs and the parity check of this file)."""

    def __init__(self, solver='spsolve'):
        from oracle import mgrit_oracle as O
        self.Heat1D = lambda **kw: O.Heat1DOracle(solver=solver, **kw)
        self.Advection1D = lambda **kw: O.Advection1DOracle(solver=solver, **kw)
        self.Heat2D = O.Heat2DOracle
        self.Dahlquist = O.DahlquistOracle


def cpu_cores():
    from oracle import mgrit_oracle_mp as OM
    return max(1, min(OM.host_cores(), 64))


def cpu_model():
    try:
        with open('/proc/cpuinfo') as f:
            for line in f:
                if line.startswith('model name'):
                    return line.split(':', 1)[1].strip()
    except OSError:
        pass
    return 'unknown'


def reference_src():
    """The unmodified reference package, importable with the single-rank mpi4py / empty matplotlib stubs of
    oracle/stubs; None on the GPU box, where /root/reference does not exist."""
    src = '/root/reference/src'
    if os.environ.get('MGRIT_BENCH_NO_REFERENCE') or not os.path.isdir(os.path.join(src, 'pymgrit')):
        return None
    return src


def cpu_sample(wl='cfg5', cores=None, kind=None):
    """Full MGRIT solve of a time window of the workload with the reference's arithmetic (per-point Python loop + SciPy
    SuperLU per step).  kind 'reference': the unmodified pymgrit.Mgrit on one core (the stub mpi4py has one rank);
    kind 'port': the oracle, time-parallel over `cores` worker processes the way the reference spreads time points over
    mpi4py ranks.  Returns a dict (DOF/s, seconds, iterations, cores, kind, description, window fraction)."""
    w = WORKLOADS[wl]
    if kind is None:
        kind = 'reference' if reference_src() and os.environ.get('MGRIT_BENCH_CPU_KIND') != 'port' else 'port'
    cores = (1 if kind == 'reference' else cpu_cores()) if cores is None else cores
    kw, ndof = dict(w['kw']), w['ndof']
    if wl == 'cfg3':
        # a 512 x 512 sparse direct solve takes 4.6 s: hours per iteration.  Timed instead: the same
        # problem on a 64 x 64 grid and the first 65 of the 4097 time points, 3 levels of coarsening 8 x 8 -- SuperLU's
        # cost per unknown grows with the grid, so this favours the CPU figure.
        kw.update(nx=64, ny=64)
        ndof = 64 * 64
        nt_sample, coarsening = 65, (8, 8)
    else:
        nt_sample, coarsening = CPU_SAMPLES[wl][1 if cores >= 4 else 0]
    if os.environ.get('MGRIT_BENCH_CPU_SAMPLE_NT'):          # tests/test_bench_cpu.py: a sample that runs in a second
        nt_sample = int(os.environ['MGRIT_BENCH_CPU_SAMPLE_NT'])
    t0_, t1_, nt_full = w['t']
    nt_sample = min(nt_sample, nt_full)
    t_stop = t0_ + (t1_ - t0_) * (nt_sample - 1) / (nt_full - 1)          # same dt as the workload
    solver_kw = dict(w['solver'])
    start = time.time()
    if kind == 'reference':
        stubs = os.path.join(ROOT, 'oracle', 'stubs')
        for p in (reference_src(), stubs):
            if p not in sys.path:
                sys.path.insert(0, p)
        import pymgrit                                           # the unmodified reference
        prob = build_levels(app_class(pymgrit, w['app']), kw, (t0_, t_stop, nt_sample), coarsening)
        mg = pymgrit.Mgrit(problem=prob, logging_lvl=logging.WARNING, **solver_kw)
        info = mg.solve()
        how = 'the unmodified reference (pymgrit.Mgrit from /root/reference/src, single-rank mpi4py stub)'
    else:
        from oracle import mgrit_oracle as O
        from oracle import mgrit_oracle_mp as OM
        prob = build_levels(app_class(_OracleNS(), w['app']), kw, (t0_, t_stop, nt_sample), coarsening)
        if cores > 1:
            mg = OM.ParallelMgritOracle(prob, workers=cores, **solver_kw)
        else:
            mg = O.MgritOracle(prob, **solver_kw)
        info = mg.solve()
        if cores > 1:
            mg.close()
        how = (f'the oracle port of the reference (per-point Python loop + SciPy SuperLU per step), time-parallel over '
               f'{cores} worker process(es) like its mpi4py time ranks')
    sec = time.time() - start
    its = len(info['conv'])
    what = (f'the first {nt_sample} of the workload\'s {nt_full} time points (t in [{t0_:g}, {t_stop:.6g}], same dt, '
            f'{len(coarsening) + 1}-level, coarsening {"x".join(str(m) for m in coarsening)}'
            + (', spatial grid reduced to 64x64' if wl == 'cfg3' else '') +
            f'; {its} iterations to {solver_kw.get("tol", 1e-7):g}): {how}; DOF/s is linear in nt')
    return dict(value=ndof * nt_sample / sec, seconds=sec, iterations=its, cores=cores, kind=kind, sample=what,
                window_fraction=nt_sample / nt_full, dofs_per_s_per_iteration=ndof * nt_sample * its / sec)


def reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    wl = args.workload
    vals, times, last = [], [], None
    warm = min(args.warmup, 1)                # a CPU loop has nothing to warm beyond imports; keeps the run in minutes
    for k in range(warm + args.steps):
        last = cpu_sample(wl)
        if k >= warm:
            vals.append(last['value'])
            times.append(last['seconds'])
    sec = float(np.mean(times))
    val = float(np.mean(vals))
    cpu = {'value': val, 'unit': UNIT, 'cores': last['cores'], 'kind': last['kind'], 'sample': last['sample'],
           'cpu_model': cpu_model(), 'host_cores': cpu_cores(), 'window_fraction': last['window_fraction'],
           'iterations': last['iterations'], 'dofs_per_s_per_iteration': last['dofs_per_s_per_iteration']}
    line = {'impl': 'reference', 'metric': metric_name(wl), 'value': val, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': describe(wl)},
            'cpu_baseline': cpu,
            'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# parity of the measured run (after the timed region)
# ------------------------------------------------------------------------------------------------
def expected_history(wl):
    path = os.path.join(ROOT, 'tests', 'golden', 'bench_conv.json')
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(wl)
    return None


def parity_check(solver, info, wl, world, rank, nsample=48):
    """Size-independent parity of the solve that was just timed (tests/test_gpu_fullsize.py does the same under pytest):
    at sampled level-0 points i -- both sides of every slab boundary, the last points, random ones -- u[i] must be
    Phi_oracle(u[i-1]) (Phi evaluated on the CPU with the reference's arithmetic): to 1e-10 relative at F-points, within
    the reported residual at C-points; and the residual history must equal the committed single-GPU history within the
    absolute floor of SURVEY.md 8c.  Every rank checks the points it owns; rank 0 gets the verdicts."""
    import torch
    import torch.distributed as dist
    w = WORKLOADS[wl]
    lv = solver._lv[0]
    t_loc = np.asarray(solver.t[0])
    nloc = len(t_loc)
    rng = np.random.default_rng(20261017 + rank)
    cset = set(int(c) for c in (lv.cpts if lv.cpts is not None else []))
    cand = {1, 2, nloc - 1, nloc - 2} | set(int(i) for i in rng.integers(1, nloc, nsample // max(world, 1) + 4))
    if lv.cpts is not None and len(lv.cpts) > 1:
        cand |= {int(lv.cpts[1]), int(lv.cpts[-1])}
    sample = sorted(i for i in cand if 1 <= i < nloc)
    if w['app'] == 'heat2d':
        sample = sample[:2] + sample[-2:]                   # a 512 x 512 sparse solve takes seconds on the CPU
    from oracle import mgrit_oracle as O
    try:
        ns = _OracleNS(solver='c')
        O.c_lib()
    except Exception:
        ns = _OracleNS()
    orc = app_class(ns, w['app'])(t_interval=t_loc, **w['kw'])
    conv_last = float(info['conv'][-1]) if len(info['conv']) else 0.0
    worst_f, worst_c, bad = 0.0, 0.0, []
    for i in sample:
        pair = lv.values(idx=np.array([i - 1, i]))
        want = np.asarray(orc.phi(pair[0].reshape(np.shape(orc.u0)), t_loc[i - 1], t_loc[i])).reshape(pair[1].shape)
        scale = max(float(np.max(np.abs(want))), 1e-300)
        if i in cset:
            defect = float(np.linalg.norm((want - pair[1]).ravel()))
            worst_c = max(worst_c, defect)
            if defect > conv_last * (1 + 1e-6) + 1e-10 * scale * np.sqrt(want.size):
                bad.append(('C', i, defect))
        else:
            err = float(np.max(np.abs(want - pair[1]))) / scale
            worst_f = max(worst_f, err)
            if err > 1e-10:
                bad.append(('F', i, err))
    mine = dict(rank=rank, points=len(sample), max_rel_f=worst_f, max_defect_c=worst_c, bad=bad[:4])
    parts = [mine]
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, mine)
    if rank != 0:
        return None
    out = {'checked': True, 'points': int(sum(p['points'] for p in parts)),
           'max_rel': max(p['max_rel_f'] for p in parts), 'max_c_defect': max(p['max_defect_c'] for p in parts),
           'reported_residual': conv_last, 'failures': [b for p in parts for b in p['bad']],
           'what': 'u[i] vs Phi_oracle(u[i-1]) at sampled level-0 points of every time rank (slab boundaries, ends, random): '
                   'relative error at F-points (<= 1e-10), 2-norm defect at C-points (<= reported residual)'}
    ref = expected_history(wl)
    if ref is not None:
        conv = np.asarray(info['conv'], dtype=float)
        refc = np.asarray(ref['conv'], dtype=float)
        floor = 1e-10 * max(refc[0], ref['scale'])
        out['conv_abs_diff'] = float(np.max(np.abs(conv - refc))) if len(conv) == len(refc) else None
        out['conv_floor'] = floor
        out['iterations_match'] = len(conv) == len(refc)
        if out['conv_abs_diff'] is None or out['conv_abs_diff'] > floor:
            out['failures'].append(('conv', list(map(float, conv)), list(map(float, refc))))
    out['ok'] = not out['failures']
    return out


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def gpu_arm(args):
    import torch
    import torch.distributed as dist
    import pymgrit_b200 as P

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    wl = args.workload
    w = WORKLOADS[wl]
    coarsening = w['coarsening']
    if args.coarsening:
        coarsening = tuple(int(x) for x in args.coarsening.split(','))
    nt, ndof = w['t'][2], w['ndof']
    solver_kw = dict(w['solver'])
    if args.max_iter:
        solver_kw['max_iter'] = args.max_iter

    def make_problem():
        return build_levels(app_class(P, w['app']), w['kw'], w['t'], coarsening)

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

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---- end to end through the public API from host inputs, result back on the host: the headline ----
    def e2e_once():
        barrier()
        t0 = time.perf_counter()
        s2 = P.Mgrit(problem=make_problem(), logging_lvl=logging.WARNING, **solver_kw)
        inf2 = s2.solve()
        last = s2.u[0][-1].get_values()             # solution at the final time (device -> host)
        conv_host = np.array(inf2['conv'])
        torch.cuda.synchronize()
        ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        return ms, s2.h2d_bytes, last.nbytes + 8 * len(conv_host)

    cold_ms, h2d, d2h = e2e_once()                  # the very first call of the process
    for _ in range(max(0, min(args.warmup, 3) - 1)):
        e2e_once()
    e2e_all = []
    for _ in range(max(3, min(args.steps, 9))):
        ms, h2d, d2h = e2e_once()
        e2e_all.append(ms)
    e2e_ms = float(np.median(e2e_all))             # the host side of a 15 ms call is at the mercy of the box: median

    # ---- device-resident timing: tables in HBM, time setup sweeps (nested iteration) + iterations ----
    solver = P.Mgrit(problem=make_problem(), logging_lvl=logging.WARNING, **solver_kw)
    info = None
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

    # ---- parity of what was just timed ----
    parity = None
    if not args.no_parity:
        parity = parity_check(solver, info, wl, world, rank)

    # ---- per-kernel roofline on level 0 (CUDA events around single launches on the solved state) ----
    hbm, which = peaks()
    kernels = solver.time_level0_sweeps(repeats=5, iterations=iters, hbm_gbs=hbm)
    clocks = sampler.stop() if rank == 0 else None
    roofline = None
    live = [k for k in kernels if k['bound'] == 'hbm' and k['share_ms'] > 0]
    if live:
        dom = max(live, key=lambda k: k['share_ms'])
        roofline = {'bound': 'hbm', 'kernel': dom['name'], 'achieved': dom['gbs'], 'peak': hbm, 'unit': 'GB/s',
                    'frac': dom['gbs'] / hbm, 'traffic': ncu_traffic(dom['name'], dom['intervals'], coarsening[0]),
                    'peak_source': which,
                    'frac_of_nominal_7700': dom['gbs'] / 7700.0,      # HGX B200 data-sheet figure (B200_PROFILING.md)
                    'frac_of_nominal_8000': dom['gbs'] / 8000.0,      # the ~8 TB/s BASELINE.json's north star quotes
                    'algorithmic_bytes': dom['algorithmic_bytes'], 'ms': dom['ms'],
                    'share_of_solve': dom['share_ms'] / ms_step,
                    'fp64_peak_source': 'self-measured issue rate (scripts/micro/fp64_latency.cu); MEASURED_PEAKS.json has '
                                        'no FP64 entry'}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            # in a fresh process: the CPU port forks its time-parallel workers, which a process that has initialised CUDA
            # and started threads should not do
            r = subprocess.run([sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--steps', '1',
                                '--warmup', '0', '--workload', wl], stdout=subprocess.PIPE,
                               stderr=subprocess.DEVNULL, text=True, timeout=900)
            lines = [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
            if lines:
                ref = json.loads(lines[-1])
                cpu = dict(ref['cpu_baseline'])
                cpu['seconds'] = ref['ms_per_step'] / 1e3
        rep = getattr(solver.problem[0], 'kind', None)
        line = {'metric': metric_name(wl), 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong',
                'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                'config': {'workload': describe(wl, coarsening), 'iterations': iters,
                           'conv': [float(c) for c in info['conv']],
                           'l2': ('working set (level 0: %.1f GB) is far larger than the 126 MB L2' if ndof * nt * 8 > 1e9
                                  else 'working set (level 0: %.2f GB) is of the order of the 126 MB L2 and is not flushed '
                                       'between steps: parity-size workload, not the headline') % (ndof * nt * 8 / 1e9),
                           'parallelism': f'time-slab x{world}',
                           'rows': 'sine coefficients (diagonal Phi)' if rep == P._lib.APP_HEAT1D_SINE else 'as the application stores them'},
                'time_to_tolerance_s': e2e_ms * 1e-3, 'time_to_tolerance_device_s': ms_step * 1e-3, 'clocks': clocks,
                'gpu_launches': launches // args.steps,
                'e2e': {'value': ndof * nt / (e2e_ms * 1e-3), 'unit': UNIT, 'ms_per_step': e2e_ms, 'cold_ms': cold_ms, 'samples_ms': [round(v, 3) for v in e2e_all],
                        'statistic': 'median of samples_ms',
                        'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                        'd2h': 'the solution at the final time point and the residual history'},
                'roofline': roofline, 'kernels': kernels, 'parity': parity, 'cpu_baseline': cpu}
        print(json.dumps(line), flush=True)
    ok = True
    if parity is not None and rank == 0:
        ok = bool(parity['ok'])
    if world > 1:
        flag = torch.tensor([1 if ok else 0], device='cuda')
        dist.broadcast(flag, 0)
        ok = bool(flag.item())
        dist.destroy_process_group()
    if not ok:
        sys.stdout.flush()
        sys.stderr.write('bench.py: parity check of the timed run FAILED\n')
        sys.exit(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--workload', default='cfg5', choices=sorted(WORKLOADS))
    ap.add_argument('--coarsening', default='', help='comma-separated coarsening factors per level (overrides the workload)')
    ap.add_argument('--max-iter', type=int, default=0, help='cap on the MGRIT iterations (default: the solver\'s 100)')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-parity', action='store_true')
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
