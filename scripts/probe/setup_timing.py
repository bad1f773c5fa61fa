"""Wall-clock sections of Mgrit.__init__ on the headline workload (monkeypatched timers, no profiler)."""
import logging, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
import pymgrit_b200 as P
from pymgrit_b200.core import mgrit as M, device_level as DL, partition as PT, rhs_tables as RT
from pymgrit_b200.heat import heat_1d as H

acc = {}


def timed(obj, name, label=None, sync=False):
    fn = getattr(obj, name)
    label = label or f'{getattr(obj, "__name__", obj)}.{name}'

    def inner(*a, **kw):
        t0 = time.perf_counter()
        r = fn(*a, **kw)
        if sync:
            torch.cuda.synchronize()
        acc[label] = acc.get(label, 0.0) + (time.perf_counter() - t0) * 1e3
        return r
    setattr(obj, name, inner)


timed(PT, 'c_point_masks')
timed(PT.Partition, '__init__', 'Partition')
timed(DL.DeviceLevel, '__init__', 'DeviceLevel.__init__')
timed(DL.DeviceLevel, 'finish_tables', 'DeviceLevel.finish_tables')
timed(H.Heat1D, 'level_tables', 'Heat1D.level_tables')
timed(RT.RhsSplit, 'coefficients', 'RhsSplit.coefficients')
timed(DL, 'dt_classes')
timed(DL, 'step_const_table')
timed(M.Mgrit, 'nested_iteration', 'nested_iteration (host side)')
timed(M.Mgrit, '_init_levels', '_init_levels')
timed(torch.cuda, 'synchronize', 'cuda.synchronize')
timed(torch, 'zeros', 'torch.zeros')
nt, co = bench.workload_grid('cfg5')
for k in range(5):
    acc.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    prob = bench.hierarchy(P.Heat1D, nt, co)
    t1 = time.perf_counter()
    s = P.Mgrit(problem=prob, logging_lvl=logging.WARNING, **bench.SOLVER_KW)
    t2 = time.perf_counter()
    print(f'run {k}: hierarchy {1e3 * (t1 - t0):.1f} ms, Mgrit() {1e3 * (t2 - t1):.1f} ms: ' +
          ', '.join(f'{k_} {v:.1f}' for k_, v in sorted(acc.items(), key=lambda kv: -kv[1])))
    del s, prob
