"""Instantiate a tests/cases.py case against the product package (GPU)."""
import logging

import numpy as np

import cases as C


def b200_apps():
    import pymgrit_b200 as P
    apps = {'heat1d': P.Heat1D, 'advection1d': P.Advection1D, 'dahlquist': P.Dahlquist, 'brusselator': P.Brusselator}
    if hasattr(P, 'Heat2D'):
        apps['heat2d'] = P.Heat2D
    if hasattr(P, 'AllenCahn'):
        apps['allencahn'] = P.AllenCahn
    if hasattr(P, 'Heat1DBDF2'):
        apps['heat1d2pts'] = lambda method, **kw: {'BDF1': P.Heat1DBDF1, 'BDF2': P.Heat1DBDF2}[method](**kw)
    return apps


def b200_problem(case):
    apps = b200_apps()
    grids = C.case_time_grids(case)
    return [apps[case['app']](t_interval=t, **C.level_app_kw(case, l)) for l, t in enumerate(grids)]


def b200_transfer(case):
    import pymgrit_b200 as P
    if 'transfer' not in case:
        return None
    return [{'space': P.GridTransferHeat1D, 'copy': P.GridTransferCopy}[k]() for k in case['transfer']]


def run_b200(name, **extra):
    import pymgrit_b200 as P
    case = C.CASES[name]
    kw = dict(case['solver'])
    kw.update(extra)
    if 'transfer' in case:
        kw['transfer'] = b200_transfer(case)
    if 'at_k' in case:
        solver = P.AtMgrit(problem=b200_problem(case), k=case['at_k'], logging_lvl=logging.WARNING, **kw)
    else:
        solver = P.Mgrit(problem=b200_problem(case), logging_lvl=logging.WARNING, **kw)
    info = solver.solve()
    return solver, info


def solution_rows(solver, idx=None):
    """Level-0 solution as host arrays: (rows at idx, per-point 2-norms of all points)."""
    lv = solver._lv[0]
    u = lv.values()                                    # [points, *vector shape], node values
    norms = np.sqrt(np.sum(u.reshape(len(u), -1) ** 2, axis=1))
    rows = u if idx is None else u[idx]
    return rows, norms
