"""Helpers that instantiate a tests/cases.py case against the CPU oracle (test infrastructure)."""
import os

import numpy as np

import cases as C
from oracle import mgrit_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

ORACLE_APPS = {'heat1d': O.Heat1DOracle, 'heat2d': O.Heat2DOracle, 'advection1d': O.Advection1DOracle,
               'dahlquist': O.DahlquistOracle, 'brusselator': O.BrusselatorOracle, 'heat1d2pts': O.Heat1D2PtsOracle,
               'allencahn': O.AllenCahnOracle}


def oracle_problem(case, solver=None):
    grids = C.case_time_grids(case)
    out = []
    for l, t in enumerate(grids):
        kw = C.level_app_kw(case, l)
        if solver is not None and case['app'] in ('heat1d', 'advection1d'):
            kw['solver'] = solver
        out.append(ORACLE_APPS[case['app']](t_interval=t, **kw))
    return out


def oracle_transfer(case):
    if 'transfer' not in case:
        return None
    return [{'space': O.Heat1DSpaceTransfer, 'copy': O.CopyTransfer}[k]() for k in case['transfer']]


def run_oracle(name, solver=None):
    case = C.CASES[name]
    if 'at_k' in case:
        mg = O.AtMgritOracle(oracle_problem(case, solver), k=case['at_k'], **case['solver'])
        return mg, mg.solve()
    mg = O.MgritOracle(oracle_problem(case, solver), transfer=oracle_transfer(case), **case['solver'])
    info = mg.solve()
    return mg, info


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'))


def assert_history_close(conv, conv_ref, scale, rtol=1e-10):
    """same length; |diff| <= rtol * max(conv_ref[0], scale)."""
    conv, conv_ref = np.asarray(conv), np.asarray(conv_ref)
    assert len(conv) == len(conv_ref), (conv, conv_ref)
    if len(conv_ref):
        floor = rtol * max(conv_ref[0], scale)
        assert np.all(np.abs(conv - conv_ref) <= floor), (conv, conv_ref, floor)


def assert_solution_close(rows, norms, gold, rtol=1e-10):
    """rows: values at gold['u_rows_idx']; norms: per-point 2-norms of all level-0 points."""
    ref_rows, ref_norms = gold['u_rows'], gold['u_norms']
    scale = max(np.max(np.abs(ref_rows)), 1e-300)
    assert rows.shape == ref_rows.shape
    assert np.max(np.abs(rows - ref_rows)) <= rtol * scale
    assert np.max(np.abs(np.asarray(norms) - ref_norms)) <= rtol * max(np.max(ref_norms), 1e-300)
