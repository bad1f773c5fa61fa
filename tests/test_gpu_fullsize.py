"""BASELINE.json configs at FULL size on the GPU, checked through size-independent properties.

The CPU oracle cannot solve these sizes in seconds, so the checks are the ones the domain offers:
  * fixed point: the converged MGRIT iterate is the sequential time-stepping solution, so at sampled time points
    u[i] must equal Phi_oracle(u[i-1]) -- Phi evaluated on the CPU by the oracle (the reference's arithmetic) --
    exactly (to rounding) at F-points and within the final residual at C-points;
  * the residual the solver reports equals the temporal 2-norm of those C-point defects;
  * a prefix of the solution equals the oracle's sequential time stepping from the initial condition;
  * the residual history contracts and ends below tol in the number of iterations a reduced-size oracle run needs.
Tolerance: float64, 1e-10 relative (BASELINE.json north_star), written at each assert.
"""
import logging

import numpy as np
import pytest

import cases as C

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def P():
    import pymgrit_b200 as P
    return P


def _hierarchy(make, t0, coarsening):
    levels = [make(t0)]
    for m in coarsening:
        levels.append(make(levels[-1].t[::m]))
    return levels


def _rows(solver, idx):
    lv = solver._lv[0]
    return lv.values(idx=np.asarray(idx))


def _check_fixed_point(solver, oracle, info, sample, tol_rel=1e-10):
    """u[i] == Phi_oracle(u[i-1]) at the sampled points: to rounding at F-points, within the last residual at
    C-points.  Returns the C-point defects seen."""
    lv = solver._lv[0]
    cset = set(int(c) for c in lv.cpts)
    t = solver.t[0]
    scale = None
    worst_f = 0.0
    for i in sample:
        pair = _rows(solver, [i - 1, i])
        want = oracle.phi(pair[0].reshape(oracle.u0.shape), t[i - 1], t[i]).reshape(pair[1].shape)
        scale = max(np.max(np.abs(want)), 1e-300)
        defect = np.linalg.norm((want - pair[1]).ravel())
        if i in cset:
            # one C-point's defect is bounded by the temporal 2-norm of all of them
            assert defect <= info['conv'][-1] * (1 + 1e-6) + tol_rel * scale * np.sqrt(want.size), (i, defect)
        else:
            worst_f = max(worst_f, np.max(np.abs(want - pair[1])) / scale)
            assert np.max(np.abs(want - pair[1])) <= tol_rel * scale, (i, np.max(np.abs(want - pair[1])), scale)
    return worst_f


def test_cfg5_heat1d_full_size(P):
    """configs[4] on one GPU: heat_1d nx=1025, nt=2^20+1, FCF V-cycles, the hierarchy bench.py uses."""
    import bench
    from oracle import mgrit_oracle as O
    nt, coarsening = bench.workload_grid('cfg5')
    solver = P.Mgrit(problem=bench.hierarchy(P.Heat1D, nt, coarsening), logging_lvl=logging.WARNING, **bench.SOLVER_KW)
    info = solver.solve()
    conv = info['conv']
    assert conv[-1] < 1e-10 and len(conv) <= 4
    assert np.all(conv[1:] < 0.1 * conv[:-1])                       # FCF V-cycle contracts by > 10x per iteration here
    orc = O.Heat1DOracle(solver='c', nt=nt, **bench.HEAT_KW)
    rng = np.random.default_rng(5)
    m0 = coarsening[0]
    sample = sorted(set(int(i) for i in rng.integers(1, nt, 48)) | {m0, nt - 1, nt - 2, 7 * m0, 1})
    _check_fixed_point(solver, orc, info, sample)
    # prefix against sequential time stepping with the oracle
    npre = 129
    got = _rows(solver, np.arange(npre))
    u = orc.u0.copy()
    for i in range(1, npre):
        u = orc.phi(u, orc.t[i - 1], orc.t[i])
        assert np.max(np.abs(got[i] - u)) <= 1e-10 * np.max(np.abs(u)) + 2 * conv[-1], i
    # the reported residual is the 2-norm over time of the C-point defects (recomputed on the host for a slice)
    sq = solver.compute_residual()[:len(solver._lv[0].cpts)].cpu().numpy()
    assert abs(np.sqrt(np.sum(sq[1:])) - conv[-1]) <= 1e-10 * max(conv[0], 1.0)
    k = 3
    c = int(solver._lv[0].cpts[k])
    pair = _rows(solver, [c - 1, c])
    r = orc.phi(pair[0], orc.t[c - 1], orc.t[c]) - pair[1]
    assert abs(np.sqrt(sq[k]) - np.linalg.norm(r)) <= 1e-10 * np.linalg.norm(pair[1])


def test_cfg4_advection_full_size(P):
    """configs[3]: advection nx=4096, nt=65537 on [0, 2], coarsening 2, 10 levels, nested iteration.  MGRIT contracts
    slowly on this hyperbolic problem (the reference's own iteration does too: same algorithm), so the run is cut after 8
    iterations and checked through what holds after ANY number of iterations: every F-point is Phi of its predecessor,
    the C-point defects are what the solver reports, the residual history decreases, and -- the exactness property of
    MGRIT with FCF-relaxation -- the first k*m points after k iterations are the sequential time-stepping solution."""
    from oracle import mgrit_oracle as O
    kw = dict(c=1, x_start=-1, x_end=1, nx=4096)
    t0 = np.linspace(0, 2, 65537)
    iters = 8
    solver = P.Mgrit(problem=_hierarchy(lambda t: P.Advection1D(t_interval=t, **kw), t0, [2] * 9),
                     logging_lvl=logging.WARNING, tol=1e-10, cf_iter=1, nested_iteration=True, max_iter=iters)
    info = solver.solve()
    conv = info['conv']
    assert len(conv) == iters and np.all(conv[1:] < conv[:-1]), conv
    orc = O.Advection1DOracle(solver='c', t_interval=t0, **kw)
    rng = np.random.default_rng(4)
    sample = sorted(set(int(i) for i in rng.integers(1, len(t0), 24)) | {1, 2, len(t0) - 1})
    _check_fixed_point(solver, orc, info, sample)
    npre = iters * 2 + 1                               # k * m points are exact after k iterations
    got = _rows(solver, np.arange(npre))
    u = orc.u0.copy()
    for i in range(1, npre):
        u = orc.phi(u, t0[i - 1], t0[i])
        assert np.max(np.abs(got[i] - u)) <= 1e-10 * np.max(np.abs(u)), i
    # the reported residual is the temporal 2-norm of the C-point defects: recompute a few of them with the oracle
    sq = solver.compute_residual()[:len(solver._lv[0].cpts)].cpu().numpy()
    assert abs(np.sqrt(np.sum(sq[1:])) - conv[-1]) <= 1e-10 * conv[0]
    for k in (1, 77, len(sq) - 1):
        c = int(solver._lv[0].cpts[k])
        pair = _rows(solver, [c - 1, c])
        r = orc.phi(pair[0], t0[c - 1], t0[c]) - pair[1]
        assert abs(np.sqrt(sq[k]) - np.linalg.norm(r)) <= 1e-10 * np.linalg.norm(pair[1])


def test_cfg4_advection_two_level_to_tolerance(P):
    """configs[3] as bench.py runs it: two levels, coarsening 2, the coarsest level (32769 points) solved time-parallel in
    Fourier space (csrc/fourier.cu).  To the tolerance 1e-10 (4 V-cycles), then: every F-point is Phi of its predecessor,
    the C-point defects are what the solver reports, the first k*m points are the time-stepping solution, and the same
    run with the chain of dependent solves on the coarsest level (MGB_ADVECTION_FOURIER=0) gives the same history."""
    import os
    from oracle import mgrit_oracle as O
    kw = dict(c=1, x_start=-1, x_end=1, nx=4096)
    t0 = np.linspace(0, 2, 65537)
    out = {}
    for mode in ('1', '0'):
        os.environ['MGB_ADVECTION_FOURIER'] = mode
        try:
            solver = P.Mgrit(problem=_hierarchy(lambda t: P.Advection1D(t_interval=t, **kw), t0, [2]),
                             logging_lvl=logging.WARNING, tol=1e-10, cf_iter=1, nested_iteration=True, max_iter=20)
            assert bool(solver._spectral) == (mode == '1')
            out[mode] = (solver, solver.solve())
        finally:
            os.environ.pop('MGB_ADVECTION_FOURIER', None)
    solver, info = out['1']
    conv = info['conv']
    assert len(conv) == 4 and conv[-1] < 1e-10 and np.all(conv[1:] < conv[:-1]), conv
    conv_chain = out['0'][1]['conv']
    assert len(conv_chain) == len(conv)
    assert np.max(np.abs(conv_chain - conv)) <= 1e-10 * max(conv[0], 1.0), (conv, conv_chain)
    orc = O.Advection1DOracle(solver='c', t_interval=t0, **kw)
    rng = np.random.default_rng(5)
    sample = sorted(set(int(i) for i in rng.integers(1, len(t0), 24)) | {1, 2, len(t0) - 1})
    _check_fixed_point(solver, orc, info, sample)
    npre = len(conv) * 2 + 1
    got = _rows(solver, np.arange(npre))
    u = orc.u0.copy()
    for i in range(1, npre):
        u = orc.phi(u, t0[i - 1], t0[i])
        assert np.max(np.abs(got[i] - u)) <= 1e-10 * np.max(np.abs(u)), i
    # both runs agree at sampled points to the accuracy of the iteration
    a, b = _rows(solver, sample), _rows(out['0'][0], sample)
    assert np.max(np.abs(a - b)) <= 1e-9 * np.max(np.abs(b))


def test_cfg3_heat2d_full_size(P):
    """configs[2]: heat_2d backward Euler 512 x 512, nt=4097 on [0, 5], 4-level F-cycle, coarsening 8."""
    from oracle import mgrit_oracle as O
    kw = dict(x_start=0, x_end=1, y_start=0, y_end=1, nx=512, ny=512, a=1, rhs=C.heat2d_rhs)
    t0 = np.linspace(0, 5, 4097)
    solver = P.Mgrit(problem=_hierarchy(lambda t: P.Heat2D(t_interval=t, **kw), t0, [8] * 3),
                     logging_lvl=logging.WARNING, tol=1e-10, cycle_type='F')
    info = solver.solve()
    conv = info['conv']
    assert conv[-1] < 1e-10, conv
    assert np.all(conv[1:] < conv[:-1])
    orc = O.Heat2DOracle(t_interval=t0, **kw)
    # sparse direct solves of the 512 x 512 system take seconds each on the CPU: three F-points, one C-point
    _check_fixed_point(solver, orc, info, [1, 2051, 4095, 4096])
