"""CPU tests of the host-side table logic (no GPU): how a NumPy right-hand-side callable becomes device tables
(pymgrit_b200/core/rhs_tables.py), the dt classes of a time grid, the reduction of the two-point BDF steps to the
coefficients the kernel takes (checked against the oracle's step), and that the applications can be built and
deep-copied without a device (core/simple_setup_problem.py:34-41 deep-copies them)."""
import copy

import numpy as np
import pytest
from scipy.linalg import solve_banded

import cases as C
from oracle import mgrit_oracle as O
from pymgrit_b200.core import device_level as dl
from pymgrit_b200.core import partition
from pymgrit_b200.core.rhs_tables import RhsSplit

X = np.linspace(0, 1, 67)[1:-1]
T = np.linspace(0, 2, 41)


def _direct(rhs, t):
    return np.stack([np.broadcast_to(np.asarray(rhs(X, tt), dtype=float), X.shape) for tt in t])


@pytest.mark.parametrize('rhs, kind, terms', [
    (lambda x, t: x * 0, 'zero', 0),
    (C.heat_rhs, 'separable', 1),
    (C.heat_rhs_rank2, 'separable', 2),
    (lambda x, t: x * (1 - x), 'separable', 1),                  # no time dependence: the callable ignores t
    (C.heat_rhs_nonsep, 'dense', None),
])
def test_split_kinds_reproduce_the_callable(rhs, kind, terms):
    sp = RhsSplit(rhs, X).analyse(T)
    assert sp.kind == kind
    if kind == 'separable':
        assert sp.basis.shape == (terms, len(X))
        want = _direct(rhs, T)
        got = sp.coefficients(T) @ sp.basis
        assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))
        assert sp.reproduces(T[::3])
        scale = np.linspace(0.5, 1.5, len(T))
        np.testing.assert_allclose(sp.coefficients(T, scale=scale), sp.coefficients(T) * scale[:, None], rtol=1e-15)
    if kind == 'dense':
        np.testing.assert_array_equal(sp.dense(T), _direct(rhs, T))


def test_callable_that_does_not_broadcast_falls_back_to_point_evaluation():
    def rhs(x, t):
        return np.sin(np.pi * x) * np.cos(float(t))           # float(t) fails for an array of times
    sp = RhsSplit(rhs, X).analyse(T)
    assert sp.kind == 'separable'
    want = _direct(rhs, T)
    assert np.max(np.abs(sp.coefficients(T) @ sp.basis - want)) <= 1e-12


def _pulse(t):
    return np.where((0.42 < t) & (t < 0.44), 1.0, 0.0)


@pytest.mark.parametrize('nt', [1025, (1 << 17) + 1])           # every (x, t) compared / sampled check points, all t
@pytest.mark.parametrize('rhs, terms', [
    (lambda x, t: np.sin(np.pi * x) * _pulse(t), 1),                                       # nothing at the sampled times
    (lambda x, t: np.sin(np.pi * x) * np.cos(t) + np.exp(-100 * (x - .3) ** 2) * _pulse(t), 2),   # a second shape, briefly
])
def test_forcing_active_only_between_the_sampled_times_is_kept(rhs, terms, nt):
    """The reference evaluates rhs(x, t_stop) in every step (heat/heat_1d.py:214): a pulse that the dozen sampled times
    miss must still reach the tables (it used to be classified 'zero' / rank 1)."""
    t = np.linspace(0, 1, nt)
    sp = RhsSplit(rhs, X).analyse(t)
    assert sp.kind == 'separable' and sp.basis.shape[0] == terms
    pick = np.unique(np.concatenate([np.flatnonzero(_pulse(t) > 0)[::max(1, nt // 2000)], np.arange(0, nt, max(1, nt // 50))]))
    want = _direct(rhs, t[pick])
    got = sp.coefficients(t)[pick] @ sp.basis
    assert np.max(np.abs(got - want)) <= 1e-12


def test_forcing_that_fits_no_short_split_becomes_dense():
    def rhs(x, t):                                              # a travelling bump switched on for a while: rank > 4
        return np.exp(-200 * (x - t) ** 2) * _pulse(t)
    t = np.linspace(0, 1, 1025)
    sp = RhsSplit(rhs, X).analyse(t)
    assert sp.kind == 'dense'


@pytest.fixture
def fresh_pool():
    """Stop the table threads after the test: later tests fork worker processes (oracle/mgrit_oracle_mp.py)."""
    from pymgrit_b200.core import rhs_tables
    yield
    if rhs_tables._POOL is not None:
        rhs_tables._POOL[0].shutdown(wait=True)
        rhs_tables._POOL = None


def test_long_grid_goes_through_the_thread_pool_unchanged(fresh_pool):
    t = np.linspace(0, 2, (1 << 16) + 1)
    sp = RhsSplit(C.heat_rhs_rank2, X).analyse(t)
    scale = np.full(len(t), t[1] - t[0])
    out = np.empty((len(t), 2))
    got = sp.coefficients(t, scale=scale, out=out)
    assert got is out
    inv = np.linalg.inv(sp.basis[:, sp.sel])
    want = (np.asarray(C.heat_rhs_rank2(X[sp.sel][None, :], t[:, None])) @ inv) * scale[:, None]
    np.testing.assert_allclose(got, want, rtol=1e-14, atol=1e-300)
    pick = [0, 12345, len(t) - 1]
    direct = _direct(C.heat_rhs_rank2, t[pick]) * scale[0]
    assert np.max(np.abs(got[pick] @ sp.basis - direct)) <= 1e-12 * np.max(np.abs(direct))


def test_dt_classes():
    d, idx = dl.dt_classes(np.linspace(0, 2, 2 ** 10 + 1))
    assert idx is None and d[0] == 2.0 ** -9
    t = np.array([0.0, 0.5, 1.0, 1.25, 1.75, 2.0])
    d, idx = dl.dt_classes(t)
    np.testing.assert_array_equal(d, [0.25, 0.5])
    np.testing.assert_array_equal(idx, [0, 1, 1, 0, 1, 0])
    d2, idx2 = dl.dt_classes(t, t[1:] - t[:-1])
    np.testing.assert_array_equal(d, d2)
    np.testing.assert_array_equal(idx, idx2)
    assert dl.dt_classes(np.array([3.0]))[1] is None


@pytest.mark.parametrize('size', [2, 3, 5])
def test_slab_bounds_from_composed_masks_match_a_search(size):
    """Slab ends are level-0 indices of coarsest-grid points: composing the C-point masks (core/partition.py) must give
    what searching the fine grid for the coarsest times gives, also for irregular index hierarchies."""
    t0 = np.linspace(0, 5, 65)
    t1 = t0[np.array([0, 3, 10, 12, 14, 17, 23, 27, 33, 34, 55, 57, 59, 61, 63, 64])]
    t2 = t1[::2]
    t3 = t2[::2]
    for grids in ([t0, t1, t2, t3], [t0, t0[::4], t0[::16]], [t0, t0[::2]]):
        if len(grids[-1]) - 1 < size:
            with pytest.raises(Exception):                # fewer coarsest intervals than ranks: refused
                partition.Partition(grids, size, 0)
            continue
        parts = [partition.Partition(grids, size, r) for r in range(size)]
        coarse_idx = np.flatnonzero(np.isin(grids[0], grids[-1]))
        for p in parts:
            assert p.window[1] == len(t0) - 1 or p.window[1] in coarse_idx
        owned = np.concatenate([np.arange(p.window[0], p.window[1] + 1) for p in parts])
        np.testing.assert_array_equal(owned, np.arange(len(t0)))


def _thomas(r, b):
    """(I + r tridiag(-1, 2, -1))^-1 b."""
    n = len(b)
    ab = np.zeros((3, n))
    ab[0, 1:] = -r
    ab[1] = 1 + 2 * r
    ab[2, :-1] = -r
    return solve_banded((1, 1), ab, b)


@pytest.mark.parametrize('method', ['BDF1', 'BDF2'])
@pytest.mark.parametrize('dt', [0.05, 0.013])
def test_two_point_coefficients_reproduce_the_reference_step(method, dt):
    """heat_1d_2pts_bdf{1,2}.py step == S(r1)(a1 first + b1 second + c1 b(t_stop)), S(r2)(a2 second + b2 tmp1 + c2 b(t_stop + dtau))
    with the coefficients pymgrit_b200/heat/heat_1d_2pts_bdf*.py hands to the kernel."""
    import pymgrit_b200 as P
    kw = dict(x_start=0, x_end=1, nx=34, a=0.7, dtau=0.004, init_cond=C.heat_init, rhs=C.heat_rhs_rank2, t_start=0,
              t_stop=1, nt=21)
    app = {'BDF1': P.Heat1DBDF1, 'BDF2': P.Heat1DBDF2}[method](**kw)
    orc = O.Heat1D2PtsOracle(method=method, **kw)
    rng = np.random.default_rng(5)
    u = rng.standard_normal((2, app.nx))
    t_start = 0.3
    r1, r2, a1, b1, a2, b2, c1, c2 = app._coefficients(dt)
    tmp1 = _thomas(r1, a1 * u[0] + b1 * u[1] + c1 * C.heat_rhs_rank2(app.x, t_start + dt))
    tmp2 = _thomas(r2, a2 * u[1] + b2 * tmp1 + c2 * C.heat_rhs_rank2(app.x, t_start + dt + app.dtau))
    want = orc.phi(u, t_start, t_start + dt)
    assert np.max(np.abs(np.stack([tmp1, tmp2]) - want)) <= 1e-12 * np.max(np.abs(want))


def test_applications_build_and_deepcopy_without_a_device():
    import pymgrit_b200 as P
    heat = P.Heat1D(nx=17, t_start=0, t_stop=1, nt=9, **C.HEAT)
    prob = P.simple_setup_problem(heat, level=3, coarsening=2)
    assert [len(p.t) for p in prob] == [9, 5, 3]
    assert prob[1]._rhs_split.kind == 'separable'
    np.testing.assert_array_equal(prob[2].vector_t_start.get_values(), C.heat_init(heat.x))
    bdf = P.Heat1DBDF2(x_start=0, x_end=1, nx=11, a=1, dtau=0.1, t_start=0, t_stop=1, nt=5)
    twin = copy.deepcopy(bdf)
    assert twin.vector_template.size == 9 and twin.vector_template.shape == (2, 9)
    assert twin.vector_t_start._lazy_second is not None          # the second start value is made on the device later
    with pytest.raises(Exception):                               # a right-hand side the pair kernels cannot take
        P.Heat1DBDF1(x_start=0, x_end=1, nx=11, a=1, dtau=0.1, rhs=C.heat_rhs_nonsep, t_start=0, t_stop=1, nt=5)
    tr = P.GridTransferHeat1D()
    tr.check(P.Heat1D(nx=17, t_start=0, t_stop=1, nt=9, **C.HEAT), P.Heat1D(nx=9, t_start=0, t_stop=1, nt=5, **C.HEAT))
    with pytest.raises(Exception):
        tr.check(P.Heat1D(nx=17, t_start=0, t_stop=1, nt=9, **C.HEAT), P.Heat1D(nx=10, t_start=0, t_stop=1, nt=5, **C.HEAT))


@pytest.mark.parametrize('nx', [65, 50])
def test_sine_space_tables_restate_the_heat1d_step(nx):
    """Heat1D with its rows in sine space (csrc/phi.cuh Heat1DSine): the host tables (eigenvalues, reciprocals, their
    thread-transposed layout) pushed through a NumPy restatement of the kernel's arithmetic give the reference's step
    (heat_1d.py:198-217, here the oracle's sparse solve)."""
    import pymgrit_b200 as P
    from oracle import mgrit_oracle as O
    from scipy.fft import dst
    kw = dict(x_start=0, x_end=1, nx=nx, a=0.7, init_cond=C.heat_init, rhs=C.heat_rhs_rank2)
    t = np.linspace(0, 1, 17)
    app = P.Heat1D(t_interval=t, **kw)
    assert app.can_sine() and app.as_sine().kind == 7 and app.kind == 1
    orc = O.Heat1DOracle(t_interval=t, **kw)
    n = nx - 2
    threads, chunk = 32, 3 if n > 32 else 1
    for grid in (t, t ** 1.5):                                   # uniform: reciprocals; non-uniform: division per step
        tab, dt = app.sine_host_tables(grid, threads, chunk)
        lam = tab['diag'][0].T.reshape(-1)[:n]
        inv = tab['diag'][1].T.reshape(-1)[:n]
        assert np.all(tab['diag'][:, :, :].transpose(0, 2, 1).reshape(2, -1)[:, n:] == 0.0)
        uniform = tab['ndt'] == 1
        assert tab['sconst'][0, 1] == (1.0 if uniform else 0.0)
        u = np.asarray(orc.u0, dtype=float)
        for i in (1, 7, 16):
            xh = dst(u, type=1, norm='ortho')
            bh = dst(np.asarray(kw['rhs'](app.x, grid[i])) * dt[i], type=1, norm='ortho')
            if uniform:
                new = (xh + bh) * inv
            else:
                row = tab['sconst'][tab['dtidx'][i]]
                assert row[0] == dt[i]
                new = (xh + bh) / (1 + row[0] * lam)
            got = dst(new, type=1, norm='ortho')
            ref = orc.phi(u, grid[i - 1], grid[i])
            assert np.max(np.abs(got - ref)) <= 1e-13 * np.max(np.abs(ref))
            u = ref


def test_native_host_passes_are_bit_identical_to_numpy():
    """csrc/host_tables.cu (no device): the step sizes of a time grid with their extrema, and the scaled transpose of the
    time factors of a right-hand side, against the NumPy expressions they replace -- every element the same IEEE
    operation, for uniform and non-uniform grids, one and several threads, strided input rows."""
    import ctypes as C
    from pymgrit_b200 import _lib
    lib = _lib.lib()
    rng = np.random.default_rng(3)
    for n, threads in ((1, 4), (2, 4), (1000, 1), (70001, 3), (300007, 8)):
        t = np.cumsum(rng.random(n) + 1e-3) if n % 2 else np.linspace(0.0, 2.0, n)
        dt = np.full(n, np.nan)
        lo, hi = C.c_double(-1.0), C.c_double(-1.0)
        assert lib.mgb_host_time_steps(t.ctypes.data, n, dt.ctypes.data, C.byref(lo), C.byref(hi), threads) == 0
        ref = np.zeros(n)
        ref[1:] = t[1:] - t[:-1]
        assert np.array_equal(dt, ref)
        if n > 1:
            assert lo.value == ref[1:].min() and hi.value == ref[1:].max()
        else:
            assert lo.value == 0.0 and hi.value == 0.0
        for q in (1, 2, 3):
            wide = rng.standard_normal((q, n + 5))
            src = wide[:, 2:2 + n]                                   # rows with a stride
            scale = rng.standard_normal(n)
            out = np.full((n, q), np.nan)
            assert lib.mgb_host_scale_rows(src.ctypes.data, src.strides[0] // 8, q, n, scale.ctypes.data, out.ctypes.data,
                                           threads) == 0
            assert np.array_equal(out, (src * scale).T)
            assert lib.mgb_host_scale_rows(src.ctypes.data, src.strides[0] // 8, q, n, None, out.ctypes.data, threads) == 0
            assert np.array_equal(out, src.T)
    # and through the host helpers that call them on long grids
    from pymgrit_b200.core import device_level as dl
    t = np.linspace(0, 2, dl._PAR_MIN + 77) ** 1.1
    dt, lo, hi = dl.time_steps(t)
    assert np.array_equal(dt[1:], t[1:] - t[:-1]) and dt[0] == 0.0 and lo == dt[1:].min() and hi == dt[1:].max()


@pytest.mark.parametrize('a,b,n', [(0, 2, 2 ** 20 + 1), (0, 5, 2 ** 17 + 3), (-1.3, 7.77, 300001), (0.1, 0.1000001, 2 ** 18),
                                   (3, 1, 2 ** 17 + 1), (0, 2, 1000), (0.0, 0.0, 2 ** 17)])
def test_long_time_grids_are_numpy_linspace_bit_for_bit(a, b, n):
    """core/application.py linspace(): long grids are filled by native threads (mgb_host_affine_ramp: i * step + start with
    the product and the sum rounded separately); the reference's grids are np.linspace (core/application.py:60), and every
    dt-dependent constant of the engine is keyed on the exact differences of the grid."""
    from pymgrit_b200.core.application import linspace
    assert np.array_equal(linspace(a, b, n), np.linspace(a, b, n))
