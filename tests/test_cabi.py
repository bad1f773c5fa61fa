"""CPU-side checks of the C ABI: the library loads and exports every symbol include/mgrit_b200.h declares, the host
helpers behave, and a sweep without a CUDA device fails loudly (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from pymgrit_b200 import build, _lib
    build.build()
    return _lib.lib()


def test_exports_every_declared_symbol(lib):
    from pymgrit_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'mgrit_b200.h')).read()
    declared = set(re.findall(r'\b(mgb_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations found'
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.mgb_abi_version() == _lib.ABI_VERSION == 10


def test_struct_layout_matches_header(lib):
    from pymgrit_b200 import _lib
    # offsets implied by include/mgrit_b200.h on LP64
    assert C.sizeof(_lib.MgbLevel) % 8 == 0
    assert _lib.MgbLevel.u_dev.offset == 16
    assert _lib.MgbLevel.p.offset + 8 * 8 == _lib.MgbLevel.ip.offset


def test_team_shape_and_consts(lib):
    from pymgrit_b200 import _lib
    t, e = C.c_int32(), C.c_int32()
    assert lib.mgb_team_shape(_lib.APP_HEAT1D, 1023, C.byref(t), C.byref(e)) == 0
    assert (t.value, e.value) == (32, 33)
    assert lib.mgb_team_shape(_lib.APP_HEAT1D, 4095, C.byref(t), C.byref(e)) == 0
    assert t.value * e.value >= 4096
    assert lib.mgb_team_shape(_lib.APP_HEAT1D, 10 ** 6, C.byref(t), C.byref(e)) == 2
    assert b'no kernel shape' in lib.mgb_last_error()
    cw = lib.mgb_step_consts_width(_lib.APP_HEAT1D, 32, 33)
    row = np.zeros(cw)
    r = 128.0
    assert lib.mgb_heat1d_step_consts(r, 1023, 32, 33, row.ctypes.data_as(_lib.c_double_p)) == 0
    beta = row[0]
    assert abs(beta + 1 / beta - (2 + 1 / r)) < 1e-14          # root of b^2 - (2 + 1/r) b + 1
    assert abs(row[1] - beta / r) < 1e-18 and abs(row[3] - beta ** 11) < 1e-15
    assert lib.mgb_heat1d_step_consts(-1.0, 1023, 32, 33, row.ctypes.data_as(_lib.c_double_p)) == 1


def test_two_point_shape_and_widths(lib):
    """MGB_APP_HEAT1D_2PTS: chunk = 2 h + 1, a thread owns h elements of each time point of the pair."""
    from pymgrit_b200 import _lib
    t, e = C.c_int32(), C.c_int32()
    assert lib.mgb_team_shape(_lib.APP_HEAT1D_2PTS, 999, C.byref(t), C.byref(e)) == 0
    assert e.value % 2 == 1 and t.value * ((e.value - 1) // 2) >= 999
    assert lib.mgb_team_shape(_lib.APP_HEAT1D_2PTS, 9, C.byref(t), C.byref(e)) == 0
    assert (t.value, e.value) == (32, 3)
    half = lib.mgb_heat1d_2pts_half_width(32, 19)
    assert half == lib.mgb_step_consts_width(_lib.APP_HEAT1D, 32, 9)
    assert lib.mgb_step_consts_width(_lib.APP_HEAT1D_2PTS, 32, 19) == 2 * half + 8
    assert lib.mgb_team_shape(_lib.APP_HEAT1D_2PTS, 10 ** 6, C.byref(t), C.byref(e)) == 2


def test_heat1d_constants_solve_the_system(lib):
    """Host-side check of the Toeplitz factorisation the kernel uses: emulate Phi with the table in numpy."""
    from pymgrit_b200 import _lib
    n, T, E = 38, 32, 3
    for r in (0.01, 2.0, 128.0, 2048.0):
        cw = lib.mgb_step_consts_width(_lib.APP_HEAT1D, T, E)
        row = np.zeros(cw)
        assert lib.mgb_heat1d_step_consts(r, n, T, E, row.ctypes.data_as(_lib.c_double_p)) == 0
        beta, cs, kappa = row[0], row[1], row[2]
        rng = np.random.default_rng(0)
        b = rng.standard_normal(n)
        y = np.zeros(n)
        acc = 0.0
        for i in range(n):
            acc = beta * acc + b[i] * cs
            y[i] = acc
        z = np.zeros(n)
        acc = 0.0
        for i in range(n - 1, -1, -1):
            acc = beta * acc + y[i]
            z[i] = acc
        i = np.arange(n)
        h = (beta ** (i + 1.0) - beta ** (2.0 * n + 1 - i)) / (1 - beta * beta)
        x = z - kappa * z[0] * h
        A = np.diag(np.full(n, 1 + 2 * r)) + np.diag(np.full(n - 1, -r), 1) + np.diag(np.full(n - 1, -r), -1)
        assert np.max(np.abs(A @ x - b)) <= 1e-11 * np.max(np.abs(b)) * (1 + 4 * r)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA device present')
    from pymgrit_b200 import _lib
    out = np.zeros(4)
    rc = lib.mgb_vec_sumsq(4, out.ctypes.data, out.ctypes.data, None)
    assert rc == 3 and b'CUDA' in lib.mgb_last_error()
    import pymgrit_b200 as P
    with pytest.raises(Exception):
        P.Mgrit(problem=[P.Dahlquist(t_start=0, t_stop=5, nt=11)]).solve()


@pytest.mark.parametrize('name', ['heat2d_bc', 'heat2d_example', 'heat2d_cfg3_small', 'heat2d_cn', 'heat2d_cn_3lvl',
                                  'heat2d_fe'])
def test_heat2d_sine_space_tables_reproduce_the_oracle(lib, name):
    """Host-side check of the Heat2D design (pymgrit_b200/heat/heat_2d.py, csrc/phi.cuh): the row layout, symbol row,
    right-hand-side factors and boundary coupling the host builds, pushed through a numpy restatement of the kernel's
    arithmetic, give the oracle's theta-method step (sparse direct solve; BE, CN and FE cases) to rounding."""
    import cases as CS
    import pymgrit_b200 as P
    from oracle import mgrit_oracle as O
    case = CS.CASES[name]
    kw = case['app_kw']
    t = CS.case_time_grids(case)[0]
    app, orc = P.Heat2D(t_interval=t, **kw), O.Heat2DOracle(t_interval=t, **kw)
    fam = app.family()
    nx, ny = app.nx, app.ny
    assert fam.pitch == fam.nsys * fam.tile and fam.boff >= fam.nint
    assert np.allclose(fam.sx @ fam.sx, np.eye(nx - 2), atol=1e-13)

    def to_rows(nodes):
        r = np.zeros(fam.pitch)
        r[:fam.nint] = (fam.sx @ nodes[1:-1, 1:-1] @ fam.sy).reshape(-1)
        r[fam.boff:fam.boff + len(fam.bnodes)] = nodes.reshape(-1)[fam.bnodes]
        return r

    def from_rows(r):
        nodes = np.zeros((nx, ny))
        nodes[1:-1, 1:-1] = fam.sx @ r[:fam.nint].reshape(nx - 2, ny - 2) @ fam.sy
        nodes.reshape(-1)[fam.bnodes] = r[fam.boff:fam.boff + len(fam.bnodes)]
        return nodes

    assert np.array_equal(app.vector_t_start.get_values(), orc.u0)
    fields = []
    if fam.split.kind == 'separable':
        for b in fam.split.basis:
            f = np.zeros((nx, ny))
            f[1:-1, 1:-1] = b.reshape(nx - 2, ny - 2)
            fields.append(f)
    if fam.coupling is not None:
        fields.append(fam.coupling)
    rx = [to_rows(f) for f in fields]
    dt = np.zeros(len(t))
    dt[1:] = np.diff(t)
    th = app.theta
    coef = fam.split.coefficients(t) if fam.split.kind == 'separable' else None
    cols = ([(th * coef + (1 - th) * np.concatenate([coef[:1], coef[:-1]])) * dt[:, None]] if coef is not None else []) + \
           ([dt[:, None]] if fam.coupling is not None else [])
    rt = np.concatenate(cols, axis=1)
    assert rt.shape[1] == len(rx) <= 3
    row, u = to_rows(orc.u0), orc.u0
    for i in range(1, 5):
        x = row * (1 - (1 - th) * dt[i] * fam.sig) + sum(rt[i, k] * rx[k] for k in range(len(rx)))
        x = x / (1 + th * dt[i] * fam.sig)
        x[fam.boff:] = fam.sig[fam.boff:]              # boundary tiles: the Dirichlet values
        x[fam.nint:fam.boff] = 0.0
        row, u = x, orc.phi(u, t[i - 1], t[i])
        assert np.max(np.abs(from_rows(row) - u)) <= 1e-12 * np.max(np.abs(u))
        assert abs(np.linalg.norm(row) - np.linalg.norm(u)) <= 1e-12 * np.linalg.norm(u)      # Parseval


@pytest.mark.parametrize('name', ['heat1d_bdf2_nonuniform', 'heat1d_bdf1_small', 'heat1d_bdf2_small_f'])
def test_two_point_tables_reproduce_the_oracle(lib, name):
    """Host-side check of the two-point design (pymgrit_b200/heat/heat_1d_2pts.py, csrc/phi.cuh Heat1D2Pts): the
    step-constant rows (two Heat1D blocks + a1 b1 a2 b2 skip1), the dt classes, the thread-transposed spatial factors
    and the [npts][2 q] time factors the host builds, pushed through a numpy restatement of the kernel's arithmetic
    (Toeplitz recurrences + Sherman-Morrison, as in test_heat1d_constants_solve_the_system), give the oracle's BDF step
    on every level of the case."""
    import cases as CS
    import pymgrit_b200 as P
    from pymgrit_b200 import _lib
    from oracle import mgrit_oracle as O
    case = CS.CASES[name]
    grids = CS.case_time_grids(case)

    def solve(row, b, n):
        beta, cs, kappa = row[0], row[1], row[2]
        y, acc = np.zeros(n), 0.0
        for i in range(n):
            acc = beta * acc + b[i] * cs
            y[i] = acc
        z, acc = np.zeros(n), 0.0
        for i in range(n - 1, -1, -1):
            acc = beta * acc + y[i]
            z[i] = acc
        i = np.arange(n)
        h = (beta ** (i + 1.0) - beta ** (2.0 * n + 1 - i)) / (1 - beta * beta)
        return z - kappa * z[0] * h

    for lvl, t in enumerate(grids):
        kw = CS.level_app_kw(case, lvl)
        method = kw.pop('method')
        app = {'BDF1': P.Heat1DBDF1, 'BDF2': P.Heat1DBDF2}[method](t_interval=t, **kw)
        orc = O.Heat1D2PtsOracle(t_interval=t, method=method, **kw)
        T, E, h = app._shape()
        n = app.nx
        assert E == 2 * h + 1 and T * h >= n and app.row_pitch() == T * E
        tab = app.level_tables(t, T, E)
        half = lib.mgb_heat1d_2pts_half_width(T, E)
        assert tab['cw'] == 2 * half + 8 == tab['sconst'].shape[1]
        q = tab['nrhs']
        rx = tab['rhs_x'].transpose(0, 2, 1).reshape(q, T * h)[:, :n]          # back from [q][h][T] to [q][n]
        assert np.array_equal(rx, app._rhs_split.basis)
        assert tab['rhs_t'].shape == (len(t), 2 * q)
        rng = np.random.default_rng(11 + lvl)
        u = rng.standard_normal((2, n))
        for i in (1, len(t) // 2, len(t) - 1):
            row = tab['sconst'][0 if tab['dtidx'] is None else tab['dtidx'][i]]
            a1, b1, a2, b2, skip1 = row[2 * half:2 * half + 5]
            s1 = a1 * u[0] + b1 * u[1] + tab['rhs_t'][i, :q] @ rx
            tmp1 = s1 if skip1 else solve(row[:half], s1, n)
            s2 = a2 * u[1] + b2 * tmp1 + tab['rhs_t'][i, q:] @ rx
            tmp2 = solve(row[half:2 * half], s2, n)
            want = orc.phi(u, t[i - 1], t[i])
            assert np.max(np.abs(np.stack([tmp1, tmp2]) - want)) <= 1e-11 * np.max(np.abs(want)), (name, lvl, i)


@pytest.mark.parametrize('n', [7, 63, 1023])
def test_fast_sine_transform_algorithm(n):
    """numpy restatement of k_rows_dst (csrc/spectral.cu), index for index: odd extension to 2N points, radix-2
    decimation-in-frequency stages with twiddles W_{2N}^(pos N / half), bit-reversed read-out, -sqrt(2/N)/2 Im(.).
    It must equal the product with the orthonormal sine matrix that mgb_rows_gemm computes."""
    N, N2 = n + 1, 2 * (n + 1)
    bits = N2.bit_length() - 1
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n)
    z = np.zeros(N2, dtype=complex)
    z[1:N] = x
    z[N2 - np.arange(1, N)] = -x
    tw = np.exp(-1j * np.pi * np.arange(N) / N)
    b = np.arange(N)
    for s in range(bits - 1, -1, -1):
        half = 1 << s
        pos = b & (half - 1)
        i = ((b >> s) << (s + 1)) + pos
        j = i + half
        p, q = z[i].copy(), z[j].copy()
        z[i] = p + q
        z[j] = (p - q) * tw[pos * (N >> s)]
    rev = np.array([int(format(k + 1, '0%db' % bits)[::-1], 2) for k in range(n)])
    got = -0.5 * np.sqrt(2.0 / N) * z[rev].imag
    k = np.arange(1, n + 1)
    want = x @ (np.sqrt(2.0 / N) * np.sin(np.pi * np.outer(k, k) / N))
    assert np.max(np.abs(got - want)) <= 1e-13 * np.sqrt(n)


def _emulate_team_solve(row, b, n, T, E):
    """numpy restatement of Heat1D::solve (csrc/phi.cuh) thread by thread: per-thread chunks of E elements in SUB
    sub-chains, the constant-ratio exclusive scans over the team truncated to `nscan` doubling steps (common.cuh
    scan_fwd / scan_bwd incl. the cross-warp carry), and the Sherman-Morrison correction with the per-thread PH / QH."""
    SUB = 3 if E % 3 == 0 else 1
    SL, PT = E // SUB, 2 + 2 * (3 if E % 3 == 0 else 1)
    beta, cs, kappa, bsl = row[0], row[1], row[2], row[3]
    Bd, B32, pw, nscan = row[4:9], row[9], row[10:10 + SL], int(row[23])
    per = row[24:24 + T * PT].reshape(T, PT)
    blf, blb, PH, QH = per[:, 0], per[:, 1], per[:, 2:2 + SUB], per[:, 2 + SUB:2 + 2 * SUB]
    x = np.zeros((T, E))
    x.reshape(-1)[:n] = b                                     # elements beyond n are 0 on entry
    nv = n - np.arange(T) * E

    def scan(a, forward):
        a = a.copy()
        lane, warp = np.arange(T) % 32, np.arange(T) // 32
        for k in range(5):
            if k >= nscan:
                break
            d = 1 << k
            t = np.zeros(T)
            for w in range(T // 32):                          # shuffles stay inside a warp
                seg = a[32 * w:32 * w + 32]
                t[32 * w:32 * w + 32] = np.concatenate([seg[:d], seg[:-d]]) if forward else np.concatenate([seg[d:], seg[-d:]])
            upd = (lane >= d) if forward else (lane + d < 32)
            a = np.where(upd, Bd[k] * t + a, a)
        excl = np.zeros(T)
        for w in range(T // 32):
            seg = a[32 * w:32 * w + 32]
            excl[32 * w:32 * w + 32] = np.concatenate([[0.0], seg[:-1]]) if forward else np.concatenate([seg[1:], [0.0]])
        if T > 32:
            W = T // 32
            tot = a[31::32] if forward else a[0::32]          # each warp's inclusive total
            for w in range(W):
                carry = 0.0
                rng = range(w) if forward else range(W - 1, w, -1)
                for v in rng:
                    carry = B32 * carry + tot[v]
                sel = warp == w
                excl[sel] = (blf if forward else blb)[sel] * carry + excl[sel]
        return excl

    e = np.zeros((T, SUB))
    for jj in range(SL):
        for s in range(SUB):
            j = s * SL + jj
            e[:, s] = beta * e[:, s] + x[:, j] * cs
            x[:, j] = e[:, s]
    a = e[:, 0].copy()
    for s in range(1, SUB):
        a = bsl * a + e[:, s]
    inflow = scan(a, True)
    if n % E == 0:
        inflow = np.where(nv > 0, inflow, 0.0)
    for s in range(SUB):
        for jj in range(SL):
            j = s * SL + jj
            y = pw[jj] * inflow + x[:, j]
            x[:, j] = y if n % E == 0 else np.where(j < nv, y, 0.0)
        inflow = bsl * inflow + e[:, s]
    f = np.zeros((T, SUB))
    for jj in range(SL - 1, -1, -1):
        for s in range(SUB):
            j = s * SL + jj
            f[:, s] = beta * f[:, s] + x[:, j]
            x[:, j] = f[:, s]
    a2 = f[:, SUB - 1].copy()
    for s in range(SUB - 2, -1, -1):
        a2 = bsl * a2 + f[:, s]
    inb = np.zeros((T, SUB))
    inb[:, SUB - 1] = scan(a2, False)
    for s in range(SUB - 2, -1, -1):
        inb[:, s] = bsl * inb[:, s + 1] + f[:, s + 1]
    z0 = pw[SL - 1] * inb[0, 0] + x[0, 0]
    gamma = kappa * z0
    for s in range(SUB):
        Bc = gamma * QH[:, s] + inb[:, s]
        Ac = -gamma * PH[:, s]
        for jj in range(SL):
            j = s * SL + jj
            x[:, j] = pw[SL - 1 - jj] * Bc + (pw[jj] * Ac + x[:, j])
    return x.reshape(-1)[:n]


@pytest.mark.parametrize('n, T, E', [(1023, 32, 33), (38, 32, 3), (15, 32, 1), (480, 32, 15), (999, 128, 9), (4095, 128, 33),
                                     (65, 32, 5)])
def test_team_solve_with_the_host_constants(lib, n, T, E):
    """The full constant row mgb_heat1d_step_consts writes (scalars, power table, scan depth, per-thread B^lane, PH, QH),
    used exactly as the kernel uses it, solves (I + r tridiag(-1, 2, -1)) x = b -- on the CPU, for the shapes of the
    headline workload, the multi-warp teams and the half-chunks of the two-point rows."""
    from pymgrit_b200 import _lib
    for r in (2.0, 128.0, 0.03, 8192.0):
        cw = lib.mgb_step_consts_width(_lib.APP_HEAT1D, T, E)
        row = np.zeros(cw)
        assert lib.mgb_heat1d_step_consts(r, n, T, E, row.ctypes.data_as(_lib.c_double_p)) == 0
        rng = np.random.default_rng(n + T)
        b = rng.standard_normal(n)
        x = _emulate_team_solve(row, b, n, T, E)
        from scipy.linalg import solve_banded
        ab = np.zeros((3, n))
        ab[0, 1:] = -r
        ab[1] = 1 + 2 * r
        ab[2, :-1] = -r
        want = solve_banded((1, 1), ab, b)
        assert np.max(np.abs(x - want)) <= 1e-11 * np.max(np.abs(want)), (n, T, E, r)


@pytest.mark.parametrize('n, T, E', [(4095, 128, 33), (128, 32, 5), (5, 32, 1), (256, 32, 9)])
def test_advection_team_solve_with_the_host_constants(lib, n, T, E):
    """Advection1D::apply (csrc/phi.cuh) restated in numpy with the row of mgb_advection1d_step_consts: the cyclic lower
    bidiagonal system (1 + nu) y_i - nu y_{i-1} = u_i (advection_1d.py:101-143) is solved by the forward recurrence, one
    exclusive scan over the team and the closure y_{n-1} = p_{n-1} / (1 - rho^n)."""
    from pymgrit_b200 import _lib
    SUB = 3 if E % 3 == 0 else 1
    SL, PT = E // SUB, 2 + 2 * SUB
    for nu in (0.0625, 1.0, 37.5):
        cw = lib.mgb_step_consts_width(_lib.APP_ADVECTION1D, T, E)
        row = np.zeros(cw)
        assert lib.mgb_advection1d_step_consts(nu, n, T, E, row.ctypes.data_as(_lib.c_double_p)) == 0
        rho, sig, dinv, rsl = row[0], row[1], row[2], row[3]
        Bd, B32, pw, nscan = row[4:9], row[9], row[10:10 + SL], int(row[23])
        per = row[24:24 + T * PT].reshape(T, PT)
        blf, RH = per[:, 0], per[:, 2:2 + SUB]
        rng = np.random.default_rng(n)
        u = rng.standard_normal(n)
        x = np.zeros((T, E))
        x.reshape(-1)[:n] = u
        e = np.zeros((T, SUB))
        for jj in range(SL):
            for s in range(SUB):
                j = s * SL + jj
                e[:, s] = rho * e[:, s] + x[:, j] * sig
                x[:, j] = e[:, s]
        a = e[:, 0].copy()
        for s in range(1, SUB):
            a = rsl * a + e[:, s]
        # exclusive forward scan with ratio B = rho^E (common.cuh scan_fwd, truncated to nscan doubling steps)
        lane, warp = np.arange(T) % 32, np.arange(T) // 32
        acc = a.copy()
        for k in range(min(nscan, 5)):
            d = 1 << k
            t = np.zeros(T)
            for w in range(T // 32):
                seg = acc[32 * w:32 * w + 32]
                t[32 * w:32 * w + 32] = np.concatenate([seg[:d], seg[:-d]])
            acc = np.where(lane >= d, Bd[k] * t + acc, acc)
        inflow = np.zeros(T)
        for w in range(T // 32):
            seg = acc[32 * w:32 * w + 32]
            inflow[32 * w:32 * w + 32] = np.concatenate([[0.0], seg[:-1]])
        if T > 32:
            tot = acc[31::32]
            for w in range(T // 32):
                carry = 0.0
                for v in range(w):
                    carry = B32 * carry + tot[v]
                inflow[warp == w] = blf[warp == w] * carry + inflow[warp == w]
        last = n - 1
        ins = np.zeros((T, SUB))
        ylast = 0.0
        inn = inflow.copy()
        for s in range(SUB):
            ins[:, s] = inn
            for jj in range(SL):
                j = s * SL + jj
                y = pw[jj] * inn + x[:, j]
                if last // E < T and last - (last // E) * E == j:
                    ylast = y[last // E]
            inn = rsl * inn + e[:, s]
        Y = ylast * dinv
        for s in range(SUB):
            cf = RH[:, s] * Y + ins[:, s]
            for jj in range(SL):
                j = s * SL + jj
                x[:, j] = pw[jj] * cf + x[:, j]
        got = x.reshape(-1)[:n]
        A = np.diag(np.full(n, 1 + nu)) - np.diag(np.full(n - 1, nu), -1)
        A[0, n - 1] -= nu
        want = np.linalg.solve(A, u)
        assert np.max(np.abs(got - want)) <= 1e-11 * np.max(np.abs(want)), (n, T, E, nu)
