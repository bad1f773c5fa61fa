"""CPU: the algorithm of csrc/fourier.cu restated in numpy -- Bluestein's transform for rows of arbitrary length and the
complex scalar recurrences per frequency -- against the oracle's chain of Advection1D steps (mgrit.py:459-486 with
advection_1d.py:129-143).  Pins the sign conventions and the step factor 1 / (1 + nu (1 - exp(-i theta)))."""
import numpy as np
import pytest

from oracle import mgrit_oracle as O


def bluestein_tables(n):
    m = 1
    while m < 2 * n - 1:
        m *= 2
    j = np.arange(n)
    chirp = np.exp(1j * np.pi * ((j * j) % (2 * n)) / n)          # m^2 reduced mod 2n in integers (k_tables)
    b = np.zeros(m, complex)
    b[:n] = chirp
    b[m - np.arange(1, n)] = chirp[1:]
    return m, chirp, np.fft.fft(b)


def rfft_rows(x, n, m, chirp, bhat):
    z = np.zeros((x.shape[0], m), complex)
    z[:, :n] = x * np.conj(chirp)
    conv = np.fft.ifft(np.fft.fft(z, axis=1) * bhat, axis=1)
    k = n // 2 + 1
    return np.conj(chirp[:k]) * conv[:, :k]


def irfft_rows(xh, n, m, chirp, bhat):
    k = n // 2 + 1
    full = np.zeros((xh.shape[0], n), complex)
    full[:, :k] = np.conj(xh)
    hi = np.arange(k, n)
    full[:, k:] = xh[:, n - hi]
    z = np.zeros((xh.shape[0], m), complex)
    z[:, :n] = full * np.conj(chirp)
    conv = np.fft.ifft(np.fft.fft(z, axis=1) * bhat, axis=1)
    return (np.conj(chirp) * conv[:, :n]).real / n


@pytest.mark.parametrize('nx,uniform', [(16, True), (17, False), (130, True), (258, False)])
def test_fourier_recurrences_reproduce_the_chain_of_advection_steps(nx, uniform):
    npts = 40
    t = np.linspace(0, 2, npts) if uniform else 2 * np.linspace(0, 1, npts) ** 1.4
    app = O.Advection1DOracle(c=1, x_start=-1, x_end=1, nx=nx, t_interval=t)
    n = app.n
    rng = np.random.default_rng(nx)
    g = rng.standard_normal((npts, n)) * 1e-2
    u = np.zeros((npts, n))
    u[0] = np.asarray(app.u0).reshape(-1)
    for i in range(1, npts):
        u[i] = g[i] + app.phi(u[i - 1], t[i - 1], t[i])
    m, chirp, bhat = bluestein_tables(n)
    w = rfft_rows(np.vstack([u[:1], g[1:]]), n, m, chirp, bhat)
    assert np.allclose(w, np.fft.rfft(np.vstack([u[:1], g[1:]]), axis=1), rtol=0, atol=1e-12)
    theta = 2 * np.pi * np.arange(n // 2 + 1) / n
    cx, sx = 1 - np.cos(theta), np.sin(theta)
    fac = app.c / app.dx
    for i in range(1, npts):
        nu = fac * (t[i] - t[i - 1])
        w[i] = w[i] + w[i - 1] / ((1 + nu * cx) + 1j * nu * sx)
    back = irfft_rows(w, n, m, chirp, bhat)
    assert np.max(np.abs(back[1:] - u[1:])) <= 1e-12 * np.max(np.abs(u))


# ---- the pass structure of csrc/fourier.cu restated: plan, radix-2^RL passes in "registers", derived twiddles ---------
def _plan(log2m):
    """make_plan(): (stages in the pass, top stage) from the top stage down."""
    out, s = [], log2m - 1
    while s + 1 > 5:
        out.append((3, s))
        s -= 3
    if s + 1 == 5:
        out.append((3, s))
        s -= 3
    out.append((s + 1, s))
    return out


def _rot16(k):
    return np.exp(-1j * np.pi * k / 8)


def _dif_pass(z, log2m, rl, s_top, tw):
    q, half_m = s_top - rl + 1, 1 << (log2m - 1)
    for gi in range(1 << (log2m - rl)):
        lo, hi = gi & ((1 << q) - 1), gi >> q
        base = (hi << (s_top + 1)) + lo
        idx = [base + (a << q) for a in range(1 << rl)]
        v = [z[i] for i in idx]
        ws = tw[lo * (half_m >> s_top)]                       # one load per pass; the next stage's twiddle is its square
        for j in range(rl):
            ha = 1 << (rl - 1 - j)
            for a in range(1 << rl):
                if a & ha:
                    continue
                w = ws * _rot16((a & (ha - 1)) * (8 // ha))
                x, y = v[a], v[a + ha]
                v[a], v[a + ha] = x + y, (x - y) * w
            ws = ws * ws
        for i, val in zip(idx, v):
            z[i] = val


def _dit_pass(z, log2m, rl, q, tw):
    s_top, half_m = q + rl - 1, 1 << (log2m - 1)
    for gi in range(1 << (log2m - rl)):
        lo, hi = gi & ((1 << q) - 1), gi >> q
        base = (hi << (s_top + 1)) + lo
        idx = [base + (a << q) for a in range(1 << rl)]
        v = [z[i] for i in idx]
        wj = [tw[lo * (half_m >> s_top)]]
        for j in range(1, rl):
            wj.append(wj[-1] * wj[-1])
        for j in range(rl - 1, -1, -1):
            ha = 1 << (rl - 1 - j)
            for a in range(1 << rl):
                if a & ha:
                    continue
                w = wj[j] * _rot16((a & (ha - 1)) * (8 // ha))
                x, y = v[a], v[a + ha] * np.conj(w)
                v[a], v[a + ha] = x + y, x - y
        for i, val in zip(idx, v):
            z[i] = val


@pytest.mark.parametrize('log2m', list(range(1, 11)))
def test_fft_pass_structure_of_the_kernels(log2m):
    """Forward passes (natural in, bit-reversed out) reproduce numpy's FFT up to the bit reversal, and the backward passes
    undo them (times M): for every transform length the plan of passes covers each radix-2 stage exactly once, and the
    twiddles derived by squaring and by the fixed 16th roots of unity are the right ones."""
    m = 1 << log2m
    rng = np.random.default_rng(log2m)
    x = rng.standard_normal(m) + 1j * rng.standard_normal(m)
    tw = np.exp(-2j * np.pi * np.arange(m // 2) / m) if m > 1 else np.ones(1, complex)
    plan = _plan(log2m)
    assert sum(rl for rl, _ in plan) == log2m and all(1 <= rl <= 4 for rl, _ in plan)
    z = x.copy()
    for rl, top in plan:
        _dif_pass(z, log2m, rl, top, tw)
    rev = np.array([int(format(i, '0%db' % log2m)[::-1], 2) for i in range(m)])
    assert np.allclose(z[rev], np.fft.fft(x), rtol=0, atol=1e-12 * m)
    for rl, top in reversed(plan):
        _dit_pass(z, log2m, rl, top - rl + 1, tw)
    assert np.allclose(z / m, x, rtol=0, atol=1e-13 * m)
