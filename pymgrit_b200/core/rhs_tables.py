"""Device representation of a user right-hand side b(x, t) given as a NumPy callable.

The reference evaluates `rhs(self.x, t_stop)` inside every step (heat/heat_1d.py:214).  A kernel
cannot call Python, so the callable is turned into tables once per level:

  separable   b(x, t) = sum_k T_k(t) X_k(x), k < q <= MAX_TERMS.  Found numerically: sample b at a
              few times, take the row space (pivoted QR), pick q well-conditioned spatial points (pivoted
              QR) and get T_k(t_i) for every time point from b at those q points only.  The kernel
              then adds dt_i * sum_k T_k(t_i) X_k(x) out of an L1-resident table: no HBM traffic.
  dense       anything else: one row b(x, t_i) * dt_i per time point, streamed like the FAS rows.

The split is verified against direct evaluations of the callable; a right-hand side that does not
reproduce to 1e-13 relative falls back to the dense table.
"""
import numpy as np

MAX_TERMS = 4
_SAMPLES = 12
_POOL = None


def _pool():
    """Worker threads for the table evaluation on long time grids, started once per process."""
    global _POOL
    if _POOL is None:
        import concurrent.futures as cf
        import os
        nthr = max(1, min(16, (os.cpu_count() or 1)))
        _POOL = (cf.ThreadPoolExecutor(max_workers=nthr, thread_name_prefix='mgb-rhs'), nthr)
    return _POOL


def _pivoted_rows(R, rtol, max_rows):
    """Orthonormal rows spanning the rows of R up to rtol (modified Gram-Schmidt, largest remaining row first)."""
    W = np.array(R, dtype=float)
    out, first = [], None
    while len(out) < max_rows:
        norms = np.sqrt(np.sum(W * W, axis=1))
        k = int(np.argmax(norms))
        if first is None:
            first = norms[k]
        if norms[k] <= rtol * first or norms[k] == 0.0:
            break
        v = W[k] / norms[k]
        for b in out:                                            # re-orthogonalise against the accepted rows
            v = v - np.dot(b, v) * b
        v = v / np.sqrt(np.dot(v, v))
        out.append(v)
        W = W - np.outer(W @ v, v)
    return np.array(out) if out else np.zeros((0, R.shape[1]))


def _pivot_columns(B):
    """Indices of len(B) columns of the q x n matrix B chosen by Gram-Schmidt with column pivoting."""
    W = np.array(B, dtype=float)
    sel = []
    for _ in range(W.shape[0]):
        k = int(np.argmax(np.sum(W * W, axis=0)))
        sel.append(k)
        c = W[:, k] / np.sqrt(np.dot(W[:, k], W[:, k]))
        W = W - np.outer(c, c @ W)
    return np.array(sel, dtype=int)


class Sampler1D:
    """b(x, t) on the 1-D grid x (heat/heat_1d.py:214: rhs(self.x, t_stop))."""

    def __init__(self, rhs, x):
        self.rhs, self.x = rhs, np.asarray(x, dtype=float)
        self.size = len(self.x)

    def full(self, t):
        val = np.asarray(self.rhs(self.x, t), dtype=float)
        return np.broadcast_to(val, np.shape(self.x)).astype(float)

    def subset(self, sel, t):
        """b at the points sel for all times t in one broadcast call: [len(t), len(sel)]."""
        return np.asarray(self.rhs(self.x[sel][None, :], t[:, None]), dtype=float)


class Sampler2D:
    """b(x, y, t) on the interior nodes, flattened row-major (heat/heat_2d.py:301-303:
    rhs(x=self.x_2d[1:-1], y=self.y_2d[:, 1:-1], t=t_stop))."""

    def __init__(self, rhs, x, y):
        self.rhs = rhs
        self.xi, self.yi = np.asarray(x, dtype=float)[1:-1], np.asarray(y, dtype=float)[1:-1]
        self.shape = (len(self.xi), len(self.yi))
        self.size = self.shape[0] * self.shape[1]

    def full(self, t):
        val = np.asarray(self.rhs(x=self.xi[:, None], y=self.yi[None, :], t=t), dtype=float)
        return np.broadcast_to(val, self.shape).astype(float).reshape(-1)

    def subset(self, sel, t):
        xs, ys = self.xi[sel // self.shape[1]], self.yi[sel % self.shape[1]]
        return np.asarray(self.rhs(x=xs[None, :], y=ys[None, :], t=t[:, None]), dtype=float)


class RhsSplit:
    """kind: 'zero' | 'separable' | 'dense'."""

    def __init__(self, rhs, x=None, sampler=None, max_terms=MAX_TERMS):
        self.sampler = sampler if sampler is not None else Sampler1D(rhs, x)
        self.max_terms = max_terms
        self.kind, self.basis, self.sel = None, None, None

    def _rows(self, times):
        if len(times) == 0:
            return np.zeros((0, self.sampler.size))
        return np.stack([self.sampler.full(float(tt)) for tt in times])

    def analyse(self, t):
        t = np.asarray(t, dtype=float)
        pick = np.unique(np.round(np.linspace(0, len(t) - 1, min(len(t), _SAMPLES))).astype(int))
        sample_t = t[pick]
        mid = 0.5 * (sample_t[:-1] + sample_t[1:]) if len(sample_t) > 1 else sample_t
        R = self._rows(sample_t)
        scale = np.max(np.abs(R)) if R.size else 0.0
        if scale == 0.0 and not np.any(self._rows(mid)):
            self.kind = 'zero'
            return self
        # row space of the samples by Gram-Schmidt with pivoting (a dozen rows: rank-revealing enough, microseconds)
        basis = _pivoted_rows(R, 1e-13, self.max_terms + 1)
        q = len(basis)
        if q > self.max_terms or q >= len(sample_t):
            self.kind = 'dense'
            return self
        basis = np.ascontiguousarray(basis)                      # orthonormal rows spanning b(., t)
        sel = np.sort(_pivot_columns(basis))                     # q points where the basis is well conditioned
        self.basis, self.sel = basis, sel
        # verify on times that were not used to build the basis
        check = self._rows(mid)
        coef = self.coefficients(mid)
        err = np.max(np.abs(coef @ basis - check)) if check.size else 0.0
        self.kind = 'separable' if err <= 1e-13 * max(scale, np.max(np.abs(check)) if check.size else 0.0) else 'dense'
        return self

    def coefficients(self, t, scale=None, out=None):
        """T_k(t_i) for every t_i (times scale[i] if given): shape (len(t), q), written to `out` if given (the caller
        may hand in page-locked memory so that the upload needs no staging copy)."""
        t = np.asarray(t, dtype=float)
        inv = np.linalg.inv(self.basis[:, self.sel])             # q x q system, q <= MAX_TERMS
        if out is None:
            out = np.empty((len(t), len(self.sel)))
        try:                                   # one broadcast call per chunk when the callable allows it
            self._broadcast_coefficients(t, inv, scale, out)
        except Exception:
            vals = np.stack([self.sampler.full(float(tt))[self.sel] for tt in t])
            np.matmul(vals, inv, out=out)
            if scale is not None:
                out *= np.asarray(scale, dtype=float)[:, None]
        return out

    def _broadcast_coefficients(self, t, inv, scale, out):
        """sampler.subset over all of t -> coefficients, or None if the callable does not broadcast over (t, x).
        Long grids are cut into chunks worked on by a persistent pool of threads (NumPy ufuncs release the GIL); each
        chunk is evaluated, solved for the coefficients and scaled in place, which matters at nt = 2^20 where this is
        the largest host cost of the setup."""
        nt, q = len(t), len(self.sel)
        probe = np.unique(np.array([0, nt // 2, nt - 1]))
        want = np.stack([self.sampler.full(float(t[i]))[self.sel] for i in probe])

        def work(a, b):
            vals = np.asarray(self.sampler.subset(self.sel, t[a:b]), dtype=float)
            if vals.shape != (b - a, q):
                vals = np.broadcast_to(vals, (b - a, q))          # e.g. a right-hand side that does not depend on t
            for k, i in enumerate(probe):                         # the broadcast call must equal the plain one
                if a <= i < b and not np.array_equal(vals[i - a], want[k]):
                    raise ValueError('right-hand side does not broadcast')
            np.matmul(vals, inv, out=out[a:b])
            if scale is not None:
                out[a:b] *= scale[a:b, None]

        if nt < (1 << 15):
            work(0, nt)
            return out
        pool, nthr = _pool()
        edges = np.linspace(0, nt, nthr + 1).astype(int)        # one chunk per thread: fewest GIL hand-overs
        list(pool.map(lambda ab: work(*ab), zip(edges[:-1], edges[1:])))
        return out

    def dense(self, t):
        return self._rows(np.asarray(t, dtype=float))

    def reproduces(self, t, samples=3):
        """True if the separable form matches direct evaluations at a few of the times t (used when a level adopts
        the split another level of the same problem has analysed)."""
        if self.kind != 'separable':
            return False
        t = np.asarray(t, dtype=float)
        pick = t[np.unique(np.round(np.linspace(0, len(t) - 1, min(len(t), samples))).astype(int))]
        direct = self._rows(pick)
        scale = max(np.max(np.abs(direct)), 1e-300)
        return bool(np.max(np.abs(self.coefficients(pick) @ self.basis - direct)) <= 1e-12 * scale)
