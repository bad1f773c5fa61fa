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
_PIECE = 8192            # time points per broadcast call of the user callable on long grids
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
        """b at the points sel for all times t in one broadcast call: [len(sel), len(t)] (time along the contiguous
        axis: long inner loops for NumPy)."""
        return np.asarray(self.rhs(self.x[sel][:, None], t[None, :]), dtype=float)


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
        return np.asarray(self.rhs(x=xs[:, None], y=ys[:, None], t=t[None, :]), dtype=float)


class RhsSplit:
    """kind: 'zero' | 'separable' | 'dense'.

    The candidate split comes from a dozen sampled times; it is then VALIDATED on every time point it will be used for
    (validate()): the reference evaluates rhs(x, t_stop) in every step (heat/heat_1d.py:214), so a forcing that is
    active only between the samples must not be dropped.  Validation evaluates b at a handful of check points (the q
    pivot points plus spread and pseudo-random ones) for ALL t in one broadcast call -- the whole grid when the problem
    is small -- and compares with the split; times that do not fit are added to the samples and the split is rebuilt
    (a pulse in time times a fixed spatial shape is still separable, its time factor is then exact at every t_i);
    what still does not fit becomes 'dense'."""
    FULL_CHECK = 1 << 21          # nt * n up to which every (x, t) is compared
    N_CHECK = 6                   # check points in space beyond that (plus the q pivot points)

    def __init__(self, rhs, x=None, sampler=None, max_terms=MAX_TERMS):
        self.sampler = sampler if sampler is not None else Sampler1D(rhs, x)
        self.max_terms = max_terms
        self.kind, self.basis, self.sel = None, None, None
        self._sample_t = np.zeros(0)
        self._valid = {}              # time-grid key -> cached [q, nt] coefficients (None for 'zero')
        self._grids = {}              # time-grid key -> the validated grid itself (slabs of it are valid too)
        n = self.sampler.size
        rng = np.random.RandomState(12345)
        spread = np.unique(np.round(np.linspace(0, n - 1, min(n, self.N_CHECK - 2) + 2)[1:-1]).astype(int))
        self._check_pts = np.unique(np.concatenate([spread, rng.randint(0, n, size=min(n, 2))]))

    def _rows(self, times):
        if len(times) == 0:
            return np.zeros((0, self.sampler.size))
        return np.stack([self.sampler.full(float(tt)) for tt in times])

    @staticmethod
    def _key(t):
        t = np.ascontiguousarray(t, dtype=float)
        step = max(1, len(t) // 61)
        return (len(t), float(t[0]), float(t[-1]), hash(t[::step].tobytes())) if len(t) else (0,)

    def _slab_of(self, t):
        """(key, offset) if t is a contiguous piece of a grid this split has been validated on (the slab of a time rank),
        else None.  Checked on the end points and a strided sample of the piece."""
        if len(t) == 0:
            return None
        for key, grid in self._grids.items():
            if len(grid) < len(t):
                continue
            off = int(np.searchsorted(grid, t[0]))
            if off + len(t) > len(grid) or grid[off] != t[0] or grid[off + len(t) - 1] != t[-1]:
                continue
            step = max(1, len(t) // 257)
            if np.array_equal(grid[off:off + len(t):step], t[::step]):
                return key, off
        return None

    def _build(self, sample_t):
        """Candidate split from the rows b(., t), t in sample_t."""
        self._sample_t = np.unique(np.asarray(sample_t, dtype=float))
        self._valid = {}
        self._grids = {}
        R = self._rows(self._sample_t)
        self._scale = float(np.max(np.abs(R))) if R.size else 0.0
        if self._scale == 0.0:
            self.kind, self.basis, self.sel = 'zero', None, None
            return
        # row space of the samples by Gram-Schmidt with pivoting (a few dozen rows: rank-revealing enough, microseconds)
        basis = _pivoted_rows(R, 1e-13, self.max_terms + 1)
        q = len(basis)
        if q > self.max_terms or q >= len(self._sample_t):
            self.kind, self.basis, self.sel = 'dense', None, None
            return
        self.basis = np.ascontiguousarray(basis)                 # orthonormal rows spanning b(., t)
        self.sel = np.sort(_pivot_columns(self.basis))           # q points where the basis is well conditioned
        self.kind = 'separable'

    def analyse(self, t):
        t = np.asarray(t, dtype=float)
        pick = np.unique(np.round(np.linspace(0, len(t) - 1, min(len(t), _SAMPLES))).astype(int))
        sample_t = t[pick]
        mid = 0.5 * (sample_t[:-1] + sample_t[1:]) if len(sample_t) > 1 else sample_t
        self._build(np.concatenate([sample_t, mid]))
        self.validate(t)
        return self

    def _chunks(self, pts, t, each):
        """each(a, b, vals) for chunks [a, b) of t, vals = b at the spatial points pts for t[a:b], shape [len(pts), b-a]:
        one broadcast call per chunk, chunks spread over the worker threads on long grids (NumPy releases the GIL);
        per-time calls if the callable does not broadcast over (x, t)."""
        nt = len(t)
        probe = np.unique(np.array([0, nt // 2, nt - 1]))
        want = np.stack([self.sampler.full(float(t[i]))[pts] for i in probe])

        def work(a, b):
            vals = np.asarray(self.sampler.subset(pts, t[a:b]), dtype=float)
            if vals.shape != (len(pts), b - a):
                vals = np.broadcast_to(vals, (len(pts), b - a))   # e.g. a right-hand side that does not depend on t
            for k, i in enumerate(probe):                         # the broadcast call must equal the plain one
                if a <= i < b and not np.array_equal(vals[:, i - a], want[k]):
                    raise ValueError('right-hand side does not broadcast')
            return vals

        try:
            if nt < (1 << 15):
                each(0, nt, work(0, nt))
            else:
                pool, nthr = _pool()
                edges = np.linspace(0, nt, nthr + 1).astype(int)

                def run(ab):                  # cache-sized pieces: the temporaries of the callable stay in L2
                    for a in range(ab[0], ab[1], _PIECE):
                        b = min(ab[1], a + _PIECE)
                        each(a, b, work(a, b))
                list(pool.map(run, zip(edges[:-1], edges[1:])))
        except ValueError:
            step = 4096
            for a in range(0, nt, step):
                b = min(nt, a + step)
                each(a, b, np.stack([self.sampler.full(float(tt))[pts] for tt in t[a:b]], axis=1))

    def validate(self, t, retries=3):
        """Make the split hold at EVERY time of t (see the class docstring); may rebuild it or turn it 'dense'.
        Returns the kind.  The coefficients found on the way are kept for coefficients(t)."""
        t = np.asarray(t, dtype=float)
        key = self._key(t)
        n = self.sampler.size
        for _ in range(retries + 1):
            if self.kind == 'dense' or key in self._valid or self._slab_of(t) is not None:
                return self.kind
            full = len(t) * n <= self.FULL_CHECK
            pts = np.arange(n) if full else (self._check_pts if self.sel is None else
                                             np.unique(np.concatenate([self.sel, self._check_pts])))
            zero = self.kind == 'zero'
            coef = None if zero else np.empty((len(self.sel), len(t)))          # [q, nt]
            bad = []                                             # (time index, size of the defect)
            if not zero:
                pos = np.searchsorted(pts, self.sel)
                invT = np.ascontiguousarray(np.linalg.inv(self.basis[:, self.sel]).T)
                bptsT = np.ascontiguousarray(self.basis[:, pts].T)             # [len(pts), q]

            def each(a, b, vals):
                top = float(max(np.max(vals), -np.min(vals))) if vals.size else 0.0
                if zero:
                    if top != 0.0:
                        cols = np.flatnonzero(np.any(vals != 0.0, axis=0))
                        bad.extend(zip((cols + a).tolist(), np.max(np.abs(vals[:, cols]), axis=0).tolist()))
                    return
                np.matmul(invT, vals[pos], out=coef[:, a:b])
                diff = bptsT[:, :1] * coef[:1, a:b]                # sum of outer products (BLAS is slow for q = 1..4)
                for k in range(1, len(self.sel)):
                    diff += bptsT[:, k:k + 1] * coef[k:k + 1, a:b]
                diff -= vals
                tol = 1e-12 * max(self._scale, top)
                if float(max(np.max(diff), -np.min(diff))) > tol:
                    err = np.max(np.abs(diff), axis=0)
                    cols = np.flatnonzero(err > tol)
                    bad.extend(zip((cols + a).tolist(), err[cols].tolist()))

            if full:
                each(0, len(t), np.ascontiguousarray(self._rows(t).T))
            else:
                self._chunks(pts, t, each)
            if not bad:
                if len(self._valid) > 8:
                    self._valid.clear()
                    self._grids.clear()
                self._valid[key] = coef
                self._grids[key] = t
                return self.kind
            # the times that fit worst join the samples (largest defects, plus the first and last offender)
            bad.sort()
            worst = sorted(bad, key=lambda r: -r[1])[:6]
            extra = t[np.unique([r[0] for r in worst] + [bad[0][0], bad[-1][0]])]
            self._build(np.concatenate([self._sample_t, extra]))
        self.kind, self.basis, self.sel = 'dense', None, None
        self._valid = {}
        self._grids = {}
        return self.kind

    def coefficients(self, t, scale=None, out=None):
        """T_k(t_i) for every t_i (times scale[i] if given): shape (len(t), q), written to `out` if given (the caller
        may hand in page-locked memory so that the upload needs no staging copy)."""
        t = np.asarray(t, dtype=float)
        q = len(self.sel)
        if out is None:
            out = np.empty((len(t), q))
        scale = None if scale is None else np.asarray(scale, dtype=float)
        cached = self._valid.get(self._key(t))
        if cached is None and self.kind == 'separable':
            hit = self._slab_of(t)                               # a time rank's slab of a grid validated as a whole
            if hit is not None:
                cached = self._valid[hit[0]][:, hit[1]:hit[1] + len(t)]
        if cached is not None and cached.shape == (q, len(t)):   # found while validating this grid
            from pymgrit_b200.core.device_level import parallel_pieces, host_threads, _PAR_MIN
            if (len(t) >= _PAR_MIN and cached.dtype == np.float64 and cached.strides[1] == 8 and cached.strides[0] % 8 == 0
                    and out.flags.c_contiguous and (scale is None or scale.flags.c_contiguous)):
                from pymgrit_b200 import _lib            # native threads, no interpreter lock (csrc/host_tables.cu)
                _lib.check(_lib.lib().mgb_host_scale_rows(cached.ctypes.data, cached.strides[0] // 8, q, len(t),
                                                          None if scale is None else scale.ctypes.data, out.ctypes.data,
                                                          host_threads()), 'host_scale_rows')
                return out

            def piece(a, b):
                for k in range(q):
                    if scale is not None:
                        np.multiply(cached[k, a:b], scale[a:b], out=out[a:b, k])
                    else:
                        out[a:b, k] = cached[k, a:b]
            parallel_pieces(len(t), piece)
            return out
        invT = np.ascontiguousarray(np.linalg.inv(self.basis[:, self.sel]).T)   # q x q system, q <= MAX_TERMS

        def each(a, b, vals):
            """One chunk: evaluated, solved for the coefficients and scaled in place (at nt = 2^20 this is the largest
            host cost of the setup)."""
            c = invT @ vals if q > 1 else vals * invT[0, 0]
            if scale is not None:
                c *= scale[None, a:b]
            out[a:b] = c.T
        self._chunks(self.sel, t, each)
        return out

    def dense(self, t):
        return self._rows(np.asarray(t, dtype=float))

    def reproduces(self, t):
        """True if the split (possibly refined on the way) holds at every time of t; False if it had to become dense."""
        return self.validate(t) != 'dense'
