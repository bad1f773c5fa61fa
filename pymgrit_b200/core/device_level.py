"""One level of the time-grid hierarchy in HBM: the arrays and the struct mgb_level handed to the
C ABI (include/mgrit_b200.h).  PyTorch owns the buffers; nothing here computes."""
import ctypes as C
import weakref

import numpy as np

from pymgrit_b200 import _lib

_SHARED_TABLES = {}        # (device, table key) -> weak reference to a device tensor several levels point to
TRACE = None               # a list while scripts/e2e_breakdown.py measures: (label, perf_counter seconds) marks of the setup


def mark(label):
    if TRACE is not None:
        import time
        TRACE.append((label, time.perf_counter()))


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise Exception('pymgrit_b200 needs a CUDA device: there is no CPU fallback')
    return torch


class _PinnedArray(np.ndarray):
    """ndarray view of a page-locked torch tensor (kept alive by the view)."""
    _owner = None
    _slot = None
    _base_tensor = None


_PINNED_POOL = {}         # bytes -> list of [page-locked tensor, CUDA event of the last upload from it or None]
import threading
_POOL_LOCK = threading.Lock()   # level 0's host tables are made on a helper thread (DeviceLevel.start_host_tables)


def pinned_array(shape):
    """Uninitialised float64 host array in page-locked memory, so that DeviceLevel uploads it with one asynchronous
    copy and no staging pass.  cudaHostAlloc costs ~5 ms per megabyte-sized buffer, more than the whole solve of the
    headline workload, so the buffers are kept in a small pool for the life of the process: a buffer is handed out again
    once the array that used it is gone and the upload queued from it has completed (its event is waited for)."""
    import sys
    torch = _torch()
    shape = tuple(int(v) for v in shape)
    nbytes = 8 * int(np.prod(shape)) if shape else 8
    ten = None
    with _POOL_LOCK:
        for ent in _PINNED_POOL.setdefault(nbytes, []):
            if sys.getrefcount(ent[0]) <= 2:            # only the pool (and this call) hold it: its array was dropped
                if ent[1] is not None:
                    ent[1].synchronize()
                    ent[1] = None
                ten, slot = ent[0], ent
                break
        if ten is None:
            ten = torch.empty((nbytes // 8,), dtype=torch.float64, pin_memory=True)
            slot = [ten, None]
            if len(_PINNED_POOL[nbytes]) < 4:
                _PINNED_POOL[nbytes].append(slot)
    view = ten.view(shape) if shape else ten
    arr = view.numpy().view(_PinnedArray)
    arr._owner = view
    arr._slot = slot
    arr._base_tensor = ten
    return arr


def _stage_small(host):
    """Copy of a small host array in page-locked memory (pooled), typed like the array."""
    nbytes = host.nbytes
    bucket = 64
    while bucket < nbytes:                 # few pool sizes: powers of two
        bucket *= 2
    buf = pinned_array((bucket // 8,))
    raw = buf._owner.view(_torch().uint8)[:nbytes]
    view = raw.view({np.dtype(np.float64): _torch().float64, np.dtype(np.int32): _torch().int32}[host.dtype]).view(host.shape)
    arr = view.numpy().view(_PinnedArray)
    arr[...] = host
    arr._owner, arr._slot, arr._base_tensor = view, buf._slot, buf._base_tensor
    return arr


def upload_small(dst, host):
    """dst (device tensor) <- host array through pooled page-locked memory, asynchronously.  The pool entry is marked
    with an event recorded after the copy: the pool hands the buffer out again only once the copy has run.  (A staging
    array that is merely dropped after queueing its copy can be overwritten by the next table before the device has read
    it -- seen as an intermittent wrong reciprocal table on two time ranks, where the device lags the host during the
    setup.)"""
    torch = _torch()
    a = _stage_small(np.ascontiguousarray(host))
    dst.copy_(a._owner, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    a._slot[1] = ev
    return dst


def team_shape(kind, n):
    import os
    forced = os.environ.get('MGB_SHAPE_%d' % kind)            # experiments: "threads,chunk" for application kind `kind`
    if forced:
        t, e = (int(v) for v in forced.split(','))
        if t * e >= n + (n & 1):
            return t, e
    t, e = C.c_int32(0), C.c_int32(0)
    _lib.check(_lib.lib().mgb_team_shape(kind, n, C.byref(t), C.byref(e)), 'team_shape')
    return t.value, e.value


def step_const_table(kind, values, n, team_threads, chunk):
    """[len(values)][cw] table of Phi constants, one row per distinct dt-dependent parameter."""
    lib = _lib.lib()
    cw = lib.mgb_step_consts_width(kind, team_threads, chunk)
    fn = {_lib.APP_HEAT1D: lib.mgb_heat1d_step_consts, _lib.APP_ADVECTION1D: lib.mgb_advection1d_step_consts}[kind]
    out = np.zeros((len(values), cw))
    for k, v in enumerate(values):
        _lib.check(fn(float(v), n, team_threads, chunk, out[k].ctypes.data_as(_lib.c_double_p)), 'step_consts')
    return out


class DeviceLevel:
    """Arrays of one level on the current CUDA device + the matching struct mgb_level."""

    def __init__(self, app, t, cpts=None, with_g=False, u_init=None, defer_tables=False, zero_u=True):
        """defer_tables: allocate the arrays now, build and upload the Phi tables when finish_tables() is called (the
        solver does that for level 0 after it has queued the coarse-level work of nested iteration, so that the host
        evaluates the long level-0 tables while the device is busy).  zero_u=False: leave u uninitialised (the caller
        guarantees that every row is written before it is read, as nested iteration + one cycle do)."""
        torch = _torch()
        dev = torch.device('cuda', torch.cuda.current_device())
        self.app = app
        self.t = np.asarray(t, dtype=float)
        self.npts = len(self.t)
        self.n = int(app.ndof)                            # doubles of a row that carry values
        tiny = app.kind in (_lib.APP_DAHLQUIST, _lib.APP_BRUSSELATOR)
        self.pitch = int(app.row_pitch()) if hasattr(app, 'row_pitch') else (self.n if tiny else self.n + (self.n & 1))
        self.batched = app.kind == _lib.APP_BATCHED          # no fused kernels, no struct mgb_level (core/batched.py)
        self.team_threads, self.chunk = (0, 0) if self.batched else team_shape(app.kind, self.n)
        # the (asynchronous) zero fill of the level arrays runs on the device while the host builds the tables
        self.zero_filled = bool(zero_u) and u_init is None
        mark('level: shape')
        if u_init is not None:
            self.u = u_init
        elif zero_u:
            self.u = torch.zeros((self.npts, self.pitch), dtype=torch.float64, device=dev)
        else:
            self.u = torch.empty((self.npts, self.pitch), dtype=torch.float64, device=dev)
        mark('level: u allocated')
        self.g = torch.zeros((self.npts, self.pitch), dtype=torch.float64, device=dev) if with_g else None
        mark('level: g allocated')
        self.cpts = None if cpts is None else np.asarray(cpts, dtype=np.int32)
        self._keep = []                                   # tensors referenced by the struct
        self.h2d_bytes = 0
        self.nsys = 1
        self.c = None
        self._host_job = None
        if not defer_tables:
            self.finish_tables()

    def start_host_tables(self):
        """Deferred tables of a long level whose application splits them into a host part (NumPy on the table threads,
        into page-locked memory) and a device part: the host part starts now on a helper thread, so that it runs while
        the constructor of the solver allocates the other levels, builds their tables and queues nested iteration;
        finish_tables() waits for it.  Nothing changes in what is computed."""
        import os
        fn = getattr(self.app, 'level_tables_host', None)
        if (fn is None or self.c is not None or self.batched or self._host_job is not None or self.npts < _PAR_MIN
                or os.environ.get('MGB_TABLES_AHEAD', '1') == '0'):
            return
        torch = _torch()
        dev, box = self.u.device, {}

        def work():
            try:
                with torch.cuda.device(dev):            # page-locked buffers belong to this rank's device context
                    box['host'] = fn(self.t, self.team_threads, self.chunk)
            except BaseException as exc:                # re-raised by finish_tables() on the constructor's thread
                box['error'] = exc
        th = threading.Thread(target=work, name='mgb-level-tables', daemon=True)
        th.start()
        self._host_job = (th, box)

    def finish_tables(self):
        if self.c is not None or self.batched:
            return
        torch = _torch()
        app, dev = self.app, self.u.device
        tiny = app.kind in (_lib.APP_DAHLQUIST, _lib.APP_BRUSSELATOR)
        mark('tables: begin')
        if self._host_job is not None:
            th, box = self._host_job
            self._host_job = None
            th.join()
            mark('tables: waited for the helper thread')
            if 'error' in box:
                raise box['error']
            tab = app.level_tables_finish(box['host'], self.team_threads, self.chunk)
        else:
            tab = app.level_tables(self.t, self.team_threads, self.chunk)
        mark('tables: host tables made')

        # The small host tables of the level travel in ONE page-locked buffer and one asynchronous copy (each separate
        # upload costs ~40 us of host time: staging buffer, copy, event); the device tensors are views of that copy.
        batch = {}
        small = [(k, np.ascontiguousarray(a, dtype=dt_)) for k, a, dt_ in (
            ('cpts', self.cpts, np.int32), ('t', self.t if tiny else None, np.float64),
            ('sconst', tab.get('sconst'), np.float64), ('dtidx', tab.get('dtidx'), np.int32),
            ('diag', tab.get('diag'), np.float64),
            ('rhs_x', tab.get('rhs_x') if tab.get('rhs_x_key') is None and tab.get('rhs_x_dev') is None else None, np.float64),
            ('rhs_t', tab.get('rhs_t'), np.float64))
            if a is not None and getattr(a, '_owner', None) is None and 0 < np.asarray(a).size * np.dtype(dt_).itemsize <= (1 << 18)]
        if len(small) > 1:
            offs, total = [], 0
            for _, a in small:
                offs.append(total)
                total += (a.nbytes + 255) & ~255
            bucket = 256
            while bucket < total:
                bucket *= 2
            stage = pinned_array((bucket // 8,))
            raw = stage._owner.view(torch.uint8)
            host_bytes = raw.numpy()
            for (k, a), off in zip(small, offs):
                host_bytes[off:off + a.nbytes] = a.reshape(-1).view(np.uint8)
            dev_raw = raw[:total].to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            stage._slot[1] = ev                               # the pool reuses the buffer after this copy has run
            self._keep += [stage._owner, dev_raw]
            self.h2d_bytes += sum(a.nbytes for _, a in small)
            tdt = {np.dtype(np.float64): torch.float64, np.dtype(np.int32): torch.int32}
            for (k, a), off in zip(small, offs):
                batch[k] = dev_raw[off:off + a.nbytes].view(tdt[a.dtype]).view(a.shape)

        def up(a, dtype, key=None):
            if a is None:
                return None
            if key in batch:
                return batch[key]
            owner = getattr(a, '_owner', None)
            if owner is None and 0 < a.size * np.dtype(dtype).itemsize <= (1 << 20):
                # small table: through page-locked memory as well -- a pageable cudaMemcpy waits for everything queued on
                # the stream (the coarse sweeps of nested iteration) before it returns
                a = _stage_small(np.ascontiguousarray(a, dtype=dtype))
                owner = a._owner
            if owner is not None and a.dtype == dtype:        # page-locked already: asynchronous copy, no staging
                ten = owner.to(dev, non_blocking=True)
                slot = getattr(a, '_slot', None)
                if slot is not None:                           # the pool reuses the buffer after this upload has run
                    ev = torch.cuda.Event()
                    ev.record()
                    slot[1] = ev
                self._keep.append(owner)                       # alive until the copy has run
                self.h2d_bytes += a.nbytes
                self._keep.append(ten)
                return ten
            host = np.ascontiguousarray(a, dtype=dtype)
            ten = torch.as_tensor(host).to(dev)
            self.h2d_bytes += host.nbytes
            self._keep.append(ten)
            return ten

        self.cpts_dev = up(self.cpts, np.int32, 'cpts')
        self.t_dev = up(self.t, np.float64, 't') if tiny else None     # only the ODE kernels read the time grid
        sconst = up(tab.get('sconst'), np.float64, 'sconst')
        dtidx = up(tab.get('dtidx'), np.int32, 'dtidx')
        # tables an application family shares between its levels are handed over as device tensors
        if tab.get('rhs_x_dev') is not None:
            rhs_x = tab['rhs_x_dev']
        elif tab.get('rhs_x_key') is not None:
            key = (dev.index,) + tuple(tab['rhs_x_key'])
            hit = _SHARED_TABLES.get(key)
            if hit is None or hit[0]() is None or hit[0]().shape != tab['rhs_x'].shape:
                ten = up(tab['rhs_x'], np.float64)
                _SHARED_TABLES[key] = (weakref.ref(ten),)
                rhs_x = ten
            else:
                rhs_x = hit[0]()
                self._keep.append(rhs_x)
        else:
            rhs_x = up(tab.get('rhs_x'), np.float64, 'rhs_x')
        rhs_t = up(tab.get('rhs_t'), np.float64, 'rhs_t')
        sig = tab.get('sig_dev')
        self._keep += [t_ for t_ in (rhs_x, sig) if t_ is not None]
        rhs_dense = None
        if tab.get('rhs_dense_dev') is not None:
            rhs_dense = tab['rhs_dense_dev']
            self._keep.append(rhs_dense)
        elif tab.get('rhs_dense') is not None:
            dense = np.zeros((self.npts, self.pitch))
            dense[:, :self.n] = tab['rhs_dense']
            rhs_dense = up(dense, np.float64)

        def ptr(ten):
            return None if ten is None else ten.data_ptr()

        c = _lib.MgbLevel()
        c.app, c.n, c.pitch, c.npts = app.kind, self.n, self.pitch, self.npts
        c.u_dev, c.g_dev = ptr(self.u), ptr(self.g)
        c.cpts_dev, c.ncpts = ptr(self.cpts_dev), 0 if self.cpts is None else len(self.cpts)
        c.team_threads, c.chunk = self.team_threads, self.chunk
        c.ndt, c.cw = int(tab.get('ndt', 1)), int(tab.get('cw', 0))
        c.dtidx_dev, c.sconst_dev = ptr(dtidx), ptr(sconst)
        c.nrhs = int(tab.get('nrhs', 0))
        c.nsys = int(tab.get('nsys', 1))
        c.sig_dev = ptr(sig)
        diag = up(tab.get('diag'), np.float64, 'diag')
        nat = tab.get('nat_dev')
        self.nat_dev = nat
        if nat is not None:
            self._keep.append(nat)
        c.nat_dev = ptr(nat)
        c.diag_dev = ptr(diag)
        self.nsys = c.nsys
        c.rhs_x_dev, c.rhs_t_dev, c.rhs_dense_dev = ptr(rhs_x), ptr(rhs_t), ptr(rhs_dense)
        c.t_dev = ptr(self.t_dev)
        for k, v in enumerate(tab.get('p', [])):
            c.p[k] = float(v)
        for k, v in enumerate(tab.get('ip', [])):
            c.ip[k] = int(v)
        self.c = c
        mark('tables: uploaded')

    @property
    def ref(self):
        return C.byref(self.c)

    def ensure_t_dev(self):
        """Upload the time grid for a kernel family that reads it (the sine-space solve; the ODE kernels always have it)."""
        if self.t_dev is None:
            torch = _torch()
            host = np.ascontiguousarray(self.t, dtype=np.float64)
            self.t_dev = torch.as_tensor(host).to(self.u.device)
            self.h2d_bytes += host.nbytes
            self.c.t_dev = self.t_dev.data_ptr()

    # -- values <-> rows (identity for every application except Heat2D, whose rows are in sine space) ----------
    def get_vector(self, arr, i):
        """Time point i of a level array as a Vector of the application (a view where the layout allows it)."""
        return self.app.vector_template._new(self.app.rows_to_values(arr[i:i + 1])[0])

    def set_vector(self, arr, i, vec):
        shape = self.app.vector_template.shape
        self.app.values_to_rows(vec.device_values.reshape((1,) + tuple(shape)), arr[i:i + 1])

    def values(self, arr=None, idx=None, chunk=64):
        """Host array [points, *vector shape] of the values at points idx (all by default)."""
        arr = self.u if arr is None else arr
        idx = np.arange(self.npts) if idx is None else np.atleast_1d(np.asarray(idx))
        shape = tuple(self.app.vector_template.shape)
        out = np.empty((len(idx),) + shape)
        torch = _torch()
        for a in range(0, len(idx), chunk):
            sel = torch.as_tensor(idx[a:a + chunk], device=arr.device, dtype=torch.long)
            out[a:a + chunk] = self.app.rows_to_values(arr.index_select(0, sel)).cpu().numpy()
        return out


_PAR_MIN = 1 << 17          # arrays at least this long are worked on in pieces by the table threads


def host_threads():
    """Threads for the native host passes (csrc/host_tables.cu): memory-bound, 8 are enough."""
    import os
    return max(1, min(8, os.cpu_count() or 1))


def parallel_pieces(n, fn):
    """fn(a, b) over pieces [a, b) of range(n) on the table threads (core/rhs_tables.py; NumPy releases the GIL), in
    this thread for short ranges.  Returns the list of results in order."""
    if n < _PAR_MIN:
        return [fn(0, n)]
    from pymgrit_b200.core.rhs_tables import _pool
    pool, nthr = _pool()
    edges = np.linspace(0, n, nthr + 1).astype(int)
    return list(pool.map(lambda ab: fn(int(ab[0]), int(ab[1])), zip(edges[:-1], edges[1:])))


def time_steps(t):
    """dt[i] = t[i] - t[i-1] (dt[0] = 0), its smallest and largest value: one pass over t in pieces (at nt = 2^20 the
    separate NumPy passes of diff / min / max are a millisecond each)."""
    t = np.asarray(t, dtype=float)
    dt = np.empty(len(t))
    if len(t) == 0:
        return dt, 0.0, 0.0
    dt[0] = 0.0
    if len(t) == 1:
        return dt, 0.0, 0.0

    if len(t) >= _PAR_MIN and t.flags.c_contiguous:
        # long grids: native threads (csrc/host_tables.cu) -- no interpreter lock held while the other levels are built
        lo, hi = C.c_double(0.0), C.c_double(0.0)
        _lib.check(_lib.lib().mgb_host_time_steps(t.ctypes.data, len(t), dt.ctypes.data, C.byref(lo), C.byref(hi),
                                                  host_threads()), 'host_time_steps')
        return dt, lo.value, hi.value

    def piece(a, b):                       # steps into points a+1 .. b
        np.subtract(t[a + 1:b + 1], t[a:b], out=dt[a + 1:b + 1])
        return float(dt[a + 1:b + 1].min()), float(dt[a + 1:b + 1].max())
    parts = parallel_pieces(len(t) - 1, piece)
    return dt, min(p[0] for p in parts), max(p[1] for p in parts)


def dt_classes(t, dt=None, lo_hi=None):
    """dt_i = t[i] - t[i-1] grouped by exact value: (distinct values, index per point or None).  dt: t[1:] - t[:-1] if the
    caller has it already (lo_hi: its minimum and maximum, from time_steps)."""
    t = np.asarray(t, dtype=float)
    if len(t) < 2:
        return np.array([1.0]), None
    if dt is None:
        dt = t[1:] - t[:-1]
    if lo_hi is None:
        lo_hi = (dt.min(), dt.max()) if dt[0] == dt[-1] == dt[len(dt) // 2] else (0.0, 1.0)
    if lo_hi[0] == lo_hi[1]:                                    # uniform grid
        return dt[:1].copy(), None
    uniq, inv = np.unique(dt, return_inverse=True)
    idx = np.zeros(len(t), dtype=np.int32)
    idx[1:] = inv
    return uniq, idx


def rhs_x_layout(basis, n, team_threads, chunk):
    """[q][n] spatial factors -> [q][chunk][team_threads] (thread-transposed, zero padded)."""
    q = basis.shape[0]
    full = np.zeros((q, team_threads * chunk))
    full[:, :n] = basis
    return np.ascontiguousarray(full.reshape(q, team_threads, chunk).transpose(0, 2, 1))


def single_step(app, u_start, t_start, t_stop):
    """Application.step on the device: out = Phi(u_start) from t_start to t_stop (one launch)."""
    torch = _torch()
    cache = app.__dict__.setdefault('_step_cache', {})
    key = (float(t_start), float(t_stop))
    lvl = cache.get(key)
    if lvl is None:
        if len(cache) > 64:
            cache.clear()
        lvl = DeviceLevel(app, np.array(key))
        cache[key] = lvl
    shape = tuple(app.vector_template.shape)
    x = u_start.device_values
    buf_in = torch.zeros((1, lvl.pitch), dtype=torch.float64, device=x.device)
    app.values_to_rows(x.reshape((1,) + shape), buf_in)
    buf_out = torch.zeros_like(buf_in)
    _lib.check(_lib.lib().mgb_step(lvl.ref, 1, buf_in.data_ptr(), buf_out.data_ptr(), _lib.current_stream_ptr()), 'step')
    return app.vector_template._new(app.rows_to_values(buf_out)[0])
