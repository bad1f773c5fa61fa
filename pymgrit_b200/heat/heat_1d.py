"""1-D heat equation with the reference's interface (heat/heat_1d.py:14-217), state and Phi on the GPU.

    u_t - a u_xx = b(x, t),  homogeneous Dirichlet,  backward Euler:
    u_i = (I + dt L)^-1 (u_{i-1} + dt b(x, t_i)),   L = (a/dx^2) tridiag(-1, 2, -1)

The solve runs in libmgrit_b200 (csrc/phi.cuh, Heat1D) as a Toeplitz factorisation with
constant-coefficient recurrences; this file only prepares its per-level tables.
"""
import numpy as np

from pymgrit_b200 import _lib
from pymgrit_b200.core.application import DeviceApplication
from pymgrit_b200.core.vector import DeviceVector
from pymgrit_b200.core import device_level as dl
from pymgrit_b200.core.rhs_tables import RhsSplit


_SPLITS = {}          # (id(rhs), grid) -> RhsSplit, a handful of entries


def _shared_split(rhs, x, t, also=()):
    """The split of `rhs` on the grid x, valid at every time of t (and of the grids in `also`).  Levels of a hierarchy
    (deep copies made by simple_setup_problem or separately constructed applications) share one split object, hence one
    table of spatial factors on the device; each level validates it on its own time points (RhsSplit.validate)."""
    key = (id(rhs), len(x), float(x[0]), float(x[-1]))
    split = _SPLITS.get(key)
    if split is None or split.sampler.rhs is not rhs or split.kind == 'dense':
        split = RhsSplit(rhs, x).analyse(t)
    for grid in (t,) + tuple(also):
        if split.validate(grid) == 'dense':
            break
    if len(_SPLITS) > 16:
        _SPLITS.clear()
    if split.kind != 'dense':
        _SPLITS[key] = split
    return split


class VectorHeat1D(DeviceVector):
    """Vector of the nx-2 interior unknowns (heat_1d.py:14-128), stored in HBM."""

    def __init__(self, size, tensor=None):
        super().__init__((int(size),), tensor)


_PI = np.longdouble('3.14159265358979323846264338327950288')
SINE_GEMM_MAX = 2048        # largest n transformed with the dense sine matrix when n + 1 is not a power of two


_EIG = {}


def heat1d_eigenvalues(n, fac):
    """lam_k = fac * 4 sin^2(pi k / (2 (n+1))), k = 1..n, of fac * tridiag(-1, 2, -1) (heat_1d.py:177-196), in extended
    precision (read-only: every level of a hierarchy asks for the same array)."""
    key = (int(n), np.longdouble(fac).tobytes())
    lam = _EIG.get(key)
    if lam is None:
        if len(_EIG) > 16:
            _EIG.clear()
        k = np.arange(1, n + 1).astype(np.longdouble)
        lam = np.longdouble(fac) * 4 * np.sin(k * _PI / (2 * (n + 1))) ** 2
        lam.setflags(write=False)
        _EIG[key] = lam
    return lam


class SineTransform:
    """rows -> rows S on the device, S the orthonormal DST-I matrix (S = S^T = S^-1): a fast sine transform when n + 1
    is a power of two (mgb_rows_dst), else the product with S (mgb_rows_gemm).  One object per (n, device)."""
    _cache = {}

    @classmethod
    def get(cls, n):
        torch = dl._torch()
        key = (int(n), torch.cuda.current_device())
        if key not in cls._cache:
            if len(cls._cache) > 8:
                cls._cache.clear()
            cls._cache[key] = cls(int(n))
        return cls._cache[key]

    def __init__(self, n):
        import os
        torch = dl._torch()
        dev = torch.device('cuda', torch.cuda.current_device())
        self.n = n
        self.fast = ((n + 1) & n) == 0 and 32 * (n + 1) <= 200 * 1024 and os.environ.get('MGB_SPECTRAL_FFT', '1') != '0'
        if self.fast:
            self.smat = None
            self.twiddles = torch.empty((n + 1, 2), dtype=torch.float64, device=dev)
            _lib.check(_lib.lib().mgb_dst_twiddles(n, self.twiddles.data_ptr(), _lib.current_stream_ptr()), 'dst_twiddles')
        else:
            self.smat = torch.empty((n, n), dtype=torch.float64, device=dev)
            _lib.check(_lib.lib().mgb_sine_matrix(n, self.smat.data_ptr(), n, _lib.current_stream_ptr()), 'sine_matrix')

    def apply(self, src, src_pitch, dst, dst_pitch, rows, row0=None):
        """dst[:rows, :n] = src[:rows, :n] S (row 0 of src taken from row0 if given); src != dst."""
        if rows <= 0:
            return
        n = self.n
        if self.fast:
            _lib.check(_lib.lib().mgb_rows_dst(rows, n, src.data_ptr(), src_pitch,
                                               None if row0 is None else row0.data_ptr(), self.twiddles.data_ptr(),
                                               dst.data_ptr(), dst_pitch, _lib.current_stream_ptr()), 'rows_dst')
            return
        _lib.check(_lib.lib().mgb_rows_gemm(rows, n, n, src.data_ptr(), src_pitch,
                                            None if row0 is None else row0.data_ptr(), self.smat.data_ptr(), n,
                                            dst.data_ptr(), dst_pitch, _lib.current_stream_ptr()), 'rows_gemm')

    def rows(self, values):
        """[count, n] contiguous device tensor -> its transform (new tensor)."""
        torch = dl._torch()
        values = values.contiguous()
        out = torch.empty_like(values)
        self.apply(values, values.shape[1], out, out.shape[1], values.shape[0])
        return out


class Heat1D(DeviceApplication):
    """Same constructor as the reference (heat_1d.py:131-175).  Two device representations of the level rows:

      kind = APP_HEAT1D        the node values; Phi = the Toeplitz tridiagonal solve of csrc/phi.cuh (Heat1D)
      kind = APP_HEAT1D_SINE   the sine coefficients u S; Phi = one FMA + one multiplication per unknown (Heat1DSine)

    Mgrit picks the sine representation for a hierarchy of Heat1D levels connected by GridTransferCopy whose
    right-hand side is zero or separable (`as_sine()` returns the twin application it then runs); `sine_space=False`
    (or the environment variable MGB_HEAT1D_SINE=0) keeps the node representation, `sine_space=True` insists.
    User-visible values (vector_t_start, Mgrit.u[l][i].get_values(), step()) are node values either way."""
    kind = _lib.APP_HEAT1D

    def __init__(self, x_start, x_end, nx, a, init_cond=lambda x: x * 0, rhs=lambda x, t: x * 0, *args,
                 sine_space=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.x_start = x_start
        self.x_end = x_end
        self.x = np.linspace(self.x_start, self.x_end, nx)[1:-1]     # heat_1d.py:154-157
        self.nx = nx - 2
        self.ndof = self.nx
        self.dx = self.x[1] - self.x[0]
        self.a = a
        self.rhs = rhs
        self.sine_space = sine_space
        self.vector_template = VectorHeat1D(self.nx)
        self.init_cond = init_cond
        self.vector_t_start = VectorHeat1D(self.nx)
        self.vector_t_start.set_values(np.asarray(self.init_cond(self.x), dtype=float))
        # device representation of rhs(x, t): analysed once per (callable, grid) so that the levels of a hierarchy --
        # deep copies made by simple_setup_problem or separately constructed applications -- share one split, hence
        # one table of spatial factors on the device
        self._rhs_split = _shared_split(self.rhs, self.x, self.t)

    # ---- representation -------------------------------------------------------------------------------------------
    def can_sine(self):
        import os
        if self.sine_space is False or (self.sine_space is None and os.environ.get('MGB_HEAT1D_SINE', '1') == '0'):
            return False
        n = self.nx
        return self._rhs_split.kind != 'dense' and n >= 1 and (((n + 1) & n) == 0 and n <= 4095 or n <= SINE_GEMM_MAX)

    def as_sine(self):
        """The same problem with its level rows in sine space (a shallow copy: the user's object is left alone)."""
        import copy
        twin = copy.copy(self)
        twin.kind = _lib.APP_HEAT1D_SINE
        return twin

    @property
    def _in_sine(self):
        return self.kind == _lib.APP_HEAT1D_SINE

    def rows_to_values(self, rows):
        if not self._in_sine:
            return super().rows_to_values(rows)
        torch = dl._torch()
        out = torch.empty((rows.shape[0], self.nx), dtype=torch.float64, device=rows.device)
        SineTransform.get(self.nx).apply(rows, rows.stride(0), out, self.nx, rows.shape[0])
        return out

    def values_to_rows(self, values, rows) -> None:
        if not self._in_sine:
            return super().values_to_rows(values, rows)
        vals = values.reshape(rows.shape[0], self.nx).contiguous()
        if rows.stride(0) > self.nx:
            rows[:, self.nx:] = 0.0
        SineTransform.get(self.nx).apply(vals, self.nx, rows, rows.stride(0), rows.shape[0])

    def _sine_device_tables(self, team_threads, chunk):
        """Device tables every sine-space level of this problem shares: eigenvalues in natural order and the transformed
        spatial right-hand-side factors (natural order [q][pitch] and thread-transposed [q][chunk][team_threads])."""
        torch = dl._torch()
        dev = torch.device('cuda', torch.cuda.current_device())
        split = self._rhs_split
        key = (dev.index, team_threads, chunk, None if split.basis is None else split.basis.shape[0],
               None if split.basis is None else hash(split.basis.tobytes()), float(self.a), float(self.dx))
        hit = getattr(split, '_sine_dev', None)
        if hit is not None and hit[0] == key:
            return hit[1]
        n, pitch = self.nx, self.nx + (self.nx & 1)
        lam = np.asarray(heat1d_eigenvalues(n, np.longdouble(self.a) / np.longdouble(self.dx) ** 2), dtype=np.float64)
        tabs = dict(lam=torch.as_tensor(lam).to(dev), h2d=8 * n, rxh=None, rhs_x=None, nrhs=0)
        if split.kind == 'separable':
            q = split.basis.shape[0]
            basis = torch.as_tensor(np.ascontiguousarray(split.basis)).to(dev)
            tabs['h2d'] += split.basis.nbytes
            rxh = torch.zeros((q, pitch), dtype=torch.float64, device=dev)
            SineTransform.get(n).apply(basis, n, rxh, pitch, q)
            full = torch.zeros((q, team_threads * chunk), dtype=torch.float64, device=dev)
            full[:, :n] = rxh[:, :n]
            tabs['rxh'] = rxh
            tabs['rhs_x'] = full.reshape(q, team_threads, chunk).transpose(1, 2).contiguous()   # [q][chunk][team_threads]
            tabs['nrhs'] = q
        split._sine_dev = (key, tabs)
        return tabs

    def sine_host_tables(self, t, team_threads, chunk):
        """Host part of a sine-space level's Phi data (include/mgrit_b200.h, MGB_APP_HEAT1D_SINE): dt classes, the
        step-constant rows [dt, use_reciprocals], diag = [eigenvalues; 1 / (1 + dt lam)] thread-transposed, and dt per
        point."""
        t = np.asarray(t, dtype=float)
        dt_full, dt_lo, dt_hi = dl.time_steps(t)
        dts, dtidx = dl.dt_classes(t, dt_full[1:], (dt_lo, dt_hi))
        uniform = len(dts) == 1
        sconst = np.zeros((len(dts), 8))
        sconst[:, 0] = dts
        sconst[:, 1] = 1.0 if uniform else 0.0
        n = self.nx
        lam_l = heat1d_eigenvalues(n, np.longdouble(self.a) / np.longdouble(self.dx) ** 2)      # heat_1d.py:185
        flat = np.zeros((2, team_threads * chunk))
        flat[0, :n] = np.asarray(lam_l, dtype=np.float64)
        if uniform:       # 1 / (1 + dt lam_k) in extended precision, rounded once; the padding acts on zeros
            flat[1, :n] = np.asarray(1 / (1 + np.longdouble(dts[0]) * lam_l), dtype=np.float64)
        diag = np.ascontiguousarray(flat.reshape(2, team_threads, chunk).transpose(0, 2, 1))   # [2][chunk][team_threads]
        pitch = n + (n & 1)
        return dict(ndt=len(dts), dtidx=dtidx, sconst=sconst, cw=8, diag=diag, nat=flat[:, :pitch].copy()), dt_full

    def level_tables(self, t, team_threads, chunk):
        if self._in_sine:
            return self._level_tables_sine(t, team_threads, chunk)
        return self._level_tables_node(t, team_threads, chunk)

    def _level_tables_node(self, t, team_threads, chunk):
        fac = self.a / self.dx ** 2                                   # heat_1d.py:185
        t = np.asarray(t, dtype=float)
        dt_full, dt_lo, dt_hi = dl.time_steps(t)
        dts, dtidx = dl.dt_classes(t, dt_full[1:], (dt_lo, dt_hi))
        tab = dict(ndt=len(dts), dtidx=dtidx,
                   sconst=dl.step_const_table(self.kind, dts * fac, self.nx, team_threads, chunk))
        tab['cw'] = tab['sconst'].shape[1]
        split = self._rhs_split
        if split.kind == 'separable':
            tab['nrhs'] = split.basis.shape[0]
            tab['rhs_x'] = dl.rhs_x_layout(split.basis, self.nx, team_threads, chunk)
            tab['rhs_x_key'] = (id(split), team_threads, chunk)      # levels with the same split share the device table
            # b * dt (heat_1d.py:214), evaluated straight into page-locked memory on long grids
            out = dl.pinned_array((len(t), tab['nrhs'])) if len(t) >= (1 << 15) else None
            tab['rhs_t'] = split.coefficients(t, scale=dt_full, out=out)
        elif split.kind == 'dense':
            tab['rhs_dense'] = split.dense(t) * dt_full[:, None]
        return tab

    def level_tables_host(self, t, team_threads, chunk):
        """The part of level_tables() that needs no device: step classes, eigenvalue tables, the time factors of the
        right-hand side (into page-locked memory on long grids).  DeviceLevel.start_host_tables runs it on a helper
        thread for level 0; level_tables_finish() adds the device tables."""
        if not self._in_sine:
            return dict(node=self._level_tables_node(t, team_threads, chunk))
        tab, dt_full = self.sine_host_tables(t, team_threads, chunk)
        dl.mark('heat1d: dt classes, diag')
        split = self._rhs_split
        if split.kind == 'separable':
            out = dl.pinned_array((len(t), split.basis.shape[0])) if len(t) >= (1 << 15) else None
            tab['rhs_t'] = split.coefficients(t, scale=dt_full, out=out)
            dl.mark('heat1d: rhs time factors')
        return dict(sine=tab)

    def level_tables_finish(self, host, team_threads, chunk):
        if 'node' in host:
            return host['node']
        torch = dl._torch()
        tab = host['sine']
        shared = self._sine_device_tables(team_threads, chunk)
        dl.mark('heat1d: shared device tables')
        # the same tables in natural mode order for the one-thread-per-mode sweeps (csrc/sine_modes.cu)
        nat_host = tab.pop('nat')
        q = shared['nrhs']
        nat = torch.empty((2 + q, nat_host.shape[1]), dtype=torch.float64, device=shared['lam'].device)
        dl.upload_small(nat[:2], nat_host)
        if q:
            nat[2:].copy_(shared['rxh'])
        tab['nat_dev'] = nat
        if self._rhs_split.kind == 'separable':
            tab['nrhs'] = shared['nrhs']
            tab['rhs_x_dev'] = shared['rhs_x']
        dl.mark('heat1d: nat upload')
        return tab

    def _level_tables_sine(self, t, team_threads, chunk):
        return self.level_tables_finish(self.level_tables_host(t, team_threads, chunk), team_threads, chunk)

    # ---- coarsest-level solve in sine space (csrc/spectral.cu) ---------------------------------------------------
    SPECTRAL_MIN_POINTS = 24     # below this the sequential Phi chain (mgb_forward_solve) is as fast

    def spectral_solver(self, level):
        """A SpectralSolve for `level` (the coarsest DeviceLevel of a hierarchy), or None when the chain of Phi
        applications is used: short levels, or a right-hand side that is not separable."""
        if self._in_sine:
            return SineLevelSolve(self, level) if level.npts >= 2 else None
        if level.npts < self.SPECTRAL_MIN_POINTS or self._rhs_split.kind == 'dense':
            return None
        return SpectralSolve(self, level)

    def spectral_min_points(self):
        return 2 if self._in_sine else self.SPECTRAL_MIN_POINTS


class SpectralSolve:
    """u_i = g_i + Phi_i(u_{i-1}) over a whole level (mgrit.py:459-486) as two sine transforms and n scalar recurrences
    (include/mgrit_b200.h, "the coarsest-level solve in sine space")."""

    def __init__(self, app, level):
        torch = dl._torch()
        dev = level.u.device
        n, pitch = app.nx, level.pitch
        self.n, self.pitch, self.level = n, pitch, level
        self.h2d_bytes = 0
        level.ensure_t_dev()                     # the scalar recurrences read dt_i = t[i] - t[i-1]
        self.xform = SineTransform.get(n)          # fast sine transform for n + 1 a power of two, else the product with S
        lam = np.asarray(heat1d_eigenvalues(n, np.longdouble(app.a) / np.longdouble(app.dx) ** 2), dtype=np.float64)
        self.lam = torch.as_tensor(lam).to(dev)
        self.h2d_bytes += lam.nbytes
        self.work = torch.zeros((level.npts, pitch), dtype=torch.float64, device=dev)
        self.rxhat = None
        split = app._rhs_split
        if split.kind == 'separable':
            basis = np.zeros((split.basis.shape[0], pitch))
            basis[:, :n] = split.basis
            rows = torch.as_tensor(basis).to(dev)
            self.h2d_bytes += basis.nbytes
            self.rxhat = torch.zeros_like(rows)
            self._gemm(rows, self.rxhat, len(basis))

    def _gemm(self, src, dst, rows, row0=None):
        """dst[:rows, :n] = src[:rows, :n] S (row 0 of src taken from row0 if given)."""
        self.xform.apply(src, self.pitch, dst, self.pitch, rows, row0=row0)

    def transform_in(self):
        """work[0] = u[0] S, work[i] = g[i] S: one product."""
        lv = self.level
        if lv.g is not None:
            self._gemm(lv.g, self.work, lv.npts, row0=lv.u)
        else:
            self._gemm(lv.u, self.work, 1)
            self.work[1:].zero_()

    def recur(self, ends=None, zero_start=False):
        lv = self.level
        _lib.check(_lib.lib().mgb_heat1d_spectral_recur(lv.ref, self.lam.data_ptr(),
                                                        None if self.rxhat is None else self.rxhat.data_ptr(),
                                                        self.work.data_ptr(), None if ends is None else ends.data_ptr(),
                                                        1 if zero_start else 0, _lib.current_stream_ptr()), 'spectral_recur')

    def recur_time_parallel(self, comm):
        """The recurrences over all time ranks without a rank-to-rank chain: local recurrences (from zero on ranks > 0),
        one all-gather of (last value, factor product) per rank, local fix-up."""
        torch = dl._torch()
        if getattr(self, '_ends', None) is None:
            self._ends = torch.zeros((2, self.pitch), dtype=torch.float64, device=self.work.device)
            self._all_ends = torch.zeros((comm.size, 2, self.pitch), dtype=torch.float64, device=self.work.device)
        self.recur(ends=self._ends, zero_start=comm.rank > 0)
        box = getattr(comm, 'mailbox', None)
        if box and box.gather and box.pitch == self.pitch:
            all_ends = box.share_rows(self._ends)            # peer-memory stores instead of a collective
        else:
            comm.all_gather_rows(self._all_ends, self._ends)
            all_ends = self._all_ends.data_ptr()
        if comm.rank > 0:
            _lib.check(_lib.lib().mgb_heat1d_spectral_fixup(self.level.ref, self.lam.data_ptr(), self.work.data_ptr(),
                                                            all_ends, comm.rank, _lib.current_stream_ptr()), 'spectral_fixup')

    def transform_out(self, first_row=1):
        """u[i] = work[i] S for i >= first_row (row 0 is the initial condition on time rank 0, the ghost otherwise)."""
        lv = self.level
        if lv.npts > first_row:
            self._gemm(self.work[first_row:], lv.u[first_row:], lv.npts - first_row)

    def solve(self, comm):
        """The whole level; returns the number of kernel launches."""
        rank, size = comm.Get_rank(), comm.Get_size()
        self.transform_in()
        if size > 1:
            self.recur_time_parallel(comm)
        else:
            self.recur()
        self.transform_out(first_row=0 if rank > 0 else 1)
        return 3 + (1 if size > 1 and rank > 0 else 0)


class SineLevelSolve:
    """The sequential solve (mgrit.py:459-486) of a level whose rows are already in sine space: no transforms, the n
    scalar recurrences run time-parallel in one launch (mgb_sine_level_solve).  Between time ranks: every rank runs from
    zero, the (last value, factor product) rows travel once, one fix-up pass -- no rank-to-rank chain
    (mgrit.py:467-484)."""

    def __init__(self, app, level):
        shared = app._sine_device_tables(level.team_threads, level.chunk)
        self.level, self.pitch = level, level.pitch
        self.lam, self.rxh = shared['lam'], shared['rxh']
        self.h2d_bytes = 0
        level.ensure_t_dev()                     # the scalar recurrences read dt_i = t[i] - t[i-1]
        self._ends = None

    def _recur(self, ends=None, zero_start=False):
        _lib.check(_lib.lib().mgb_sine_level_solve(self.level.ref, self.lam.data_ptr(),
                                                   None if self.rxh is None else self.rxh.data_ptr(),
                                                   None if ends is None else ends.data_ptr(), 1 if zero_start else 0,
                                                   _lib.current_stream_ptr()), 'sine_level_solve')

    def solve(self, comm):
        rank, size = comm.Get_rank(), comm.Get_size()
        if size == 1:
            self._recur()
            return 1
        torch = dl._torch()
        if self._ends is None:
            dev = self.level.u.device
            self._ends = torch.zeros((2, self.pitch), dtype=torch.float64, device=dev)
            self._all_ends = torch.zeros((size, 2, self.pitch), dtype=torch.float64, device=dev)
        self._recur(ends=self._ends, zero_start=rank > 0)
        box = getattr(comm, 'mailbox', None)
        if box and box.gather and box.pitch == self.pitch:
            all_ends = box.share_rows(self._ends)            # peer-memory stores instead of a collective
        else:
            comm.all_gather_rows(self._all_ends, self._ends)
            all_ends = self._all_ends.data_ptr()
        if rank > 0:
            _lib.check(_lib.lib().mgb_heat1d_spectral_fixup(self.level.ref, self.lam.data_ptr(), self.level.u.data_ptr(),
                                                            all_ends, rank, _lib.current_stream_ptr()), 'spectral_fixup')
        return 1 + (1 if rank > 0 else 0)
