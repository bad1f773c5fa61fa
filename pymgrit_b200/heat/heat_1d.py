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


class Heat1D(DeviceApplication):
    kind = _lib.APP_HEAT1D

    def __init__(self, x_start, x_end, nx, a, init_cond=lambda x: x * 0, rhs=lambda x, t: x * 0, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.x_start = x_start
        self.x_end = x_end
        self.x = np.linspace(self.x_start, self.x_end, nx)[1:-1]     # heat_1d.py:154-157
        self.nx = nx - 2
        self.ndof = self.nx
        self.dx = self.x[1] - self.x[0]
        self.a = a
        self.rhs = rhs
        self.vector_template = VectorHeat1D(self.nx)
        self.init_cond = init_cond
        self.vector_t_start = VectorHeat1D(self.nx)
        self.vector_t_start.set_values(np.asarray(self.init_cond(self.x), dtype=float))
        # device representation of rhs(x, t): analysed once per (callable, grid) so that the levels of a hierarchy --
        # deep copies made by simple_setup_problem or separately constructed applications -- share one split, hence
        # one table of spatial factors on the device
        self._rhs_split = _shared_split(self.rhs, self.x, self.t)

    def level_tables(self, t, team_threads, chunk):
        fac = self.a / self.dx ** 2                                   # heat_1d.py:185
        t = np.asarray(t, dtype=float)
        dt_full = np.empty(len(t))
        dt_full[0] = 0.0
        np.subtract(t[1:], t[:-1], out=dt_full[1:])
        dts, dtidx = dl.dt_classes(t, dt_full[1:])
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

    # ---- coarsest-level solve in sine space (csrc/spectral.cu) ---------------------------------------------------
    SPECTRAL_MIN_POINTS = 24     # below this the sequential Phi chain (mgb_forward_solve) is as fast

    def spectral_solver(self, level):
        """A SpectralSolve for `level` (the coarsest DeviceLevel of a hierarchy), or None when the chain of Phi
        applications is used: short levels, or a right-hand side that is not separable."""
        if level.npts < self.SPECTRAL_MIN_POINTS or self._rhs_split.kind == 'dense':
            return None
        return SpectralSolve(self, level)


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
        import os
        # n + 1 a power of two (nx = 2^k + 1): fast sine transform, one CTA per row; otherwise the product with S
        self.fast = ((n + 1) & n) == 0 and 32 * (n + 1) <= 200 * 1024 and os.environ.get('MGB_SPECTRAL_FFT', '1') != '0'
        if self.fast:
            self.smat = None
            self.twiddles = torch.empty((n + 1, 2), dtype=torch.float64, device=dev)
            _lib.check(_lib.lib().mgb_dst_twiddles(n, self.twiddles.data_ptr(), _lib.current_stream_ptr()), 'dst_twiddles')
        else:
            self.smat = torch.empty((n, n), dtype=torch.float64, device=dev)
            _lib.check(_lib.lib().mgb_sine_matrix(n, self.smat.data_ptr(), n, _lib.current_stream_ptr()), 'sine_matrix')
        k = np.arange(1, n + 1).astype(np.longdouble)
        pi = np.longdouble('3.14159265358979323846264338327950288')
        fac = np.longdouble(app.a) / np.longdouble(app.dx) ** 2                       # heat_1d.py:185
        lam = np.asarray(fac * 4 * np.sin(k * pi / (2 * (n + 1))) ** 2, dtype=np.float64)
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
        if self.fast:
            _lib.check(_lib.lib().mgb_rows_dst(rows, self.n, src.data_ptr(), self.pitch,
                                               None if row0 is None else row0.data_ptr(), self.twiddles.data_ptr(),
                                               dst.data_ptr(), self.pitch, _lib.current_stream_ptr()), 'rows_dst')
            return
        _lib.check(_lib.lib().mgb_rows_gemm(rows, self.n, self.n, src.data_ptr(), self.pitch,
                                            None if row0 is None else row0.data_ptr(), self.smat.data_ptr(), self.n,
                                            dst.data_ptr(), self.pitch, _lib.current_stream_ptr()), 'rows_gemm')

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
