"""1-D advection with the reference's interface (advection/advection_1d.py:14-143) on the GPU:
u_t + c u_x = 0, periodic, first-order upwind, backward Euler:  u_i = (I + dt L)^-1 u_{i-1}."""
import numpy as np

from pymgrit_b200 import _lib
from pymgrit_b200.core.application import DeviceApplication
from pymgrit_b200.core.vector import DeviceVector
from pymgrit_b200.core import device_level as dl


class VectorAdvection1D(DeviceVector):
    def __init__(self, size, tensor=None):
        super().__init__((int(size),), tensor)


class Advection1D(DeviceApplication):
    kind = _lib.APP_ADVECTION1D

    def __init__(self, c, x_start, x_end, nx, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.c = c
        self.x_start = x_start
        self.x_end = x_end
        self.x = np.linspace(self.x_start, self.x_end, nx)[0:-1]      # advection_1d.py:85-88
        self.nx = nx - 1
        self.ndof = self.nx
        self.dx = self.x[1] - self.x[0]
        self.vector_template = VectorAdvection1D(self.nx)
        self.vector_t_start = VectorAdvection1D(self.nx)
        self.initialise()

    def initialise(self):
        self.vector_t_start.set_values(np.exp(-self.x ** 2))           # advection_1d.py:122-127

    def level_tables(self, t, team_threads, chunk):
        fac = self.c / self.dx                                         # advection_1d.py:108
        dts, dtidx = dl.dt_classes(t)
        tab = dict(ndt=len(dts), dtidx=dtidx,
                   sconst=dl.step_const_table(self.kind, dts * fac, self.nx, team_threads, chunk))
        tab['cw'] = tab['sconst'].shape[1]
        return tab

    # ---- coarsest-level solve in Fourier space (csrc/fourier.cu) -------------------------------------------------
    SPECTRAL_MIN_POINTS = 96     # below this the sequential Phi chain (mgb_forward_solve) is as fast
    spectral_single_rank_only = True
    SPECTRAL_MAX_N = 4096        # one row's convolution (2^p >= 2n - 1 complex doubles) must fit in shared memory

    def spectral_solver(self, level):
        """A FourierSolve for `level` (the coarsest DeviceLevel of a hierarchy), or None when the chain of Phi
        applications is used: short levels, long rows, or MGB_ADVECTION_FOURIER=0."""
        import os
        if (os.environ.get('MGB_ADVECTION_FOURIER', '1') == '0' or level.npts < self.SPECTRAL_MIN_POINTS
                or self.nx > self.SPECTRAL_MAX_N):
            return None
        return FourierSolve(self, level)

    def spectral_min_points(self):
        return self.SPECTRAL_MIN_POINTS


_FFT_TABLES = {}         # (device, n) -> (tw, chirp, bhat)


def circ_fft_tables(n, dev):
    """Twiddles, chirp and the chirp's spectrum for rows of length n (include/mgrit_b200.h, mgb_circ_fft_tables)."""
    torch = dl._torch()
    key = (dev.index, int(n))
    hit = _FFT_TABLES.get(key)
    if hit is None:
        m = _lib.lib().mgb_circ_fft_length(int(n))
        mk = lambda cnt: torch.empty((cnt, 2), dtype=torch.float64, device=dev)
        tw, chirp, bhat, work = mk(m // 2), mk(n), mk(m), mk(m)
        _lib.check(_lib.lib().mgb_circ_fft_tables(int(n), tw.data_ptr(), chirp.data_ptr(), bhat.data_ptr(), work.data_ptr(),
                                                  _lib.current_stream_ptr()), 'circ_fft_tables')
        hit = (tw, chirp, bhat, work)            # `work` stays alive until the stream has used it
        _FFT_TABLES[key] = hit
    return hit[:3]


class FourierSolve:
    """u_i = g_i + Phi_i(u_{i-1}) over a whole level (mgrit.py:459-486) as a real Fourier transform of all rows, n/2 + 1
    complex scalar recurrences that run time-parallel, and the inverse transform (include/mgrit_b200.h, "the
    coarsest-level solve in Fourier space").  One time rank; with several ranks the solver keeps the chain."""
    def __init__(self, app, level):
        torch = dl._torch()
        dev = level.u.device
        self.level, self.n = level, int(app.nx)
        self.fac = float(app.c) / float(app.dx)                      # advection_1d.py:108
        self.h2d_bytes = 0
        level.ensure_t_dev()                                         # the recurrences read dt_i = t[i] - t[i-1]
        self.tw, self.chirp, self.bhat = circ_fft_tables(self.n, dev)
        self.ldw = 2 * (self.n // 2 + 1)
        self.work = torch.empty((level.npts, self.ldw), dtype=torch.float64, device=dev)

    def solve(self, comm):
        """The whole level; returns the number of kernel launches."""
        lv, lib, st = self.level, _lib.lib(), _lib.current_stream_ptr()
        if lv.npts < 2:
            return 0
        src = lv.g if lv.g is not None else lv.u                     # a one-level "hierarchy" has no g: rows 1.. are zero
        if lv.g is None:
            lv.u[1:].zero_()
        _lib.check(lib.mgb_rows_rfft(lv.npts, self.n, src.data_ptr(), lv.pitch, lv.u.data_ptr(), self.tw.data_ptr(),
                                     self.chirp.data_ptr(), self.bhat.data_ptr(), self.work.data_ptr(), self.ldw, st),
                   'rows_rfft')
        _lib.check(lib.mgb_advection1d_spectral_recur(self.n, lv.npts, lv.t_dev.data_ptr(), self.fac, self.work.data_ptr(),
                                                      self.ldw, st), 'advection1d_spectral_recur')
        _lib.check(lib.mgb_rows_irfft(lv.npts, 1, self.n, self.work.data_ptr(), self.ldw, self.tw.data_ptr(),
                                      self.chirp.data_ptr(), self.bhat.data_ptr(), lv.u.data_ptr(), lv.pitch, st),
                   'rows_irfft')
        return 3
