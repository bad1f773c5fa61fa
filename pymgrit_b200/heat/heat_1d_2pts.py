"""Common part of the two-point heat applications (heat/heat_1d_2pts_bdf1.py, heat/heat_1d_2pts_bdf2.py) on the GPU.

A time point of these applications is the pair (u(t_i), u(t_i + dtau)); one step is two tridiagonal solves.  Both
methods are written as

    tmp1 = (I + r1 K)^-1 (a1 first + b1 second + c1 b(x, t_i)),
    tmp2 = (I + r2 K)^-1 (a2 second + b2 tmp1  + c2 b(x, t_i + dtau)),        K = tridiag(-1, 2, -1)

and run in libmgrit_b200 (csrc/phi.cuh, Heat1D2Pts; include/mgrit_b200.h "two-point rows"); this file only prepares the
per-level tables: the coefficients above per distinct dt, and the time factors of the separable right-hand side at
t_i and t_i + dtau.  The method is a property of the level, so a BDF2 fine level over BDF1 coarse levels
(examples/example_heat_1d_bdf2.py) is one hierarchy of the same kernels.
"""
import numpy as np

from pymgrit_b200 import _lib
from pymgrit_b200.core.application import DeviceApplication
from pymgrit_b200.core import device_level as dl
from pymgrit_b200.heat.heat_1d import _shared_split
from pymgrit_b200.heat.vector_heat_1d_2pts import VectorHeat1D2Pts


class Heat1D2Pts(DeviceApplication):
    kind = _lib.APP_HEAT1D_2PTS
    method = None                 # 'BDF1' | 'BDF2', set by the subclasses

    def __init__(self, x_start, x_end, nx, dtau, a, init_cond=lambda x: x * 0, rhs=lambda x, t: x * 0, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.x_start = x_start
        self.x_end = x_end
        self.x = np.linspace(self.x_start, self.x_end, nx)[1:-1]
        self.nx = nx - 2
        self.ndof = self.nx                   # unknowns of ONE of the two time points (struct mgb_level.n)
        self.dx = self.x[1] - self.x[0]
        self.a = a
        self.dtau = dtau
        self.rhs = rhs
        self.init_cond = init_cond
        self.vector_template = VectorHeat1D2Pts(self.nx, dtau)
        self.vector_t_start = VectorHeat1D2Pts(self.nx, dtau)
        self._rhs_split = _shared_split(self.rhs, self.x, self.t, also=(np.asarray(self.t, dtype=float) + dtau,))
        if self._rhs_split.kind == 'dense':
            raise Exception('the two-point heat applications need a right-hand side that is a short sum of products '
                            'X(x) T(t) (at most 4 terms); this one is not')
        first = np.asarray(self.init_cond(self.x), dtype=float) * np.ones(self.nx)
        self.vector_t_start.set_lazy_second(first, self._second_start_value)

    # ---- the value at t_start + dtau (heat_1d_2pts_bdf1.py:64-66, heat_1d_2pts_bdf2.py:65-68) ---------------------
    def _second_start_value(self, first):
        raise NotImplementedError

    def _heat1d(self, t_a, t_b):
        """Backward-Euler Heat1D application on the two time points (t_a, t_b): its device step is the solve
        (I + (t_b - t_a) L)^-1 (u + (t_b - t_a) b(x, t_b)) both start-up formulas need."""
        from pymgrit_b200.heat.heat_1d import Heat1D
        return Heat1D(x_start=self.x_start, x_end=self.x_end, nx=self.nx + 2, a=self.a, rhs=self.rhs,
                      t_interval=np.array([t_a, t_b], dtype=float))

    # ---- row layout -------------------------------------------------------------------------------------------------
    def _shape(self):
        t, e = dl.team_shape(self.kind, self.nx)
        return t, e, (e - 1) // 2

    def row_pitch(self):
        t, e, _ = self._shape()
        return t * e

    def _row_index(self, device):
        """Positions of first[0..n) and second[0..n) inside a row (include/mgrit_b200.h, "two-point rows")."""
        cache = self.__dict__.setdefault('_row_index_cache', {})
        key = str(device)
        if key not in cache:
            torch = dl._torch()
            t, e, h = self._shape()
            k = np.arange(self.nx)
            first = (k // h) * e + (k % h)
            cache[key] = (torch.as_tensor(first, device=device, dtype=torch.long),
                          torch.as_tensor(first + h, device=device, dtype=torch.long))
        return cache[key]

    def rows_to_values(self, rows):
        i1, i2 = self._row_index(rows.device)
        torch = dl._torch()
        return torch.stack([rows.index_select(1, i1), rows.index_select(1, i2)], dim=1)

    def values_to_rows(self, values, rows) -> None:
        i1, i2 = self._row_index(rows.device)
        vals = values.reshape(rows.shape[0], 2, self.nx)
        rows.zero_()
        rows.index_copy_(1, i1, vals[:, 0])
        rows.index_copy_(1, i2, vals[:, 1])

    def __getstate__(self):
        state = super().__getstate__()
        state.pop('_row_index_cache', None)
        return state

    # ---- per-level tables ---------------------------------------------------------------------------------------------
    def _coefficients(self, dt):
        """(r1, r2, a1, b1, a2, b2, c1, c2) for the step over dt = t_stop - t_start."""
        raise NotImplementedError

    def level_tables(self, t, team_threads, chunk):
        lib = _lib.lib()
        h = (chunk - 1) // 2
        half = lib.mgb_heat1d_2pts_half_width(team_threads, chunk)
        cw = lib.mgb_step_consts_width(self.kind, team_threads, chunk)
        dts, dtidx = dl.dt_classes(t)
        sconst = np.zeros((len(dts), cw))
        c1 = np.zeros(len(dts))
        c2 = np.zeros(len(dts))
        for k, dt in enumerate(dts):
            r1, r2, a1, b1, a2, b2, c1[k], c2[k] = self._coefficients(float(dt))
            row = sconst[k]
            if r1 > 0.0:
                _lib.check(lib.mgb_heat1d_step_consts(float(r1), self.nx, team_threads, h,
                                                      row[:half].ctypes.data_as(_lib.c_double_p)), 'step_consts')
            _lib.check(lib.mgb_heat1d_step_consts(float(r2), self.nx, team_threads, h,
                                                  row[half:2 * half].ctypes.data_as(_lib.c_double_p)), 'step_consts')
            row[2 * half:2 * half + 5] = (a1, b1, a2, b2, 0.0 if r1 > 0.0 else 1.0)
        tab = dict(ndt=len(dts), dtidx=dtidx, sconst=sconst, cw=cw)
        split = self._rhs_split
        if split.kind == 'separable':
            t = np.asarray(t, dtype=float)
            q = split.basis.shape[0]
            per_point = np.zeros(len(t), dtype=np.int64) if dtidx is None else np.asarray(dtidx, dtype=np.int64)
            s1, s2 = c1[per_point], c2[per_point]
            s1[0] = s2[0] = 0.0                                   # point 0 is never produced by a step
            tab['nrhs'] = q
            tab['rhs_x'] = dl.rhs_x_layout(split.basis, self.nx, team_threads, h)
            tab['rhs_x_key'] = ('2pts', id(split), team_threads, h)
            rhs_t = np.empty((len(t), 2 * q))
            rhs_t[:, :q] = split.coefficients(t, scale=s1)              # c1 b(x, t_i)
            rhs_t[:, q:] = split.coefficients(t + self.dtau, scale=s2)  # c2 b(x, t_i + dtau)
            tab['rhs_t'] = rhs_t
        return tab
