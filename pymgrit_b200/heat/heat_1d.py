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
        # device representation of rhs(x, t), analysed once here so that the deep copies made by
        # simple_setup_problem share it
        self._rhs_split = RhsSplit(self.rhs, self.x).analyse(self.t)

    def level_tables(self, t, team_threads, chunk):
        fac = self.a / self.dx ** 2                                   # heat_1d.py:185
        dts, dtidx = dl.dt_classes(t)
        tab = dict(ndt=len(dts), dtidx=dtidx,
                   sconst=dl.step_const_table(self.kind, dts * fac, self.nx, team_threads, chunk))
        tab['cw'] = tab['sconst'].shape[1]
        split = self._rhs_split
        dt_full = np.zeros(len(t))
        dt_full[1:] = np.diff(t)
        if split.kind == 'separable':
            tab['nrhs'] = split.basis.shape[0]
            tab['rhs_x'] = dl.rhs_x_layout(split.basis, self.nx, team_threads, chunk)
            tab['rhs_t'] = split.coefficients(t) * dt_full[:, None]  # b * dt, heat_1d.py:214
        elif split.kind == 'dense':
            tab['rhs_dense'] = split.dense(t) * dt_full[:, None]
        return tab
