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
