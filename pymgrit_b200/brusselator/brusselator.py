"""Brusselator system with the reference's interface (brusselator/brusselator.py:14-132) on the GPU:
x' = A + x^2 y - (B + 1) x,  y' = B x - x^2 y,  A = 1, B = 3, classical RK4."""
import numpy as np

from pymgrit_b200 import _lib
from pymgrit_b200.core.application import DeviceApplication
from pymgrit_b200.core.vector import DeviceVector


class VectorBrusselator(DeviceVector):
    def __init__(self, tensor=None):
        super().__init__((2,), tensor)

    @property
    def value(self):
        return self.get_values()


class Brusselator(DeviceApplication):
    kind = _lib.APP_BRUSSELATOR
    ndof = 2

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.vector_template = VectorBrusselator()
        self.vector_t_start = VectorBrusselator()
        self.vector_t_start.set_values(np.array([0.0, 1.0]))
        self.a = 1
        self.b = 3

    def level_tables(self, t, team_threads, chunk):
        return dict()
