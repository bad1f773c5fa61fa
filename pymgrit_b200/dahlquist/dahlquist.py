"""Dahlquist test problem with the reference's interface (dahlquist/dahlquist.py:12-111) on the GPU."""
import numpy as np

from pymgrit_b200 import _lib
from pymgrit_b200.core.application import DeviceApplication
from pymgrit_b200.core.vector import DeviceVector


class VectorDahlquist(DeviceVector):
    """Scalar state (dahlquist.py:12-58)."""

    def __init__(self, value=0.0, tensor=None):
        super().__init__((), tensor)
        if tensor is None:
            self.set_values(value)

    @property
    def value(self):
        return float(self.get_values())

    def clone_zero(self):
        return VectorDahlquist(0.0)

    def clone_rand(self):
        return VectorDahlquist(np.random.rand(1)[0])


class Dahlquist(DeviceApplication):
    kind = _lib.APP_DAHLQUIST
    ndof = 1

    def __init__(self, constant_lambda=-1, method='BE', *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.vector_template = VectorDahlquist(0)
        self.vector_t_start = VectorDahlquist(1)
        self.lambda_value = constant_lambda
        if method in ('BE', 'FE', 'TR', 'MR'):
            self.method = method
        else:
            raise Exception('Unknown method. Choose BE (Backward Euler), FE (Forward Euler), TR (Trapezoidal rule) ' +
                            'or MR (implicit mid-point rule)')

    def level_tables(self, t, team_threads, chunk):
        return dict(p=[float(self.lambda_value)], ip=[_lib.DAHLQUIST_METHODS[self.method]])
