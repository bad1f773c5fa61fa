"""Vector of two consecutive time points with the reference's interface (heat/vector_heat_1d_2pts.py:12-140), in HBM.

The reference keeps two ndarrays and the spacing dtau; here the pair is one (2, size) float64 CUDA tensor (row 0 = first
time point, row 1 = second), so +, -, * and norm() are the DeviceVector launches: the norm of the pair is the 2-norm of
both halves appended (vector_heat_1d_2pts.py:68-74) = the 2-norm of the whole tensor.
"""
import numpy as np

from pymgrit_b200.core.vector import DeviceVector


class VectorHeat1D2Pts(DeviceVector):
    def __init__(self, size, dtau, tensor=None):
        super().__init__((2, int(size)), tensor)
        self.size_per_point = int(size)
        self.dtau = dtau
        self._lazy_second = None

    # the reference's `size` is the number of spatial unknowns of ONE time point (vector_heat_1d_2pts.py:26)
    @property
    def size(self):
        return self.size_per_point

    @property
    def device_values(self):
        if self._lazy_second is not None:          # second time point still to be produced on the device
            fn, self._lazy_second = self._lazy_second, None
            first = np.array(self._host[0], dtype=float)
            second = fn(first)
            torch = _torch()
            self.values = torch.stack([torch.as_tensor(first).to(second.device), second.reshape(-1)]).contiguous()
            self._host = None
        return DeviceVector.device_values.fget(self)

    def set_lazy_second(self, first, fn):
        """First time point given on the host; the second one is fn(first) -> CUDA tensor, evaluated on first use (so
        that constructing an application needs no device)."""
        self._host = np.zeros(self.shape)
        self._host[0] = np.asarray(first, dtype=float)
        self.values = None
        self._lazy_second = fn

    def _new(self, tensor=None):
        out = super()._new(tensor)
        out._lazy_second = None
        return out

    def clone(self):
        self.device_values if self._lazy_second is not None else None
        return super().clone()

    def clone_rand(self):
        out = self._new()
        out._host = np.stack([np.random.rand(self.size_per_point), np.random.rand(self.size_per_point)])
        return out

    def set_values(self, first_time_point, second_time_point=None, dtau=None):
        """set_values(first, second, dtau) as in the reference (vector_heat_1d_2pts.py:113-123); a single (2, size)
        array or CUDA tensor is accepted too (what the engine hands back)."""
        self._lazy_second = None
        if second_time_point is None:
            super().set_values(first_time_point)
            return
        torch = _torch()
        if isinstance(first_time_point, torch.Tensor) and isinstance(second_time_point, torch.Tensor):
            super().set_values(torch.stack([first_time_point.reshape(-1), second_time_point.reshape(-1)]))
        else:
            super().set_values(np.stack([_host(first_time_point), _host(second_time_point)]))
        if dtau is not None:
            self.dtau = dtau

    def get_values(self):
        """(values at the first time point, values at the second time point, dtau)."""
        if self._lazy_second is not None:
            self.device_values
        vals = super().get_values()
        return vals[0], vals[1], self.dtau

    def pack(self):
        first, second, _ = self.get_values()
        return np.array([first, second])

    def unpack(self, values):
        self.set_values(values[0], values[1], self.dtau)

    def __getstate__(self):
        state = dict(self.__dict__)
        if self._lazy_second is None or self.values is not None:
            state['_host'] = np.array(DeviceVector.get_values(self), copy=True)
        else:
            state['_host'] = np.array(self._host, copy=True)      # lazy second point: the closure travels with the copy
        state['values'] = None
        return state


def _host(a):
    torch = _torch()
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy().reshape(-1)
    return np.asarray(a, dtype=float).reshape(-1)


def _torch():
    import torch
    return torch
