"""Vector API of the reference (core/vector.py:19-151) with the data held in HBM.

`Vector` is the abstract interface, unchanged.  `DeviceVector` is the GPU base class every
application vector of this package derives from: one float64 torch tensor on the CUDA device, all
arithmetic done by the C ABI (mgb_vec_axpby / mgb_vec_sumsq), no host round trip except
`get_values()` / `pack()`, which return NumPy data like the reference does.
"""
from abc import ABC, abstractmethod

import numpy as np


class Vector(ABC):
    """Abstract vector class (same abstract methods as core/vector.py:38-110)."""

    def __init__(self):
        pass

    @abstractmethod
    def __add__(self, other): ...

    @abstractmethod
    def __sub__(self, other): ...

    @abstractmethod
    def __mul__(self, other): ...

    @abstractmethod
    def norm(self): ...

    @abstractmethod
    def clone(self): ...

    @abstractmethod
    def clone_zero(self): ...

    @abstractmethod
    def clone_rand(self): ...

    @abstractmethod
    def set_values(self, values): ...

    @abstractmethod
    def get_values(self): ...

    @abstractmethod
    def pack(self): ...

    @abstractmethod
    def unpack(self, values): ...

    # derived operators, core/vector.py:113-151
    def __rmul__(self, other):
        return self * other

    def __imul__(self, other):
        return self * other

    def __iadd__(self, other):
        return self + other

    def __isub__(self, other):
        return self - other


def _torch():
    import torch
    return torch


class DeviceVector(Vector):
    """A vector whose values live in a CUDA tensor of shape `shape` (float64)."""

    def __init__(self, shape, tensor=None):
        super().__init__()
        self.shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        self._host = None
        if tensor is not None:
            self.values = tensor
        else:
            self._host = np.zeros(self.shape)      # materialised on the device on first use
            self.values = None

    # -- storage ---------------------------------------------------------------------------------
    def _new(self, tensor=None):
        out = self.__class__.__new__(self.__class__)
        out.__dict__.update({k: v for k, v in self.__dict__.items() if k not in ('values', '_host')})
        out.values, out._host = tensor, None
        if tensor is None:
            out._host = np.zeros(self.shape)
        return out

    @property
    def device_values(self):
        """The CUDA tensor (created from pending host data on first access)."""
        if self.values is None:
            torch = _torch()
            if not torch.cuda.is_available():
                raise Exception('pymgrit_b200 needs a CUDA device: there is no CPU fallback')
            self.values = torch.as_tensor(np.ascontiguousarray(self._host, dtype=np.float64)).cuda()
            self._host = None
        return self.values

    @property
    def _numel(self):
        return int(np.prod(self.shape)) if self.shape else 1

    @property
    def size(self):
        return self._numel

    # -- arithmetic (C ABI) ------------------------------------------------------------------------
    def _axpby(self, a, other, b):
        from pymgrit_b200 import _lib
        torch = _torch()
        x = self.device_values.contiguous()
        y = other.device_values.contiguous() if other is not None else None
        out = torch.empty_like(x)
        _lib.check(_lib.lib().mgb_vec_axpby(self._numel, a, x.data_ptr(), b, y.data_ptr() if y is not None else None,
                                            out.data_ptr(), _lib.current_stream_ptr()), 'vec_axpby')
        return self._new(out)

    def __add__(self, other):
        return self._axpby(1.0, other, 1.0)

    def __sub__(self, other):
        return self._axpby(1.0, other, -1.0)

    def __mul__(self, other):
        return self._axpby(float(other), None, 0.0)

    def norm(self):
        from pymgrit_b200 import _lib
        torch = _torch()
        x = self.device_values.contiguous()
        out = torch.empty(1, dtype=torch.float64, device=x.device)
        _lib.check(_lib.lib().mgb_vec_sumsq(self._numel, x.data_ptr(), out.data_ptr(), _lib.current_stream_ptr()),
                   'vec_sumsq')
        return float(np.sqrt(out.item()))

    # -- clone family ------------------------------------------------------------------------------
    def clone(self):
        if self.values is None:
            out = self._new()
            out._host = np.array(self._host, dtype=float, copy=True)
            return out
        return self._new(self.values.clone())

    def clone_zero(self):
        return self._new()

    def clone_rand(self):
        out = self._new()
        out._host = np.random.rand(*self.shape) if self.shape else np.array(np.random.rand())
        return out

    # -- host access -------------------------------------------------------------------------------
    def set_values(self, values):
        torch = _torch()
        if isinstance(values, torch.Tensor):
            self.values, self._host = values.to(dtype=torch.float64).reshape(self.shape), None
        else:
            self._host = np.array(values, dtype=float).reshape(self.shape)
            self.values = None

    def get_values(self):
        if self.values is None:
            return self._host
        return self.values.detach().cpu().numpy().reshape(self.shape)

    def pack(self):
        return self.get_values()

    def unpack(self, values):
        self.set_values(values)

    def __getstate__(self):          # pickling / deepcopy: carry host data only
        state = dict(self.__dict__)
        state['_host'] = np.array(self.get_values(), copy=True)
        state['values'] = None
        return state
