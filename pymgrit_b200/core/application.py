"""Application API of the reference (core/application.py:17-107), plus the device-side contract.

`Application` is the abstract plugin class, unchanged: time grid from (t_start, t_stop, nt) or
t_interval, required attributes vector_template / vector_t_start, abstract step().
`DeviceApplication` adds what the batched engine needs: which Phi kernel family the application
maps to and the per-level tables of that kernel (built on the host, uploaded once per level).
"""
from abc import ABCMeta, abstractmethod

import numpy as np

from pymgrit_b200.core.vector import Vector


class _Required:
    """Descriptor for an attribute every application has to assign in its constructor (stored as `_<name>`).  Reading it
    before it was assigned raises AttributeError, which is what the check after construction looks for."""

    def __set_name__(self, owner, name):
        self.name, self.slot = name, '_' + name

    def __get__(self, obj, objtype=None):
        if obj is None:
            return self
        try:
            return obj.__dict__[self.slot]
        except KeyError:
            raise AttributeError(self.name) from None

    def __set__(self, obj, value):
        obj.__dict__[self.slot] = value


class MetaApplication(ABCMeta):
    """After an application object has been built, every name in `required_attributes` must exist
    (core/application.py:17-29 raises ValueError otherwise)."""
    required_attributes = []

    def __call__(cls, *args, **kwargs):
        obj = super().__call__(*args, **kwargs)
        missing = [name for name in obj.required_attributes if not hasattr(obj, name)]
        if missing:
            raise ValueError('required attribute (%s) not set' % missing[0])
        return obj


_INCREASING = {}         # id(array) -> weak reference: time grids made by linspace() below with t_stop > t_start


def known_increasing(t) -> bool:
    """True if `t` is a grid this module built with increasing values, or a forward slice of one (t[::m], what
    simple_setup_problem and the reference's examples hand to the coarse levels): lets the solver skip an O(nt) pass."""
    for arr in (t, getattr(t, 'base', None)):
        if arr is None:
            continue
        ref = _INCREASING.get(id(arr))
        if ref is not None and ref() is arr:
            return arr is t or (t.ndim == 1 and t.strides[0] > 0)
    return False


def linspace(t_start, t_stop, nt):
    """np.linspace(t_start, t_stop, nt), bit for bit (i * step + t_start, last point = t_stop), filled in pieces by the
    table threads when the grid is long: at nt = 2^20 NumPy's three passes over 8 MB are a third of the host time of the
    setup."""
    import weakref
    nt = int(nt)
    if nt < (1 << 17) or not np.isscalar(t_start) or not np.isscalar(t_stop):
        t = np.linspace(t_start, t_stop, nt)
    else:
        from pymgrit_b200.core.device_level import parallel_pieces
        start, stop = float(t_start), float(t_stop)
        step = (stop - start) / (nt - 1)
        if step == 0:
            t = np.linspace(t_start, t_stop, nt)
        else:
            t = np.empty(nt)
            try:
                from pymgrit_b200 import _lib
                from pymgrit_b200.core.device_level import host_threads
                native = _lib.lib().mgb_host_affine_ramp(start, step, nt, t.ctypes.data, host_threads()) == 0
            except Exception:
                native = False

            def piece(a, b):
                np.multiply(np.arange(a, b, dtype=float), step, out=t[a:b])
                t[a:b] += start
            if not native:
                parallel_pieces(nt, piece)
            t[-1] = stop
    if nt > 1 and t[-1] > t[0]:
        if len(_INCREASING) > 64:
            for k in [k for k, r in _INCREASING.items() if r() is None]:
                del _INCREASING[k]
        try:
            _INCREASING[id(t)] = weakref.ref(t)
        except TypeError:
            pass
    return t


def _time_grid(t_start, t_stop, nt, t_interval):
    """(grid, first, last, count) from either form of the constructor arguments (core/application.py:45-68)."""
    if t_interval is not None:
        if not isinstance(t_interval, np.ndarray):
            raise Exception('t_interval has the wrong type. Should be a numpy array')
        return t_interval, t_interval[0], t_interval[-1], len(t_interval)
    if t_start is None or t_stop is None or nt is None:
        raise Exception('Specify an interval by t_start, t_stop and nt or by t_interval')
    return linspace(t_start, t_stop, nt), t_start, t_stop, nt


class Application(object, metaclass=MetaApplication):
    """Time integrator plug-in: a time grid, a template vector, the state at t_start and step()."""
    required_attributes = ['vector_template', 'vector_t_start']
    vector_template = _Required()        # a Vector that clone_zero() / clone_rand() can be called on
    vector_t_start = _Required()         # the solution at the first time point

    def __init__(self, t_start: float = None, t_stop: float = None, nt: int = None,
                 t_interval: np.ndarray = None) -> None:
        self.t, self.t_start, self.t_end, self.nt = _time_grid(t_start, t_stop, nt, t_interval)

    @abstractmethod
    def step(self, u_start: Vector, t_start: float, t_stop: float) -> Vector:
        """Time integration from t_start to t_stop."""


class DeviceApplication(Application):
    """An application whose Phi exists as a kernel family of libmgrit_b200.

    Subclasses set
      kind          MGB_APP_* constant
      ndof          spatial unknowns per time point
    and implement level_tables(t), the host arrays of struct mgb_level for a time grid t.
    step() is provided here: one launch of mgb_step on a two-point level.
    """
    kind = 0
    ndof = 0

    def level_tables(self, t: np.ndarray, team_threads: int, chunk: int) -> dict:
        raise NotImplementedError

    # values <-> level rows.  rows: [count, pitch] tensor; values: [count, *vector shape] tensor.  The 1-D and ODE
    # applications store the values themselves at the start of a row.
    def rows_to_values(self, rows):
        shape = tuple(self.vector_template.shape)
        return rows[:, :self.ndof].view((rows.shape[0],) + shape)

    def values_to_rows(self, values, rows) -> None:
        rows[:, :self.ndof].copy_(values.reshape(rows.shape[0], self.ndof))

    def step(self, u_start, t_start: float, t_stop: float):
        from pymgrit_b200.core.device_level import single_step
        return single_step(self, u_start, t_start, t_stop)

    def __getstate__(self):          # simple_setup_problem deep-copies applications
        state = dict(self.__dict__)
        state.pop('_step_cache', None)
        return state
