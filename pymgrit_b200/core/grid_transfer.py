"""GridTransfer API of the reference (core/grid_transfer.py:15-55)."""
from abc import ABC, abstractmethod

from pymgrit_b200.core.vector import Vector


class GridTransfer(ABC):
    def __init__(self):
        pass

    @abstractmethod
    def restriction(self, u: Vector) -> Vector:
        """Restrict u to the next coarser spatial grid."""

    @abstractmethod
    def interpolation(self, u: Vector) -> Vector:
        """Interpolate u to the next finer spatial grid."""


class DeviceGridTransfer(GridTransfer):
    """A spatial grid transfer that exists as row-wise kernels of libmgrit_b200, so that the solver can apply it to all
    C-points of a level in one launch.  A transfer written in Python on `Vector.get_values()` (how users of the reference
    write theirs) cannot run inside a sweep; pymgrit_b200.Mgrit raises for those.

    Subclasses implement (rows are level arrays [points][pitch] on the device, see include/mgrit_b200.h):
      check(fine_app, coarse_app)                                          raise if the two levels do not fit
      restrict_rows(nrows, src, src_index, dst, fine_app)                  dst[j] = R(src[src_index[j]])
      interpolate_rows(nrows, first, a, b, dst, dst_index, accumulate, coarse_app)
                                                                           dst[dst_index[j]] (+)= P(a[j] - b[j]), j >= first
    """

    def check(self, fine_app, coarse_app) -> None:
        raise NotImplementedError

    def restrict_rows(self, nrows, src, src_index, dst, fine_app) -> None:
        raise NotImplementedError

    def interpolate_rows(self, nrows, first, a, b, dst, dst_index, accumulate, coarse_app) -> None:
        raise NotImplementedError
