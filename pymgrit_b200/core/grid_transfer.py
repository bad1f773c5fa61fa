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
