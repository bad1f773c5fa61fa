"""Identity transfer (core/grid_transfer_copy.py:25-47): the only transfer the batched engine fuses
into its restriction / correction kernels (injection in time, copy in space)."""
from pymgrit_b200.core.grid_transfer import GridTransfer
from pymgrit_b200.core.vector import Vector


class GridTransferCopy(GridTransfer):
    def __init__(self):
        super().__init__()

    def restriction(self, u: Vector) -> Vector:
        return u.clone()

    def interpolation(self, u: Vector) -> Vector:
        return u.clone()
