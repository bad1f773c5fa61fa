"""pymgrit_b200 -- the MGRIT hot path of pymgrit as batched sm_100a CUDA sweeps behind the reference's API.

    from pymgrit_b200 import Heat1D, simple_setup_problem, Mgrit
    info = Mgrit(problem=simple_setup_problem(Heat1D(...), level=3, coarsening=4), tol=1e-10).solve()

Same names and arguments as `pymgrit` (reference: src/pymgrit/__init__.py); see DESIGN.md for what is on the device.
"""
from pymgrit_b200.core.application import Application, DeviceApplication
from pymgrit_b200.core.vector import Vector, DeviceVector
from pymgrit_b200.core.grid_transfer import GridTransfer, DeviceGridTransfer
from pymgrit_b200.core.grid_transfer_copy import GridTransferCopy
from pymgrit_b200.core.simple_setup_problem import simple_setup_problem
from pymgrit_b200.core.mgrit import Mgrit
from pymgrit_b200.core.at_mgrit import AtMgrit
from pymgrit_b200.heat.heat_1d import Heat1D, VectorHeat1D
from pymgrit_b200.heat.heat_2d import Heat2D, VectorHeat2D
from pymgrit_b200.heat.grid_transfer_heat_1d import GridTransferHeat1D, GridTransferHeat
from pymgrit_b200.heat.vector_heat_1d_2pts import VectorHeat1D2Pts
from pymgrit_b200.heat.heat_1d_2pts_bdf1 import Heat1DBDF1
from pymgrit_b200.heat.heat_1d_2pts_bdf2 import Heat1DBDF2
from pymgrit_b200.advection.advection_1d import Advection1D, VectorAdvection1D
from pymgrit_b200.dahlquist.dahlquist import Dahlquist, VectorDahlquist
from pymgrit_b200.brusselator.brusselator import Brusselator, VectorBrusselator
from pymgrit_b200.allen_cahn.allen_cahn import AllenCahn, VectorAllenCahn2D
from pymgrit_b200.core.batched import BatchedApplication
from pymgrit_b200.core.split import split_communicator

__version__ = '0.1.0'
