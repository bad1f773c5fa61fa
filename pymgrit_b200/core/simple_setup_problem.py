"""Level hierarchy by temporal coarsening (the reference's helper, core/simple_setup_problem.py:15-43): level l is a deep
copy of the finest application restricted to every coarsening**l-th time point."""
import copy
import warnings
from typing import List

from pymgrit_b200.core.application import Application

_ONE_POINT = ("This choice leads to a coarsest grid with only one time point, which is the initial point. "
              "It is recommended to choose a structure with at least two points on the coarsest grid.")


def _on_grid(problem: Application, t) -> Application:
    """A deep copy of `problem` whose time grid is t (device tables are rebuilt per level, not copied)."""
    twin = copy.deepcopy(problem)
    twin.t, twin.nt = t, len(t)
    twin.t_start, twin.t_end = t[0], t[-1]
    return twin


def simple_setup_problem(problem: Application, level: int, coarsening: int) -> List[Application]:
    if len(problem.t[::coarsening * level]) == 1:        # the reference's own test for the warning
        warnings.warn(_ONE_POINT)
    grids = [problem.t]
    while len(grids) < level:
        grids.append(grids[-1][::coarsening])
    return [problem] + [_on_grid(problem, t) for t in grids[1:]]
