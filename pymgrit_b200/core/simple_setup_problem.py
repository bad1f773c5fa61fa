"""simple_setup_problem of the reference (core/simple_setup_problem.py:15-43): deep copies of the
finest problem on every `coarsening`-th time point."""
import copy
import warnings
from typing import List

from pymgrit_b200.core.application import Application


def simple_setup_problem(problem: Application, level: int, coarsening: int) -> List[Application]:
    if len(problem.t[::coarsening * level]) == 1:
        warnings.warn(
            "This choice leads to a coarsest grid with only one time point, which is the initial point. "
            "It is recommended to choose a structure with at least two points on the coarsest grid.")
    hierarchy = [problem]
    for _ in range(level - 1):
        t_coarse = hierarchy[-1].t[::coarsening]
        nxt = copy.deepcopy(problem)
        nxt.t = t_coarse
        nxt.nt = len(t_coarse)
        nxt.t_start, nxt.t_end = t_coarse[0], t_coarse[-1]
        hierarchy.append(nxt)
    return hierarchy
