"""Problem data shared by the examples: the 1-D heat equation of examples/example_heat_1d.py:20-41 of the reference,
u_t - u_xx = -sin(pi x)(sin t - pi^2 cos t), u(x, 0) = sin(pi x), exact solution sin(pi x) cos t."""
import numpy as np


def rhs(x, t):
    return -np.sin(np.pi * x) * (np.sin(t) - np.pi ** 2 * np.cos(t))


def init_cond(x):
    return np.sin(np.pi * x)
