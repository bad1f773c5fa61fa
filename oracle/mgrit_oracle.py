"""CPU oracle for the MGRIT hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain numpy/scipy restatement of the serial (one time rank) algorithm of the reference
``pymgrit.core.mgrit.Mgrit`` and of the time integrators ``Application.step`` that lie on the path
named by BASELINE.json.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module; ``pymgrit_b200`` never does.

Parity status: PINNED.  ``tests/test_oracle.py`` checks this restatement against
  * the reference's own known-answer vectors (tests/heat/test_heat_1d.py:31-42,
    tests/heat/test_heat_2d.py:230-249, tests/advection/test_advection_1d.py:33-45,
    tests/dahlquist/test_dahlquist.py:55-62, tests/brusselator/test_brusselator.py),
  * the reference's whole-solver values (tests/core/test_mgrit.py:59-70, tests/mpi/results/*), and
  * outputs of the unmodified reference run in the build container (tests/golden/*.npz, made by
    tests/golden/make_golden.py with the import stubs in oracle/stubs/).

Every level is held as one ndarray ``u[l]`` of shape (N_l, *vector_shape) instead of a list of
Vector objects; all arithmetic follows the order of operations of the reference lines cited.
"""
from __future__ import annotations

import ctypes
import os
import time
from typing import Callable, List, Optional, Sequence

import numpy as np
from scipy import sparse as sp
from scipy.sparse.linalg import spsolve

_HERE = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------------------------
# Optional C accelerator for Phi (oracle/phi_oracle.c, built by oracle/Makefile).  Used only to
# make large parity cases finish in seconds; the default arithmetic is scipy's SuperLU like the
# reference.
# --------------------------------------------------------------------------------------------
_CLIB = None


def c_lib():
    """Load oracle/_build/libphi_oracle.so (None if it was not built)."""
    global _CLIB
    if _CLIB is None:
        path = os.path.join(_HERE, "_build", "libphi_oracle.so")
        if not os.path.exists(path):
            return None
        lib = ctypes.CDLL(path)
        dp = ctypes.POINTER(ctypes.c_double)
        lib.oracle_heat1d_be.argtypes = [ctypes.c_int, ctypes.c_double, dp, dp, dp]
        lib.oracle_heat1d_be.restype = None
        lib.oracle_advection1d_be.argtypes = [ctypes.c_int, ctypes.c_double, dp, dp]
        lib.oracle_advection1d_be.restype = None
        _CLIB = lib
    return _CLIB


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


# --------------------------------------------------------------------------------------------
# Problems (Phi).  Each has: t (time grid), u0 (state at t[0]), phi(u, t_start, t_stop), norm(v)
# --------------------------------------------------------------------------------------------
class OracleProblem:
    """Time grid as in core/application.py:45-68."""

    def __init__(self, t_start=None, t_stop=None, nt=None, t_interval=None):
        if t_interval is None:
            if t_start is None or t_stop is None or nt is None:
                raise Exception('Specify an interval by t_start, t_stop and nt or by t_interval')
            self.t = np.linspace(t_start, t_stop, nt)
        else:
            self.t = np.asarray(t_interval, dtype=float)
        self.nt = len(self.t)

    def coarsen(self, t_new):
        """Same problem on another time grid (core/simple_setup_problem.py:34-41)."""
        import copy
        other = copy.copy(self)
        other.t = np.asarray(t_new, dtype=float)
        other.nt = len(other.t)
        return other

    @staticmethod
    def norm(v) -> float:
        return float(np.linalg.norm(v))

    def phi(self, u, t_start, t_stop):
        raise NotImplementedError


class Heat1DOracle(OracleProblem):
    """heat/heat_1d.py:131-217: backward Euler, 2nd-order central differences, zero Dirichlet."""

    def __init__(self, x_start, x_end, nx, a, init_cond=lambda x: x * 0, rhs=lambda x, t: x * 0,
                 solver='spsolve', **kw):
        super().__init__(**kw)
        x = np.linspace(x_start, x_end, nx)          # heat_1d.py:154-157
        self.x = x[1:-1]
        self.n = nx - 2
        self.dx = self.x[1] - self.x[0]
        self.a = a
        fac = a / self.dx ** 2                         # heat_1d.py:185-194
        self.fac = fac
        self.L = sp.diags([np.ones(self.n) * 2 * fac, np.ones(self.n - 1) * -fac, np.ones(self.n - 1) * -fac],
                          [0, -1, 1], shape=(self.n, self.n), format='csr')
        self.I = sp.identity(self.n, dtype='float', format='csr')
        self.rhs = rhs
        self.u0 = np.asarray(init_cond(self.x), dtype=float)   # heat_1d.py:173-175
        self.solver = solver

    def phi(self, u, t_start, t_stop):
        dt = t_stop - t_start
        b = u + self.rhs(self.x, t_stop) * dt          # heat_1d.py:214 (order: rhs*dt, then +u)
        if self.solver == 'spsolve':
            return spsolve(dt * self.L + self.I, b)    # heat_1d.py:213
        lib = c_lib()
        if lib is None:
            raise RuntimeError("oracle C library not built (make -C oracle)")
        b = np.ascontiguousarray(b, dtype=float)
        out = np.empty_like(b)
        work = np.empty_like(b)
        lib.oracle_heat1d_be(self.n, dt * self.fac, _dptr(b), _dptr(out), _dptr(work))
        return out


class Heat1D2PtsOracle(OracleProblem):
    """heat/heat_1d_2pts_bdf1.py:16-117 (method='BDF1') and heat/heat_1d_2pts_bdf2.py:17-138 (method='BDF2'): the state
    of a time point is the PAIR of values at t and t + dtau (heat/vector_heat_1d_2pts.py:12-140), held here as one
    (2, n) array, so the vector norm (2-norm of both halves appended, vector_heat_1d_2pts.py:68-74) is the Frobenius
    norm of that array."""

    def __init__(self, x_start, x_end, nx, dtau, a, init_cond=lambda x: x * 0, rhs=lambda x, t: x * 0,
                 method='BDF2', **kw):
        super().__init__(**kw)
        if method not in ('BDF1', 'BDF2'):
            raise Exception('Unknown method')
        x = np.linspace(x_start, x_end, nx)
        self.x = x[1:-1]
        self.n = nx - 2
        self.dx = self.x[1] - self.x[0]
        self.a = a
        self.dtau = dtau
        fac = a / self.dx ** 2
        self.L = sp.diags([np.ones(self.n) * 2 * fac, np.ones(self.n - 1) * -fac, np.ones(self.n - 1) * -fac],
                          [0, -1, 1], shape=(self.n, self.n), format='csr')
        self.I = sp.identity(self.n, dtype='float', format='csr')
        self.rhs = rhs
        self.method = method
        tmp1 = np.asarray(init_cond(self.x), dtype=float)
        if method == 'BDF1':                             # one BDF1 step, heat_1d_2pts_bdf1.py:64-66
            tmp2 = spsolve(dtau * self.L + self.I, tmp1 + self.rhs(self.x, self.t[0] + dtau) * dtau)
        else:                                            # trapezoidal rule, heat_1d_2pts_bdf2.py:65-68
            tmp2 = spsolve((dtau / 2) * self.L + self.I,
                           (self.I - (dtau / 2) * self.L) * tmp1 +
                           (dtau / 2) * (self.rhs(self.x, self.t[0]) + self.rhs(self.x, self.t[0] + dtau)))
        self.u0 = np.stack([tmp1, tmp2])

    def phi(self, u, t_start, t_stop):
        first, second, dtau = u[0], u[1], self.dtau
        if self.method == 'BDF1':                        # heat_1d_2pts_bdf1.py:108-113
            tmp1 = spsolve((t_stop - t_start - dtau) * self.L + self.I,
                           second + self.rhs(self.x, t_stop) * (t_stop - t_start - dtau))
            tmp2 = spsolve(dtau * self.L + self.I, tmp1 + self.rhs(self.x, t_stop + dtau) * dtau)
            return np.stack([tmp1, tmp2])
        # BDF2 on the variably spaced grid, heat_1d_2pts_bdf2.py:110-133
        tau_i = t_stop - t_start - dtau
        tau_im1 = dtau
        r_i = tau_i / tau_im1
        coeffm2 = (r_i ** 2) / (tau_i * (1 + r_i))
        coeffm1 = (1 + r_i) / tau_i
        coeff = (1 + 2 * r_i) / (tau_i * (1 + r_i))
        rhs = self.rhs(self.x, t_stop) - coeffm2 * first + coeffm1 * second
        tmp1 = spsolve(self.L + coeff * self.I, rhs)
        tau_im1 = tau_i
        tau_i = dtau
        r_i = tau_i / tau_im1
        coeffm2 = (r_i ** 2) / (tau_i * (1 + r_i))
        coeffm1 = (1 + r_i) / tau_i
        coeff = (1 + 2 * r_i) / (tau_i * (1 + r_i))
        rhs = self.rhs(self.x, t_stop + dtau) - coeffm2 * second + coeffm1 * tmp1
        tmp2 = spsolve(self.L + coeff * self.I, rhs)
        return np.stack([tmp1, tmp2])


class Heat2DOracle(OracleProblem):
    """heat/heat_2d.py:139-366: theta method (BE, CN, FE), Dirichlet values as given."""

    def __init__(self, x_start, x_end, y_start, y_end, nx, ny, a,
                 rhs=lambda x, y, t: 0 * x * y, init_cond=lambda x, y: x * y * 0, method='BE',
                 bc_left=0, bc_right=0, bc_bottom=0, bc_top=0, **kw):
        super().__init__(**kw)
        if method not in ('BE', 'FE', 'CN'):                     # heat_2d.py:192-200
            raise Exception("Unknown method. Choose BE (Backward Euler), FE (Forward Euler) or CN (Crank-Nicolson")
        self.theta = {'BE': 1, 'FE': 0, 'CN': 0.5}[method]
        self.x = np.linspace(x_start, x_end, nx)
        self.y = np.linspace(y_start, y_end, ny)
        self.x_2d = self.x[:, np.newaxis]
        self.y_2d = self.y[np.newaxis, :]
        self.nx, self.ny = nx, ny
        self.dx = self.x[1] - self.x[0]
        self.dy = self.y[1] - self.y[0]
        self.a = a
        self.rhs = rhs

        def as_fn(v):
            return v if callable(v) else (lambda s, _v=v: _v)
        self.bc_left, self.bc_right = as_fn(bc_left), as_fn(bc_right)
        self.bc_bottom, self.bc_top = as_fn(bc_bottom), as_fn(bc_top)
        self.L = self._matrix()
        self.I = sp.identity(nx * ny, dtype='float', format='csr')
        init = np.array(init_cond(self.x_2d, self.y_2d), dtype=float)     # heat_2d.py:243-248
        self._apply_bc(init)
        self.u0 = init

    def _apply_bc(self, b):                              # heat_2d.py:315-319 (same order of writes)
        b[:, 0] = self.bc_left(self.x)
        b[:, -1] = self.bc_right(self.x)
        b[-1, :] = self.bc_bottom(self.y)
        b[0, :] = self.bc_top(self.y)

    def _matrix(self):
        """5-point stencil with zero rows at boundary nodes (heat_2d.py:250-287); built row-wise."""
        nx, ny = self.nx, self.ny
        fx = self.a / self.dx ** 2
        fy = self.a / self.dy ** 2
        idx = np.arange(nx * ny).reshape(nx, ny)
        interior = idx[1:-1, 1:-1].ravel()
        rows = np.concatenate([interior] * 5)
        cols = np.concatenate([interior, interior - 1, interior + 1, interior - ny, interior + ny])
        vals = np.concatenate([np.full(interior.size, 2 * (fx + fy)), np.full(interior.size, -fy),
                               np.full(interior.size, -fy), np.full(interior.size, -fx),
                               np.full(interior.size, -fx)])
        return sp.csr_matrix((vals, (rows, cols)), shape=(nx * ny, nx * ny))

    def phi(self, u, t_start, t_stop):
        dt = t_stop - t_start
        xi, yi = self.x_2d[1:-1], self.y_2d[:, 1:-1]
        if self.theta == 0:                              # FE, heat_2d.py:341-358 (boundary values are ADDED, as there)
            new = np.zeros((self.nx, self.ny))
            self._apply_bc(new)
            new += ((self.I - dt * self.L) * u.flatten()).reshape(self.nx, self.ny)
            new[1:-1, 1:-1] += dt * self.rhs(x=xi, y=yi, t=t_start)
            return new
        b = np.zeros((self.nx, self.ny))
        if self.theta == 1:                              # BE, heat_2d.py:300-303
            b[1:-1, 1:-1] = u[1:-1, 1:-1] + dt * self.rhs(x=xi, y=yi, t=t_stop)
        else:                                            # CN, heat_2d.py:305-313
            b += ((self.I - self.theta * dt * self.L) * u.flatten()).reshape(self.nx, self.ny)
            b[1:-1, 1:-1] += self.theta * dt * self.rhs(x=xi, y=yi, t=t_stop) + \
                (1 - self.theta) * dt * self.rhs(x=xi, y=yi, t=t_start)
        self._apply_bc(b)                                # heat_2d.py:315-319
        new = spsolve(dt * self.theta * self.L + self.I, b.flatten())  # heat_2d.py:363
        return new.reshape(self.nx, self.ny)


class AllenCahnOracle(OracleProblem):
    """allen_cahn/allen_cahn.py:136-270, IMEX branch (allen_cahn.py:191-197): reaction explicit, periodic 5-point
    Laplacian implicit through a sparse direct solve."""

    def __init__(self, nx=128, nu=2, eps=0.04, radius=0.25, method='IMEX', **kw):
        super().__init__(**kw)
        if method != 'IMEX':
            raise Exception('the oracle restates the IMEX branch only')
        self.nx = self.ny = nx
        self.nu, self.eps, self.radius = nu, eps, radius
        self.dx = 1.0 / nx                                           # allen_cahn.py:167
        self.x = np.linspace(start=-0.5, stop=0.5, num=nx)           # allen_cahn.py:170
        one = sp.lil_matrix((nx, nx))                                # circulant second difference (allen_cahn.py:175-189)
        for i in range(nx):
            one[i, i] += -2.0
            one[i, (i + 1) % nx] += 1.0
            one[i, (i - 1) % nx] += 1.0
        one = one.tocsc()
        self.space_disc = (sp.kron(one, sp.eye(nx)) + sp.kron(sp.eye(nx), one)) * (1.0 / self.dx ** 2)
        self.id = sp.eye(nx * nx)
        r = np.sqrt(self.x[:, None] ** 2 + self.x[None, :] ** 2)     # allen_cahn.py:238-243
        self.u0 = np.tanh((radius - r) / (np.sqrt(2) * eps))

    def phi(self, u, t_start, t_stop):
        new = np.asarray(u, dtype=float).flatten()
        rhs = new + (t_stop - t_start) * (1 / self.eps ** 2 * new * (1.0 - new ** self.nu))      # allen_cahn.py:194
        new = spsolve((self.id - (t_stop - t_start) * self.space_disc).tocsc(), rhs)             # allen_cahn.py:195
        return new.reshape(self.nx, self.ny)


class Advection1DOracle(OracleProblem):
    """advection/advection_1d.py:68-143: implicit Euler + first-order upwind, periodic."""

    def __init__(self, c, x_start, x_end, nx, solver='spsolve', **kw):
        super().__init__(**kw)
        x = np.linspace(x_start, x_end, nx)
        self.x = x[0:-1]
        self.n = nx - 1
        self.dx = self.x[1] - self.x[0]
        self.c = c
        fac = c / self.dx
        self.fac = fac
        m = sp.diags([np.ones(self.n) * fac, np.ones(self.n) * -fac], [0, -1], shape=(self.n, self.n), format='lil')
        m[0, self.n - 1] = -fac                         # advection_1d.py:117-118
        self.L = sp.csr_matrix(m)
        self.I = sp.identity(self.n, dtype='float', format='csr')
        self.u0 = np.exp(-self.x ** 2)                  # advection_1d.py:122-127
        self.solver = solver

    def phi(self, u, t_start, t_stop):
        dt = t_stop - t_start
        if self.solver == 'spsolve':
            return spsolve(dt * self.L + self.I, u)     # advection_1d.py:140
        lib = c_lib()
        if lib is None:
            raise RuntimeError("oracle C library not built (make -C oracle)")
        b = np.ascontiguousarray(u, dtype=float)
        out = np.empty_like(b)
        lib.oracle_advection1d_be(self.n, dt * self.fac, _dptr(b), _dptr(out))
        return out


class DahlquistOracle(OracleProblem):
    """dahlquist/dahlquist.py:61-111."""

    def __init__(self, constant_lambda=-1, method='BE', **kw):
        super().__init__(**kw)
        if method not in ('BE', 'FE', 'TR', 'MR'):
            raise Exception('Unknown method')
        self.lam = constant_lambda
        self.method = method
        self.u0 = np.array(1.0)

    @staticmethod
    def norm(v) -> float:
        return float(np.linalg.norm(v))                 # dahlquist.py:36-37

    def phi(self, u, t_start, t_stop):
        z = (t_stop - t_start) * self.lam
        if self.method == 'BE':
            return 1 / (1 - z) * u
        if self.method == 'FE':
            return (1 + z) * u
        if self.method == 'TR':
            return (1 + z / 2) / (1 - z / 2) * u
        k1 = -1 / (1 - z / 2) * u                       # 'MR', dahlquist.py:107-109
        return u + (t_stop - t_start) * k1


class BrusselatorOracle(OracleProblem):
    """brusselator/brusselator.py:68-132: classical RK4 on the 2-component Brusselator."""

    def __init__(self, **kw):
        super().__init__(**kw)
        self.u0 = np.array([0.0, 1.0])

    @staticmethod
    def _f(y):
        a, b = 1, 3
        return np.array([a + (y[0] ** 2) * y[1] - (b + 1) * y[0], b * y[0] - (y[0] ** 2) * y[1]], dtype=float)

    def phi(self, u, t_start, t_stop):
        dt = t_stop - t_start
        k1 = self._f(u)
        k2 = self._f(u + dt / 2 * k1)
        k3 = self._f(u + dt / 2 * k2)
        k4 = self._f(u + dt * k3)
        return u + dt / 6 * (k1 + 2 * k2 + 2 * k3 + k4)


def simple_hierarchy(problem: OracleProblem, level: int, coarsening: int) -> List[OracleProblem]:
    """core/simple_setup_problem.py:15-43."""
    out = [problem]
    for _ in range(level - 1):
        out.append(problem.coarsen(out[-1].t[::coarsening]))
    return out


# --------------------------------------------------------------------------------------------
# Serial MGRIT (FAS) on arrays
# --------------------------------------------------------------------------------------------
class CopyTransfer:
    """core/grid_transfer_copy.py:25-47: the identity in both directions."""

    @staticmethod
    def restriction(u):
        return u

    @staticmethod
    def interpolation(u):
        return u


class Heat1DSpaceTransfer:
    """The spatial grid transfer of examples/example_spatial_coarsening.py:18-79 (user code of the reference's example,
    restated on arrays): full-weighting restriction and linear interpolation between 1-D grids of 2 n + 1 and n interior
    points with homogeneous Dirichlet boundaries.  Sums are taken in the example's order."""

    @staticmethod
    def restriction(u):
        return u[0:-2:2] * 1 / 4 + u[1:-1:2] * 1 / 2 + u[2::2] * 1 / 4

    @staticmethod
    def interpolation(u):
        out = np.zeros(2 * len(u) + 1)
        out[0:-2:2] += 1 / 2 * u
        out[1:-1:2] += u
        out[2::2] += 1 / 2 * u
        return out


class MgritOracle:
    """One-rank restatement of core/mgrit.py.  transfer: one object per level pair with restriction(array) /
    interpolation(array) (core/grid_transfer.py:31-55); default = the identity (core/grid_transfer_copy.py)."""

    def __init__(self, problem: Sequence[OracleProblem], weight_c: float = 1.0, max_iter: int = 100,
                 tol: float = 1e-7, nested_iteration: bool = True, cf_iter=1, cycle_type: str = 'V',
                 t_norm: int = 2, conv_crit: int = 0, phi_counter: Optional[list] = None, transfer=None):
        if cycle_type not in ('V', 'F'):
            raise Exception("Cycle-type " + str(cycle_type) + " is not implemented. Choose 'V' or 'F'")
        if t_norm not in (1, 2, 3):
            raise Exception('Unknown norm.')
        if conv_crit not in (0, 1, 2, 3):
            raise Exception('Unknown convergence criterion')
        # one time rank: the local criteria 2 / 3 (mgrit.py:434-454) stop on the same per-point norms as 0 / 1
        conv_crit = conv_crit % 2
        t0 = time.time()
        self.problem = list(problem)
        self.L = len(self.problem)
        self.transfer = list(transfer) if transfer is not None else [CopyTransfer() for _ in range(self.L - 1)]
        if len(self.transfer) != self.L - 1:
            raise Exception('There should be exactly one transfer operator for each level except the coarsest grid')
        self.cf_iter = [cf_iter] * self.L if isinstance(cf_iter, int) else list(cf_iter)
        self.weight_c = weight_c
        self.tol = tol
        self.max_iter = max_iter
        self.cycle_type = cycle_type
        self.ord = {1: 1, 2: None, 3: np.inf}[t_norm]            # mgrit.py:182
        self.conv_crit = conv_crit
        self.conv = np.zeros(max_iter + 1)
        self.nphi = phi_counter if phi_counter is not None else [0] * self.L
        self.t = [np.copy(p.t) for p in self.problem]
        # C-point index tables (mgrit.py:212, 768-770): level-l points that are also level-(l+1) points
        self.cpts = []
        for l in range(self.L):
            if l < self.L - 1:
                self.cpts.append(np.where(np.isin(self.t[l], self.t[l + 1]))[0])
            else:
                self.cpts.append(np.arange(len(self.t[l])))
        self.u = [np.zeros((len(self.t[l]),) + np.shape(self.problem[l].u0)) for l in range(self.L)]    # mgrit.py:846-858
        self.g = [None] + [np.zeros_like(self.u[l]) for l in range(1, self.L)]
        self.v = [None] + [np.zeros_like(self.u[l]) for l in range(1, self.L)]
        for l in range(self.L):
            self.u[l][0] = self.problem[l].u0
        if nested_iteration:
            self.nested_iteration()
        if conv_crit == 1:
            self.last = self.u[0].copy()
        self.time_setup = time.time() - t0

    # -- Phi wrapper ----------------------------------------------------------------------
    def phi(self, l, u, i):
        """Phi_l from point i-1 to point i of level l."""
        self.nphi[l] += 1
        return self.problem[l].phi(u, self.t[l][i - 1], self.t[l][i])

    def _fpts(self, l):
        is_c = np.zeros(len(self.t[l]), dtype=bool)
        is_c[self.cpts[l]] = True
        return np.where(~is_c)[0]

    # -- sweeps ---------------------------------------------------------------------------
    def f_relax(self, l):                                # mgrit.py:312-327
        u, g = self.u[l], self.g[l]
        for i in self._fpts(l):                          # ascending is valid in serial: each interval ascends
            if l == 0:
                u[i] = self.phi(l, u[i - 1], i)
            else:
                u[i] = g[i] + self.phi(l, u[i - 1], i)

    def c_relax(self, l):                                # mgrit.py:354-368
        u, g, w = self.u[l], self.g[l], self.weight_c
        for i in self.cpts[l]:
            if i == 0:
                continue
            if l == 0:
                u[i] = self.phi(l, u[i - 1], i) * w + u[i] * (1.0 - w)
            else:
                u[i] = (g[i] + self.phi(l, u[i - 1], i)) * w + u[i] * (1.0 - w)

    def fas_residual(self, l):                           # mgrit.py:497-547
        u, g = self.u[l], self.g[l]
        c = self.cpts[l]
        R = self.transfer[l].restriction
        for j in range(len(c)):
            self.u[l + 1][j] = R(u[c[j]])                # restriction of every C-point (injection for the identity)
        self.v[l + 1] = self.u[l + 1].copy()
        v = self.v[l + 1]
        for j in range(1, len(c)):
            if l == 0:
                fine = self.phi(l, u[c[j] - 1], c[j]) - u[c[j]]
            else:
                fine = g[c[j]] - u[c[j]] + self.phi(l, u[c[j] - 1], c[j])
            self.g[l + 1][j] = R(fine) + v[j] - self.phi(l + 1, v[j - 1], j)

    def error_correction(self, l):                       # mgrit.py:722-726
        c = self.cpts[l]
        for j in range(1, len(c)):
            e = self.transfer[l].interpolation(self.u[l + 1][j] - self.v[l + 1][j])
            self.u[l][c[j]] = self.u[l][c[j]] + e

    def forward_solve(self, l):                          # mgrit.py:471-481
        u, g = self.u[l], self.g[l]
        for i in range(1, len(u)):
            if l == 0:
                u[i] = self.phi(l, u[i - 1], i)
            else:
                u[i] = g[i] + self.phi(l, u[i - 1], i)

    def iteration(self, l, cycle_type, iteration, first_f):   # mgrit.py:261-290
        if l == self.L - 1:
            self.forward_solve(l)
            return
        if (l > 0 or (iteration == 0 and l == 0)) and first_f:
            self.f_relax(l)
        for _ in range(self.cf_iter[l]):
            self.c_relax(l)
            self.f_relax(l)
        self.fas_residual(l)
        self.iteration(l + 1, cycle_type, iteration, True)
        self.error_correction(l)
        self.f_relax(l)
        if l != 0 and cycle_type == 'F':
            self.iteration(l, 'V', iteration, False)

    def nested_iteration(self):                          # mgrit.py:551-566
        self.forward_solve(self.L - 1)
        for l in range(self.L - 2, -1, -1):
            c = self.cpts[l]
            for j in range(1, len(c)):
                self.u[l][c[j]] = self.transfer[l].interpolation(self.u[l + 1][j])
            if l > 0:
                self.iteration(l, 'V', 0, True)

    def residual_norms(self):                            # mgrit.py:405-413
        u = self.u[0]
        p = self.problem[0]
        return [p.norm(self.phi(0, u[i - 1], i) - u[i]) for i in self.cpts[0] if i != 0]

    def jump_norms(self):                                # mgrit.py:372-385
        p = self.problem[0]
        out = [p.norm(self.u[0][i] - self.last[i]) for i in self.cpts[0] if i != 0]
        self.last = self.u[0].copy()
        return out

    def solve(self):                                     # mgrit.py:590-646
        t0 = time.time()
        for it in range(self.max_iter):
            self.iteration(0, self.cycle_type, it, True)
            val = self.residual_norms() if self.conv_crit == 0 else self.jump_norms()
            self.conv[it + 1] = np.linalg.norm(np.array(val), ord=self.ord) if len(val) else 0.0
            if self.conv[it + 1] < self.tol:
                break
        return {'conv': self.conv[np.where(self.conv != 0)], 'time_setup': self.time_setup,
                'time_solve': time.time() - t0}


class AtMgritOracle(MgritOracle):
    """core/at_mgrit.py:16-88 on one time rank: the coarsest level is not solved sequentially; every coarsest point is the
    end of its own local coarse grid of at most k points, started from the previous iterate (at_mgrit.py:75-86)."""

    def __init__(self, problem, k, **kw):
        self.k = k
        if kw.get('conv_crit', 0) not in (0, 1):
            raise Exception('Local convergence criteria are not implemented for AT-MGRIT. Please select a global criterion.')
        super().__init__(problem, **kw)

    def forward_solve(self, l):                          # at_mgrit.py:37-88 (comm_time_size == 1 branch)
        if self.L == 1:
            return
        u, g = self.u[l], self.g[l]
        old = u.copy()
        for point in range(len(u)):
            tmp = old[max(0, point - self.k + 1)]
            for i in range(max(1, point - self.k + 2), point + 1):
                tmp = g[i] + self.phi(l, tmp, i)
            u[point] = tmp


def time_stepping(problem: OracleProblem) -> np.ndarray:
    """Sequential time stepping u_i = Phi(u_{i-1}) (the 1-level case, tests/core/test_mgrit.py:72-84)."""
    u = np.zeros((len(problem.t),) + np.shape(problem.u0))
    u[0] = problem.u0
    for i in range(1, len(problem.t)):
        u[i] = problem.phi(u[i - 1], problem.t[i - 1], problem.t[i])
    return u
