"""2-D Allen-Cahn equation with the reference's interface (allen_cahn/allen_cahn.py:136-270), Phi on the GPU.

    u_t = u_xx + u_yy + (1/eps^2) u (1 - u^nu)   on [-0.5, 0.5]^2, periodic,   u(x, 0) = tanh((R0 - |x|) / (sqrt(2) eps))

Built: the IMEX branch (allen_cahn.py:191-197) -- reaction explicit, diffusion implicit:
    rhs = u + dt (1/eps^2) u (1 - u^nu),      (I - dt L) y = rhs,
L the periodic 5-point Laplacian.  The reference calls spsolve on the (nx^2 x nx^2) matrix every step; L is
L1 (x) I + I (x) L1 with the circulant second difference L1, which the real orthonormal Fourier basis Q diagonalises
exactly, so the same linear system is solved directly as  y = Q [(Q^T rhs Q) / (1 + dt (mu_i + mu_j))] Q^T:
four batched FP64 products per step for ALL coarse intervals of a sweep at once (mgb_allen_cahn_imex_rows).  The
application runs on the batched path (core/batched.py): the nonlinear reaction term keeps it out of the fused,
transform-once sweeps of the heat equations.

Not built: the fully implicit and Crank-Nicolson branches (allen_cahn.py:198-229): Newton iterations whose Jacobian
I - fac (L + diag(...)/eps^2) changes with the iterate and is not diagonalised by Q -- a sparse direct solve per Newton
step and time point in the reference.  Constructing them raises.
"""
import numpy as np

from pymgrit_b200 import _lib
from pymgrit_b200.core.batched import BatchedApplication, _torch
from pymgrit_b200.core.vector import DeviceVector

_PI = np.longdouble('3.14159265358979323846264338327950288')


class VectorAllenCahn2D(DeviceVector):
    """(nx, ny) node values (allen_cahn.py:16-133), stored in HBM."""

    def __init__(self, nx, ny, tensor=None):
        super().__init__((int(nx), int(ny)), tensor)
        self.nx, self.ny = int(nx), int(ny)


def periodic_fourier_basis(n):
    """(Q, mu): Q [n, n] real orthonormal, columns = eigenvectors of the circulant second difference
    tridiag_periodic(1, -2, 1) = -Q diag(mu) Q^T, mu_m = 4 sin^2(pi f_m / n) with f_m the frequency of column m
    (constant, cos/sin pairs, and the alternating vector for even n).  Extended precision, rounded once."""
    j = np.arange(n, dtype=np.int64)
    cols, freq = [np.full(n, 1 / np.sqrt(np.longdouble(n)))], [0]
    for f in range(1, (n - 1) // 2 + 1):
        ang = ((j * f) % n).astype(np.longdouble) * (2 * _PI / n)          # exact argument reduction in integers
        s = np.sqrt(np.longdouble(2) / n)
        cols += [s * np.cos(ang), s * np.sin(ang)]
        freq += [f, f]
    if n % 2 == 0 and n > 1:
        cols.append(np.where(j % 2 == 0, 1, -1).astype(np.longdouble) / np.sqrt(np.longdouble(n)))
        freq.append(n // 2)
    q = np.stack(cols, axis=1)
    mu = 4 * np.sin(_PI * np.asarray(freq, dtype=np.longdouble) / n) ** 2
    return np.asarray(q, dtype=np.float64), np.asarray(mu, dtype=np.longdouble)


class AllenCahn(BatchedApplication):
    """Same constructor as the reference (allen_cahn.py:146-173); method='IMEX' runs on the device."""

    def __init__(self, nx=128, nu=2, eps=0.04, newton_maxiter=100, newton_tol=1e-12, lin_tol=1e-12, lin_maxiter=100,
                 radius=0.25, method='IMPL', *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.nu = nu
        self.eps = eps
        self.newton_maxiter = newton_maxiter
        self.newton_tol = newton_tol
        self.lin_tol = lin_tol
        self.lin_maxiter = lin_maxiter
        self.radius = radius
        self.nx = nx
        self.ny = nx
        self.method = method
        if self.method not in ('IMPL', 'IMEX', 'CN'):
            raise Exception("Unknown method. Choose IMPL (implicit), IMEX (implicit-explicit) or CN (Crank-Nicolson")
        if self.method != 'IMEX':
            raise Exception("pymgrit_b200.AllenCahn: only method='IMEX' has device kernels; the Newton branches (IMPL, CN) "
                            "are not built")
        if int(nu) != nu or nu < 1:
            raise Exception('pymgrit_b200.AllenCahn: nu must be a positive integer')
        self.dx = 1.0 / self.nx
        self.x = np.linspace(start=-0.5, stop=0.5, num=self.nx)
        self.ndof = self.nx * self.ny
        self.vector_template = VectorAllenCahn2D(nx=self.nx, ny=self.ny)
        self.vector_t_start = self.initial_guess()
        self._q_host, mu = periodic_fourier_basis(self.nx)
        self._mu_host = np.asarray(mu / np.longdouble(self.dx) ** 2, dtype=np.float64)       # eigenvalues of -L1 / dx^2
        self._dev = {}

    def initial_guess(self):
        """allen_cahn.py:231-245, vectorised."""
        r = np.sqrt(self.x[:, None] ** 2 + self.x[None, :] ** 2)
        initial = VectorAllenCahn2D(nx=self.nx, ny=self.ny)
        initial.set_values(np.tanh((self.radius - r) / (np.sqrt(2) * self.eps)))
        return initial

    def exact_radius(self, t):
        return np.sqrt(max(self.radius ** 2 - 2.0 * t, 0))

    def compute_radius(self, u):
        return np.sqrt(np.count_nonzero(u.get_values() >= 0.0) / np.pi) * self.dx

    # ---- device side --------------------------------------------------------------------------------------------------
    def _tables(self):
        torch = _torch()
        d = torch.cuda.current_device()
        if d not in self._dev:
            dev = torch.device('cuda', d)
            self._dev[d] = dict(q=torch.as_tensor(self._q_host).to(dev),
                                qt=torch.as_tensor(np.ascontiguousarray(self._q_host.T)).to(dev),
                                mu=torch.as_tensor(self._mu_host).to(dev), work={})
        return self._dev[d]

    def step_rows(self, src, src_idx, dst, dst_idx, t_start, t_stop) -> None:
        torch = _torch()
        tab = self._tables()
        count = int(len(t_start))
        if count == 0:
            return
        nn = self.nx * self.ny
        work = tab['work'].get(count)
        if work is None:
            if len(tab['work']) > 16:
                tab['work'].clear()
            work = tab['work'][count] = (torch.empty((count, nn), dtype=torch.float64, device=src.device),
                                         torch.empty((count, nn), dtype=torch.float64, device=src.device),
                                         torch.empty(count, dtype=torch.float64, device=src.device))
        dt_host = torch.as_tensor(np.asarray(t_stop, dtype=float) - np.asarray(t_start, dtype=float))
        work[2].copy_(dt_host, non_blocking=False)
        _lib.check(_lib.lib().mgb_allen_cahn_imex_rows(self.nx, count, src.data_ptr(), src.stride(0), src_idx.data_ptr(),
                                                       dst.data_ptr(), dst.stride(0), dst_idx.data_ptr(), work[2].data_ptr(),
                                                       1.0 / self.eps ** 2, int(self.nu), tab['q'].data_ptr(),
                                                       tab['qt'].data_ptr(), tab['mu'].data_ptr(), work[0].data_ptr(),
                                                       work[1].data_ptr(), _lib.current_stream_ptr()), 'allen_cahn_imex_rows')

    def __getstate__(self):
        state = super().__getstate__()
        state['_dev'] = {}
        return state
