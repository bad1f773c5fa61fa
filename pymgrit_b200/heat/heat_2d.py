"""2-D heat equation with the reference's interface (heat/heat_2d.py:20-366, backward Euler), on the GPU in sine space.

    u_t - a (u_xx + u_yy) = b(x, y, t)  on [x_start, x_end] x [y_start, y_end],  Dirichlet data on the boundary,
    u_i = (I + dt L)^-1 rhs_i           5-point Laplacian on the interior nodes, identity rows on the boundary nodes
                                        (heat_2d.py:250-320, 360-366).

L = fx Tx (x) I + fy I (x) Ty with Toeplitz tridiag(-1, 2, -1) factors, which the orthonormal sine transforms
diagonalise exactly.  The solver keeps every level in sine space (csrc/phi.cuh, Heat2D): there Phi is one division per
coefficient, and sums, differences, injection and 2-norms are unchanged because the transform is orthogonal.  Values
are transformed only where they enter or leave: the initial condition, the spatial factors of the right-hand side,
`Mgrit.u[lvl][i]` / `get_values()`, and `step()` on a stand-alone vector (csrc/heat2d.cu).  This is a direct solver
like the reference's SuperLU call, not an iteration: results agree to rounding (tests/test_gpu_parity.py).
"""
from collections import OrderedDict
from typing import Callable, Union

import numpy as np

from pymgrit_b200 import _lib
from pymgrit_b200.core.application import DeviceApplication
from pymgrit_b200.core.vector import DeviceVector
from pymgrit_b200.core import device_level as dl
from pymgrit_b200.core.rhs_tables import RhsSplit, Sampler2D

MAX_TERMS = 3            # register-resident right-hand-side factors of the Heat2D kernels (csrc/phi.cuh)
_PI = np.longdouble('3.14159265358979323846264338327950288')


class VectorHeat2D(DeviceVector):
    """(nx, ny) node values including the boundary nodes (heat_2d.py:20-136), stored in HBM."""

    def __init__(self, nx, ny, tensor=None):
        super().__init__((int(nx), int(ny)), tensor)
        self.nx, self.ny = int(nx), int(ny)


def sine_matrix(n):
    """S[j][k] = sqrt(2/(n+1)) sin(pi (j+1)(k+1)/(n+1)): symmetric, orthogonal, diagonalises tridiag(-1, 2, -1)."""
    j = np.arange(1, n + 1, dtype=np.int64)
    red = np.outer(j, j) % (2 * (n + 1))                    # exact argument reduction in integers
    ang = red.astype(np.longdouble) * (_PI / (n + 1))
    return np.asarray(np.sqrt(np.longdouble(2) / (n + 1)) * np.sin(ang), dtype=np.float64)


def laplace_eigenvalues(n):
    """4 sin^2(pi (k+1) / (2 (n+1))), k < n, in extended precision."""
    k = np.arange(1, n + 1).astype(np.longdouble)
    return 4 * np.sin(k * _PI / (2 * (n + 1))) ** 2


class _Family:
    """What all levels (and deep copies) of one Heat2D problem share: layout, sine matrices, symbol row,
    right-hand-side factors.  Host tables are built once; device tensors once per CUDA device."""

    def __init__(self, app):
        nx, ny = app.nx, app.ny
        self.nx, self.ny = nx, ny
        tile, nsys, fb, pitch = (np.zeros(1, dtype=np.int32) for _ in range(4))
        p = lambda a: a.ctypes.data_as(_lib.c_int32_p)
        _lib.check(_lib.lib().mgb_heat2d_layout(nx, ny, p(tile), p(nsys), p(fb), p(pitch)), 'heat2d_layout')
        self.tile, self.nsys, self.first_boundary, self.pitch = int(tile[0]), int(nsys[0]), int(fb[0]), int(pitch[0])
        mx, my = nx - 2, ny - 2
        self.nint = mx * my
        self.boff = self.first_boundary * self.tile
        fx = np.longdouble(app.a) / np.longdouble(app.dx) ** 2          # heat_2d.py:260-261
        fy = np.longdouble(app.a) / np.longdouble(app.dy) ** 2
        # boundary nodes in row order: i = 0, i = nx-1, j = 0 (i = 1..nx-2), j = ny-1
        ii = np.concatenate([np.zeros(ny, int), np.full(ny, nx - 1), np.arange(1, nx - 1), np.arange(1, nx - 1)])
        jj = np.concatenate([np.arange(ny), np.arange(ny), np.zeros(mx, int), np.full(mx, ny - 1)])
        self.bnodes = ii * ny + jj
        bcv = app.boundary_values()                                     # (nx, ny), zero in the interior
        sig = np.zeros(self.pitch)
        sym = fx * laplace_eigenvalues(mx)[:, None] + fy * laplace_eigenvalues(my)[None, :]
        sig[:self.nint] = np.asarray(sym, dtype=np.float64).reshape(-1)
        sig[self.boff:self.boff + len(self.bnodes)] = bcv.reshape(-1)[self.bnodes]
        self.sig = sig
        # coupling of the interior equations to the (constant) boundary values, heat_2d.py:250-287: the interior rows
        # keep their off-diagonal entries towards boundary nodes
        coup = np.zeros((nx, ny))
        coup[1, 1:-1] += float(fx) * bcv[0, 1:-1]
        coup[-2, 1:-1] += float(fx) * bcv[-1, 1:-1]
        coup[1:-1, 1] += float(fy) * bcv[1:-1, 0]
        coup[1:-1, -2] += float(fy) * bcv[1:-1, -1]
        self.coupling = coup if np.any(coup) else None
        self.split = RhsSplit(app.rhs, sampler=Sampler2D(app.rhs, app.x, app.y),
                              max_terms=MAX_TERMS - (1 if self.coupling is not None else 0)).analyse(app.t)
        self.sx, self.sy = sine_matrix(mx), sine_matrix(my)
        self._dev = {}

    # ---- device side ------------------------------------------------------------------------------------------
    def dev(self):
        torch = dl._torch()
        d = torch.cuda.current_device()
        if d not in self._dev:
            dev = torch.device('cuda', d)
            st = dict(sx=torch.as_tensor(self.sx).to(dev), sy=torch.as_tensor(self.sy).to(dev),
                      sig=torch.as_tensor(self.sig).to(dev), rhs_x=None, nterms=0,
                      h2d=self.sx.nbytes + self.sy.nbytes + self.sig.nbytes)
            self._dev[d] = st
            fields = []
            if self.split.kind == 'separable':
                nodes = np.zeros((self.split.basis.shape[0], self.nx, self.ny))
                nodes[:, 1:-1, 1:-1] = self.split.basis.reshape(-1, self.nx - 2, self.ny - 2)
                fields.append(nodes)
            if self.coupling is not None:
                fields.append(self.coupling[None])
            if fields:
                nodes = np.concatenate(fields)
                rows = self.to_rows(torch.as_tensor(nodes).to(dev))
                rows[:, self.boff:] = 0.0                 # factors act on the interior coefficients only
                st['rhs_x'], st['nterms'] = rows, len(nodes)
                st['h2d'] += nodes.nbytes
        return self._dev[d]

    def _transform(self, fn, src, dst_cols, what):
        torch = dl._torch()
        st = self.dev()
        count = src.shape[0]
        src = src.contiguous()
        dst = torch.empty((count, dst_cols), dtype=torch.float64, device=src.device)
        work = torch.empty((count, self.nint), dtype=torch.float64, device=src.device)
        _lib.check(fn(self.nx, self.ny, st['sx'].data_ptr(), st['sy'].data_ptr(), src.data_ptr(), dst.data_ptr(), count,
                      work.data_ptr(), _lib.current_stream_ptr()), what)
        return dst

    def to_rows(self, nodes):
        """[count, nx, ny] node values -> [count, pitch] level rows."""
        return self._transform(_lib.lib().mgb_heat2d_to_rows, nodes.reshape(nodes.shape[0], -1), self.pitch,
                               'heat2d_to_rows')

    def from_rows(self, rows):
        out = self._transform(_lib.lib().mgb_heat2d_from_rows, rows, self.nx * self.ny, 'heat2d_from_rows')
        return out.view(rows.shape[0], self.nx, self.ny)


def in_time_dense(vals, dts, theta):
    """Rows dt_i (theta b(t_i) + (1 - theta) b(t_{i-1})) of a block of consecutive time points (first row: no predecessor
    inside the block, only used when the block starts at point 0, whose step does not exist)."""
    prev = np.concatenate([vals[:1], vals[:-1]])
    return (theta * vals + (1 - theta) * prev) * dts[:, None]


_FAMILIES = OrderedDict()


class Heat2D(DeviceApplication):
    """Same constructor as the reference (heat_2d.py:147-248): theta method (BE, CN, FE) in sine space."""
    kind = _lib.APP_HEAT2D

    def __init__(self, x_start: float, x_end: float, y_start: float, y_end: float, nx: int, ny: int, a: float,
                 rhs: Callable = lambda x, y, t: 0 * x * y, init_cond: Callable = lambda x, y: x * y * 0,
                 method: str = 'BE', bc_left: Union[int, float, Callable] = 0, bc_right: Union[int, float, Callable] = 0,
                 bc_bottom: Union[int, float, Callable] = 0, bc_top: Union[int, float, Callable] = 0, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.x_start, self.x_end, self.y_start, self.y_end = x_start, x_end, y_start, y_end
        self.x = np.linspace(x_start, x_end, nx)
        self.y = np.linspace(y_start, y_end, ny)
        self.x_2d = self.x[:, np.newaxis]
        self.y_2d = self.y[np.newaxis, :]
        self.nx, self.ny = nx, ny
        self.dx = self.x[1] - self.x[0]
        self.dy = self.y[1] - self.y[0]
        self.a = a
        self.rhs = rhs
        if method == 'BE':                                    # heat_2d.py:192-200
            self.theta = 1
        elif method == 'FE':
            self.theta = 0
        elif method == 'CN':
            self.theta = 1 / 2
        else:
            raise Exception("Unknown method. Choose BE (Backward Euler), FE (Forward Euler) or CN (Crank-Nicolson")
        self.method = method

        def as_fn(v, name):                                  # heat_2d.py:205-231
            if isinstance(v, (float, int)):
                return lambda s, _v=v: _v
            if callable(v):
                return v
            raise Exception('Choose float, int or function for boundary condition ' + name)
        self.bc_left, self.bc_right = as_fn(bc_left, 'bc_left'), as_fn(bc_right, 'bc_right')
        self.bc_bottom, self.bc_top = as_fn(bc_bottom, 'bc_bottom'), as_fn(bc_top, 'bc_top')

        self.ndof = None                                     # set from the family's layout below
        self.vector_template = VectorHeat2D(nx, ny)
        self.init_cond = init_cond
        self.vector_t_start = VectorHeat2D(nx, ny)
        init = np.array(np.broadcast_to(self.init_cond(self.x_2d, self.y_2d), (nx, ny)), dtype=float)   # heat_2d.py:243
        self._apply_bc(init)
        self.vector_t_start.set_values(init)
        # the Dirichlet data as the four boundary lines (what _apply_bc writes): a key of 4 (nx + ny) doubles instead of the
        # nx x ny array, whose bytes took a millisecond per level to copy and to hash
        edges = tuple(np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=float), shape))
                      for v, shape in ((self.bc_left(self.x), (nx,)), (self.bc_right(self.x), (nx,)),
                                       (self.bc_bottom(self.y), (ny,)), (self.bc_top(self.y), (ny,))))
        if self.theta == 0 and any(np.any(e) for e in edges):
            # the reference's forward-Euler branch ADDS the Dirichlet values to the boundary nodes in every step
            # (heat_2d.py:346-353), so they grow with the step count: not reproduced on the device
            raise Exception("pymgrit_b200.Heat2D: method 'FE' is available for homogeneous Dirichlet data only")
        self._family_key = (nx, ny, float(x_start), float(x_end), float(y_start), float(y_end), float(a), id(rhs),
                            b''.join(e.tobytes() for e in edges))
        self.ndof = self.family().pitch

    def _apply_bc(self, b):                                  # heat_2d.py:244-247, 315-319: same order of writes
        b[:, 0] = self.bc_left(self.x)
        b[:, -1] = self.bc_right(self.x)
        b[-1, :] = self.bc_bottom(self.y)
        b[0, :] = self.bc_top(self.y)

    def boundary_values(self):
        b = np.zeros((self.nx, self.ny))
        self._apply_bc(b)
        return b

    def family(self) -> _Family:
        fam = self.__dict__.get('_fam')
        if fam is None:
            fam = _FAMILIES.get(self._family_key)
            # a split analysed by another level of the same problem is adopted if it reproduces b at this level's times
            if fam is None or (fam.split.kind == 'separable' and not fam.split.reproduces(self.t)):
                fam = _Family(self)
                _FAMILIES[self._family_key] = fam
                while len(_FAMILIES) > 8:
                    _FAMILIES.popitem(last=False)
            self._fam = fam
        return fam

    # ---- DeviceApplication contract ---------------------------------------------------------------------------
    def row_pitch(self):
        return self.family().pitch

    def rows_to_values(self, rows):
        return self.family().from_rows(rows)

    def values_to_rows(self, values, rows) -> None:
        rows.copy_(self.family().to_rows(values))

    def level_tables(self, t, team_threads, chunk):
        fam = self.family()
        st = fam.dev()
        t = np.asarray(t, dtype=float)
        dts, dtidx = dl.dt_classes(t)
        th = self.theta
        sconst = np.zeros((len(dts), 8))
        sconst[:, 0] = th * dts                                # implicit part, heat_2d.py:363
        sconst[:, 1] = (1 - th) * dts                          # explicit part, heat_2d.py:306, 352
        dt_full = np.zeros(len(t))
        dt_full[1:] = np.diff(t)
        tab = dict(ndt=len(dts), dtidx=dtidx, sconst=sconst, cw=8, nsys=fam.nsys, sig_dev=st['sig'],
                   ip=[fam.first_boundary, 1 if th == 1 else 0, 0, 0])

        def in_time(vals):
            """dt_i (theta b(t_i) + (1 - theta) b(t_{i-1})) from b at every point of t (heat_2d.py:302, 309-313, 356)."""
            prev = np.concatenate([vals[:1], vals[:-1]])
            return (th * vals + (1 - th) * prev) * dt_full.reshape((-1,) + (1,) * (vals.ndim - 1))
        cols = []
        if fam.split.kind == 'separable':
            cols.append(in_time(fam.split.coefficients(t)))
        if fam.coupling is not None:
            cols.append(dt_full[:, None])
        if cols:
            tab['nrhs'], tab['rhs_x_dev'], tab['rhs_t'] = st['nterms'], st['rhs_x'], np.concatenate(cols, axis=1)
        if fam.split.kind == 'dense':
            torch = dl._torch()
            dense = torch.empty((len(t), fam.pitch), dtype=torch.float64, device=st['sig'].device)
            for a in range(0, len(t), 16):                                  # transform in batches
                nodes = np.zeros((len(t[a:a + 16]), self.nx, self.ny))
                lo = max(a - 1, 0)                                           # one point back for the explicit part
                vals = in_time_dense(fam.split.dense(t[lo:a + 16]), dt_full[lo:a + 16], th)[a - lo:]
                nodes[:, 1:-1, 1:-1] = vals.reshape(-1, self.nx - 2, self.ny - 2)
                rows = fam.to_rows(torch.as_tensor(nodes).to(dense.device))
                rows[:, fam.boff:] = 0.0
                dense[a:a + 16] = rows
            tab['rhs_dense_dev'] = dense
        return tab

    def __getstate__(self):                                  # deep copies (simple_setup_problem) share the family
        state = super().__getstate__()
        state.pop('_fam', None)
        return state
