"""The batched path: MGRIT sweeps for applications that bring their own device Phi.

The reference's plug-in contract is `Application.step(u_start, t_start, t_stop)` -- any time integrator, called once per
time point (core/application.py:99, core/mgrit.py:321-323).  The fused sweeps of libmgrit_b200 cover the applications
whose Phi fits a register-resident team kernel (csrc/phi.cuh).  Everything else -- a 2-D periodic solve (AllenCahn), a
user's own CUDA / torch code -- derives from `BatchedApplication` and implements ONE method,

    step_rows(src, src_idx, dst, dst_idx, t_start, t_stop)     dst[dst_idx[k]] = Phi(src[src_idx[k]], t_start[k] -> t_stop[k])

on level arrays in HBM ([points, pitch] float64 tensors, a row = the flattened values of a time point).  `BatchedSweeps`
then runs every sweep of mgrit.py with all coarse intervals of a level per call: the host walks the positions inside an
interval (m - 1 steps for an F-relaxation), each position is one `step_rows` over all intervals, and the row arithmetic
around it is libmgrit_b200's mgb_rows_lincomb / mgb_rows_sumsq (csrc/generic.cu), operation by operation the reference's.
To add a new Phi: INTEGRATION.md section 3.
"""
import numpy as np

from pymgrit_b200 import _lib
from pymgrit_b200.core.application import DeviceApplication


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise Exception('pymgrit_b200 needs a CUDA device: there is no CPU fallback')
    return torch


class BatchedApplication(DeviceApplication):
    """An application without fused team kernels.  Subclasses set `ndof` (doubles per time point; rows hold the flattened
    vector values) and implement step_rows(); `step()` on a single Vector is provided on top of it."""
    kind = _lib.APP_BATCHED

    def step_rows(self, src, src_idx, dst, dst_idx, t_start, t_stop) -> None:
        """dst[dst_idx[k]] = Phi(src[src_idx[k]]) from t_start[k] to t_stop[k], for all k at once.

        src, dst: [points, pitch] float64 CUDA tensors (may be the same array: the rows read and written by one call never
        overlap); src_idx, dst_idx: int32 CUDA tensors of row numbers; t_start, t_stop: float64 HOST arrays.  Asynchronous
        on the current CUDA stream."""
        raise NotImplementedError

    def level_tables(self, t, team_threads, chunk):
        return {}

    def step(self, u_start, t_start: float, t_stop: float):
        torch = _torch()
        x = u_start.device_values.reshape(1, -1).contiguous()
        out = torch.empty_like(x)
        idx = torch.zeros(1, dtype=torch.int32, device=x.device)
        self.step_rows(x, idx, out, idx, np.array([float(t_start)]), np.array([float(t_stop)]))
        return self.vector_template._new(out.reshape(self.vector_template.shape))


class _Steps:
    """One batched Phi call prepared at setup: row indices on the device, the times on the host."""

    def __init__(self, torch, dev, src, dst, t):
        self.n = len(src)
        self.src = torch.as_tensor(np.asarray(src, dtype=np.int32)).to(dev)
        self.dst = torch.as_tensor(np.asarray(dst, dtype=np.int32)).to(dev)
        self.t0 = np.ascontiguousarray(t[np.asarray(src, dtype=np.int64)], dtype=float) if self.n else np.zeros(0)
        self.t1 = np.ascontiguousarray(t[np.asarray(dst, dtype=np.int64)], dtype=float) if self.n else np.zeros(0)


class BatchedSweeps:
    """The sweeps of one Mgrit object on the batched path (one time rank)."""

    def __init__(self, solver):
        torch = _torch()
        self.s = solver
        self.lib = _lib.lib()
        lv0 = solver._lv[0]
        dev = lv0.u.device
        self.dev = dev
        self.plans = []
        for lvl, lv in enumerate(solver._lv):
            t = np.asarray(lv.t, dtype=float)
            plan = {}
            if lv.cpts is not None:
                cp = np.asarray(lv.cpts, dtype=np.int64)
                ends = np.append(cp[1:], lv.npts)
                lens = ends - cp
                # F-relaxation: position d of every interval that is longer than d  (mgrit.py:312-327)
                plan['f'] = [_Steps(torch, dev, cp[lens > d] + d - 1, cp[lens > d] + d, t) for d in range(1, int(lens.max()))]
                last = ends - 1                                  # last point of every interval (the one C-relaxation reads)
                plan['f_last'] = torch.as_tensor(last[lens > 1].astype(np.int32)).to(dev)
                # C-relaxation (mgrit.py:354-368): runs of adjacent C-points are visited in ascending order
                k = np.arange(1, len(cp))
                depth = np.zeros(len(cp), dtype=np.int64)
                for j in k:                                      # point 0 is never relaxed: a C-point next to it waits for nobody
                    depth[j] = depth[j - 1] + 1 if (j >= 2 and cp[j] - cp[j - 1] == 1) else 0
                waves = []
                for d in range(int(depth[1:].max()) + 1 if len(cp) > 1 else 0):
                    sel = k[depth[1:] == d]
                    waves.append((_Steps(torch, dev, cp[sel] - 1, cp[sel], t), torch.as_tensor(sel.astype(np.int32)).to(dev)))
                plan['c'] = waves
                plan['cp'] = torch.as_tensor(cp.astype(np.int32)).to(dev)              # C-points, all
                plan['cp1'] = torch.as_tensor(cp[1:].astype(np.int32)).to(dev)         # C-points j >= 1
                plan['cpm1'] = torch.as_tensor((cp[1:] - 1).astype(np.int32)).to(dev)
                plan['cprev'] = torch.as_tensor(cp[:-1].astype(np.int32)).to(dev)      # c_{j-1}, j >= 1
                plan['j1'] = torch.arange(1, len(cp), dtype=torch.int32, device=dev)
                plan['res'] = _Steps(torch, dev, cp[1:] - 1, cp[1:], t)                # Phi(u[c-1]) at every C-point >= 1
                plan['ncp'] = len(cp)
                plan['tmp'] = torch.zeros((max(len(cp), 1), lv.pitch), dtype=torch.float64, device=dev)
                plan['tmp2'] = torch.zeros((max(len(cp), 1), lv.pitch), dtype=torch.float64, device=dev)
            plan['seq'] = None                                   # forward solve: built on first use
            self.plans.append(plan)
        for lvl in range(len(solver._lv) - 1):                   # the coarse step of the FAS restriction: v[j-1] -> j
            nc = self.plans[lvl]['ncp']
            tc = np.asarray(solver._lv[lvl + 1].t, dtype=float)
            self.plans[lvl]['coarse'] = _Steps(torch, dev, np.arange(0, nc - 1), np.arange(1, nc), tc)

    # -- row helpers ------------------------------------------------------------------------------------------------
    def _lin(self, count, lv, out, oi, a, x, xi, b=0.0, y=None, yi=None, c=0.0, z=None, zi=None):
        if count <= 0:
            return
        p = lambda t_: None if t_ is None else t_.data_ptr()
        n = lv.n
        _lib.check(self.lib.mgb_rows_lincomb(count, n, out.data_ptr(), out.stride(0), p(oi), float(a), x.data_ptr(), x.stride(0),
                                             p(xi), float(b), p(y), 0 if y is None else y.stride(0), p(yi), float(c), p(z),
                                             0 if z is None else z.stride(0), p(zi), self.s._stream()), 'rows_lincomb')

    def _phi(self, lvl, steps, src, dst, dst_idx=None):
        if steps.n:
            self.s.problem[lvl].step_rows(src, steps.src, dst, steps.dst if dst_idx is None else dst_idx, steps.t0, steps.t1)
            self.s.launches += 1

    # -- sweeps -----------------------------------------------------------------------------------------------------
    def f_relax(self, lvl, last_only=False):
        lv, plan = self.s._lv[lvl], self.plans[lvl]
        for st in plan['f']:
            self._phi(lvl, st, lv.u, lv.u)
            if lv.g is not None:
                self._lin(st.n, lv, lv.u, st.dst, 1.0, lv.g, st.dst, 1.0, lv.u, st.dst)           # g[i] + Phi(u[i-1])

    def c_relax(self, lvl):
        lv, plan = self.s._lv[lvl], self.plans[lvl]
        w = float(self.s.weight_c)
        for st, _ in plan['c']:
            if st.n == 0:
                continue
            tmp = plan['tmp']
            rows = self._arange(st.n)
            self._phi(lvl, st, lv.u, tmp, rows)
            if lv.g is not None:
                self._lin(st.n, lv, tmp, rows, 1.0, lv.g, st.dst, 1.0, tmp, rows)                  # g[c] + Phi(u[c-1])
            if w == 1.0:
                self._lin(st.n, lv, lv.u, st.dst, 1.0, tmp, rows)
            else:
                self._lin(st.n, lv, lv.u, st.dst, w, tmp, rows, 1.0 - w, lv.u, st.dst)             # mgrit.py:364-366

    def _arange(self, n):
        torch = _torch()
        cache = self.__dict__.setdefault('_ar', {})
        if n not in cache:
            cache[n] = torch.arange(n, dtype=torch.int32, device=self.dev)
        return cache[n]

    def fas_residual(self, lvl):
        """mgrit.py:497-547 with the identity transfer: injection, then
        G.g[j] = ((Phi_f(u[c-1]) - u[c] [+ g[c]]) + u[c]) - Phi_c(u[c_{j-1}])."""
        fine, coarse, plan = self.s._lv[lvl], self.s._lv[lvl + 1], self.plans[lvl]
        ncp = plan['ncp']
        n1 = ncp - 1
        tmp, tmp2 = plan['tmp'], plan['tmp2']
        rows = self._arange(n1)
        self._lin(ncp, fine, coarse.u, None, 1.0, fine.u, plan['cp'])                              # injection of every C-point
        if n1 <= 0:
            return
        self._phi(lvl, plan['res'], fine.u, tmp, rows)                                             # Phi_f(u[c-1])
        if fine.g is not None:
            self._lin(n1, fine, tmp2, rows, 1.0, fine.g, plan['cp1'], -1.0, fine.u, plan['cp1'])   # g[c] - u[c]
            self._lin(n1, fine, tmp, rows, 1.0, tmp2, rows, 1.0, tmp, rows)                        # ... + Phi_f
        else:
            self._lin(n1, fine, tmp, rows, 1.0, tmp, rows, -1.0, fine.u, plan['cp1'])              # Phi_f - u[c]
        self._lin(n1, fine, tmp, rows, 1.0, tmp, rows, 1.0, fine.u, plan['cp1'])                   # ... + v[j] (= u[c])
        self._phi(lvl + 1, plan['coarse'], coarse.u, tmp2, rows)                                   # Phi_c(v[j-1])
        self._lin(n1, fine, coarse.g, plan['j1'], 1.0, tmp, rows, -1.0, tmp2, rows)

    def error_correction(self, lvl, f_relax=False, last_only=False):
        fine, coarse, plan = self.s._lv[lvl], self.s._lv[lvl + 1], self.plans[lvl]
        n1 = plan['ncp'] - 1
        tmp = plan['tmp']
        rows = self._arange(max(n1, 1))
        if n1 > 0:                                                                                 # u[c] + (G.u[j] - u[c])
            self._lin(n1, fine, tmp, rows, 1.0, coarse.u, plan['j1'], -1.0, fine.u, plan['cp1'])
            self._lin(n1, fine, fine.u, plan['cp1'], 1.0, fine.u, plan['cp1'], 1.0, tmp, rows)
        if f_relax:
            self.f_relax(lvl)

    def forward_solve(self, lvl):
        torch = _torch()
        lv, plan = self.s._lv[lvl], self.plans[lvl]
        if plan['seq'] is None:
            t = np.asarray(lv.t, dtype=float)
            plan['seq'] = [_Steps(torch, self.dev, [i - 1], [i], t) for i in range(1, lv.npts)]
        for st in plan['seq']:
            self._phi(lvl, st, lv.u, lv.u)
            if lv.g is not None:
                self._lin(1, lv, lv.u, st.dst, 1.0, lv.g, st.dst, 1.0, lv.u, st.dst)

    def inject_up(self, lvl):
        fine, coarse, plan = self.s._lv[lvl], self.s._lv[lvl + 1], self.plans[lvl]
        self._lin(plan['ncp'] - 1, fine, fine.u, plan['cp1'], 1.0, coarse.u, plan['j1'])

    def residual_norms(self, out_sq):
        lv, plan = self.s._lv[0], self.plans[0]
        n1 = plan['ncp'] - 1
        out_sq[:1].zero_()
        if n1 <= 0:
            return
        tmp = plan['tmp']
        rows = self._arange(n1)
        self._phi(0, plan['res'], lv.u, tmp, rows)
        self._lin(n1, lv, tmp, rows, 1.0, tmp, rows, -1.0, lv.u, plan['cp1'])                      # Phi(u[c-1]) - u[c]
        _lib.check(self.lib.mgb_rows_sumsq(n1, lv.n, tmp.data_ptr(), tmp.stride(0), None, out_sq[1:].data_ptr(),
                                           self.s._stream()), 'rows_sumsq')

    def jump_norms(self, last, out_sq):
        lv, plan = self.s._lv[0], self.plans[0]
        n1 = plan['ncp'] - 1
        out_sq[:1].zero_()
        if n1 > 0:
            tmp = plan['tmp']
            rows = self._arange(n1)
            self._lin(n1, lv, tmp, rows, 1.0, lv.u, plan['cp1'], -1.0, last, plan['cp1'])
            _lib.check(self.lib.mgb_rows_sumsq(n1, lv.n, tmp.data_ptr(), tmp.stride(0), None, out_sq[1:].data_ptr(),
                                               self.s._stream()), 'rows_sumsq')
        last.copy_(lv.u)
