"""Time-parallel CPU oracle -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

The reference parallelises over `comm_time` ranks with mpi4py (core/mgrit.py:728-858): every rank owns a contiguous
block of time points and runs the same per-point Python loops on it.  mpi4py and mpiexec are not available in this
image, so this module restates that decomposition with forked worker processes: the level arrays of
``MgritOracle`` live in shared memory, and every sweep hands each worker one contiguous block of coarse intervals
(F-relaxation), C-points (C-relaxation, FAS residual, residual norms) -- exactly the work a time rank would do, with the
rank-to-rank messages replaced by the shared arrays.  The arithmetic per point is the serial oracle's (and therefore
the reference's: SciPy SuperLU per step); results are identical to ``MgritOracle`` bit for bit
(tests/test_oracle.py::test_parallel_oracle_is_bitwise_serial).

Used by ``bench.py --impl reference`` and the ``cpu_baseline`` leg to time the reference algorithm on all host cores.
"""
from __future__ import annotations

import multiprocessing as mp
import os
from multiprocessing import shared_memory

import numpy as np

from oracle.mgrit_oracle import MgritOracle


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _blocks(n: int, parts: int):
    """Contiguous blocks [a, b) of range(n), sizes as in Mgrit.split_into (mgrit.py:829-838)."""
    base, rem = divmod(n, parts)
    out, a = [], 0
    for p in range(parts):
        b = a + base + (1 if p < rem else 0)
        out.append((a, b))
        a = b
    return out


class ParallelMgritOracle(MgritOracle):
    """MgritOracle whose sweeps run on `workers` forked processes over shared-memory level arrays."""

    def __init__(self, problem, workers=None, **kw):
        self.workers = max(1, int(workers or host_cores()))
        self._shm, self._pipes, self._procs = [], [], []
        self._nested = kw.get('nested_iteration', True)
        kw = dict(kw)
        kw['nested_iteration'] = False                   # run it after the workers exist
        super().__init__(problem, **kw)
        import time
        t0 = time.time()
        self.u = [self._share(a) for a in self.u]
        self.g = [None] + [self._share(a) for a in self.g[1:]]
        self.v = [None] + [self._share(a) for a in self.v[1:]]
        ctx = mp.get_context('fork')
        for w in range(self.workers):
            parent, child = ctx.Pipe()
            proc = ctx.Process(target=self._worker_loop, args=(w, child), daemon=True)
            proc.start()
            self._pipes.append(parent)
            self._procs.append(proc)
        if self._nested:
            self.nested_iteration()
        if self.conv_crit == 1:
            self.last = self.u[0].copy()
        self.time_setup += time.time() - t0

    # -- shared memory ------------------------------------------------------------------------
    def _share(self, arr):
        shm = shared_memory.SharedMemory(create=True, size=max(arr.nbytes, 8))
        self._shm.append(shm)
        out = np.ndarray(arr.shape, dtype=arr.dtype, buffer=shm.buf)
        out[...] = arr
        return out

    def close(self):
        for p in self._pipes:
            try:
                p.send(None)
            except Exception:
                pass
        for pr in self._procs:
            pr.join(timeout=5)
        self._pipes, self._procs = [], []
        self.u = [np.array(a) for a in self.u]           # detach from the shared blocks before unlinking them
        self.g = [None] + [np.array(a) for a in self.g[1:]]
        self.v = [None] + [np.array(a) for a in self.v[1:]]
        for shm in self._shm:
            try:
                shm.close()
                shm.unlink()
            except Exception:
                pass
        self._shm = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- workers ------------------------------------------------------------------------------
    def _worker_loop(self, w, pipe):
        while True:
            msg = pipe.recv()
            if msg is None:
                return
            name, l, a, b = msg
            pipe.send(getattr(self, '_part_' + name)(l, a, b))

    def _run(self, name, l, n):
        """Split range(n) into one block per worker and run `_part_<name>` on each; results in block order."""
        blocks = [(a, b) for a, b in _blocks(n, self.workers) if b > a]
        for pipe, (a, b) in zip(self._pipes, blocks):
            pipe.send((name, l, a, b))
        out = [pipe.recv() for pipe, _ in zip(self._pipes, blocks)]
        self.nphi[l] += sum(r[0] for r in out)
        return [r[1] for r in out]

    # blocks of work: the same loops as the serial oracle, restricted to [a, b)
    def _part_f_relax(self, l, a, b):                    # intervals a..b-1 (interval k starts at C-point k)
        u, g, c = self.u[l], self.g[l], self.cpts[l]
        n0 = self.nphi[l]
        for k in range(a, b):
            stop = c[k + 1] if k + 1 < len(c) else len(u)
            for i in range(c[k] + 1, stop):
                u[i] = self.phi(l, u[i - 1], i) if l == 0 else g[i] + self.phi(l, u[i - 1], i)
        return self.nphi[l] - n0, None

    def _part_c_relax(self, l, a, b):                    # C-points 1+a .. b
        u, g, w, c = self.u[l], self.g[l], self.weight_c, self.cpts[l]
        n0 = self.nphi[l]
        for j in range(1 + a, 1 + b):
            i = c[j]
            if l == 0:
                u[i] = self.phi(l, u[i - 1], i) * w + u[i] * (1.0 - w)
            else:
                u[i] = (g[i] + self.phi(l, u[i - 1], i)) * w + u[i] * (1.0 - w)
        return self.nphi[l] - n0, None

    def _part_fas(self, l, a, b):
        u, g, c, v = self.u[l], self.g[l], self.cpts[l], self.v[l + 1]
        n0, n1 = self.nphi[l], self.nphi[l + 1]
        for j in range(1 + a, 1 + b):
            if l == 0:
                fine = self.phi(l, u[c[j] - 1], c[j]) - u[c[j]]
            else:
                fine = g[c[j]] - u[c[j]] + self.phi(l, u[c[j] - 1], c[j])
            self.g[l + 1][j] = fine + v[j] - self.phi(l + 1, v[j - 1], j)
        return (self.nphi[l] - n0) + (self.nphi[l + 1] - n1), None

    def _part_residual(self, l, a, b):
        u, p, c = self.u[0], self.problem[0], self.cpts[0]
        n0 = self.nphi[0]
        out = [p.norm(self.phi(0, u[c[j] - 1], c[j]) - u[c[j]]) for j in range(1 + a, 1 + b)]
        return self.nphi[0] - n0, out

    # -- sweeps (parent side) -----------------------------------------------------------------
    def _have_workers(self):
        return bool(self._pipes)

    def f_relax(self, l):
        if not self._have_workers():
            return super().f_relax(l)
        self._run('f_relax', l, len(self.cpts[l]))

    def c_relax(self, l):
        c = self.cpts[l]
        if not self._have_workers() or np.any(np.diff(c) == 1):     # adjacent C-points are order-dependent: serial
            return super().c_relax(l)
        self._run('c_relax', l, len(c) - 1)

    def fas_residual(self, l):
        if not self._have_workers():
            return super().fas_residual(l)
        c = self.cpts[l]
        self.u[l + 1][:len(c)] = self.u[l][c]
        self.v[l + 1][...] = self.u[l + 1]
        self._run('fas', l, len(c) - 1)

    def residual_norms(self):
        if not self._have_workers():
            return super().residual_norms()
        parts = self._run('residual', 0, len(self.cpts[0]) - 1)
        return [x for part in parts for x in part]
