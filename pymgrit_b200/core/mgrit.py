"""MGRIT (FAS) solver with the reference's interface (core/mgrit.py:20-858), executed as batched GPU sweeps.

The constructor arguments, the attributes user code reads (u, t, index_local*, conv, solve_iter, comm_time_rank...),
the log lines and the return value of solve() are the reference's.  What differs is how a sweep runs: where the
reference loops over time points and calls Application.step once per point (mgrit.py:314-327, 356-368, 407-412,
472-481, 524-547), each method below makes ONE call into libmgrit_b200 (include/mgrit_b200.h) that covers every
coarse interval of the level; the state of a level is one [points x dofs] array in HBM.

Supported on the device path: applications derived from DeviceApplication, the identity GridTransferCopy, the global
convergence criteria (conv_crit 0 and 1; the local ones, 2 and 3, on one time rank).  Anything else raises: there is no
per-point Python fallback.
"""
import logging
import sys
import time
from typing import List

import numpy as np

from pymgrit_b200 import _lib
from pymgrit_b200.core.application import Application, DeviceApplication
from pymgrit_b200.core.comm import SerialComm, as_time_comm
from pymgrit_b200.core.device_level import DeviceLevel
from pymgrit_b200.core.grid_transfer import GridTransfer, DeviceGridTransfer
from pymgrit_b200.core.grid_transfer_copy import GridTransferCopy
from pymgrit_b200.core import partition


class _LevelVectors:
    """`mgrit.u[lvl]`: sequence of Vector views onto the rows of a level array (row i = local point i)."""

    def __init__(self, level: DeviceLevel, which: str, before_read=None):
        self._level, self._which, self._before_read = level, which, before_read

    def _array(self):
        if self._before_read is not None:
            self._before_read()              # level 0: F-points that were left out while iterating are computed now
        return getattr(self._level, self._which)

    def __len__(self):
        return self._level.npts

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        i = int(i)
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        return self._level.get_vector(self._array(), i)

    def __setitem__(self, i, vec):
        i = int(i)
        if i < 0:
            i += len(self)
        self._level.set_vector(self._array(), i, vec)

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class Mgrit:
    """
    MGRIT solver class (FAS formulation) for time-stepping problems u_i = Phi(u_{i-1}).
    """

    def __init__(self, problem: List[Application], transfer: List[GridTransfer] = None, weight_c: float = 1.0,
                 max_iter: int = 100, tol: float = 1e-7, nested_iteration: bool = True, cf_iter: int = 1,
                 cycle_type: str = 'V', comm_time=None, comm_space=None,
                 logging_lvl: int = logging.INFO, output_fcn=None, output_lvl=1, t_norm=2,
                 random_init_guess: bool = False, conv_crit: int = 0) -> None:
        logging.basicConfig(format='%(levelname)s - %(asctime)s - %(message)s', datefmt='%d-%m-%y %H:%M:%S',
                            level=logging_lvl, stream=sys.stdout)
        self._log_lvl = logging_lvl
        self.setup_phases = []                 # (name, host milliseconds) of the constructor's phases, for profiles/
        _t_phase = [time.perf_counter()]

        def phase(name):
            now = time.perf_counter()
            self.setup_phases.append((name, 1e3 * (now - _t_phase[0])))
            _t_phase[0] = now

        if transfer is None:
            transfer = [GridTransferCopy() for _ in range(len(problem) - 1)]

        # argument checks, mgrit.py:79-128
        if len(problem) != (len(transfer) + 1):
            raise Exception('There should be exactly one transfer operator for each level except the coarsest grid')
        for i in range(len(problem) - 1):
            if len(problem[i].t) < len(problem[i + 1].t):
                raise Exception(
                    'The time grid on level ' + str(i + 1) + ' contains more time points than level ' + str(i))
        if cycle_type != 'V' and cycle_type != 'F':
            raise Exception("Cycle-type " + str(cycle_type) + " is not implemented. Choose 'V' or 'F'")
        if output_lvl not in [0, 1, 2]:
            raise Exception("Unknown output level. Choose 0, 1 or 2.")
        masks = partition.c_point_masks([p.t for p in problem])
        for lvl in range(1, len(problem)):
            if getattr(masks[lvl - 1], 'stride', 0) > 0:
                continue           # t_coarse == t_fine[::m] on an increasing grid (partition.c_point_masks): nothing to count
            if int(np.count_nonzero(masks[lvl - 1])) != len(np.unique(problem[lvl].t)):
                raise Exception('Some points from level ' + str(lvl - 1) + ' are not points of level ' + str(lvl))
        if t_norm not in [1, 2, 3]:
            raise Exception('Unknown norm. Please choose 1 (one norm), 2 (two-norm) or 3 (inf-norm)')
        if conv_crit not in [0, 1, 2, 3]:
            raise Exception('Unknown convergence criterion. Please choose: '
                            '0 (global space-time residual), 1 (global jump)'
                            '2 (local space-time residual)3 (local jump)')
        if isinstance(cf_iter, int):
            cf_iter = [cf_iter for _ in range(len(problem))]
        elif isinstance(cf_iter, list):
            if len(cf_iter) < len(problem) - 1:
                raise Exception('Too few cf_iter. '
                                'Specify a list of values for all but the coarsest level or an integer '
                                '(used for all levels).')
        else:
            raise Exception('Incorrect datatype cf_iter. '
                            'Specify a list of values for all but the coarsest level or an integer '
                            '( used for all levels).')

        problem = self._choose_representation(problem, transfer)
        # what the device path does not cover fails loudly (no per-point fallback)
        for p in problem:
            if not isinstance(p, DeviceApplication):
                raise Exception('pymgrit_b200.Mgrit runs applications derived from DeviceApplication only; '
                                + type(p).__name__ + ' has no device kernels')
        # grid transfers: the identity is fused into the sweeps; a DeviceGridTransfer (spatial coarsening as row-wise
        # kernels) splits the FAS restriction and the correction around it; a transfer written in Python cannot run
        # inside a sweep
        batched = [p.kind == _lib.APP_BATCHED for p in problem]
        if any(batched) and not (all(batched) and all(type(tr) is GridTransferCopy for tr in transfer)):
            raise Exception('applications on the batched path (BatchedApplication) need every level to be one and the '
                            'identity transfer GridTransferCopy between the levels')
        for lvl, tr in enumerate(transfer):
            if type(tr) is GridTransferCopy:
                if (problem[lvl].kind, problem[lvl].ndof) != (problem[lvl + 1].kind, problem[lvl + 1].ndof):
                    raise Exception('levels connected by GridTransferCopy must use the same application kind and '
                                    'spatial size')
            elif isinstance(tr, DeviceGridTransfer):
                tr.check(problem[lvl], problem[lvl + 1])
            else:
                raise Exception('only GridTransferCopy and device grid transfers (DeviceGridTransfer, e.g. '
                                'GridTransferHeat1D) run in the device sweeps; ' + type(tr).__name__ + ' does not')
        # conv_crit 2 / 3 (local criteria, mgrit.py:434-454) let a time rank stop once its own points and all earlier
        # ranks have converged.  On one rank that is the global test on the same per-point norms; the rank-by-rank
        # shutdown protocol (message kind 6, sender_finished) is not built.

        phase('argument checks, C-point masks, representation')
        self.comm_time = as_time_comm(comm_time)
        if comm_space is not None and hasattr(comm_space, 'Get_size') and comm_space.Get_size() > 1:
            # the reference hands comm_space to applications that are parallel in space (PETSc, mgrit.py:129-140); every
            # application of this package lives on one GPU
            raise Exception('comm_space with more than one process is not supported: the applications of pymgrit_b200 are '
                            'not parallel in space (one GPU per time rank)')
        self.comm_space = comm_space
        self.comm_time_rank = self.comm_time.Get_rank()
        self.comm_time_size = self.comm_time.Get_size()
        if self.comm_time_size > len(problem[0].t):
            raise Exception('More processors than time points. Not useful and not implemented yet')
        self.spatial_parallel = False
        self.comm_space_rank = -99
        self.comm_space_size = 1

        if any(batched) and self.comm_time_size > 1:
            raise Exception('the batched path (BatchedApplication) runs on one time rank')
        if conv_crit in (2, 3) and self.comm_time_size > 1:
            raise Exception('local convergence criteria (conv_crit 2, 3) are available on one time rank only; '
                            'use conv_crit 0 or 1 with several ranks')
        self.comm_time.barrier()
        runtime_setup_start = time.time()
        self.log_info(f"Start setup")
        if conv_crit in (2, 3):
            self.log_info(f"A local criterion is used. The following output describes only the convergence of "
                          f"the time points of one process")

        self.launches = 0
        self._f_stale = False
        self._sweep_counts = {}                # level-0 sweeps launched by the last solve(), by name (time_level0_sweeps)
        self._l0_store_hook = None             # set by _solve_queued_ahead: is the cycle being queued expected to be the last?
        self._f_uninit = False
        import os as _os
        self._lazy_f = _os.environ.get('MGB_LAZY_F', '1') != '0'
        self.problem = problem
        self.weight_c = weight_c
        self.lvl_max = len(problem)
        self.step = [p.step for p in problem]
        self.t = []
        self.m = []
        self.restriction = [tr.restriction for tr in transfer]
        self.interpolation = [tr.interpolation for tr in transfer]
        self.tol = tol
        self.conv = np.zeros(max_iter + 1)
        self.cf_iter = cf_iter
        self.cycle_type = cycle_type
        self.random_init_guess = random_init_guess
        self.iter_max = max_iter
        self.solve_iter = 0
        self.nes_it = nested_iteration
        self.runtime_solve = 0
        self.runtime_setup = 0
        self.t_norm = 1 if t_norm == 1 else None if t_norm == 2 else np.inf
        self._t_norm_id = t_norm
        self.conv_crit = conv_crit
        self.global_conv_crit = conv_crit in (0, 1)
        self._jump_crit = conv_crit in (1, 3)
        self.save_values_last_iter = None
        self.output_lvl = output_lvl
        self.output_fcn = output_fcn if output_fcn is not None and callable(output_fcn) else None

        # index tables (mgrit.py:206-222, 742-827) and level storage in HBM (mgrit.py:840-858)
        self.global_t = [np.asarray(p.t, dtype=float) for p in problem]      # read-only here: no 8 MB copies at nt = 2^20
        part = partition.Partition(self.global_t, self.comm_time_size, self.comm_time_rank, masks=masks)
        self._part = part
        self.m = part.m
        self.cpts = part.cpts
        self.index_local = part.index_local
        self.index_local_c = part.index_local_c
        self.t = part.t_local
        self.int_start, self.int_stop = part.int_start, part.int_stop
        self.send_to, self.get_from = part.send_to, part.get_from
        phase('partition')
        self._lv = []
        for lvl in range(self.lvl_max):
            cp = part.sweep_cpts[lvl] if lvl < self.lvl_max - 1 else None
            # level 0's tables (the long ones) are built after the coarse part of nested iteration has been queued
            defer = lvl == 0 and self.lvl_max > 1 and bool(nested_iteration) and not random_init_guess
            # ... and its 8.6 GB array is not zero-filled when nested iteration and at least one cycle follow: the
            # C-points come from the interpolation, the F-points from the F-relaxations, before anything reads them
            # (restart() relies on the same fact).  Only a user who looks at F-points of level 0 between the
            # constructor and the first iteration would see the difference, so a per-iteration output_fcn keeps the fill.
            self._lv.append(DeviceLevel(problem[lvl], part.t_local[lvl], cpts=cp, with_g=lvl > 0, defer_tables=defer,
                                        zero_u=not (defer and max_iter > 0 and not (callable(output_fcn) and
                                                                                    output_lvl == 2))
                                        or _os.environ.get('MGB_ZERO_U') == '1'))
            if defer:
                self._lv[0].start_host_tables()      # the long host tables of level 0 are made while the rest is set up
        self._f_uninit = self.lvl_max > 1 and not self._lv[0].zero_filled
        phase('level arrays + coarse-level tables')
        # levels whose down-sweep runs as one fused launch (mgb_down_sweep): unweighted C-relaxation, at least one
        # F-point in every interval (on every time rank), team kernels
        import os
        self._xfer = [None if type(tr) is GridTransferCopy else tr for tr in transfer] + [None]
        flags = []
        for lvl in range(self.lvl_max - 1):
            cp = self._lv[lvl].cpts
            ok = (weight_c == 1.0 and self._xfer[lvl] is None and not batched[lvl]
                  and problem[lvl].kind in (_lib.APP_HEAT1D, _lib.APP_ADVECTION1D, _lib.APP_HEAT2D, _lib.APP_HEAT1D_2PTS,
                                            _lib.APP_HEAT1D_SINE)
                  and (cp is None or len(cp) < 2 or int(np.min(np.diff(cp))) >= 2)
                  and os.environ.get('MGB_FUSED_DOWN', '1') != '0')
            flags.append(ok)
        # the sequential solve on the coarsest level in sine space where the application offers it (Heat1D)
        self._spectral = {}
        last = self._lv[-1]
        maker = getattr(problem[-1], 'spectral_solver', None)
        min_pts = problem[-1].spectral_min_points() if hasattr(problem[-1], 'spectral_min_points') else 1 << 30
        if maker is not None and getattr(problem[-1], 'spectral_single_rank_only', False) and self.comm_time_size > 1:
            maker = None                     # (Advection1D's Fourier solve: the rank-to-rank chain stays)
        flags.append(maker is not None and last.npts >= min_pts)
        flags = self.comm_time.all_true(flags)          # every rank must take the same path: one small all-reduce
        self._fused_down = flags[:-1] + [False]
        if flags[-1]:
            sp = maker(last) if last.npts > 0 else None
            if sp is not None:
                self._spectral[self.lvl_max - 1] = sp
        self.comm_time.setup_peer_exchange(self)         # ghost rows over peer memory where the ranks can (core/comm.py)
        import weakref
        me = weakref.ref(self)               # no reference cycle: a dropped solver frees its 8.6 GB level at once

        def before_read():
            solver = me()
            if solver is not None:
                solver._materialise_f_points()
        self.u = [_LevelVectors(lv, 'u', before_read if k == 0 else None) for k, lv in enumerate(self._lv)]
        self.g = [None] + [_LevelVectors(lv, 'g') for lv in self._lv[1:]]
        self.v = [None] * self.lvl_max       # never materialised: identical to the fine level's C-point rows ...
        self._rres = [None] * self.lvl_max   # ... except below a spatial transfer: v, fine residual rows, their restriction
        self._rresc = [None] * self.lvl_max
        torch = _lib_torch()
        for lvl in range(self.lvl_max - 1):
            if self._xfer[lvl] is not None:
                fine, coarse = self._lv[lvl], self._lv[lvl + 1]
                ncp = len(fine.cpts)
                coarse.v = torch.zeros_like(coarse.u)
                self.v[lvl + 1] = _LevelVectors(coarse, 'v')
                self._rres[lvl] = torch.zeros((ncp, fine.pitch), dtype=torch.float64, device=fine.u.device)
                self._rresc[lvl] = torch.zeros((ncp, coarse.pitch), dtype=torch.float64, device=fine.u.device)
        self._batched = None
        if any(batched):
            from pymgrit_b200.core.batched import BatchedSweeps
            self._batched = BatchedSweeps(self)
        phase('path flags, coarsest solver, peer exchange')
        self._init_levels()
        phase('initial condition')
        if nested_iteration:
            self.nested_iteration()
        self._lv[0].finish_tables()
        phase('nested iteration queued + level-0 tables')
        dev = self._lv[0].u.device
        ncp0 = len(self._lv[0].cpts) if self._lv[0].cpts is not None else 1
        nsys0 = self._lv[0].nsys
        self._sq = torch.zeros(max(ncp0, 1) * (1 + nsys0 if nsys0 > 1 else 1), dtype=torch.float64, device=dev)
        self._norm_out = torch.zeros(1, dtype=torch.float64, device=dev)
        self._norm_host = torch.zeros(1, dtype=torch.float64).pin_memory()

        if self._jump_crit:
            self.save_values_last_iter = self._lv[0].u.clone()

        if self.iter_max == 0:
            self.comm_time.barrier()
        phase('norm buffers')
        if self._log_lvl <= logging.INFO or self.iter_max == 0 or (self.output_fcn is not None and self.output_lvl == 2):
            torch.cuda.synchronize()         # "Setup took ..." is printed: make it the time the setup really took
        # otherwise the sweeps queued by the setup (nested iteration) are still running when the constructor returns and
        # solve() queues behind them: time_setup + time_solve is unchanged, the host does not idle in between
        phase('wait for the device')
        self.runtime_setup = time.time() - runtime_setup_start

        if self.output_fcn is not None and self.output_lvl == 2:
            self.output_fcn(self)

        self.log_info(f"Setup took {self.runtime_setup} s")

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _choose_representation(problem, transfer):
        """A hierarchy of Heat1D levels of one spatial size connected by the identity transfer runs with its rows in sine
        space (heat/heat_1d.py, csrc/phi.cuh Heat1DSine): Phi is then diagonal and every sweep HBM-bound.  Returns the
        list of applications the solver runs (twins of the user's objects, or the objects themselves)."""
        from pymgrit_b200.heat.heat_1d import Heat1D
        if not all(type(tr) is GridTransferCopy for tr in transfer):
            return problem
        if not all(isinstance(p, Heat1D) and p.kind == _lib.APP_HEAT1D for p in problem):
            return problem
        if len({(p.nx, float(p.a), float(p.dx)) for p in problem}) != 1:
            return problem
        want = [p.sine_space for p in problem]
        if not all(p.can_sine() for p in problem):
            if any(w is True for w in want):
                raise Exception('sine_space=True needs a zero or separable right-hand side on every level and a supported '
                                'spatial size')
            return problem
        return [p.as_sine() for p in problem]

    def _materialise_f_points(self):
        """Level 0 while iterating keeps only the C-points and the last F-point of every interval up to date (all that the
        residual and the next C-relaxation read, mgb_error_correction MGB_CORRECT_LAST_ONLY).  Before anybody looks at
        the level -- the end of solve(), an output function, Mgrit.u[0][i] -- one F-relaxation computes the others: the
        values the reference holds at that moment (its F-points are the Phi chains from the C-points, mgrit.py:287)."""
        if self._f_uninit:
            # level 0 was allocated without a fill (nested iteration and the first cycle write every row before a sweep
            # reads it); somebody looks at it before that first cycle: the F-points hold the reference's zero initial guess
            self._f_uninit = False
            lv = self._lv[0]
            if lv.cpts is not None and lv.npts > 0:
                torch = _lib_torch()
                is_f = torch.ones(lv.npts, dtype=torch.bool, device=lv.u.device)
                is_f[torch.as_tensor(np.asarray(lv.cpts, dtype=np.int64), device=lv.u.device)] = False
                lv.u[is_f] = 0.0
        if self._f_stale:
            self._f_stale = False
            self.f_relax(lvl=0)

    def _init_levels(self):
        """mgrit.py:846-858: zero (or random) initial guess, initial condition at the first point of rank 0."""
        if self.random_init_guess:
            lv = self._lv[0]
            shape = lv.app.vector_template.shape
            vals = np.stack([np.random.rand(*shape) if shape else np.random.rand(1) for _ in range(lv.npts)])
            for a in range(0, lv.npts, 64):
                lv.app.values_to_rows(_lib_torch().as_tensor(vals[a:a + 64]).to(lv.u.device), lv.u[a:a + 64])
        if self.comm_time_rank == 0:
            for lv in self._lv:
                if lv.npts > 0:
                    lv.set_vector(lv.u, 0, lv.app.vector_t_start)

    def _all_ranks(self, flag: bool) -> bool:
        """True if `flag` holds on every time rank (all ranks must take the same path through a chain exchange)."""
        return self.comm_time.all_true([flag])[0]

    def log_info(self, message: str) -> None:
        """Only the last time rank logs (mgrit.py:247-259)."""
        if self.comm_time_rank == self.comm_time_size - 1:
            logging.info(message)

    def _stream(self):
        self.launches += 1                      # every C-ABI sweep below is one kernel launch of this library
        return _lib.current_stream_ptr()

    def _count(self, lvl, name):
        if lvl == 0:
            self._sweep_counts[name] = self._sweep_counts.get(name, 0) + 1

    def restart(self) -> None:
        """Forget the current iterate and redo the setup sweeps (initial guess, nested iteration) with the level tables
        that are already in HBM.  Not in the reference (which rebuilds everything); used by bench.py."""
        for lv in self._lv:
            if not self.nes_it:
                lv.u.zero_()         # with nested iteration every point is written before it is read
            if lv.g is not None:
                lv.g.zero_()
        self.conv = np.zeros(self.iter_max + 1)
        self.solve_iter = 0
        self._f_stale = False
        self._init_levels()
        if self.nes_it:
            self.nested_iteration()
        if self._jump_crit:
            self.save_values_last_iter = self._lv[0].u.clone()

    @property
    def h2d_bytes(self):
        """Bytes copied host -> device while building the levels (tables, time grids, initial condition)."""
        return (sum(lv.h2d_bytes for lv in self._lv) + 8 * self._lv[0].n * self.lvl_max +
                sum(sp.h2d_bytes for sp in self._spectral.values()))

    def time_level0_sweeps(self, repeats: int = 5, iterations: int = 1, hbm_gbs: float = None):
        """CUDA-event timing of each level-0 sweep alone (bench.py roofline).  Per sweep: the algorithmic bytes (every
        row an interval must read or write, counted once), ms, GB/s, how often it runs per
        iteration / per solve of this solver, and what bounds it.  Two lower bounds are computed per sweep: the
        algorithmic bytes at the HBM peak, and -- for the heat_1d kernels, whose FP64 instruction count per Phi is known
        from the SASS (251 per 33 unknowns for the Toeplitz solve; 1 + nrhs per unknown in sine space) -- the FP64 warp
        instructions at the issue rate of the pipe measured by scripts/micro/fp64_latency.cu (one warp instruction per
        2.07 cycles and scheduler; self-measured, MEASURED_PEAKS.json has no FP64 entry).  `bound` names the larger one,
        `frac` = that bound / measured time."""
        torch = _lib_torch()
        lv = self._lv[0]
        if self.lvl_max < 2 or lv.cpts is None:
            return []
        counts = dict(self._sweep_counts)      # what the last solve() launched on level 0 (the calls below count too)
        k0 = len(lv.cpts) - 1
        m = self.m[0]
        row = 8.0 * lv.n
        fused = self._fused_down[0]
        cf = self.cf_iter[0]
        lazy = self._lazy_f
        its = max(1, int(iterations))
        first_f = 1                     # the F-relaxation that opens iteration 0 (mgrit.py:274-275)
        sweeps = [
            # name, launch, rows per interval, Phi per interval, launches per iteration, launches per solve
            ('f_relax', lambda: self.f_relax(0), m, m - 1, 0, 1 if lazy else 0),
            ('f_relax(last point only)', lambda: self.f_relax(0, last_only=True), 2, m - 1,
             cf - 1, first_f),
            ('c_relax', lambda: self.c_relax(0), 2 + (1 if self.weight_c != 1.0 else 0), 1,
             cf - (1 if fused else 0), 0),
            ('fas_residual', lambda: self.fas_residual(0), 4, 2, 0 if fused else 1, 0),
            ('error_correction+f_relax', lambda: self.error_correction(0, f_relax=True), m + 2, m - 1,
             0 if lazy else 1, 0),
            ('error_correction+f_relax(last point only)',
             lambda: self.error_correction(0, f_relax=True, last_only=True), 4, m - 1, 1 if lazy else 0, 0),
            ('residual_norms', lambda: self.compute_residual(), 2, 1, 1, 0),
        ]
        if fused:
            sweeps.insert(4, ('down_sweep(c_relax+f_relax+fas_residual)', lambda: self.down_sweep(0), 4, m + 3, 1, 0))
        else:
            sweeps[1] = sweeps[1][:4] + (cf, first_f)
        props = torch.cuda.get_device_properties(lv.u.device)
        sms = props.multi_processor_count
        clock_hz = float(getattr(props, 'clock_rate', 1965000)) * 1e3          # maximum SM clock (kHz -> Hz)
        fp64_peak = sms * 4 * clock_hz / 2.07            # FP64 warp instructions per second (scripts/micro/fp64_latency.cu)
        kind = self.problem[0].kind
        nrhs = int(lv.c.nrhs)
        if kind == _lib.APP_HEAT1D:
            per_elem = 251.0 / 33.0
        elif kind == _lib.APP_HEAT1D_SINE:
            per_elem = 1.0 + nrhs
        else:
            per_elem = None
        team_elems = lv.team_threads * lv.chunk * max(1, lv.nsys)
        out = []
        for name, fn, rows, nphi, per_iter, per_solve in sweeps:
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(repeats):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / repeats
            nbytes = rows * k0 * row
            ent = {'name': name, 'rows_per_interval': rows, 'phi_per_interval': nphi, 'intervals': k0,
                   'algorithmic_bytes': nbytes, 'ms': ms, 'gbs': nbytes / (ms * 1e-3) / 1e9,
                   'launches_per_iteration': per_iter, 'launches_per_solve': per_solve,
                   'share_ms': ms * (per_iter * its + per_solve)}
            if counts:
                # counted during the last solve: the cycle predicted to be the last stores every F-point in its correction
                # and the final F-relaxation does not run (_solve_queued_ahead), which the static columns do not know
                ent['launches_last_solve'] = counts.get(name, 0)
                ent['share_ms'] = ms * counts.get(name, 0)
            t_hbm = nbytes / (hbm_gbs * 1e9) * 1e3 if hbm_gbs else None
            t_fp = None
            if per_elem is not None:
                fp64_instr = nphi * k0 * per_elem * team_elems / 32.0       # warp instructions
                t_fp = fp64_instr / fp64_peak * 1e3
                ent['fp64_frac'] = t_fp / ms
            if t_hbm is not None:
                ent['hbm_frac'] = t_hbm / ms
            ent['bound'] = 'fp64' if (t_fp is not None and t_hbm is not None and t_fp > t_hbm) else 'hbm'
            ent['frac'] = max(v for v in (t_hbm, t_fp) if v is not None) / ms if (t_hbm or t_fp) else None
            out.append(ent)
        self.restart_needed = True
        return out

    # ------------------------------------------------------------------------------------------
    def iteration(self, lvl: int, cycle_type: str, iteration: int, first_f: bool) -> None:
        """One MGRIT cycle from level lvl downwards (mgrit.py:261-290)."""
        if lvl == self.lvl_max - 1:
            self.forward_solve(lvl=lvl)
            return
        # Down-sweep: C-relaxation and the FAS restriction read only the last F-point of an interval, and the
        # F-relaxation fused into error_correction() below rewrites every F-point, so the F-relaxations here keep
        # only those last points (same values, the dead stores are skipped).
        if lvl == 0:
            self._f_uninit = False           # from here on the F-points are the cycle's business (_f_stale)
        if (lvl > 0 or (iteration == 0 and lvl == 0)) and first_f:
            self.f_relax(lvl=lvl, last_only=True)
        fused = False
        for k in range(self.cf_iter[lvl]):
            if k == self.cf_iter[lvl] - 1 and self._fused_down[lvl]:
                self.down_sweep(lvl=lvl)            # C-relaxation, F-relaxation and FAS restriction in one launch
                fused = True
            else:
                self.c_relax(lvl=lvl)
                self.f_relax(lvl=lvl, last_only=True)
        if not fused:
            self.fas_residual(lvl=lvl)
        self.iteration(lvl=lvl + 1, cycle_type=cycle_type, iteration=iteration, first_f=True)
        # correction + the F-relaxation of mgrit.py:287, one launch; on level 0 only the last F-point of every interval
        # is stored while the solver iterates (_materialise_f_points)
        lazy = lvl == 0 and self._lazy_f and self.lvl_max > 1 and self._batched is None
        # F-cycle: a V-cycle from this level follows at once.  Its down-sweep reads only the last F-point of every interval
        # and its own correction + F-relaxation rewrites all of them, so the F-points stored here would be dead: same values
        # with last_only (on cfg 3 that is 6 of the 17 rows a level-1 interval moves).
        again = lvl != 0 and cycle_type == 'F'
        dead_f = again and self._lazy_f and self._batched is None and self._xfer[lvl] is None
        stale = lazy
        if lazy and self._l0_store_hook is not None and self._l0_store_hook():
            lazy = False                 # expected to be the last cycle: its correction stores every F-point (1.55 instead of
                                         # 0.22 ms on cfg 5) and the final F-relaxation (1.47 ms + a host round trip) is not needed
        self.error_correction(lvl=lvl, f_relax=True, last_only=lazy or dead_f)
        if stale:
            self._f_stale = lazy
        if again:
            self.iteration(lvl=lvl, cycle_type='V', iteration=iteration, first_f=False)

    def f_relax(self, lvl: int, last_only: bool = False) -> None:
        """F-relaxation (mgrit.py:292-333): one launch over all coarse intervals.  last_only: store only the last
        F-point of every interval (down-sweep of a cycle, where the other F-points are never read)."""
        if self._batched is not None:
            return self._batched.f_relax(lvl, last_only)
        flags = _lib.F_RELAX_LAST_ONLY if last_only else 0
        self._count(lvl, 'f_relax(last point only)' if last_only else 'f_relax')
        _lib.check(_lib.lib().mgb_f_relax(self._lv[lvl].ref, flags, self._stream()), 'f_relax')

    def c_relax(self, lvl: int) -> None:
        """C-relaxation (mgrit.py:335-370)."""
        if self._batched is not None:
            return self._batched.c_relax(lvl)
        self._count(lvl, 'c_relax')
        _lib.check(_lib.lib().mgb_c_relax(self._lv[lvl].ref, float(self.weight_c), self._stream()), 'c_relax')
        self._exchange_ghost(lvl)

    def fas_residual(self, lvl: int) -> None:
        """Injection + FAS right-hand side of the next coarser level (mgrit.py:488-549)."""
        if self._batched is not None:
            return self._batched.fas_residual(lvl)
        fine, coarse = self._lv[lvl], self._lv[lvl + 1]
        xfer = self._xfer[lvl]
        if xfer is not None:
            # spatial transfer R: fine residual rows, R of every C-point and of the residuals, v = copy, coarse FAS rhs
            ncp = len(fine.cpts)
            lib = _lib.lib()
            _lib.check(lib.mgb_residual_rows(fine.ref, self._rres[lvl].data_ptr(), self._stream()), 'residual_rows')
            xfer.restrict_rows(ncp, fine.u, fine.cpts_dev, coarse.u, fine.app)
            coarse.v[:ncp].copy_(coarse.u[:ncp])
            xfer.restrict_rows(ncp, self._rres[lvl], None, self._rresc[lvl], fine.app)
            self.launches += 3
            _lib.check(lib.mgb_fas_coarse_rhs(coarse.ref, coarse.v.data_ptr(), self._rresc[lvl].data_ptr(), ncp,
                                              self._stream()), 'fas_coarse_rhs')
            return
        if coarse.npts > 0 and fine.npts > 0 and len(fine.cpts) < 2:
            coarse.u[0].copy_(fine.u[0])             # no interval: the kernel (which injects point 0 too) has no work
        self._count(lvl, 'fas_residual')
        _lib.check(_lib.lib().mgb_fas_residual(fine.ref, coarse.ref, self._stream()), 'fas_residual')

    def down_sweep(self, lvl: int) -> None:
        """c_relax + f_relax(last_only) + fas_residual of level lvl in one pass over the level (same values)."""
        fine, coarse = self._lv[lvl], self._lv[lvl + 1]
        if self.comm_time_size > 1:
            # the next rank's ghost is my last C-point after its C-relaxation: relax that one point first, send it, and
            # let the fused pass recompute it
            if self.comm_time_rank + 1 < self.comm_time_size and fine.npts > 1:
                _lib.check(_lib.lib().mgb_c_relax_last(fine.ref, 1.0, self._stream()), 'c_relax_last')
            self._exchange_ghost(lvl)
        if coarse.npts > 0 and fine.npts > 0 and len(fine.cpts) < 2:
            coarse.u[0].copy_(fine.u[0])             # no interval: the kernel (which injects point 0 too) has no work
        self._count(lvl, 'down_sweep(c_relax+f_relax+fas_residual)')
        _lib.check(_lib.lib().mgb_down_sweep(fine.ref, coarse.ref, self._stream()), 'down_sweep')

    def error_correction(self, lvl: int, f_relax: bool = False, last_only: bool = False) -> None:
        """Coarse-grid correction of the C-points (mgrit.py:715-726), optionally fused with the next F-relaxation
        (last_only: of which only the last point of every interval is stored)."""
        if self._batched is not None:
            return self._batched.error_correction(lvl, f_relax, last_only)
        xfer = self._xfer[lvl]
        if xfer is not None:
            fine, coarse = self._lv[lvl], self._lv[lvl + 1]
            first = 0 if self.comm_time_rank > 0 else 1           # ranks > 0 correct their ghost copy themselves
            xfer.interpolate_rows(len(fine.cpts), first, coarse.u, coarse.v, fine.u, fine.cpts_dev, True, coarse.app)
            self.launches += 1
            if f_relax:
                self.f_relax(lvl, last_only=last_only)
            return
        flags = ((_lib.CORRECT_F_RELAX if f_relax else 0) | (_lib.CORRECT_GHOST if self.comm_time_rank > 0 else 0) |
                 (_lib.CORRECT_LAST_ONLY if f_relax and last_only else 0))
        if f_relax:
            self._count(lvl, 'error_correction+f_relax' + ('(last point only)' if last_only else ''))
        _lib.check(_lib.lib().mgb_error_correction(self._lv[lvl].ref, self._lv[lvl + 1].ref, flags, self._stream()),
                   'error_correction')
        # no exchange: every rank corrects its ghost copy itself (MGB_CORRECT_GHOST)

    def forward_solve(self, lvl: int) -> None:
        """Sequential time stepping on level lvl (mgrit.py:459-486).  Heat1D levels that are long enough are solved
        in sine space (csrc/spectral.cu): two transforms and n independent scalar recurrences instead of a chain of
        tridiagonal solves; the time ranks exchange one all-gather of two rows each instead of the rank-to-rank
        chain (mgrit.py:467-484)."""
        if self._batched is not None:
            return self._batched.forward_solve(lvl)
        lv = self._lv[lvl]
        sp = self._spectral.get(lvl) if lv.npts > 0 else None
        if sp is None:
            self.comm_time.recv_chain(self, lvl)
            _lib.check(_lib.lib().mgb_forward_solve(lv.ref, self._stream()), 'forward_solve')
            self.comm_time.send_chain(self, lvl)
            return
        self.launches += sp.solve(self.comm_time)

    def nested_iteration(self) -> None:
        """Coarsest solve, then interpolate upwards with a V-cycle per level (mgrit.py:551-566)."""
        self.forward_solve(self.lvl_max - 1)
        for lvl in range(self.lvl_max - 2, -1, -1):
            if lvl == 0:
                self._lv[0].finish_tables()      # host work of the setup, overlapped with the sweeps queued above
            if self._xfer[lvl] is not None:
                fine, coarse = self._lv[lvl], self._lv[lvl + 1]
                self._xfer[lvl].interpolate_rows(len(fine.cpts), 1, coarse.u, None, fine.u, fine.cpts_dev, False,
                                                 coarse.app)
                self.launches += 1
            elif self._batched is not None:
                self._batched.inject_up(lvl)
            else:
                _lib.check(_lib.lib().mgb_inject_up(self._lv[lvl].ref, self._lv[lvl + 1].ref, self._stream()),
                           'inject_up')
            self._exchange_ghost(lvl)
            if lvl > 0:
                self.iteration(lvl=lvl, cycle_type='V', iteration=0, first_f=True)

    def _exchange_ghost(self, lvl: int) -> None:
        """After the C-points of a level changed: my last point becomes the next rank's ghost (mgrit.py:305-310)."""
        self.comm_time.exchange_ghost(self, lvl)

    # ------------------------------------------------------------------------------------------
    def compute_residual(self):
        """Squared residual norms at the local C-points of level 0 (mgrit.py:387-413), left on the device."""
        if self._batched is not None:
            self._batched.residual_norms(self._sq)
            return self._sq
        self._count(0, 'residual_norms')
        _lib.check(_lib.lib().mgb_residual_norms(self._lv[0].ref, self._sq.data_ptr(), self._stream()), 'residual_norms')
        return self._sq

    def compute_jump(self):
        """Squared jump norms at the local C-points (mgrit.py:372-385)."""
        if self._batched is not None:
            self._batched.jump_norms(self.save_values_last_iter, self._sq)
            return self._sq
        _lib.check(_lib.lib().mgb_jump_norms(self._lv[0].ref, self.save_values_last_iter.data_ptr(), self._sq.data_ptr(),
                                             self._stream()), 'jump_norms')
        return self._sq

    def convergence_criterion(self, iteration: int) -> None:
        """Global criterion (mgrit.py:415-432): temporal norm of the per-point norms, reduced over the time ranks."""
        lv0 = self._lv[0]
        ncp = 0 if lv0.cpts is None else len(lv0.cpts)
        if ncp > 0:
            sq = self.compute_jump() if self._jump_crit else self.compute_residual()
            _lib.check(_lib.lib().mgb_temporal_norm(sq.data_ptr(), ncp, self._t_norm_id, self._norm_out.data_ptr(),
                                                    self._stream()), 'temporal_norm')
        else:
            self._norm_out.zero_()
        part = self.comm_time.reduce_norm(self._norm_out, self._t_norm_id)
        self._norm_host.copy_(part, non_blocking=False)
        val = float(self._norm_host[0])
        self.conv[iteration] = np.sqrt(val) if self._t_norm_id == 2 else val

    # ------------------------------------------------------------------------------------------
    def ouput_run_information(self) -> None:
        if self._log_lvl > logging.INFO:        # nothing would be printed: skip the O(nt) max-dt scan
            return
        msg = ['Run parameter overview',
               '  ' + '{0: <25}'.format(f'time interval') + ' : ' + '[' + str(self.problem[0].t[0]) + ', ' + str(
                   self.problem[0].t[-1]) + ']',
               '  ' + '{0: <25}'.format(f'number of time points ') + ' : ' + str(len(self.problem[0].t)),
               '  ' + '{0: <25}'.format(f'max dt ') + ' : ' + str(
                   np.max(self.problem[0].t[1:] - self.problem[0].t[:-1])),
               '  ' + '{0: <25}'.format(f'number of levels') + ' : ' + str(self.lvl_max),
               '  ' + '{0: <25}'.format(f'coarsening factors') + ' : ' + str(self.m[:-1]),
               '  ' + '{0: <25}'.format(f'relaxation weight') + ' : ' + str(self.weight_c),
               '  ' + '{0: <25}'.format(f'cf_iter') + ' : ' + str(self.cf_iter[:self.lvl_max - 1]),
               '  ' + '{0: <25}'.format(f'nested iteration') + ' : ' + str(self.nes_it),
               '  ' + '{0: <25}'.format(f'cycle type') + ' : ' + str(self.cycle_type),
               '  ' + '{0: <25}'.format(f'stopping tolerance') + ' : ' + str(self.tol),
               '  ' + '{0: <25}'.format(f'time communicator size') + ' : ' + str(self.comm_time_size),
               '  ' + '{0: <25}'.format(f'space communicator size') + ' : ' + str(self.comm_space_size),
               '  ' + '{0: <25}'.format(f'convergence criterion') + ' : ' + str(self.conv_crit)]
        self.log_info(message='\n'.join(msg))

    def _conv_text(self, iteration: int) -> str:
        """Residual column of the iteration log line (mgrit.py:611-624)."""
        if self.global_conv_crit:
            return '{0: <32}'.format(f" | conv: {self.conv[iteration + 1]}")
        return '{0: <32}'.format(f" | conv on process {self.comm_time_size - 1}: {self.conv[iteration + 1]}")

    def _can_queue_ahead(self) -> bool:
        """The stopping test can run on the device, one iteration behind the host (include/mgrit_b200.h,
        mgb_convergence_flag), when nobody needs the residual of an iteration before the next one is queued: no
        per-iteration log line (logging level above INFO), no per-iteration output function, the global residual
        criterion, and only sweeps that honour the stop flag (no spatial grid transfer)."""
        import os
        return (os.environ.get('MGB_QUEUE_AHEAD', '1') != '0' and self._log_lvl > logging.INFO
                and not (self.output_fcn is not None and self.output_lvl == 2)
                and self.conv_crit == 0 and all(x is None for x in self._xfer) and self.lvl_max > 1
                and self._batched is None)

    def _solve_queued_ahead(self) -> None:
        """solve()'s loop with the host one iteration ahead of the device.  Iteration k+1 is queued before the residual of
        iteration k is known; if that residual meets the tolerance the device raises the stop flag and every sweep of
        iteration k+1 returns at once, so the iterates, the iteration count and the residual history are those of the
        plain loop -- without the device idling while the host reads one double and queues the next cycle."""
        torch = _lib_torch()
        lib = _lib.lib()
        dev = self._lv[0].u.device
        nslot = self.iter_max + 1
        if getattr(self, '_hist_dev', None) is None or len(self._hist_dev) < nslot:
            self._hist_dev = torch.zeros(nslot, dtype=torch.float64, device=dev)
            self._hist_host = torch.zeros(nslot, dtype=torch.float64).pin_memory()
            self._flag = torch.zeros(1, dtype=torch.int32, device=dev)
        flag_ptr = self._flag.data_ptr()
        _lib.check(lib.mgb_write_flag(flag_ptr, 0, self._stream()), 'write_flag')
        _lib.check(lib.mgb_set_stop_flag(flag_ptr), 'set_stop_flag')
        events = {}
        try:
            def read(k):
                """Residual of iteration k (1-based) once its copy has arrived; True if it stops the iteration."""
                events.pop(k).synchronize()
                val = float(self._hist_host[k])
                self.conv[k] = np.sqrt(val) if self._t_norm_id == 2 else val
                self.solve_iter = k
                return self.conv[k] < self.tol

            st = {'done': False, 'last_read': 0, 'iteration': 0}

            def read_up_to(k):
                while st['last_read'] < k and not st['done']:
                    st['last_read'] += 1
                    st['done'] = read(st['last_read'])

            def expect_last_cycle():
                """Called by iteration() before the level-0 correction of the cycle being queued (number k, 0-based).  The
                residual of cycle k - 1 arrives while the device works on what is already queued of cycle k (down-sweep and
                coarse levels): wait for it, and say whether linear convergence at the rate of the last two residuals
                makes cycle k the last one."""
                k = st['iteration']
                if k < 2 or st['done']:
                    return False
                read_up_to(k)
                return (not st['done']) and predicts_convergence(self.conv, k, k + 1, self.tol)

            lv0 = self._lv[0]
            # short cycles: the wait would leave the device without work.  Decided from the GLOBAL size of level 0, so that
            # every time rank takes the same path (a rank that leaves the loop inside a cycle queues one reduction less)
            big = len(self.global_t[0]) * lv0.pitch * 8 >= (1 << 30) * self.comm_time_size
            import os
            self._l0_store_hook = expect_last_cycle if (big and os.environ.get('MGB_PREDICT_LAST', '1') != '0') else None
            queued = 0
            for iteration in range(self.iter_max):
                st['iteration'] = iteration
                self.iteration(lvl=0, cycle_type=self.cycle_type, iteration=iteration, first_f=True)
                if st['done']:
                    break                            # the hook saw that the previous cycle met the tolerance: this one is void
                self._queue_convergence(iteration + 1)
                ev = torch.cuda.Event()
                self._hist_host[iteration + 1:iteration + 2].copy_(self._hist_dev[iteration + 1:iteration + 2], non_blocking=True)
                ev.record()
                events[iteration + 1] = ev
                queued = iteration + 1
                if iteration >= 1:
                    read_up_to(iteration)
                    if st['done']:
                        break                        # iteration `iteration + 1` is queued but will not run
                # If the two residuals known so far say that the iteration just queued will meet the tolerance (linear
                # convergence), wait for its residual instead of queueing one more cycle of sweeps that would return at
                # once: same result, a cycle's worth of launches less.  A wrong guess costs one host round trip.
                if predicts_convergence(self.conv, st['last_read'], queued, self.tol):
                    read_up_to(queued)
                    if st['done']:
                        break
            if not st['done']:
                read_up_to(queued)
        finally:
            self._l0_store_hook = None
            _lib.check(lib.mgb_write_flag(flag_ptr, 0, self._stream()), 'write_flag')
            _lib.check(lib.mgb_set_stop_flag(None), 'set_stop_flag')

    def _queue_convergence(self, slot: int) -> None:
        """Residual norms, temporal norm, reduction over the time ranks, and the device-side stopping test."""
        lv0 = self._lv[0]
        ncp = 0 if lv0.cpts is None else len(lv0.cpts)
        if ncp > 0 and self.comm_time_size == 1:
            sq = self.compute_residual()             # one rank: norm and stopping test in one launch
            _lib.check(_lib.lib().mgb_temporal_norm_flag(sq.data_ptr(), ncp, self._t_norm_id, self._norm_out.data_ptr(),
                                                         float(self.tol), self._hist_dev[slot:].data_ptr(),
                                                         self._flag.data_ptr(), self._stream()), 'temporal_norm_flag')
            return
        if ncp > 0:
            sq = self.compute_residual()
            _lib.check(_lib.lib().mgb_temporal_norm(sq.data_ptr(), ncp, self._t_norm_id, self._norm_out.data_ptr(),
                                                    self._stream()), 'temporal_norm')
        else:
            self._norm_out.zero_()
        part = self.comm_time.reduce_norm(self._norm_out, self._t_norm_id)
        _lib.check(_lib.lib().mgb_convergence_flag(part.data_ptr(), self._t_norm_id, float(self.tol),
                                                   self._hist_dev[slot:].data_ptr(), self._flag.data_ptr(),
                                                   self._stream()), 'convergence_flag')

    def solve(self) -> dict:
        """Iterate until the stopping criterion is met (mgrit.py:590-646)."""
        torch = _lib_torch()
        self.log_info("Start solve")
        self._sweep_counts = {}
        runtime_solve_start = time.time()
        if self._can_queue_ahead():
            self._solve_queued_ahead()
            iterations = ()
        else:
            iterations = range(self.iter_max)
        for iteration in iterations:
            self.solve_iter = iteration + 1
            time_it_start = time.time()
            self.iteration(lvl=0, cycle_type=self.cycle_type, iteration=iteration, first_f=True)
            if self._log_lvl <= logging.INFO:
                torch.cuda.synchronize()        # so that the per-iteration runtime in the log is the device time
            time_it_stop = time.time()
            self.convergence_criterion(iteration=iteration + 1)

            if iteration == 0:
                self.log_info('{0: <7}'.format(f"iter {iteration + 1}") +
                              self._conv_text(iteration) +
                              '{0: <37}'.format(f" | conv factor: -") +
                              '{0: <35}'.format(f" | runtime: {time_it_stop - time_it_start} s"))
            else:
                self.log_info('{0: <7}'.format(f"iter {iteration + 1}") +
                              self._conv_text(iteration) +
                              '{0: <37}'.format(f" | conv factor: {self.conv[iteration + 1] / self.conv[iteration]}") +
                              '{0: <35}'.format(f" | runtime: {time_it_stop - time_it_start} s"))

            if self.output_fcn is not None and self.output_lvl == 2:
                self._materialise_f_points()
                self.output_fcn(self)

            if self.conv[iteration + 1] < self.tol or iteration == self.iter_max - 1:
                break
        self._materialise_f_points()
        torch.cuda.synchronize()
        self.comm_time.check_peers()             # a bounded device-side wait that gave up (csrc/peer.cu) raises here
        self.comm_time.barrier()
        self.runtime_solve = time.time() - runtime_solve_start
        self.log_info(f"Solve took {self.runtime_solve} s")

        if self.output_fcn is not None and self.output_lvl == 1:
            self.output_fcn(self)

        self.ouput_run_information()
        return {'conv': self.conv[np.where(self.conv != 0)], 'time_setup': self.runtime_setup,
                'time_solve': self.runtime_solve}

    @property
    def index_local_f(self):
        """Local indices of the F-points per level, in the reference's visiting order (mgrit.py:171, 802)."""
        return self._part.index_local_f

    # helpers kept for API compatibility (mgrit.py:728-740, 829-838)
    def split_into(self, number_points: int, number_processes: int) -> np.ndarray:
        return partition.split_into(number_points, number_processes)

    def split_points(self, length: int, size: int, rank: int):
        return partition.split_points(length, size, rank)


def predicts_convergence(conv, last_read: int, queued: int, tol: float) -> bool:
    """True if the residuals read so far (conv[1 .. last_read]) say that the cycle queued last (number `queued`) will meet
    the tolerance, assuming linear convergence at the rate of the last two of them."""
    if last_read < 2 or queued <= last_read or not conv[last_read - 1] > 0:
        return False
    rate = conv[last_read] / conv[last_read - 1]
    return bool(0 < rate < 1 and conv[last_read] * rate ** (queued - last_read) < tol)


def _lib_torch():
    import torch
    if not torch.cuda.is_available():
        raise Exception('pymgrit_b200 needs a CUDA device: there is no CPU fallback')
    return torch
