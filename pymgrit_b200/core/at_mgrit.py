"""AT-MGRIT with the reference's interface (core/at_mgrit.py:16-249) on the batched GPU sweeps, one time rank.

AT-MGRIT replaces the sequential solve on the coarsest level by local coarse grids: every coarsest point is computed
from the previous iterate at most k-1 points back (at_mgrit.py:75-86).  Those chains are independent of each other, so
the whole level is ONE launch of mgb_local_coarse_solve (csrc/sweeps.cuh, k_window) instead of a dependent chain over
the level.  Everything else is Mgrit.

Several time ranks: the reference gives every process one coarsest point and lets it solve that point's local grid from
data gathered over black / green sub-communicators (at_mgrit.py:48-73) -- the same update rule, point by point.  Here
the coarsest level is small, so every rank assembles the whole level (its own rows summed with zeros from the others:
one all-reduce each for u and g), solves ALL local grids in one launch and keeps its own rows and its ghost row.  The
result does not depend on the number of ranks and needs no restriction on points per process.
"""
import numpy as np

from pymgrit_b200 import _lib
from pymgrit_b200.core.device_level import DeviceLevel
from pymgrit_b200.core.mgrit import Mgrit


class AtMgrit(Mgrit):
    def __init__(self, k, conv_crit=0, *args, **kwargs):
        """
        :param k: Distance of the local coarse grids
        """
        self.k = k
        if conv_crit not in [0, 1]:
            raise Exception(
                'Local convergence criteria are not implemented for AT-MGRIT. Please select a global criterion.')
        if int(k) < 1:
            raise Exception('The distance k of the local coarse grids must be at least 1')
        self._whole = None
        super().__init__(conv_crit=conv_crit, *args, **kwargs)

    def forward_solve(self, lvl: int) -> None:
        """Local coarse grid problems on the coarsest level (at_mgrit.py:37-88)."""
        if self.lvl_max == 1:
            return                                  # at_mgrit.py:45: nothing happens on a one-level hierarchy
        lv = self._lv[lvl]
        if self.comm_time_size == 1:
            old = lv.u.clone()                      # tmp_u_arr, at_mgrit.py:76
            _lib.check(_lib.lib().mgb_local_coarse_solve(lv.ref, old.data_ptr(), int(self.k), self._stream()),
                       'local_coarse_solve')
            self._keep_old = old                    # alive until the launch has run
            return
        # several time ranks: assemble the whole coarsest level on every rank, solve all local grids, keep my rows
        if self._whole is None:
            self._whole = DeviceLevel(self.problem[lvl], self.global_t[lvl], cpts=None, with_g=True)
            own = np.asarray(self._part.owned[lvl])
            self._own = (int(own[0]), int(own[-1]) + 1) if len(own) else (0, 0)
        whole = self._whole
        a, b = self._own
        off = 1 if self.comm_time_rank > 0 and b > a else 0           # my row 0 is the ghost (previous rank's last point)
        whole.u.zero_()
        whole.g.zero_()
        if b > a:
            whole.u[a:b].copy_(lv.u[off:off + b - a])
            whole.g[a:b].copy_(lv.g[off:off + b - a])
        self.comm_time.sum_rows(whole.u)
        self.comm_time.sum_rows(whole.g)
        old = whole.u.clone()
        _lib.check(_lib.lib().mgb_local_coarse_solve(whole.ref, old.data_ptr(), int(self.k), self._stream()),
                   'local_coarse_solve')
        self._keep_old = old
        if b > a:
            lv.u[:off + b - a].copy_(whole.u[a - off:b])

    def ouput_run_information(self) -> None:
        super().ouput_run_information()
        if self._log_lvl <= 20:
            self.log_info(message='  ' + '{0: <25}'.format('distance') + ' : ' + str(self.k))
