"""AT-MGRIT with the reference's interface (core/at_mgrit.py:16-249) on the batched GPU sweeps, one time rank.

AT-MGRIT replaces the sequential solve on the coarsest level by local coarse grids: every coarsest point is computed
from the previous iterate at most k-1 points back (at_mgrit.py:75-86).  Those chains are independent of each other, so
the whole level is ONE launch of mgb_local_coarse_solve (csrc/sweeps.cuh, k_window) instead of a dependent chain over
the level.  Everything else is Mgrit.

The reference's multi-rank variant (at_mgrit.py:48-73: one local grid per process, black / green sub-communicators) is
a different update rule whose result depends on the number of processes; it is not built, and the constructor raises
for more than one time rank.
"""
from pymgrit_b200 import _lib
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
        super().__init__(conv_crit=conv_crit, *args, **kwargs)
        self._check_ranks()

    def _check_ranks(self):
        if self.comm_time_size != 1:
            raise Exception('pymgrit_b200.AtMgrit runs on one time rank (the process-local coarse grids of '
                            'at_mgrit.py:48-73 are not built); use Mgrit for time-parallel runs')

    def forward_solve(self, lvl: int) -> None:
        """Local coarse grid problems on the coarsest level (at_mgrit.py:37-88)."""
        self._check_ranks()
        if self.lvl_max == 1:
            return                                  # at_mgrit.py:45: nothing happens on a one-level hierarchy
        lv = self._lv[lvl]
        old = lv.u.clone()                          # tmp_u_arr, at_mgrit.py:76
        _lib.check(_lib.lib().mgb_local_coarse_solve(lv.ref, old.data_ptr(), int(self.k), self._stream()),
                   'local_coarse_solve')
        self._keep_old = old                        # alive until the launch has run

    def ouput_run_information(self) -> None:
        super().ouput_run_information()
        if self._log_lvl <= 20:
            self.log_info(message='  ' + '{0: <25}'.format('distance') + ' : ' + str(self.k))
