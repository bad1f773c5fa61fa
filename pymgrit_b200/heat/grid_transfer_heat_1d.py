"""Spatial coarsening for Heat1D on the device: full-weighting restriction and linear interpolation between grids of
2 n + 1 and n interior points (homogeneous Dirichlet boundaries).

The reference ships no such class; its users write it in Python (examples/example_spatial_coarsening.py:18-79,
docs/source/usage/advanced.rst:160-390) and pass it as `Mgrit(transfer=[...])`.  This is that transfer as kernels
(csrc/transfer.cu), same arithmetic in the same order, usable vector by vector like any GridTransfer and row-wise by the
solver (core/mgrit.py: fas_residual, error_correction, nested_iteration).
"""
import numpy as np

from pymgrit_b200 import _lib
from pymgrit_b200.core.grid_transfer import DeviceGridTransfer
from pymgrit_b200.heat.heat_1d import VectorHeat1D


def _ptr(t):
    return None if t is None else t.data_ptr()


class GridTransferHeat1D(DeviceGridTransfer):
    def __init__(self):
        super().__init__()

    # ---- GridTransfer API on single vectors (core/grid_transfer.py:31-55) ----------------------------------------
    def restriction(self, u: VectorHeat1D) -> VectorHeat1D:
        import torch
        src = u.device_values.contiguous().reshape(1, -1)
        n_fine = src.shape[1]
        if n_fine < 3 or n_fine % 2 == 0:
            raise Exception('full weighting needs 2 n + 1 interior points, got ' + str(n_fine))
        dst = torch.empty((1, (n_fine - 1) // 2), dtype=torch.float64, device=src.device)
        _lib.check(_lib.lib().mgb_heat1d_restrict_rows(1, src.data_ptr(), n_fine, None, n_fine, dst.data_ptr(),
                                                       dst.shape[1], _lib.current_stream_ptr()), 'restrict_rows')
        return VectorHeat1D(dst.shape[1], dst[0])

    def interpolation(self, u: VectorHeat1D) -> VectorHeat1D:
        import torch
        src = u.device_values.contiguous().reshape(1, -1)
        n_coarse = src.shape[1]
        dst = torch.empty((1, 2 * n_coarse + 1), dtype=torch.float64, device=src.device)
        _lib.check(_lib.lib().mgb_heat1d_interp_rows(1, 0, src.data_ptr(), None, n_coarse, n_coarse, dst.data_ptr(),
                                                     dst.shape[1], None, 0, _lib.current_stream_ptr()), 'interp_rows')
        return VectorHeat1D(dst.shape[1], dst[0])

    # ---- row-wise, for the solver -----------------------------------------------------------------------------------
    def check(self, fine_app, coarse_app) -> None:
        if fine_app.kind != _lib.APP_HEAT1D or coarse_app.kind != _lib.APP_HEAT1D:
            raise Exception('GridTransferHeat1D connects two Heat1D levels')
        if fine_app.ndof != 2 * coarse_app.ndof + 1:
            raise Exception('GridTransferHeat1D needs nx_fine - 1 = 2 (nx_coarse - 1); got ' + str(fine_app.ndof + 2) +
                            ' and ' + str(coarse_app.ndof + 2) + ' points')

    def restrict_rows(self, nrows, src, src_index, dst, fine_app) -> None:
        _lib.check(_lib.lib().mgb_heat1d_restrict_rows(int(nrows), src.data_ptr(), src.shape[1], _ptr(src_index),
                                                       int(fine_app.ndof), dst.data_ptr(), dst.shape[1],
                                                       _lib.current_stream_ptr()), 'restrict_rows')

    def interpolate_rows(self, nrows, first, a, b, dst, dst_index, accumulate, coarse_app) -> None:
        _lib.check(_lib.lib().mgb_heat1d_interp_rows(int(nrows), int(first), a.data_ptr(), _ptr(b), a.shape[1],
                                                     int(coarse_app.ndof), dst.data_ptr(), dst.shape[1], _ptr(dst_index),
                                                     1 if accumulate else 0, _lib.current_stream_ptr()), 'interp_rows')


GridTransferHeat = GridTransferHeat1D      # the name the reference's example gives its class
