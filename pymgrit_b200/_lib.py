"""ctypes binding of libmgrit_b200.so (include/mgrit_b200.h).

The library is the product: there is no Python or CPU fallback for any sweep.  Importing this module
without the built library raises; calling a sweep without a CUDA device returns MGB_ECUDA, which
`check` turns into an Exception.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'lib', 'libmgrit_b200.so')

APP_BATCHED = 0            # no fused team kernels: the batched path (core/batched.py, csrc/generic.cu)
APP_HEAT1D, APP_ADVECTION1D, APP_DAHLQUIST, APP_BRUSSELATOR, APP_HEAT2D, APP_HEAT1D_2PTS, APP_HEAT1D_SINE = 1, 2, 3, 4, 5, 6, 7
TNORM_ONE, TNORM_TWO, TNORM_INF = 1, 2, 3
ABI_VERSION = 10
F_RELAX_LAST_ONLY = 1
CORRECT_F_RELAX, CORRECT_GHOST, CORRECT_LAST_ONLY = 1, 2, 4
DAHLQUIST_METHODS = {'BE': 0, 'FE': 1, 'TR': 2, 'MR': 3}

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class MgbLevel(C.Structure):
    """struct mgb_level (include/mgrit_b200.h)."""
    _fields_ = [
        ('app', C.c_int32), ('n', C.c_int32), ('pitch', C.c_int32), ('npts', C.c_int32),
        ('u_dev', C.c_void_p), ('g_dev', C.c_void_p), ('cpts_dev', C.c_void_p),
        ('ncpts', C.c_int32), ('team_threads', C.c_int32), ('chunk', C.c_int32),
        ('ndt', C.c_int32), ('cw', C.c_int32),
        ('dtidx_dev', C.c_void_p), ('sconst_dev', C.c_void_p),
        ('nrhs', C.c_int32), ('nsys', C.c_int32),
        ('rhs_x_dev', C.c_void_p), ('rhs_t_dev', C.c_void_p), ('rhs_dense_dev', C.c_void_p),
        ('t_dev', C.c_void_p),
        ('p', C.c_double * 8), ('ip', C.c_int32 * 4),
        ('sig_dev', C.c_void_p), ('diag_dev', C.c_void_p), ('nat_dev', C.c_void_p),
    ]


# every symbol include/mgrit_b200.h declares: name -> (restype, argtypes)
_LP = C.POINTER(MgbLevel)
SYMBOLS = {
    'mgb_abi_version': (C.c_int, []),
    'mgb_last_error': (C.c_char_p, []),
    'mgb_team_shape': (C.c_int, [C.c_int32, C.c_int32, c_int32_p, c_int32_p]),
    'mgb_step_consts_width': (C.c_int, [C.c_int32, C.c_int32, C.c_int32]),
    'mgb_heat1d_step_consts': (C.c_int, [C.c_double, C.c_int32, C.c_int32, C.c_int32, c_double_p]),
    'mgb_advection1d_step_consts': (C.c_int, [C.c_double, C.c_int32, C.c_int32, C.c_int32, c_double_p]),
    'mgb_heat1d_2pts_half_width': (C.c_int, [C.c_int32, C.c_int32]),
    'mgb_f_relax': (C.c_int, [_LP, C.c_int32, C.c_void_p]),
    'mgb_c_relax': (C.c_int, [_LP, C.c_double, C.c_void_p]),
    'mgb_c_relax_last': (C.c_int, [_LP, C.c_double, C.c_void_p]),
    'mgb_fas_residual': (C.c_int, [_LP, _LP, C.c_void_p]),
    'mgb_down_sweep': (C.c_int, [_LP, _LP, C.c_void_p]),
    'mgb_error_correction': (C.c_int, [_LP, _LP, C.c_int32, C.c_void_p]),
    'mgb_forward_solve': (C.c_int, [_LP, C.c_void_p]),
    'mgb_local_coarse_solve': (C.c_int, [_LP, C.c_void_p, C.c_int32, C.c_void_p]),
    'mgb_residual_norms': (C.c_int, [_LP, C.c_void_p, C.c_void_p]),
    'mgb_jump_norms': (C.c_int, [_LP, C.c_void_p, C.c_void_p, C.c_void_p]),
    'mgb_residual_rows': (C.c_int, [_LP, C.c_void_p, C.c_void_p]),
    'mgb_fas_coarse_rhs': (C.c_int, [_LP, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    'mgb_heat1d_restrict_rows': (C.c_int, [C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                           C.c_void_p]),
    'mgb_heat1d_interp_rows': (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                         C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    'mgb_temporal_norm': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    'mgb_set_stop_flag': (C.c_int, [C.c_void_p]),
    'mgb_write_flag': (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    'mgb_temporal_norm_flag': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                         C.c_void_p]),
    'mgb_convergence_flag': (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    'mgb_inject_up': (C.c_int, [_LP, _LP, C.c_void_p]),
    'mgb_step': (C.c_int, [_LP, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    'mgb_heat2d_layout': (C.c_int, [C.c_int32, C.c_int32, c_int32_p, c_int32_p, c_int32_p, c_int32_p]),
    'mgb_heat2d_to_rows': (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                     C.c_void_p, C.c_void_p]),
    'mgb_heat2d_from_rows': (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                       C.c_void_p, C.c_void_p]),
    'mgb_sine_matrix': (C.c_int, [C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    'mgb_rows_gemm': (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                C.c_void_p, C.c_int32, C.c_void_p]),
    'mgb_dst_twiddles': (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p]),
    'mgb_rows_dst': (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                               C.c_void_p]),
    'mgb_heat1d_spectral_recur': (C.c_int, [_LP, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    'mgb_sine_level_solve': (C.c_int, [_LP, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    'mgb_heat1d_spectral_fixup': (C.c_int, [_LP, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    'mgb_host_time_steps': (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, c_double_p, c_double_p, C.c_int32]),
    'mgb_host_affine_ramp': (C.c_int, [C.c_double, C.c_double, C.c_int64, C.c_void_p, C.c_int32]),
    'mgb_host_scale_rows': (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32]),
    'mgb_circ_fft_length': (C.c_int, [C.c_int32]),
    'mgb_circ_fft_tables': (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'mgb_rows_rfft': (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_int64, C.c_void_p]),
    'mgb_rows_irfft': (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_int64, C.c_void_p]),
    'mgb_advection1d_spectral_recur': (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_double, C.c_void_p, C.c_int64,
                                                 C.c_void_p]),
    'mgb_rows_lincomb': (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_double, C.c_void_p, C.c_int32,
                                   C.c_void_p, C.c_double, C.c_void_p, C.c_int32, C.c_void_p, C.c_double, C.c_void_p, C.c_int32,
                                   C.c_void_p, C.c_void_p]),
    'mgb_rows_sumsq': (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    'mgb_allen_cahn_imex_rows': (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                           C.c_void_p, C.c_void_p, C.c_double, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p]),
    'mgb_peer_status': (C.c_int, [c_int32_p, C.c_int32]),
    'mgb_peer_set_timeout': (C.c_int, [C.c_double]),
    'mgb_peer_put_row': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    'mgb_peer_wait_row': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    'mgb_peer_put_rows': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                    C.POINTER(C.c_uint64), C.c_uint64, C.c_void_p]),
    'mgb_peer_wait_flags': (C.c_int, [C.c_int32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_uint64, C.c_void_p]),
    'mgb_vec_axpby': (C.c_int, [C.c_int32, C.c_double, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    'mgb_vec_sumsq': (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib = None


def lib():
    """The loaded library (raises if it has not been built: python -m pymgrit_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f'{LIB_PATH} not found: build it with `python -m pymgrit_b200.build` '
                              '(there is no CPU fallback)')
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.mgb_abi_version() != ABI_VERSION:
            raise ImportError('libmgrit_b200.so has an unexpected ABI version')
        _lib = handle
    return _lib


def check(rc, what=''):
    """Turn a non-zero status of the C ABI into the bare Exception the reference raises."""
    if rc != 0:
        msg = lib().mgb_last_error().decode(errors='replace')
        raise Exception(f'libmgrit_b200 {what} failed (code {rc}): {msg}')


_raw_stream = None


def current_stream_ptr():
    """cudaStream_t of torch's current stream on the current device (as the void* the C ABI takes).  Through torch's raw
    accessor: `torch.cuda.current_stream().cuda_stream` builds a Python Stream object per call, ~15 us -- a millisecond per
    solve at 60 launches."""
    global _raw_stream
    import torch
    if _raw_stream is None:
        _raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', False)
    if _raw_stream:
        return C.c_void_p(_raw_stream(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
