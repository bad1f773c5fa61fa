"""Time-parallel parity on 2 GPUs (NCCL): the same cases, the same fixtures, the same tolerances as on one GPU."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = ['heat1d_small_v', 'heat1d_small_f_cf2', 'heat1d_cfg2_nt1025', 'heat1d_small_jump', 'heat1d_small_tnorminf',
         'heat1d_small_weight', 'heat1d_trailing_f', 'dahlquist_cfg1', 'dahlquist_ml1', 'advection_example',
         'brusselator_example', 'heat1d_example', 'heat1d_bdf2_example', 'heat1d_bdf1_small', 'heat2d_cn_3lvl',
         'heat1d_spatial_example', 'heat1d_spatial_large', 'heat1d_atmgrit_k8', 'heat1d_atmgrit_k5_f']


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


@pytest.mark.parametrize('nranks', [2])
def test_cases_on_time_ranks(nranks):
    import torch
    if torch.cuda.device_count() < nranks:
        pytest.skip(f'needs {nranks} GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={nranks}',
           '--master-addr', '127.0.0.1', '--master-port', str(_free_port()), os.path.join(HERE, 'mp_gpu_case.py')] + CASES
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:]
    assert r.stdout.count('OK  ') == len(CASES)
