"""Host-side logic of the time-parallel path on CPU: the aligned slab partition, and the communicator over gloo with
world_size 2 and 3 (no kernels involved)."""
import os
import socket
import sys

import numpy as np
import pytest

from pymgrit_b200.core import partition as P


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


@pytest.mark.parametrize('size', [2, 3])
def test_comm_over_gloo(size, tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from mp_comm_worker import worker
    mp.spawn(worker, args=(size, _free_port(), str(tmp_path)), nprocs=size, join=True)
    windows = np.load(tmp_path / 'windows_0.npy')
    assert windows[0][0] == 0 and windows[-1][1] == 256
    assert all(windows[r + 1][0] == windows[r][1] + 1 for r in range(size - 1))


@pytest.mark.parametrize('nts,size', [([257, 65, 17], 1), ([257, 65, 17], 2), ([257, 65, 17], 4), ([257, 65, 17], 8),
                                      ([129, 33, 9], 3), ([1025, 257, 65, 17], 5), ([48, 16, 6], 2)])
def test_aligned_partition_covers_every_point_once(nts, size):
    if nts == [48, 16, 6]:
        t0 = np.linspace(0, 2, 48)
        ts = [t0, t0[::3], t0[::9]]
    else:
        ts = [np.linspace(0, 2, n) for n in nts]
    parts = [P.Partition(ts, size, r) for r in range(size)]
    masks = P.c_point_masks(ts)
    for l, t in enumerate(ts):
        owned = np.concatenate([p.owned[l] for p in parts])
        np.testing.assert_array_equal(owned, np.arange(len(t)))            # every point exactly once, in order
        for r, p in enumerate(parts):
            tl = p.t_local[l]
            if r > 0:
                assert tl[0] == parts[r - 1].t_local[l][-1]                 # ghost = previous rank's last point
                assert masks[l][p.owned[l][0] - 1]                          # ... which is a C-point of this level
            assert p.sweep_cpts[l][0] == 0
            if l + 1 < len(ts):
                # local C-point table addresses exactly the owned C-points (plus the leading ghost / initial point)
                glob = (p.owned[l][0] - (1 if r > 0 else 0)) + p.sweep_cpts[l]
                assert np.all(masks[l][glob])
                assert len(p.sweep_cpts[l]) == len(parts[r].t_local[l + 1])  # one coarse point per local C-point
                assert masks[l][p.owned[l][-1]] or r == size - 1             # slabs end on C-points


def test_partition_matches_reference_split_for_power_of_two():
    ts = [np.linspace(0, 2, 2 ** 12 + 1), np.linspace(0, 2, 2 ** 10 + 1), np.linspace(0, 2, 2 ** 8 + 1)]
    for size in (2, 4, 8):
        split = P.split_into(len(ts[0]), size)
        ends = np.cumsum(split) - 1
        for r in range(size):
            assert P.Partition(ts, size, r).window[1] == ends[r]


def test_too_many_ranks_raises():
    ts = [np.linspace(0, 2, 17), np.linspace(0, 2, 5)]
    with pytest.raises(Exception):
        P.Partition(ts, 8, 0)


def test_reference_decomposition_tables():
    """core/partition.reference_decomposition against tables dumped from the reference (tests/core/test_mgrit.py:86-218
    checks the same tables for nt = 65/17/5 on 7 ranks)."""
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'decomposition.npz'))
    cfgs = {'nt65_17_5_p7': ([65, 17, 5], 7), 'nt65_17_5_p4': ([65, 17, 5], 4), 'nt129_33_9_p3': ([129, 33, 9], 3),
            'nt33_17_9_5_p5': ([33, 17, 9, 5], 5), 'nt48_16_6_p4': ([48, 16, 6], 4), 'nt65_17_5_p1': ([65, 17, 5], 1),
            'nt257_65_17_p8': ([257, 65, 17], 8)}
    for key, (nts, size) in cfgs.items():
        if key == 'nt48_16_6_p4':
            t0 = np.linspace(0, 2, 48)
            ts = [t0, t0[::3], t0[::9]]
        else:
            ts = [np.linspace(0, 2, n) for n in nts]
        for r in range(size):
            for l, d in enumerate(P.reference_decomposition(ts, size, r)):
                pre = f'{key}/r{r}/l{l}/'
                for nm in ('t', 'cpts', 'index_local', 'index_local_c', 'index_local_f'):
                    np.testing.assert_array_equal(np.asarray(d[nm], dtype=float), np.asarray(gold[pre + nm], dtype=float))
                flags = [d['comm_front'], d['comm_back'], d['first_is_c_point'], d['first_is_f_point'],
                         d['last_is_c_point'], d['last_is_f_point']]
                np.testing.assert_array_equal(np.array(flags, dtype=int), gold[pre + 'flags'])
                np.testing.assert_array_equal(np.array([d['send_to'], d['get_from']]), gold[pre + 'send_get'])


def test_split_helpers():          # tests/core/test_mgrit.py:33-56
    np.testing.assert_equal(np.array([4, 3, 3]), P.split_into(10, 3))
    assert tuple(int(v) for v in P.split_points(10, 3, 0)) == (4, 0)
    assert tuple(int(v) for v in P.split_points(10, 3, 1)) == (3, 4)
    assert tuple(int(v) for v in P.split_points(10, 3, 2)) == (3, 7)
