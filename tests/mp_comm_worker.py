"""Worker for tests/test_multirank_cpu.py: exercises the time communicator over gloo with CPU tensors
(world_size 2+, launched with torch.multiprocessing).  A fake solver stands in for the engine: the communicator only
touches `solver._lv[lvl].u` and `.npts`."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class FakeLevel:
    def __init__(self, npts, pitch, fill):
        self.npts = npts
        self.u = torch.full((npts, pitch), float(fill), dtype=torch.float64)


class FakeSolver:
    def __init__(self, levels):
        self._lv = levels


def worker(rank, size, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=size)
    from pymgrit_b200.core.comm import TorchDistComm, as_time_comm
    from pymgrit_b200.core.partition import Partition
    comm = as_time_comm(None)
    assert isinstance(comm, TorchDistComm) and comm.Get_rank() == rank and comm.Get_size() == size
    # partition of a 3-level hierarchy: every rank builds its own view
    ts = [np.linspace(0, 2, 257), np.linspace(0, 2, 65), np.linspace(0, 2, 17)]
    part = Partition(ts, size, rank)
    # ghost exchange: rows carry (rank, level) so the receiver can check provenance
    levels = []
    for l in range(3):
        lv = FakeLevel(len(part.t_local[l]), 4, fill=-1.0)
        lv.u[lv.npts - 1] = 100.0 * rank + l
        levels.append(lv)
    solver = FakeSolver(levels)
    for l in range(3):
        comm.exchange_ghost(solver, l)
    for l in range(3):
        want = 100.0 * (rank - 1) + l if rank > 0 else -1.0
        assert torch.all(levels[l].u[0] == want), (rank, l, levels[l].u[0])
    # coarsest chain: running sum handed from rank to rank
    lv = levels[2]
    lv.u[0] = 0.0
    comm.recv_chain(solver, 2)
    lv.u[lv.npts - 1] = lv.u[0] + (rank + 1)
    comm.send_chain(solver, 2)
    assert torch.all(lv.u[lv.npts - 1] == sum(range(1, rank + 2)))
    # norm reduction
    part_sum = torch.tensor([float(rank + 1)], dtype=torch.float64)
    assert comm.reduce_norm(part_sum.clone(), 2).item() == size * (size + 1) / 2
    assert comm.reduce_norm(part_sum.clone(), 3).item() == size
    # path flags: element-wise AND over the ranks in one all-reduce
    assert comm.all_true([True, rank == 0, False, rank < size]) == [True, False, False, True]
    assert comm.all_true([True]) == [True]
    # no CUDA tensors over gloo: the peer-memory mailbox is not set up, NCCL/gloo send-recv stays (checked above)
    comm.setup_peer_exchange(solver)
    assert comm.mailbox is None
    # split_communicator (reference core/split.py): space groups of consecutive ranks, time groups across them
    from pymgrit_b200.core.split import split_communicator
    cx, ct = split_communicator(None, 1)
    assert dist.get_world_size(cx) == 1 and dist.get_world_size(ct) == size and dist.get_rank(ct) == rank
    assert as_time_comm(ct).Get_size() == size
    if size % 2 == 0:
        cx, ct = split_communicator(None, 2)
        assert dist.get_world_size(cx) == 2 and dist.get_rank(cx) == rank % 2
        assert dist.get_world_size(ct) == size // 2 and dist.get_rank(ct) == rank // 2
    gathered = comm.allgather(part.window)
    np.save(os.path.join(out_dir, f'windows_{rank}.npy'), np.array(gathered))
    comm.barrier()
    dist.destroy_process_group()
