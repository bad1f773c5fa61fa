"""Run test cases on N time ranks (one GPU each) and check them against the reference fixtures.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/mp_gpu_case.py case [case ...]

Every rank solves its slab; the level-0 solution is gathered on rank 0 and compared with tests/golden/<case>.npz with the
same tolerances as the single-GPU parity tests (results must not depend on the number of ranks)."""
import logging
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(HERE), HERE]

import cases as C                                    # noqa: E402
from b200_util import b200_problem, b200_transfer    # noqa: E402
from oracle_util import load_golden, assert_history_close, assert_solution_close   # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import pymgrit_b200 as P
    failed = []
    for name in sys.argv[1:]:
        case = C.CASES[name]
        if 'at_k' in case:
            solver = P.AtMgrit(problem=b200_problem(case), k=case['at_k'], logging_lvl=logging.WARNING, **case['solver'])
        else:
            solver = P.Mgrit(problem=b200_problem(case), transfer=b200_transfer(case), logging_lvl=logging.WARNING,
                             **case['solver'])
        info = solver.solve()
        lv = solver._lv[0]
        own = lv.values(idx=np.arange(1 if rank > 0 else 0, lv.npts))                       # drop the ghost row
        own = own.reshape(len(own), -1)
        parts = [None] * world
        dist.all_gather_object(parts, own)
        if rank == 0:
            u = np.concatenate(parts)
            gold = load_golden(name)
            try:
                assert u.shape[0] == len(gold['u_norms']), (u.shape, len(gold['u_norms']))
                norms = np.sqrt(np.sum(u * u, axis=1))
                rows = u[gold['u_rows_idx']].reshape(gold['u_rows'].shape)
                scale = np.max(gold['u_norms']) * np.sqrt(len(gold['u_norms']))
                assert_history_close(info['conv'], gold['conv'], scale=scale)
                assert_solution_close(rows, norms, gold)
                print(f'OK   {name} on {world} ranks: {len(info["conv"])} iterations', flush=True)
            except AssertionError as e:
                failed.append(name)
                print(f'FAIL {name} on {world} ranks: {str(e)[:400]}', flush=True)
        del solver
    flag = torch.tensor([len(failed)], device='cuda')
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(1 if flag.item() else 0)


if __name__ == '__main__':
    main()
