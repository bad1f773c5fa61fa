"""Probe: does torch's symmetric memory give peer pointers between the ranks of one box?
    python -m torch.distributed.run --nproc-per-node 2 scripts/probe/symm_probe.py
"""
import os
import torch
import torch.distributed as dist

rank, local = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty((4, 1024), dtype=torch.float64, device=torch.device('cuda', local))
    hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
    print(rank, 'rendezvous ok', type(hdl).__name__, 'buffer_ptrs', [hex(p) for p in hdl.buffer_ptrs],
          'signal_pad_ptrs', [hex(p) for p in hdl.signal_pad_ptrs], 'signal pad size', getattr(hdl, 'signal_pad_size', None),
          flush=True)
    t.fill_(rank + 1)
    dist.barrier()
    torch.cuda.synchronize()
    peer = hdl.get_buffer((rank + 1) % dist.get_world_size(), (4, 1024), torch.float64)
    print(rank, 'peer value', float(peer[0, 0]), flush=True)
    peer[1].fill_(10 * (rank + 1))          # store into the peer's memory
    torch.cuda.synchronize()
    dist.barrier()
    print(rank, 'my row 1 after peer store', float(t[1, 0]), flush=True)
except Exception as e:
    import traceback
    traceback.print_exc()
    print(rank, 'symmetric memory unavailable:', repr(e)[:300], flush=True)
print(rank, 'can_device_access_peer', torch.cuda.can_device_access_peer(local, (local + 1) % torch.cuda.device_count()))
dist.destroy_process_group()
