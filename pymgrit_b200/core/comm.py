"""Time communicator: the few collective operations the batched engine needs, over torch.distributed.

The reference exchanges pickled vectors with mpi4py isend/recv and eight message kinds (core/mgrit.py:693-713).  With the
aligned slab partition (core/partition.py) three exchanges remain:
  exchange_ghost  my last row of a level -> ghost row (point 0) of the next rank     (reference kinds 0 and 4)
  chain           coarsest-level forward solve: receive ghost, solve, send last row   (reference kind 5)
  reduce_norm     one double per rank, SUM (1-/2-norm) or MAX (inf-norm)              (gather + bcast, mgrit.py:428-432)
One process drives one GPU; rows move GPU to GPU through NCCL P2P (NVLink) without touching the host.  `SerialComm` is the
single-rank case (no torch.distributed needed).  The same class runs over gloo with CPU tensors (tests).
"""


class SerialComm:
    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def barrier(self):
        return None

    def allgather(self, value):
        return [value]

    def all_true(self, flags):
        return [bool(f) for f in flags]

    # engine hooks -----------------------------------------------------------------------------
    def exchange_ghost(self, solver, lvl):
        return None

    def recv_chain(self, solver, lvl, row=None):
        return None

    def send_chain(self, solver, lvl, row=None):
        return None

    def setup_peer_exchange(self, solver):
        return None

    def sum_rows(self, rows):
        return None

    def reduce_norm(self, partial, t_norm):
        return partial

    def check_peers(self):
        return None


class PeerMailbox:
    """Symmetric memory every time rank's neighbours can address directly (torch.distributed._symmetric_memory: the
    same allocation on every rank, peer-mapped over NVLink).  Layout in 8-byte words, identical on all ranks:
        rows  [levels][2][pitch]   the ghost row of a level in flight (two slots, sequence parity)
        flags [levels][2]          sequence number of the row in the slot, written by the predecessor
        acks  [levels]             last sequence number my successor has consumed, written by the successor
    include/mgrit_b200.h (mgb_peer_put_row / mgb_peer_wait_row) has the protocol."""

    def __init__(self, dist, group, levels, pitch, device):
        import torch
        import torch.distributed._symmetric_memory as symm_mem
        self.levels, self.pitch = levels, pitch
        self.size = dist.get_world_size(group)
        words = levels * 2 * pitch + levels * 2 + levels
        words += words & 1
        # gather region (the sine-space coarsest solve): gath [2][size][2][pitch], gflag [2][size], gack [size]
        self.gather = pitch <= 16384
        self.gbase = words
        if self.gather:
            words += 2 * self.size * 2 * pitch + 2 * self.size + self.size
            words += words & 1
        self.gseq = 0
        self.buf = symm_mem.empty((words,), dtype=torch.int64, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, group=dist.group.WORLD if group is None else group)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.rank = dist.get_rank(group)
        self.seq = [0] * levels
        torch.cuda.synchronize()
        dist.barrier(group=group)            # every rank's mailbox is zeroed before anybody stores into one

    def row(self, rank, lvl, slot):
        return self.ptrs[rank] + 8 * ((lvl * 2 + slot) * self.pitch)

    def flag(self, rank, lvl, slot):
        return self.ptrs[rank] + 8 * (self.levels * 2 * self.pitch + lvl * 2 + slot)

    def ack(self, rank, lvl):
        return self.ptrs[rank] + 8 * (self.levels * 2 * self.pitch + self.levels * 2 + lvl)

    # gather region of `rank`: rows written by `src`, the flag `src` sets, the acknowledgement `consumer` writes
    def gath(self, rank, slot, src):
        return self.ptrs[rank] + 8 * (self.gbase + ((slot * self.size + src) * 2) * self.pitch)

    def gflag(self, rank, slot, src):
        return self.ptrs[rank] + 8 * (self.gbase + 2 * self.size * 2 * self.pitch + slot * self.size + src)

    def gack(self, rank, consumer):
        return self.ptrs[rank] + 8 * (self.gbase + 2 * self.size * 2 * self.pitch + 2 * self.size + consumer)

    def share_rows(self, mine):
        """Every rank above me gets my `mine` ([2][pitch] doubles) in its gather buffer; returns the address of my
        gather buffer [size][2][pitch] once the rows of all ranks below me have arrived (device-side waits only)."""
        import ctypes as C
        from pymgrit_b200 import _lib
        self.gseq += 1
        seq, slot, me = self.gseq, self.gseq & 1, self.rank
        arr = lambda vals: (C.c_uint64 * max(len(vals), 1))(*vals)
        up = list(range(me + 1, self.size))
        down = list(range(me))
        stream = _lib.current_stream_ptr()
        _lib.check(_lib.lib().mgb_peer_put_rows(mine.data_ptr(), 2 * self.pitch, len(up),
                                                arr([self.gath(r, slot, me) for r in up]),
                                                arr([self.gflag(r, slot, me) for r in up]),
                                                arr([self.gack(me, r) for r in up]), seq, stream), 'peer_put_rows')
        _lib.check(_lib.lib().mgb_peer_wait_flags(len(down), arr([self.gflag(me, slot, r) for r in down]),
                                                  arr([self.gack(r, me) for r in down]), seq, stream), 'peer_wait_flags')
        return self.gath(me, slot, 0)


_MAILBOXES = {}


class TorchDistComm:
    """torch.distributed process group as the time communicator (NCCL on GPUs, gloo in CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size

    def barrier(self):
        self.dist.barrier(group=self.group)

    def allgather(self, value):
        out = [None] * self.size
        self.dist.all_gather_object(out, value, group=self.group)
        return out

    def all_true(self, flags):
        """Element-wise AND of a short list of booleans over the ranks: one all-reduce of a small integer tensor (the
        pickle-based all_gather_object costs a millisecond per call on NCCL)."""
        import torch
        dev = torch.device('cuda', torch.cuda.current_device()) if self.dist.get_backend(self.group) == 'nccl' else 'cpu'
        t = torch.tensor([1 if f else 0 for f in flags], dtype=torch.int32, device=dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)
        return [bool(v) for v in t.tolist()]

    def _global(self, r):
        return self.dist.get_global_rank(self.group, r) if self.group is not None else r

    def shift_rows(self, send_row, recv_row):
        """send_row -> next rank, previous rank -> recv_row (either may be None at the ends)."""
        ops = []
        if send_row is not None and self.rank + 1 < self.size:
            ops.append(self.dist.P2POp(self.dist.isend, send_row, self._global(self.rank + 1), group=self.group))
        if recv_row is not None and self.rank > 0:
            ops.append(self.dist.P2POp(self.dist.irecv, recv_row, self._global(self.rank - 1), group=self.group))
        if ops:
            for req in self.dist.batch_isend_irecv(ops):
                req.wait()

    def setup_peer_exchange(self, solver):
        """Use direct peer-memory stores for the ghost rows if every rank can: CUDA tensors over NCCL, symmetric memory
        available, every rank holds points on every level.  Decided collectively; otherwise NCCL send/recv stays."""
        import os
        self.mailbox = None
        ok = os.environ.get('MGB_PEER_EXCHANGE', '1') != '0' and self.dist.get_backend(self.group) == 'nccl'
        ok = ok and all(lv.npts > 1 and lv.u.is_cuda for lv in solver._lv)
        if ok:
            try:
                import torch.distributed._symmetric_memory  # noqa: F401
            except Exception:
                ok = False
        if not self.all_true([ok])[0]:
            return
        levels, pitch = len(solver._lv), max(lv.pitch for lv in solver._lv)
        # a stable key: the ranks of the group, not id(group) (which Python may hand to a later group object)
        ranks = tuple(self.dist.get_process_group_ranks(self.group)) if self.group is not None else ('world', self.size)
        key = (ranks, levels, pitch, solver._lv[0].u.device.index)
        hit = _MAILBOXES.get(key)
        if hit is None:                       # every rank creates (or fails to create) its mailbox at the same call
            try:
                box = PeerMailbox(self.dist, self.group, levels, pitch, solver._lv[0].u.device)
            except Exception:
                box = False
            hit = (box, self.all_true([bool(box)])[0])
            _MAILBOXES[key] = hit
        box, agreed = hit
        if agreed:
            self.mailbox = box

    def check_peers(self):
        """After a solve: did every device-side wait on a neighbour get its answer (csrc/peer.cu)?"""
        if not getattr(self, 'mailbox', None):
            return
        import ctypes as C
        from pymgrit_b200 import _lib
        err = C.c_int32(0)
        _lib.check(_lib.lib().mgb_peer_status(C.byref(err), 1), 'peer_status')
        if err.value:
            what = {1: 'a ghost row did not arrive', 2: 'the next rank never consumed a ghost row',
                    3: 'rows of the coarsest-level gather did not arrive', 4: 'a rank never consumed a gather'}
            raise Exception(f'time rank {self.rank}: {what.get(err.value, "a peer wait")} within the timeout -- another time '
                            'rank died or took a different code path; the results of this solve are invalid')

    def exchange_ghost(self, solver, lvl):
        lv = solver._lv[lvl]
        if lv.npts == 0:
            return
        box = getattr(self, 'mailbox', None)
        if not box:
            self.shift_rows(lv.u[lv.npts - 1], lv.u[0])
            return
        from pymgrit_b200 import _lib
        box.seq[lvl] += 1
        seq, slot = box.seq[lvl], box.seq[lvl] & 1
        stream = _lib.current_stream_ptr()
        if self.rank + 1 < self.size:
            _lib.check(_lib.lib().mgb_peer_put_row(lv.u[lv.npts - 1].data_ptr(), box.row(self.rank + 1, lvl, slot), lv.pitch,
                                                   box.flag(self.rank + 1, lvl, slot), box.ack(self.rank, lvl), seq, stream),
                       'peer_put_row')
        if self.rank > 0:
            _lib.check(_lib.lib().mgb_peer_wait_row(box.row(self.rank, lvl, slot), lv.u[0].data_ptr(), lv.pitch,
                                                    box.flag(self.rank, lvl, slot), box.ack(self.rank - 1, lvl), seq, stream),
                       'peer_wait_row')

    def recv_chain(self, solver, lvl, row=None):
        """row: where the incoming row goes (default: the ghost row u[0] of the level)."""
        lv = solver._lv[lvl]
        if self.rank > 0 and lv.npts > 0:
            self.dist.recv(lv.u[0] if row is None else row, self._global(self.rank - 1), group=self.group)

    def send_chain(self, solver, lvl, row=None):
        lv = solver._lv[lvl]
        if self.rank + 1 < self.size and lv.npts > 0:
            self.dist.send(lv.u[lv.npts - 1] if row is None else row, self._global(self.rank + 1), group=self.group)

    def all_gather_rows(self, out, mine):
        """out[r] = rank r's `mine` (device tensors; the sine-space coarsest solve, heat/heat_1d.py)."""
        self.dist.all_gather_into_tensor(out, mine, group=self.group)

    def sum_rows(self, rows):
        """In-place sum of a device array over the ranks (AT-MGRIT: every rank contributes its own rows, zeros elsewhere)."""
        self.dist.all_reduce(rows, op=self.dist.ReduceOp.SUM, group=self.group)

    def reduce_norm(self, partial, t_norm):
        op = self.dist.ReduceOp.MAX if t_norm == 3 else self.dist.ReduceOp.SUM
        self.dist.all_reduce(partial, op=op, group=self.group)
        return partial


def as_time_comm(comm):
    """None -> the torch.distributed world if one is initialised (the reference defaults to MPI.COMM_WORLD), else serial."""
    if comm is None:
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                return TorchDistComm(None)
        except ImportError:
            pass
        return SerialComm()
    if isinstance(comm, (SerialComm, TorchDistComm)):
        return comm
    if hasattr(comm, 'Get_rank') and hasattr(comm, 'Get_size'):
        if comm.Get_size() == 1:
            return SerialComm()
        raise Exception('pass a torch.distributed process group (or TorchDistComm) as comm_time for more than one rank')
    return TorchDistComm(comm)
