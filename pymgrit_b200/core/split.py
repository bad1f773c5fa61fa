"""Communicators for space and time parallelism (reference: core/split.py) over torch.distributed process groups."""
from typing import Tuple


def split_communicator(comm=None, splitting: int = 1) -> Tuple[object, object]:
    """Splits the processes of `comm` (a torch.distributed process group; None = the world) like the reference does with
    MPI_Comm_split (core/split.py:12-32): processes with the same rank // splitting share a space communicator, processes
    with the same rank % splitting a time communicator.  Returns (comm_x, comm_t) -- process groups; pass comm_t as
    Mgrit(comm_time=...).  The applications of this package are not parallel in space, so `splitting` is normally 1:
    comm_x then holds the calling process alone and comm_t every process of `comm`.

    Every process of `comm` must call this function (torch.distributed creates groups collectively)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        raise Exception('split_communicator needs an initialised torch.distributed process group')
    splitting = int(splitting)
    size = dist.get_world_size(comm)
    if splitting < 1 or size % splitting != 0:
        raise Exception('the splitting factor must divide the number of processes')
    ranks = dist.get_process_group_ranks(comm) if comm is not None else list(range(size))
    me = dist.get_rank(comm)
    comm_x = comm_t = None
    for color in range(size // splitting):                       # space groups: consecutive blocks of `splitting` ranks
        grp = dist.new_group([ranks[r] for r in range(size) if r // splitting == color])
        if me // splitting == color:
            comm_x = grp
    for color in range(splitting):                               # time groups: ranks with the same position in a block
        grp = dist.new_group([ranks[r] for r in range(size) if r % splitting == color])
        if me % splitting == color:
            comm_t = grp
    return comm_x, comm_t
