"""Time-parallel decomposition: which time points of every level a time rank owns.

Two things live here.

`reference_decomposition` restates, as index arithmetic on boolean masks, the tables the reference builds in
Mgrit.setup_points_and_comm_info (core/mgrit.py:742-827, pinned by tests/core/test_mgrit.py:86-218): block split of the
level-0 points (split_into, mgrit.py:829-838), one ghost point on ranks > 0, local C/F index lists (F-points grouped by
interval, last interval first), the eight communication flags and send_to / get_from.

`Partition` is what the batched engine uses.  It differs from the reference's split in one respect: slab boundaries are
moved onto points of the coarsest grid, so that on every level the slab of a rank starts right after a C-point (its ghost)
and ends on a C-point.  Then every rank sees an ordinary serial problem whose "initial condition" is the ghost row, no
coarse interval is ever split across ranks (reference message kinds 1, 2, 3 and 7 disappear, mgrit.py:316-331, 347-352,
398-403, 503-508) and only the ghost exchange (kind 0/4) and the coarsest chain (kind 5) remain.  For nt = 2^k + 1 and a
power-of-two number of ranks the two partitions coincide.  Results do not depend on the partition.
"""
import numpy as np


def split_into(number_points: int, number_processes: int) -> np.ndarray:
    """Block sizes: the first (points % processes) ranks get one extra point (mgrit.py:829-838)."""
    base, extra = divmod(int(number_points), int(number_processes))
    return np.array([base + 1] * extra + [base] * (number_processes - extra))


def split_points(length: int, size: int, rank: int):
    """(block size, index of first point) of a rank (mgrit.py:728-740)."""
    split = split_into(length, size)
    return split[rank], (np.sum(split[:rank]) if split[rank] > 0 else 0)


class _Mask(np.ndarray):
    """Boolean C-point mask that remembers when it is regular (every `stride`-th point from 0): index lists then come
    from arithmetic instead of a scan over up to 2^20 flags."""
    stride = 0


def _strided(mask, stride):
    out = mask.view(_Mask)
    out.stride = int(stride)
    return out


def _increasing(t):
    """t[i] < t[i+1] for all i (in pieces on the table threads for long grids)."""
    if len(t) < 2:
        return True
    from pymgrit_b200.core.application import known_increasing
    if known_increasing(t):
        return True
    from pymgrit_b200.core.device_level import parallel_pieces
    return all(parallel_pieces(len(t) - 1, lambda a, b: bool(np.all(t[a + 1:b + 1] > t[a:b]))))


def c_point_masks(global_t):
    """is_c[l][i]: point i of level l is also a point of level l+1 (mgrit.py:212, 768-770); all True on the coarsest."""
    masks = []
    for l, t in enumerate(global_t):
        if l + 1 == len(global_t):
            masks.append(np.ones(len(t), dtype=bool))
            continue
        tc = global_t[l + 1]
        mask = np.zeros(len(t), dtype=bool)
        stride = (len(t) - 1) // (len(tc) - 1) if len(tc) > 1 and (len(t) - 1) % (len(tc) - 1) == 0 else 0
        increasing = _increasing(t)
        if increasing and stride > 0 and np.array_equal(t[::stride], tc):
            mask[::stride] = True                        # the usual case t_coarse = t[::m]: no search needed
            mask = _strided(mask, stride)
        elif increasing:                       # sorted grid: binary search instead of np.isin's sort
            idx = np.minimum(np.searchsorted(t, tc), len(t) - 1)
            mask[idx[t[idx] == tc]] = True
        else:
            mask = np.isin(t, tc)
        masks.append(mask)
    return masks


def coarsening_factors(global_t, masks):
    """m[l] = distance between the first two C-points (mgrit.py:213-214); 1 on the coarsest level."""
    out = []
    for l in range(len(global_t)):
        if l + 1 < len(global_t):
            mask = np.asarray(masks[l])
            first = int(np.argmax(mask))                 # the first two C-points; argmax stops at the first True
            rest = mask[first + 1:]
            out.append(1 + int(np.argmax(rest)) if mask[first] and rest.any() else 1)
        else:
            out.append(1)
    return out


_SET_ORDER_LIMIT = 1 << 16


def _f_groups_reversed(fpts):
    """F-points grouped into runs of consecutive indices, runs in reverse order (mgrit.py:774-776)."""
    if len(fpts) == 0:
        return np.array([], dtype=float)
    run_id = np.zeros(len(fpts), dtype=np.int64)
    run_id[1:] = np.cumsum(np.diff(fpts) != 1)
    return fpts[np.argsort(-run_id, kind='stable')]


def _level_tables(global_t, masks, lvl, size, rank, window, lazy_f=False):
    """Tables of one level for the owner of the level-0 index window [first, last] (inclusive)."""
    t = global_t[lvl]
    n = len(t)
    is_c = masks[lvl]
    t0 = global_t[0]
    first0, last0 = window
    if first0 > last0:
        own = np.array([], dtype=int)
    elif lvl == 0:
        own = np.arange(first0, last0 + 1)
    else:
        own = np.flatnonzero((t >= t0[first0]) & (t <= t0[last0]))
    cpts = own[is_c[own]] if len(own) else own
    fpts = own[~is_c[own]] if len(own) else own
    if 0 < len(own) <= _SET_ORDER_LIMIT:
        # The reference builds the F-point list from a Python set difference (mgrit.py:773), so the order in which
        # it visits the F-intervals is CPython's set iteration order.  Results do not depend on it, but
        # index_local_f is part of the de-facto API (tests/core/test_mgrit.py:171-177): reproduce it when cheap.
        fpts = np.array(list(set(own) - set(cpts)), dtype=own.dtype)
    ghost = rank != 0 and len(own) > 0
    with_ghost = np.concatenate([[own[0] - 1], own]) if ghost else own
    off = 1 if ghost else 0
    def f_list():
        return (_f_groups_reversed(fpts) - own[0] + off) if len(fpts) else np.array([], dtype=float)

    pos = {'index_local': np.arange(len(own)) + off,
           'index_local_c': (cpts - own[0] + off) if len(own) else cpts,
           'index_local_f': f_list if lazy_f else f_list()}

    def is_f(i):
        return 0 <= i < n and not is_c[i]

    def is_cc(i):
        return 0 <= i < n and bool(is_c[i])

    has = len(own) > 0
    flags = dict(
        comm_front=bool(len(fpts) > 0 and is_f(fpts.min() - 1)),
        comm_back=bool(len(fpts) > 0 and is_f(fpts.max() + 1)),
        first_is_c_point=bool(has and is_c[own[0]] and own[0] != 0 and is_f(own[0] - 1)),
        first_is_f_point=bool(has and not is_c[own[0]] and is_cc(own[0] - 1)),
        last_is_c_point=bool(has and is_c[own[-1]] and own[-1] != n - 1 and is_f(own[-1] + 1)),
        last_is_f_point=bool(has and not is_c[own[-1]] and own[-1] != n - 1 and is_cc(own[-1] + 1)),
    )
    return own, with_ghost, cpts, pos, flags


def reference_decomposition(global_t, size, rank):
    """Per-level dicts with the reference's tables for (size, rank)."""
    masks = c_point_masks(global_t)
    split = split_into(len(global_t[0]), size)
    block, first = split[rank], int(np.sum(split[:rank])) if split[rank] > 0 else 0
    window = (first, first + block - 1)
    split_t = global_t[0][np.cumsum(split) - 1]
    out = []
    for lvl, t in enumerate(global_t):
        own, with_ghost, cpts, pos, flags = _level_tables(global_t, masks, lvl, size, rank, window)
        t_local = t[with_ghost] if len(with_ghost) else np.array([])
        send_to = get_from = -99
        if len(with_ghost) > 0:
            if t_local[-1] != t[-1]:
                send_to = int(np.searchsorted(split_t, t[with_ghost[-1] + 1]))
            if len(with_ghost) > len(own) or t_local[0] != global_t[0][0]:
                get_from = int(np.searchsorted(split_t, t_local[0]))
        d = dict(t=t_local, cpts=cpts, send_to=send_to, get_from=get_from, **pos, **flags)
        out.append(d)
    return out


class _LazyArrays:
    """A list of per-level index arrays that are built on first access (they are O(nt) and the sweeps never read them;
    user code and tests do: mgrit.index_local[lvl])."""

    def __init__(self):
        self._make, self._val = [], []

    def append(self, make):
        self._make.append(make)
        self._val.append(None)

    def __len__(self):
        return len(self._make)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        if self._val[i] is None:
            self._val[i] = self._make[i]()
        return self._val[i]

    def __iter__(self):
        return (self[i] for i in range(len(self)))


def _slab_tables(global_t, masks, lvl, rank, window):
    """What Partition needs of _level_tables, with slices instead of index arrays: a rank's points of a level are a
    contiguous range [a, b], so the ghost, the local time grid and the C-points come from views (at nt = 2^20 the index
    arithmetic of _level_tables costs 10 ms of host time per construction)."""
    t = global_t[lvl]
    is_c = masks[lvl]
    t0 = global_t[0]
    first0, last0 = window
    if first0 > last0:
        a, b = 0, -1
    elif lvl == 0:
        a, b = first0, last0
    else:
        a = int(np.searchsorted(t, t0[first0], side='left'))
        b = int(np.searchsorted(t, t0[last0], side='right')) - 1
    n_own = max(b - a + 1, 0)
    ghost = rank != 0 and n_own > 0
    off = 1 if ghost else 0
    own = lambda: np.arange(a, b + 1)
    sub = np.asarray(is_c[a:b + 1])
    stride = getattr(is_c, 'stride', 0)
    if stride > 0:                                   # regular coarsening: every stride-th point from 0
        local_c = np.arange((-a) % stride, n_own, stride)
    else:
        local_c = np.flatnonzero(sub)                # positions inside the owned range
    t_local = t[a - off:b + 1] if n_own else t[0:0]

    def f_list():
        fpts = a + np.flatnonzero(~sub)
        if 0 < n_own <= _SET_ORDER_LIMIT:            # the reference's set-iteration order, see _level_tables
            fpts = np.array(list(set(own()) - set(a + local_c)), dtype=np.int64)
        return (_f_groups_reversed(fpts) - a + off) if len(fpts) else np.array([], dtype=float)

    pos = {'index_local': lambda: np.arange(off, n_own + off), 'index_local_c': local_c + off, 'index_local_f': f_list}
    return own, t_local, a + local_c, pos


class Partition:
    """Aligned slab partition used by the engine (see module docstring)."""

    def __init__(self, global_t, size, rank, masks=None):
        self.size, self.rank = size, rank
        L = len(global_t)
        masks = c_point_masks(global_t) if masks is None else masks
        self.m = coarsening_factors(global_t, masks)
        t0 = global_t[0]
        n0 = len(t0)
        if size == 1:
            bounds = np.array([n0 - 1])
        else:
            # slab ends must be points of the coarsest grid: spread its intervals over the ranks
            # (level-0 indices of the coarsest points by composing the C-point masks: level l+1 is the C-points of
            # level l in order, so no search or sort over the 2^20 fine points is needed)
            strides = [getattr(masks[l], 'stride', 0) for l in range(L - 1)]
            if L > 1 and all(st > 0 for st in strides):       # regular coarsening on every level: plain arithmetic
                coarse_idx = np.arange(0, n0, int(np.prod(strides)))
            else:
                coarse_idx = np.flatnonzero(masks[0]) if L > 1 else np.arange(n0)
                for l in range(1, L - 1):
                    coarse_idx = coarse_idx[masks[l]]
            if len(coarse_idx) != len(global_t[-1]) or not np.array_equal(t0[coarse_idx], global_t[-1]):
                coarse_idx = np.flatnonzero(np.isin(t0, global_t[-1]))      # grids with repeated or unsorted points
            nc = len(coarse_idx)
            if nc - 1 < size:
                raise Exception(f'{size} time ranks need at least {size} intervals on the coarsest grid '
                                f'(it has {nc - 1}); use fewer ranks or a finer coarsest level')
            cuts = np.cumsum(split_into(nc - 1, size))            # coarsest-interval index at which each slab ends
            bounds = coarse_idx[cuts]
            bounds[-1] = n0 - 1                                    # trailing points after the last coarse point
        self.bounds = bounds
        first = 0 if rank == 0 else int(bounds[rank - 1]) + 1
        last = int(bounds[rank])
        self.window = (first, last)
        self.int_start, self.int_stop = t0[first], t0[last]
        self.t_local, self.cpts, self.index_local, self.index_local_c, self._f_lists = [], [], _LazyArrays(), [], []
        self.masks = masks
        self.sweep_cpts, self.send_to, self.get_from, self.owned = [], [], [], _LazyArrays()
        for lvl in range(L):
            own, t_local, cpts, pos = _slab_tables(global_t, masks, lvl, rank, (first, last))
            self.owned.append(own)
            self.t_local.append(t_local)
            self.cpts.append(cpts)
            self.index_local.append(pos['index_local'])
            self.index_local_c.append(pos['index_local_c'])
            self._f_lists.append(pos['index_local_f'])
            # C-points as the sweeps want them: local indices, starting with point 0 (initial condition or ghost)
            lc = np.asarray(pos['index_local_c'], dtype=np.int64)
            if rank != 0:
                lc = np.concatenate([[0], lc])
            self.sweep_cpts.append(lc.astype(np.int32))
            self.send_to.append(rank + 1 if rank + 1 < size else -99)
            self.get_from.append(rank - 1 if rank > 0 else -99)

    @property
    def index_local_f(self):
        """Local F-point lists in the reference's visiting order; built on first use (the sweeps do not need them)."""
        self._f_lists = [f() if callable(f) else f for f in self._f_lists]
        return self._f_lists
