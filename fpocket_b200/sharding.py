"""Who takes what when one process per GPU is used (torchrun; bench.py): structures and trajectory frames are independent units, so the
ranks only need to agree on the assignment, without communication. The one exchange step of the design - the sum of mdpocket's two
grids - is a plain all-reduce on fpk_md_device_grids() done by the caller; inside ONE process the library shards over its device list
itself (fpk_init, csrc/capi.cu). Mirrors the loops of reference src/fpmain.c:61-76 (-F list) and src/mdpocket.c:235-271."""
import numpy as np


def shard_structures(n_sites, rank, world):
    """Indices of the structures rank `rank` processes: longest-first greedy assignment by heavy-atom count (the cost of a
    structure is ~linear in it), deterministic on every rank without communication."""
    n_sites = np.asarray(n_sites)
    order = np.argsort(-n_sites, kind="stable")
    load = np.zeros(world, np.int64)
    owner = np.empty(len(n_sites), np.int32)
    for i in order:
        r = int(np.argmin(load))
        owner[i] = r
        load[r] += int(n_sites[i])
    return np.nonzero(owner == rank)[0]


def shard_frames(n_frames, rank, world):
    """Contiguous block [lo, hi) of trajectory frames; frame 0 (which fixes the grid, mdpbase.c:444-502) is on rank 0."""
    per = (n_frames + world - 1) // world
    lo = min(n_frames, rank * per)
    return lo, min(n_frames, lo + per)
