"""Multi-GPU plumbing of the path: structures and trajectory frames are independent units, so ranks only need to agree on who
takes what; the one exchange step of the whole design is the sum of mdpocket's two grids (NCCL all-reduce on the GPU box, any
torch.distributed backend in tests). Mirrors the loops of reference src/fpmain.c:61-76 (-F list) and src/mdpocket.c:235-271."""
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


def broadcast_grid_geometry(t, dist, src=0):
    """origin[3] (and dims) decided by rank 0 from frame 0 -> every rank (24 bytes)."""
    dist.broadcast(t, src=src)
    return t


def sum_grids(grids, dist):
    """density and frequency grids: one all-reduce(sum) each, at the end of the run (before normalize_grid, mdpbase.c:354)."""
    for g in grids:
        dist.all_reduce(g)
    return grids
