"""Host-side block partitioning for multi-GPU runs: blocks share no state, so ranks /
devices take contiguous block-index ranges and there is no collective on the data path.  The only
cross-rank traffic is bookkeeping: per-shard sizes (to place shards in the global output) and timings."""
import numpy as np


def shard_range(n_blocks, world, rank):
    """Contiguous range [first, first+count) of rank `rank` of `world`; same rule as the C front end
    (redux_capi.cu make_shards): first = n*rank//world."""
    a = n_blocks * rank // world
    b = n_blocks * (rank + 1) // world
    return a, b - a


def weak_first_block(n_blocks_per_rank, rank):
    """Weak scaling (bench.py): every rank codes its own batch of distinct synthetic blocks."""
    return rank * n_blocks_per_rank


def global_offsets(shard_sizes):
    """shard_sizes: list (one per rank, in rank order) of per-block compressed sizes.
    Returns (offsets uint64[n+1], shard_bases uint64[world]) of the back-to-back global output."""
    sizes = np.concatenate([np.asarray(s, dtype=np.uint64) for s in shard_sizes]) if shard_sizes else np.zeros(0, np.uint64)
    off = np.zeros(sizes.size + 1, dtype=np.uint64)
    np.cumsum(sizes, out=off[1:])
    bases, pos = [], 0
    for s in shard_sizes:
        bases.append(int(off[pos]))
        pos += len(s)
    return off, np.asarray(bases, dtype=np.uint64)


def max_over_ranks(values, dist=None, device=None):
    """Element-wise max of a list of floats over all ranks (timings are reported as the slowest rank's)."""
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.cpu()]


def gather_sizes(local_sizes, dist=None):
    """All-gather of variable-length per-block size arrays (bookkeeping only). Returns list per rank."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [np.asarray(local_sizes, dtype=np.uint64)]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, np.asarray(local_sizes, dtype=np.uint64))
    return out
