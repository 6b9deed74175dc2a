"""Multi-GPU plumbing for the block compressor: units (compressBlock calls) shard across ranks with no
data-path collective; the only exchange is the per-unit compressed size (4 B/unit) so that every rank
knows the global archive offsets.  torch.distributed is plumbing: NCCL on GPUs, gloo in
the CPU tests."""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous, balanced [lo, hi) slice of `total` units for `rank` (sizes differ by at most 1)."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_by_cost(costs, world):
    """Greedy longest-processing-time assignment of units with unequal predicted cost (bytes x method
    weight).  Returns a list of index arrays, one per rank; deterministic."""
    order = np.argsort(-np.asarray(costs, dtype=np.float64), kind="stable")
    load = np.zeros(world)
    owner = np.empty(len(order), dtype=np.int64)
    for i in order:
        r = int(np.argmin(load))
        owner[i] = r
        load[r] += costs[i]
    return [np.nonzero(owner == r)[0] for r in range(world)]


def exchange_sizes(local_sizes, total, device="cpu"):
    """All ranks contribute the compressed sizes of their contiguous shard; returns (all_sizes[total],
    global byte offsets[total]) on every rank.  One all_gather of 4 B/unit."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    base, rem = divmod(int(total), world)
    width = base + (1 if rem else 0)
    buf = torch.zeros(width, dtype=torch.int64, device=device)
    lo, hi = shard_range(total, rank, world)
    buf[: hi - lo] = torch.as_tensor(np.asarray(local_sizes, dtype=np.int64), device=device)
    if world > 1:
        out = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(out, buf)
    else:
        out = [buf]
    sizes = np.concatenate([out[r][: shard_range(total, r, world)[1] - shard_range(total, r, world)[0]].cpu().numpy()
                            for r in range(world)])
    offsets = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    return sizes, offsets


def max_over_ranks(value, device="cpu"):
    """Timing reduction used by bench.py: the slowest rank defines the step time."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
