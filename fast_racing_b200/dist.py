"""Multi-GPU plumbing: the batch of independent candidate trajectories shards across ranks with no
data-path collective; the only exchange is the final arg-min gather of one (cost, global index) pair
per rank (16 B/rank — latency-bound, so plain NCCL all-gather over NVLink)."""
import numpy as np


def shard_range(total, rank, world):
    """Block partition: rank r owns trajectories [lo, hi)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_argmin(local_cost, local_global_index, device=None):
    """All-gather (cost, index) from every rank; deterministic arg-min (ties -> lowest global index).
    Returns the same (cost, global_index) on every rank.  Works with NCCL (device tensors) and gloo."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(local_cost), int(local_global_index)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    mine = torch.tensor([float(local_cost), float(local_global_index)], dtype=torch.float64, device=device)
    allv = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(allv, mine)
    arr = torch.stack(allv).cpu().numpy()
    j = int(np.lexsort((arr[:, 1], arr[:, 0]))[0])
    return float(arr[j, 0]), int(arr[j, 1])
