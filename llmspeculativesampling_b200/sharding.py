"""Request sharding across the GPUs of one box: one process per GPU, full model replicas, no
collective in the data path (requests are independent — SURVEY.md §8e).  The only cross-rank
traffic is the end-of-run statistics reduction."""
from __future__ import annotations

from typing import List, Sequence


def shard_requests(n_requests: int, world_size: int, rank: int) -> List[int]:
    """Global request ids served by `rank` (round-robin, as in `r -> GPU r mod G`)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    return list(range(rank, n_requests, world_size))


def batches(ids: Sequence[int], batch: int) -> List[List[int]]:
    return [list(ids[i:i + batch]) for i in range(0, len(ids), batch)]


def reduce_stats(local: dict, device=None) -> dict:
    """Sum the numeric entries of `local` over all ranks and take the max of keys ending in '_max'.
    Works with both the nccl (GPU) and gloo (CPU tests) backends; a no-op without torch.distributed."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(local)
    keys = sorted(local)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    sums = torch.tensor([float(local[k]) for k in keys if not k.endswith("_max")], dtype=torch.float64, device=device)
    maxs = torch.tensor([float(local[k]) for k in keys if k.endswith("_max")], dtype=torch.float64, device=device)
    if sums.numel():
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    if maxs.numel():
        dist.all_reduce(maxs, op=dist.ReduceOp.MAX)
    out, si, mi = {}, 0, 0
    for k in keys:
        if k.endswith("_max"):
            out[k] = float(maxs[mi]); mi += 1
        else:
            out[k] = float(sums[si]); si += 1
    return out
