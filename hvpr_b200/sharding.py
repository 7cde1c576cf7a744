"""Frame-wise sharding across GPUs: frames are independent, so frame i goes to rank i mod W, weights are
replicated and NO collective sits on the data path.  torch.distributed (NCCL on GPUs, gloo in CPU tests) is used only
to agree on timings / checksums after the timed region."""
from __future__ import annotations

import os


def dist_env():
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when launched plainly."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)))


def frames_of_rank(n_frames: int, rank: int, world: int):
    """Round-robin partition: frame i -> rank i mod W (every frame owned by exactly one rank)."""
    return list(range(rank, n_frames, world))


def weak_scaling_frames(frames_per_rank: int, rank: int):
    """Weak-scaling bench: every rank owns `frames_per_rank` distinct frames (distinct RNG seeds)."""
    return list(range(rank * frames_per_rank, (rank + 1) * frames_per_rank))


def init_process_group(backend: str, device_id=None):
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {"device_id": device_id} if (device_id is not None and backend == "nccl") else {}
        dist.init_process_group(backend=backend, **kw)
    return dist


def max_over_ranks(value: float, device=None) -> float:
    """MAX-reduce a scalar over all ranks (elapsed time of the slowest rank)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_scalars(value: float, device=None):
    """all_gather one scalar per rank -> list (rank order)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(value)]
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [float(x.item()) for x in out]
