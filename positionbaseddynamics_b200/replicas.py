"""Multi-GPU = independent scene replicas, one engine per rank ("replicas only").

A scene's coloured Gauss-Seidel sweep is a chain of globally ordered colour phases with no spatial decomposition in the
reference, so the path does not shard inside a scene; it shards across scene instances.  No data-path collective exists:
torch.distributed is used only to agree on the start (barrier) and to gather per-rank timings / checksums.
Works with the NCCL backend (one B200 per rank) and with gloo (CPU tests, world_size 2).
"""
import os


def env_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend, device_index=None):
    """Returns the torch.distributed module when WORLD_SIZE > 1, else None."""
    rank, world, local = env_world()
    if world <= 1:
        return None
    import torch
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    kw = {}
    if backend == "nccl" and device_index is not None:
        kw["device_id"] = torch.device("cuda", device_index)
    dist.init_process_group(backend, **kw)
    return dist


def gather(dist, value, device="cpu"):
    """All-gather one float per rank (timings, checksums)."""
    if dist is None:
        return [float(value)]
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]


def whole_job_throughput(units_per_rank, seconds_per_rank):
    """value = units all ranks processed / the slowest rank's time (weak scaling over replicas)."""
    return len(seconds_per_rank) * units_per_rank / max(seconds_per_rank)


def checksum(x):
    """Order-sensitive fp64 checksum of a position array: replicas of the same scene must agree bit for bit."""
    import numpy as np
    a = np.ascontiguousarray(x, dtype=np.float64).ravel()
    w = np.arange(1, a.size + 1, dtype=np.float64)
    return float((a * w).sum())
