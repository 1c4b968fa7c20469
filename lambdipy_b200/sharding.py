"""Multi-GPU plumbing of the strip path: files are independent units, so ranks take disjoint
shards and never exchange payload; the only collective is an allgather of per-rank counters
(SURVEY.md 8e).  Works on any torch.distributed backend (nccl on the GPUs, gloo in CPU tests)."""
import numpy as np


def shard_indices(sizes, rank, world):
    """Size-sorted round-robin deal: indices (ascending) of the files rank `rank` owns.
    Balances bytes even with 128 MB outliers; deterministic; a partition for any world size."""
    sizes = np.asarray(sizes)
    order = np.argsort(-sizes.astype(np.int64), kind="stable")
    mine = np.sort(order[rank::world])
    return mine


def gather_counts(values, device=None):
    """Allgather a short vector of int64 counters (in_bytes, out_bytes, n_files, n_fallback ...).
    Returns an array [world, len(values)].  Without an initialised process group: [1, k]."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(v) for v in values], dtype=torch.int64, device=device)
    if not (dist.is_available() and dist.is_initialized()):
        return t.cpu().numpy().reshape(1, -1)
    out = torch.zeros(dist.get_world_size() * t.numel(), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, t)
    return out.cpu().numpy().reshape(dist.get_world_size(), -1)
