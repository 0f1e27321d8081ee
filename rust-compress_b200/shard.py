"""Multi-GPU sharding of independent units (LZ4 / BWT blocks, DEFLATE / ARI streams) — SURVEY.md §8(e).

The reference has no parallelism at all; what makes the path shard is the formats: blocks and streams are independent, so
rank r of G owns a contiguous range of units, decodes them with its own `Context` (one process per GPU), and nothing is
exchanged while kernels run.  The only collective is the final gather of the decoded shards (`gather_shards`): NCCL
all-gather over NVLink when the tensors live on GPUs, gloo in the CPU tests.
"""
import numpy as np


def partition(weights, world):
    """Contiguous ranges [lo, hi) per rank, balanced by `weights` (e.g. compressed bytes per unit).

    Unit i goes to the rank whose share of the total weight contains the unit's midpoint — deterministic, identical on
    every rank, no communication.  Returns a list of (lo, hi) of length `world` covering 0..len(weights)."""
    w = np.asarray(weights, dtype=np.float64)
    n = len(w)
    if n == 0:
        return [(0, 0)] * world
    if w.sum() <= 0:
        w = np.ones(n)
    mid = np.cumsum(w) - w / 2.0
    owner = np.minimum((mid * world / w.sum()).astype(np.int64), world - 1)
    bounds = np.searchsorted(owner, np.arange(world + 1), side="left")
    return [(int(bounds[r]), int(bounds[r + 1])) for r in range(world)]


def my_range(weights, world, rank):
    return partition(weights, world)[rank]


def gather_shards(local, local_len, dist=None, group=None):
    """All-gather one variable-length byte shard per rank.  `local`: 1-D uint8 torch tensor (CPU for gloo, CUDA for NCCL);
    `local_len`: valid bytes in it.  Returns (list of per-rank uint8 tensors trimmed to their lengths)."""
    import torch
    import torch.distributed as tdist
    dist = dist or tdist
    world = dist.get_world_size(group)
    lens = torch.zeros(world, dtype=torch.int64, device=local.device)
    mine = torch.tensor([int(local_len)], dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(lens, mine, group=group) if hasattr(dist, "all_gather_into_tensor") and local.is_cuda else \
        dist.all_gather(list(lens.split(1)), mine, group=group)
    lens_h = [int(x) for x in lens.cpu().tolist()]
    cap = max(lens_h) if lens_h else 0
    pad = torch.zeros(cap, dtype=torch.uint8, device=local.device)
    pad[: int(local_len)] = local[: int(local_len)]
    if local.is_cuda and hasattr(dist, "all_gather_into_tensor"):
        flat = torch.empty(cap * world, dtype=torch.uint8, device=local.device)
        dist.all_gather_into_tensor(flat, pad, group=group)           # uniform shards: every NVLink port is used
        parts = list(flat.split(cap))
    else:
        parts = [torch.empty(cap, dtype=torch.uint8, device=local.device) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
    return [p[:n] for p, n in zip(parts, lens_h)]
