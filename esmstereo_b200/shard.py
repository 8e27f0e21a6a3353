"""Multi-GPU sharding of stereo pairs: every pair is independent, so pair i goes to
rank i mod N with replicated weights and NO data-path collective; the only communication is the
gather of the per-rank disparities (and metric sums) -- NCCL over NVLink on GPUs, gloo in CPU tests.
The reference has nothing to port here: it only wraps the model in a single-process
`nn.DataParallel` pinned to one visible GPU (test_kitti.py:18,53).
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def shard_indices(n_pairs: int, rank: int, world: int) -> List[int]:
    """Indices of the pairs rank `rank` processes (round robin)."""
    return list(range(rank, n_pairs, world))


def gather_disparities(local: torch.Tensor, n_pairs: int, rank: int, world: int, flat: bool = True) -> torch.Tensor:
    """All-gather per-rank outputs [n_local, H, W] into dataset order.
    Ranks may hold different counts (n_pairs % world != 0): shorter ranks are padded for the
    collective and the padding is dropped afterwards.

    flat=True returns [n_pairs, H, W] (pair i at index i; one extra copy of the gathered block).  flat=False returns the
    gathered block itself as a VIEW [per_rank, world, H, W] with pair j * world + r at [j, r] -- dataset order without
    touching the data again (what a sweep loop wants: bench.py, a metric reduction, a writer that walks the pairs)."""
    if world == 1:
        return local if flat else local.unsqueeze(1)
    per_rank = (n_pairs + world - 1) // world
    shape = (per_rank,) + tuple(local.shape[1:])
    if local.shape[0] == per_rank and local.is_contiguous():
        padded = local  # the usual case: no padding, no copy
    else:
        padded = torch.zeros(shape, dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
    out = torch.empty((world,) + shape, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(world * per_rank, *shape[1:]), padded)
    # pair i = j * world + r sits at out[r, j]: a transpose puts the pairs in dataset order and the padding of the
    # shorter ranks (i >= n_pairs) at the tail -- no index tensors, no host synchronisation
    ordered = out.transpose(0, 1)
    if not flat:
        return ordered
    return ordered.reshape((world * per_rank,) + tuple(local.shape[1:]))[:n_pairs]
