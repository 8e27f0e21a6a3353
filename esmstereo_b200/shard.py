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


class PeerGather:
    """The same gather without NCCL kernels: every rank keeps its block of outputs in SYMMETRIC memory (peer-mapped over
    NVLink: torch.distributed._symmetric_memory) and pulls the blocks of its peers with plain device copies, which the
    copy engines execute -- no SM is taken from the forward that keeps running on the compute stream (an NCCL all-gather of
    8 x 15 MB holds its channels' SMs for ~350 us per sweep; the persistent one-CTA-per-SM conv kernels then wait for them:
    +30 us per step at 8 GPUs).  Two stream-ordered barriers bracket the pulls: blocks complete before anyone reads,
    everyone done reading before a block is overwritten.

        pg = PeerGather(n_local, (H, W), device);  pg.stash[j] = disparity_j ...;  view = pg.gather()   # on a side stream

    gather() returns the dataset-ordered view [n_local, world, H, W] (pair j * world + r at [j, r]) of a buffer that the
    next gather() overwrites."""

    def __init__(self, n_local: int, tail, device, dtype=torch.float32) -> None:
        import torch.distributed._symmetric_memory as symm_mem
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        group = dist.group.WORLD
        try:
            symm_mem.enable_symm_mem_for_group(group.group_name)
        except Exception:  # noqa: BLE001 -- newer torch enables every group implicitly
            pass
        self.stash = symm_mem.empty((n_local,) + tuple(tail), dtype=dtype, device=device)
        self.hdl = symm_mem.rendezvous(self.stash, group.group_name)
        self.out = torch.empty((self.world, n_local) + tuple(tail), dtype=dtype, device=device)

    def gather(self) -> torch.Tensor:
        self.hdl.barrier(channel=0)
        for step in range(self.world):
            r = (self.rank - step) % self.world  # staggered: every link carries one pull at a time
            self.out[r].copy_(self.hdl.get_buffer(r, self.stash.shape, self.stash.dtype), non_blocking=True)
        self.hdl.barrier(channel=1)
        return self.out.transpose(0, 1)
