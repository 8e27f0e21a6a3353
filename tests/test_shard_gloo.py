"""Multi-process host logic (pairs sharded by rank, gather of outputs): gloo on CPU, and -- where two GPUs are present -- the
copy-engine gather out of symmetric memory."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_pairs, out_dir):
    sys.path.insert(0, ROOT)
    from esmstereo_b200.shard import gather_disparities, shard_indices
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_indices(n_pairs, rank, world)
    # stand-in for the per-rank forward: "disparity" of pair i is a constant image of value i
    local = torch.stack([torch.full((4, 6), float(i)) for i in mine]) if mine else torch.zeros(0, 4, 6)
    full = gather_disparities(local, n_pairs, rank, world)
    view = gather_disparities(local, n_pairs, rank, world, flat=False)  # [per_rank, world, H, W]: pair j * world + r at [j, r]
    if rank == 0:
        torch.save(full, os.path.join(out_dir, "full.pt"))
        torch.save(view.clone(), os.path.join(out_dir, "view.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_pairs_shard_and_gather_world2(tmp_path):
    n_pairs, world = 5, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_pairs, str(tmp_path)), nprocs=world, join=True)
    full = torch.load(os.path.join(str(tmp_path), "full.pt"))
    assert full.shape == (n_pairs, 4, 6)
    for i in range(n_pairs):  # every pair exactly once, in dataset order
        assert torch.all(full[i] == float(i))
    view = torch.load(os.path.join(str(tmp_path), "view.pt"))  # the copy-free form: same order through a two-level index
    assert view.shape == ((n_pairs + world - 1) // world, world, 4, 6)
    for i in range(n_pairs):
        assert torch.all(view[i // world, i % world] == float(i))


def test_shard_indices_cover_everything_once():
    sys.path.insert(0, ROOT)
    from esmstereo_b200.shard import shard_indices
    for n in (0, 1, 7, 8, 9):
        for world in (1, 2, 3, 8):
            seen = sorted(i for r in range(world) for i in shard_indices(n, r, world))
            assert seen == list(range(n))


def _peer_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from esmstereo_b200.shard import PeerGather
    pg = PeerGather(3, (4, 6), dev)
    for sweep in range(2):  # the buffers are reused: two sweeps check the barriers, not just the copies
        for j in range(3):
            pg.stash[j].fill_(float(100 * sweep + j * world + rank))  # "disparity" of pair j * world + rank
        view = pg.gather()
        torch.cuda.synchronize()
        got = view.clone().cpu()
        if rank == 0:
            torch.save(got, os.path.join(out_dir, "peer%d.pt" % sweep))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_peer_gather_symmetric_memory_world2(tmp_path):
    """shard.PeerGather (copy-engine pulls out of symmetric memory): dataset order through the [j, r] view, two sweeps."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    mp.spawn(_peer_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for sweep in range(2):
        view = torch.load(os.path.join(str(tmp_path), "peer%d.pt" % sweep))
        assert view.shape == (3, world, 4, 6)
        for i in range(3 * world):
            assert torch.all(view[i // world, i % world] == float(100 * sweep + i))
