"""Host-side logic of the multi-GPU path on CPU: seed partitioning (main.py:575-576) and the
statistics all-gather, exercised with a world_size-2 gloo process group."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_partition_seeds_matches_reference_rule():
    from oac_explore_b200.seed_group import partition_seeds
    for total in (1, 7, 8, 64):
        for world in (1, 2, 4, 8):
            parts = [partition_seeds(total, r, world) for r in range(world)]
            flat = sorted(s for p in parts for s in p)
            assert flat == list(range(total))
            for r, p in enumerate(parts):
                assert all(s % world == r for s in p)          # gpu = seed % n_gpus
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oac_explore_b200.seed_group import partition_seeds, allgather_stats
    ids = partition_seeds(total, rank, world)
    local = torch.tensor([[float(s), 10.0 * s, -float(s)] for s in ids]).reshape(len(ids), 3)
    out = allgather_stats(local, ids, total)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [5, 8])
def test_allgather_stats_gloo_world2(total):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = torch.tensor([[float(s), 10.0 * s, -float(s)] for s in range(total)])
    for r in range(world):
        assert torch.equal(res[r], expect)


def test_allgather_stats_single_process():
    from oac_explore_b200.seed_group import allgather_stats
    out = allgather_stats(torch.tensor([[1.0, 2.0], [3.0, 4.0]]), [2, 0], 3)
    assert torch.equal(out, torch.tensor([[3.0, 4.0], [0.0, 0.0], [1.0, 2.0]]))
