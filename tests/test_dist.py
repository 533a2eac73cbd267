"""-m "not gpu": the N > 1 host path on CPU -- world_size 2 and 3 over gloo.  The shard arithmetic of
rrt_b200/dist.py is checked against the oracle's sharded renders (the oracle stands in for the GPU kernel:
same (rank, world, shard_mode) contract), and the reduce of the integer accumulators over a real process
group must reproduce the single-rank image BIT FOR BIT."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def test_owner_map_partitions_the_image():
    from rrt_b200.dist import choose_shard_mode, owner_map, shard_paths

    for (W, H) in ((1200, 800), (100, 52), (17, 9)):
        for world in (1, 2, 3, 8):
            om = owner_map(W, H, world)
            assert om.shape == (H, W) and om.min() == 0 and om.max() == min(world, om.max() + 1) - 1 or world == 1
            counts = [int((om == r).sum()) for r in range(world)]
            assert sum(counts) == W * H
            for mode in (0, 1):
                assert sum(shard_paths(W, H, 10, r, world, mode) for r in range(world)) == W * H * 10
    assert choose_shard_mode(1200, 800, 500, 8) == 0
    assert choose_shard_mode(32, 16, 4096, 8) == 1
    # tile interleave balances pixel counts to within one tile row on the headline image
    om = owner_map(1200, 800, 8)
    c = np.bincount(om.ravel(), minlength=8)
    assert c.max() - c.min() <= 32 * 150


def _worker(rank, world, port, mode, W, H, spp, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import load_golden
    from oracle_lib import Oracle
    from rrt_b200.dist import owner_map, reduce_accumulators, shard_paths

    scene, _ = load_golden("test2")
    orc = Oracle(scene)
    _, fixed, cnt = orc.render(W, H, spp, 50, 7, rank=rank, world=world, shard_mode=mode)
    assert cnt["paths"] == shard_paths(W, H, spp, rank, world, mode)
    if mode == 0:  # a rank only ever touches the pixels it owns
        assert np.all(fixed[owner_map(W, H, world) != rank] == 0)
    acc = torch.from_numpy(fixed.astype(np.int64).reshape(-1).copy())
    reduce_accumulators(acc, 0)
    if rank == 0:
        q.put(acc.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,mode", [(2, 0), (2, 1), (3, 0)])
def test_gloo_reduce_reproduces_single_rank_image(world, mode):
    from conftest import load_golden
    from oracle_lib import Oracle

    W, H, spp = 44, 26, 4
    scene, _ = load_golden("test2")
    _, full, _ = Oracle(scene).render(W, H, spp, 50, 7)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world * 7 + mode
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, W, H, spp, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert np.array_equal(got.astype(np.uint64), full.reshape(-1))
