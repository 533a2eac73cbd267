"""-m gpu: the in-library multi-GPU path (include/rrtb.h "multi-GPU").  Every test also runs on a box
with ONE GPU: contexts (or processes) then share device 0, which exercises the same library code -- frame ownership,
tile-store / atomic-add epilogues, CUDA IPC mapping, pinned download -- minus the NVLink hop."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu


def _devices(n):
    import torch

    have = torch.cuda.device_count()
    return [k % have for k in range(n)]


@pytest.mark.parametrize("world", [2, 3, 8])
def test_render_group_equals_single_gpu(world):
    """rrtb_render_group over `world` contexts: tile shards STORED, sample shards ADDED into the owner's frame; both
    bit-identical to the one-GPU image, float and double frames, pageable and pinned destinations."""
    from rrt_b200 import Context, PinnedBuffer, render_group

    scene, _ = load_golden("final")
    W, H, spp = 200, 120, 8  # 25 x 30 tiles; W not a multiple of... (200 = 25 * 8), H = 30 * 4
    ctxs = [Context(d) for d in _devices(world)]
    try:
        for c in ctxs:
            c.set_scene(scene, use_bvh=True)
        want, st1 = ctxs[0].render(W, H, spp, 50, seed=5, count_rays=True)
        for mode in (0, 1):
            got, st = render_group(ctxs, W, H, spp, 50, seed=5, shard_mode=mode, count_rays=True)
            assert got.tobytes() == want.tobytes(), (world, mode)
            assert st["paths"] == W * H * spp and st["rays"] == st1["rays"]
        pin = PinnedBuffer((H, W, 3), np.float32)
        got, _ = render_group(ctxs, W, H, spp, 50, seed=5, out=pin.array)
        assert got.tobytes() == want.tobytes()
        pin.free()
        want64, _ = ctxs[0].render(W, H, spp, 50, seed=5, dtype=np.float64, precision="f64")
        got64, _ = render_group(ctxs, W, H, spp, 50, seed=5, dtype=np.float64, precision="f64")
        assert got64.tobytes() == want64.tobytes()
        # ragged image: edge tiles are partly outside, the frame is re-created for the new size
        W2, H2 = 61, 35
        want2, _ = ctxs[0].render(W2, H2, 4, 50, seed=6)
        got2, _ = render_group(ctxs, W2, H2, 4, 50, seed=6)
        assert got2.tobytes() == want2.tobytes()
    finally:
        for c in ctxs:
            c.close()


def test_frame_calls_reject_misuse():
    from rrt_b200 import Context, RrtbError

    scene, _ = load_golden("test1")
    with Context(0) as a, Context(0) as b:
        a.set_scene(scene)
        b.set_scene(scene)
        p = a.params(64, 40, 2, 50, 1, 0, 2, 0)
        with pytest.raises(RrtbError):  # no frame yet
            a.render_shard(p)
        with pytest.raises(RrtbError):  # attach before the owner has a frame
            b.frame_attach(a)
        a.frame_create(64, 40)
        b.frame_attach(a)
        with pytest.raises(RrtbError):  # size differs from the frame's
            a.render_shard(a.params(32, 40, 2, 50, 1, 0, 2, 0))
        with pytest.raises(RrtbError):  # rank out of range
            a.render_shard(a.params(64, 40, 2, 50, 1, 2, 2, 0))
        with pytest.raises(RrtbError):  # the download is the owner's call
            b.frame_download(np.empty((40, 64, 3), np.float32))
        a.render_shard(p)
        b.render_shard(b.params(64, 40, 2, 50, 1, 1, 2, 0))
        out = a.frame_download(np.empty((40, 64, 3), np.float32))
        want, _ = a.render(64, 40, 2, 50, seed=1)
        assert out.tobytes() == want.tobytes()


def _ipc_worker(rank, world, port, mode, W, H, spp, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import load_golden
    from rrt_b200 import Context
    from rrt_b200.dist import DistributedRenderer

    scene, _ = load_golden("test2")
    ctx = Context(rank % torch.cuda.device_count())
    ctx.set_scene(scene, use_bvh=True)
    dr = DistributedRenderer(ctx, rank, world)
    for frame in range(2):  # two frames through the same mapping
        img, st = dr.render(W, H, spp, 50, seed=7 + frame, shard_mode=mode)
        if rank == 0:
            q.put(img.copy())
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,mode", [(2, 0), (2, 1), (3, 0)])
def test_one_process_per_gpu_frame_over_cuda_ipc(world, mode):
    """The bench.py / torchrun layout: one process per rank, rank 0 exports its frame (CUDA IPC), the others import it
    and their epilogues write into it; gloo carries the handle and the barriers."""
    import torch.multiprocessing as mp

    from rrt_b200 import Context

    W, H, spp = 120, 68, 4
    scene, _ = load_golden("test2")
    with Context(0) as c:
        c.set_scene(scene, use_bvh=True)
        want = [c.render(W, H, spp, 50, seed=7 + f)[0] for f in range(2)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() % 1500) + world * 7 + mode
    procs = [ctx.Process(target=_ipc_worker, args=(r, world, port, mode, W, H, spp, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    for g, w in zip(got, want):
        assert g.tobytes() == w.tobytes()
