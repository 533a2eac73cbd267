"""Multi-GPU host for the render path: one process per GPU, `torch.distributed` for the plumbing.

The reference has no multi-GPU path at all (`-D <n>` selects ONE device, main.cpp:107-110).
The path shards trivially -- pixels and samples are independent -- so there is no data-path collective
inside the render.  The combine is done by the library's own kernels over NVLink peer memory
(include/rrtb.h "multi-GPU"): rank 0 owns the frame, the other ranks map it once (CUDA IPC handle sent with one
`broadcast_object_list`), and every rank's resolve epilogue writes its shard straight into it:

  RRTB_SHARD_TILES    interleaved 8x4-pixel tiles, tile t -> rank t % world  (default): the tiles are
                      disjoint, each rank STORES its own as float3 -- 12 B/pixel/world over the link, no reduce
  RRTB_SHARD_SAMPLES  sample s -> rank s % world   (very high spp on small images): 64-bit fixed-point
                      partial sums ADDED with integer atomics; associative, so bit-identical for any order

torch.distributed only carries the handle and the two barriers per frame.  `reduce_accumulators` (an NCCL / gloo
sum of full-size accumulators) remains for callers that keep their shards in torch tensors, and for the CPU tests.
"""
import os

import numpy as np

from .types import SHARD_SAMPLES, SHARD_TILES

TILE_W, TILE_H = 8, 4


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def choose_shard_mode(width, height, spp, world):
    """Tiles unless there are too few tiles to keep `world` GPUs busy (then split the samples)."""
    tiles = ((width + TILE_W - 1) // TILE_W) * ((height + TILE_H - 1) // TILE_H)
    if tiles >= 64 * world or spp < world:
        return SHARD_TILES
    return SHARD_SAMPLES


def owner_map(width, height, world):
    """[H, W] int array: the rank that owns each pixel under RRTB_SHARD_TILES (row 0 = bottom scanline)."""
    tiles_x = (width + TILE_W - 1) // TILE_W
    j, i = np.meshgrid(np.arange(height), np.arange(width), indexing="ij")
    return ((j // TILE_H) * tiles_x + (i // TILE_W)) % world


def shard_paths(width, height, spp, rank, world, mode):
    """Number of camera paths rank `rank` traces."""
    if world <= 1:
        return width * height * spp
    if mode == SHARD_TILES:
        return int((owner_map(width, height, world) == rank).sum()) * spp
    return width * height * len(range(rank, spp, world))


def reduce_accumulators(acc, dst=0, group=None):
    """Sum the per-rank uint64 (stored as int64) accumulators onto rank `dst`.  `acc` is a torch tensor on
    this rank's device (CUDA + NCCL in production, CPU + gloo in the CPU tests)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(acc, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return acc


class DistributedRenderer:
    """One instance per rank.  render() returns the [H, W, 3] sums on rank 0 (None elsewhere)."""

    def __init__(self, ctx, rank=None, world=None, group=None):
        r, w, _ = env_rank_world()
        self.ctx = ctx
        self.rank = r if rank is None else rank
        self.world = w if world is None else world
        self.group = group
        self._frame = None  # (width, height, f64) the mapping was made for

    def _barrier(self):
        import torch.distributed as dist

        if self.world > 1:
            dist.barrier(group=self.group)

    def ensure_frame(self, width, height, f64=False):
        """Rank 0 creates (or re-uses) the frame, the others map it.  Collective."""
        import torch.distributed as dist

        if self._frame == (width, height, f64):
            return
        box = [None]
        if self.rank == 0:
            self.ctx.frame_create(width, height, f64)
            box[0] = self.ctx.frame_export()
        if self.world > 1:
            dist.broadcast_object_list(box, src=0, group=self.group)
            if self.rank != 0:
                self.ctx.frame_import(box[0])
        self._frame = (width, height, f64)

    def render(self, width, height, spp, max_depth=50, seed=1984, shard_mode=None, out=None, precision="f32", dtype=np.float32):
        mode = choose_shard_mode(width, height, spp, self.world) if shard_mode is None else shard_mode
        f64 = np.dtype(dtype) == np.dtype(np.float64)
        self.ensure_frame(width, height, f64)
        p = self.ctx.params(width, height, spp, max_depth, seed, self.rank, self.world, mode, precision=precision)
        self._barrier()  # the previous frame has been downloaded
        stats = self.ctx.render_shard(p)
        self._barrier()  # every rank's stores have landed in the owner's frame
        if self.rank != 0:
            return None, stats
        if out is None:
            out = np.empty((height, width, 3), dtype)
        return self.ctx.frame_download(out, mode), stats
