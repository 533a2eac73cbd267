"""Multi-GPU host for the render path: one process per GPU, `torch.distributed` for the plumbing.

The reference has no multi-GPU path at all (`-D <n>` selects ONE device, main.cpp:107-110).
The path shards trivially -- pixels and samples are independent -- so there is no data-path collective
inside the render: every rank renders its shard of the SAME image into a full-size 64-bit fixed-point
accumulator (zero where it owns nothing) and the only exchange is ONE sum of those accumulators onto
rank 0.  Integer addition is associative, so the reduced image is bit-identical to the single-GPU image
for any world size, either shard mode and any reduction order.

  RRTB_SHARD_TILES    interleaved 8x4-pixel tiles, tile t -> rank t % world  (default)
  RRTB_SHARD_SAMPLES  sample s -> rank s % world   (very high spp on small images)

The reduce is NCCL over NVLink/NVSwitch (`reduce_accumulators`); the message is 24 B/pixel (23 MB at
1200x800, 199 MB at 4K) against seconds of rendering.
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
    """One instance per rank.  render() returns the float32 [H, W, 3] sums on rank 0 (None elsewhere)."""

    def __init__(self, ctx, rank=None, world=None, group=None):
        r, w, _ = env_rank_world()
        self.ctx = ctx
        self.rank = r if rank is None else rank
        self.world = w if world is None else world
        self.group = group

    def render(self, width, height, spp, max_depth=50, seed=1984, shard_mode=None, to_host=True):
        import torch

        mode = choose_shard_mode(width, height, spp, self.world) if shard_mode is None else shard_mode
        dev = torch.device("cuda", self.ctx.device)
        n = 3 * width * height
        acc = torch.zeros(n, dtype=torch.int64, device=dev)
        torch.cuda.synchronize(dev)
        p = self.ctx.params(width, height, spp, max_depth, seed, self.rank, self.world, mode)
        stats = self.ctx.render_device(p, acc.data_ptr())
        reduce_accumulators(acc, 0, self.group)
        torch.cuda.synchronize(dev)
        if self.rank != 0:
            return None, stats
        out = torch.empty(n, dtype=torch.float32, device=dev)
        self.ctx.resolve_device(acc.data_ptr(), out.data_ptr(), n)
        if to_host:
            return out.cpu().numpy().reshape(height, width, 3), stats
        return out.view(height, width, 3), stats
