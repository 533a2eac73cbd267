"""Frame batches: many frames of one scene that differ only in the camera.

The reference renders its 261-frame camera dolly (`scenes/final_anim/anim.py:996-1004`, one scene file per
frame, `scenes/final_anim/Makefile:9-10`: `rrt -s 50 -w 1280 -h 720`) as 261 separate processes, each of
which re-parses the scene, rebuilds the world + BVH in one device thread and re-initialises curand; its
README quotes 3.3 hours for the set.  Here the scene and its LBVH are uploaded ONCE per GPU and each frame is
`rrtb_camera_set` + one render-kernel launch; across GPUs frames are dealt round-robin (frame f -> rank
f % N), which needs no collective at all.  A frame rendered this way is bit-identical to the same frame
rendered alone (tests/test_gpu_anim.py).
"""
import math

import numpy as np

from .api import camera_derive


def final_anim_cameras(width, height, n_frames=261):
    """The camera path of the reference's animation: lookfrom = (13 - i/10, 2, 3), lookat the origin,
    vfov 30, aperture 0.1, focus = |lookfrom|  (scenes/final_anim/anim.py:6-11,996-1000)."""
    cams = []
    aspect = float(np.float32(float(width) / height))
    for i in range(n_frames):
        fx, fy, fz = 13 - i / 10.0, 2.0, 3.0
        focus = math.sqrt(fx * fx + fy * fy + fz * fz)
        cams.append(camera_derive((np.float32(fx), np.float32(fy), np.float32(fz)), (0, 0, 0), (0, 1, 0), 30.0, aspect,
                                  float(np.float32(0.1)), float(np.float32(focus))))
    return cams


def frames_of_rank(n_frames, rank, world):
    """Round-robin frame ownership: no data-path collective, every rank writes its own frames."""
    return list(range(rank, n_frames, max(world, 1)))


def render_frames(ctx, cameras, width, height, spp, max_depth=50, seed=1984, rank=0, world=1, on_frame=None, count_rays=False):
    """Render this rank's frames of a camera-only animation with ONE uploaded scene (ctx.set_scene first).
    on_frame(index, image[H, W, 3] float32 sums, stats) is called per frame; returns summed stats."""
    total = dict(frames=0, rays=0, paths=0, seconds_render=0.0)
    out = np.empty((height, width, 3), np.float32)
    for f in frames_of_rank(len(cameras), rank, world):
        ctx.set_camera(cameras[f])
        img, st = ctx.render(width, height, spp, max_depth, seed, count_rays=count_rays, out=out)
        total["frames"] += 1
        total["rays"] += st["rays"]
        total["paths"] += st["paths"]
        total["seconds_render"] += st["seconds_render"]
        if on_frame is not None:
            on_frame(f, img, st)
    return total
