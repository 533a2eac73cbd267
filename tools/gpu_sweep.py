#!/usr/bin/env python
"""GPU tuning aid: time the render kernel under env-var overrides on final.txt; check the image is unchanged.
usage: gpu_sweep.py <spp> VAR=v1,v2,... [VAR2=...]"""
import os, sys, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import final_scene, W, H
from rrt_b200 import Context

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
axes = [(a.split("=")[0], a.split("=")[1].split(",")) for a in sys.argv[2:]]
scene, _ = final_scene()
ctx = Context(0)
ctx.set_scene(scene, True)
ref, st = ctx.render(W, H, spp, 50, 1984, count_rays=True, scheduler=1)
rays = st["rays"]
def run(tag, sched):
    best = 1e9
    for _ in range(3):
        img, st = ctx.render(W, H, spp, 50, 1984, scheduler=sched)
        best = min(best, st["seconds_render"])
    print("%-40s %8.2f ms  %8.1f Mrays/s  same=%s" % (tag, best * 1e3, rays / best / 1e6, img.tobytes() == ref.tobytes()), flush=True)
run("simple", 1)
run("pool default", 2)
for combo in itertools.product(*[v for _, v in axes]):
    for (k, _), v in zip(axes, combo):
        os.environ[k] = v
    run("pool " + " ".join("%s=%s" % (k[5:], v) for (k, _), v in zip(axes, combo)), 2)
