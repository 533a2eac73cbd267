#!/usr/bin/env python
"""GPU tuning aid: time the render kernel variants on final.txt and check they give the same image."""
import os, sys, json, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bench import final_scene, W, H
from rrt_b200 import Context

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
scene, _ = final_scene()
ctx = Context(0)
ctx.set_scene(scene, True)
ref, st = ctx.render(W, H, spp, 50, 1984, count_rays=True, scheduler=1)
rays = st["rays"]
def run(tag, sched, env=None):
    for k, v in (env or {}).items():
        os.environ[k] = str(v)
    best = 1e9
    for _ in range(3):
        img, st = ctx.render(W, H, spp, 50, 1984, scheduler=sched)
        best = min(best, st["seconds_render"])
    same = img.tobytes() == ref.tobytes()
    print("%-28s %8.2f ms  %8.1f Mrays/s  same=%s" % (tag, best * 1e3, rays / best / 1e6, same), flush=True)
run("simple", 1)
run("pool default", 2)
for tf in (1, 2, 4, 8, 12, 16, 24):
    for it in (2, 4):
        run("pool tf=%d iters=%d" % (tf, it), 2, {"RRTB_TH_FETCH": tf, "RRTB_STEP_ITERS": it})
