#!/usr/bin/env python
"""Small end-to-end workload for compute-sanitizer (one tool per call):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import load_golden
from rrt_b200 import Context, Scene
from rrt_b200.synthetic import write_synthetic_scene

ctx = Context(0)
for name in ("final", "test2", "test3"):
    scene, d = load_golden(name)
    ctx.set_scene(scene, True)
    ids, t = ctx.trace(d["rays"][:2000], 0.001, "bvh")
    ids2, t2 = ctx.trace(d["rays"][:200], 0.001, "scan")
    for sched in (1, 2):
        img, st = ctx.render(50, 34, 3, 50, 7, scheduler=sched, count_rays=True)
    img, st = ctx.render(33, 17, 2, 50, 7, rank=1, world=3, shard_mode=1)
    print(name, "ok", st["paths"], flush=True)
p = "/tmp/_san_synth.txt"
write_synthetic_scene(p, n_spheres=1500, ico_level=2, grid=3)   # 1504 spheres + 2880 triangles: 5 sort segments
sc = Scene.from_file(p, 64, 36)
ctx.set_scene(sc, True)
ctx.bvh_arrays()
img, st = ctx.render(64, 36, 2, 50, 3, count_rays=True)
print("synthetic ok", st["rays"], flush=True)
ctx.close()
