#!/usr/bin/env python
"""profiling aid: one render with each scheduler (final.txt, 1200x800, spp from argv)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import final_scene, W, H
from rrt_b200 import Context
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 32
scheds = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2]
scene, _ = final_scene()
ctx = Context(0)
ctx.set_scene(scene, True)
for sc in scheds:
    for _ in range(2):
        img, st = ctx.render(W, H, spp, 50, 1984, scheduler=sc)
    print("sched", sc, "%.2f ms" % (st["seconds_render"] * 1e3), flush=True)
