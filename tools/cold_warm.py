#!/usr/bin/env python
"""GPU check: device time of the first render of a fresh process against the following ones (what a one-shot CLI run pays
for the clock ramp and the kernel's first launch).  usage: cold_warm.py [spp]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import final_scene, W, H
from rrt_b200 import Context

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 500
scene, _ = final_scene()
ctx = Context(0)
ctx.set_scene(scene, True)
for i in range(4):
    _, st = ctx.render(W, H, spp, 50, 1984)
    print("render %d of this process: %.4f ms" % (i + 1, st["seconds_render"] * 1e3), flush=True)
