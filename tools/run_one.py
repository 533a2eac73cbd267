#!/usr/bin/env python
"""Profiling aid: one render of a bench workload.  usage: run_one.py <workload> <spp> [f32|f64] [scheduler]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_workload, WORKLOADS
from rrt_b200 import Context
import numpy as np

wl, spp = sys.argv[1], int(sys.argv[2])
prec = sys.argv[3] if len(sys.argv) > 3 else "f32"
sched = int(sys.argv[4]) if len(sys.argv) > 4 else 0
W, H = WORKLOADS[wl]["W"], WORKLOADS[wl]["H"]
scene, _ = load_workload(wl, W, H)
ctx = Context(0)
ctx.set_scene(scene, True)
for _ in range(2):
    img, st = ctx.render(W, H, spp, 50, 1984, count_rays=False, scheduler=sched, precision=prec, dtype=np.float64 if prec == "f64" else np.float32)
    print("%s %s spp=%d  %.2f ms" % (wl, prec, spp, st["seconds_render"] * 1e3), flush=True)
