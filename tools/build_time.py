#!/usr/bin/env python
"""LBVH build time (upload + all build kernels, CUDA events inside rrtb_scene_set) per bench workload."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_workload, WORKLOADS
from rrt_b200 import Context

ctx = Context(0)
for wl in sys.argv[1:] or ["final", "test2", "synthetic_small", "synthetic"]:
    W, H = WORKLOADS[wl]["W"], WORKLOADS[wl]["H"]
    scene, _ = load_workload(wl, W, H)
    best = 1e9
    for _ in range(4):
        ctx.set_scene(scene, True)
        _, st = ctx.render(64, 36, 1, 2, 1)
        best = min(best, st["seconds_build"])
    n = scene.n_objects
    print("%-16s %8d prims  build %.3f ms  %.1f Mprims/s" % (wl, n, best * 1e3, n / best / 1e6), flush=True)
