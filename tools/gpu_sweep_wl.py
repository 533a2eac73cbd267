#!/usr/bin/env python
"""GPU tuning aid: time the render kernel on a bench workload under env-var overrides.
usage: gpu_sweep_wl.py <workload> <spp> VAR=v1,v2 ..."""
import os, sys, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_workload, WORKLOADS
from rrt_b200 import Context
wl, spp = sys.argv[1], int(sys.argv[2])
axes = [(a.split("=")[0], a.split("=")[1].split(",")) for a in sys.argv[3:]]
W, H = WORKLOADS[wl]["W"], WORKLOADS[wl]["H"]
scene, _ = load_workload(wl, W, H)
ctx = Context(0)
ctx.set_scene(scene, True)
ref, st = ctx.render(W, H, spp, 50, 1984, count_rays=True)
rays = st["rays"]
def run(tag):
    best = 1e9
    for _ in range(2):
        img, st = ctx.render(W, H, spp, 50, 1984)
        best = min(best, st["seconds_render"])
    print("%-40s %8.2f ms  %8.1f Mrays/s  same=%s" % (tag, best * 1e3, rays / best / 1e6, img.tobytes() == ref.tobytes()), flush=True)
run("default")
for combo in itertools.product(*[v for _, v in axes]):
    for (k, _), v in zip(axes, combo):
        os.environ[k] = v
    run(" ".join("%s=%s" % (k[5:], v) for (k, _), v in zip(axes, combo)))
