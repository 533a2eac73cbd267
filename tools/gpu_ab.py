#!/usr/bin/env python
"""GPU A/B aid: time the render kernel of ONE library build (RRTB_LIB selects it) on the headline image and on the
1.1 M-primitive scene, and print a hash of each image -- every build must print the same hashes (the closest hit does not
depend on the traversal).  usage: gpu_ab.py <tag> [final_spp [synthetic_spp]] [VAR=v1,v2 ...] (variables: tuning builds)"""
import hashlib
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS, load_workload  # noqa: E402
from rrt_b200 import Context  # noqa: E402

tag = sys.argv[1]
nums = [a for a in sys.argv[2:] if "=" not in a]
axes = [(a.split("=")[0], a.split("=")[1].split(",")) for a in sys.argv[2:] if "=" in a]
spp = {"final": int(nums[0]) if nums else 64, "synthetic": int(nums[1]) if len(nums) > 1 else 8}
ctx = Context(0)
for wl in ("final", "synthetic"):
    if spp[wl] <= 0:
        continue
    W, H = WORKLOADS[wl]["W"], WORKLOADS[wl]["H"]
    scene, _ = load_workload(wl, W, H)
    ctx.set_scene(scene, True)
    _, st = ctx.render(W, H, spp[wl], 50, 1984, count_rays=True)
    rays = st["rays"]

    def run(label):
        best = 1e9
        for _ in range(3):
            img, st = ctx.render(W, H, spp[wl], 50, 1984)
            best = min(best, st["seconds_render"])
        print("%-14s %-10s %-34s %9.3f ms %9.1f Mrays/s  img %s" % (tag, wl, label, best * 1e3, rays / best / 1e6,
                                                                  hashlib.sha1(img.tobytes()).hexdigest()[:12]), flush=True)

    run("default")
    for combo in itertools.product(*[v for _, v in axes]):
        for (k, _), v in zip(axes, combo):
            os.environ[k] = v
        run(" ".join("%s=%s" % (k[5:], v) for (k, _), v in zip(axes, combo)))
    for k, _ in axes:
        os.environ.pop(k, None)
