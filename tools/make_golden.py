#!/usr/bin/env python
"""Generate tests/golden/* by EXECUTING the unmodified reference (oracle/_ref, built by oracle/Makefile
from /root/reference).  The reference ships no golden vectors of its own, so these
fixtures are what pins parity on machines where /root/reference does not exist (the GPU box).

    python tools/make_golden.py            # needs oracle/_ref/ (make -C oracle ref)

Outputs
  scene_<name>.npz   the scene exactly as the reference's float parser holds it (scene.h:212-452),
                     pixel-centre primary rays, and for those rays the reference hittable_list scan
                     (hittable_list.h:95-117) in DOUBLE on the float-rounded scene: object id, t, hit
                     record; per-object bounding boxes (float build, camera shutter interval).
  kat_math.npz       reflect / refract / Schlick reflectance / convert_color / dielectric + metal
                     scatter known answers from the reference's float build.
  <scene>_<WxH>_s<spp>_rrto.png   reference renders (rrto, double, OMP_NUM_THREADS=1 => deterministic).
"""
import ctypes as C
import os
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle_lib import REF_DIR, RefScene, RefWorld, pinhole_rays, ref_lib, ref_scene_path  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# scene -> (W, H) of the BASELINE.json config it appears in, and the ray-grid step for the fixture
SCENES = {"test1": (1200, 800, 10), "test2": (1920, 1080, 16), "test3": (1920, 1080, 16), "final": (1200, 800, 8)}

RENDERS = {  # name: (scene, W, H, spp)
    "final_600x400_s500_rrto.png": ("final", 600, 400, 500),
    "test1_480x320_s256_rrto.png": ("test1", 480, 320, 256),
    "test2_480x270_s256_rrto.png": ("test2", 480, 270, 256),
    "test3_480x270_s256_rrto.png": ("test3", 480, 270, 256),
}


def make_scene(name, W, H, step):
    rs = RefScene(ref_scene_path(name + ".txt"), W, H, "f")
    sc = rs.arrays()
    rays = pinhole_rays(sc, W, H, step)
    if sc.camera["time0"][0] != sc.camera["time1"][0]:
        rng = np.random.default_rng(20221005)
        rays[:, 6] = rng.uniform(sc.camera["time0"][0], sc.camera["time1"][0], len(rays)).astype(np.float32)
    wd = RefWorld(sc, "d")
    ids, t = wd.trace_scan(rays)
    rec = wd.trace_world(wd.list, rays)
    wf = RefWorld(sc, "f")
    t0, t1 = float(sc.camera["time0"][0]), float(sc.camera["time1"][0])
    boxes = np.stack([wf.bounding_box(i, t0, t1) for i in range(sc.n_objects)]).astype(np.float32)
    counts = np.array([rs.counts[k] for k in ("materials", "spheres", "mspheres", "triangles", "objs", "obj_insts")], np.int32)
    np.savez_compressed(
        os.path.join(GOLD, "scene_%s.npz" % name),
        W=W, H=H, step=step, counts=counts, rays=rays, ref_id=ids, ref_t=t, ref_rec=rec, ref_boxes=boxes,
        **sc.to_npz_dict(),
    )
    print(name, rs.counts, "rays", len(rays), "hits", int((ids >= 0).sum()))


def make_kat():
    lib = ref_lib("f")
    rng = np.random.default_rng(1984)
    n = 256
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    nn = rng.normal(size=(n, 3))
    nn /= np.linalg.norm(nn, axis=1, keepdims=True)
    nn[np.sum(v * nn, axis=1) > 0] *= -1  # face-forwarded normals
    v32, n32 = v.astype(np.float32).astype(np.float64), nn.astype(np.float32).astype(np.float64)
    refl = np.zeros((n, 3))
    refr = np.zeros((n, 3))
    eta = np.where(rng.uniform(size=n) < 0.5, 1.5, 1.0 / 1.5).astype(np.float32).astype(np.float64)
    for i in range(n):
        lib.ref_reflect(C.c_void_p(v32[i].ctypes.data), C.c_void_p(n32[i].ctypes.data), C.c_void_p(refl[i].ctypes.data))
        lib.ref_refract(C.c_void_p(v32[i].ctypes.data), C.c_void_p(n32[i].ctypes.data), C.c_double(eta[i]), C.c_void_p(refr[i].ctypes.data))
    cosv = np.linspace(0, 1, 65).astype(np.float32).astype(np.float64)
    schlick = np.array([[lib.ref_reflectance(c, e) for c in cosv] for e in (1.5, 1.0 / 1.5, 1.33, 2.4)])
    # tonemap
    sums = np.concatenate([rng.uniform(0, 12, size=(200, 3)), [[0, 0, 0], [10, 10, 10], [9.99, 5, 1e-3], [20, 0.1, 10.0]]]).astype(np.float32)
    rgb = np.zeros((len(sums), 3), np.int32)
    for i in range(len(sums)):
        lib.ref_convert_color(C.c_double(sums[i, 0]), C.c_double(sums[i, 1]), C.c_double(sums[i, 2]), 10, C.c_void_p(rgb[i].ctypes.data))
    # dielectric scatter: the reference picks reflect or refract with probability = reflectance; record
    # both outcome directions and the empirical reflect frequency over many calls (its own mt19937).
    lib.ref_material_create.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]
    die = C.c_void_p(lib.ref_material_create(2, 0, 0, 0, 1.5))
    met = C.c_void_p(lib.ref_material_create(1, 0.8, 0.6, 0.2, 0.0))
    nd = 48
    d_in = (v32[:nd] * rng.uniform(0.5, 6.0, size=(nd, 1))).astype(np.float32).astype(np.float64)
    front = (rng.uniform(size=nd) < 0.5).astype(np.int32)
    p = rng.uniform(-1, 1, size=(nd, 3)).astype(np.float32).astype(np.float64)
    die_dirs = np.zeros((nd, 2, 3))  # [reflect, refract] (nan if never seen)
    die_dirs[:] = np.nan
    die_freq = np.zeros(nd)
    met_dir = np.zeros((nd, 3))
    met_ok = np.zeros(nd, np.int32)
    trials = 4000
    for i in range(nd):
        ray7 = np.concatenate([p[i] - d_in[i], d_in[i], [0.0]])
        out = np.zeros(9)
        lib.ref_reflect(C.c_void_p((d_in[i] / np.linalg.norm(d_in[i])).ctypes.data), C.c_void_p(n32[i].ctypes.data), C.c_void_p(out.ctypes.data))
        nrefl = 0
        for _ in range(trials):
            lib.ref_scatter(die, C.c_void_p(ray7.ctypes.data), C.c_void_p(p[i].ctypes.data), C.c_void_p(n32[i].ctypes.data), int(front[i]), C.c_void_p(out.ctypes.data))
            d = out[3:6].copy()
            is_refl = np.dot(d, n32[i]) > 0
            nrefl += is_refl
            die_dirs[i, 0 if is_refl else 1] = d
        die_freq[i] = nrefl / trials
        met_ok[i] = lib.ref_scatter(met, C.c_void_p(ray7.ctypes.data), C.c_void_p(p[i].ctypes.data), C.c_void_p(n32[i].ctypes.data), 1, C.c_void_p(out.ctypes.data))
        met_dir[i] = out[3:6]
    np.savez_compressed(
        os.path.join(GOLD, "kat_math.npz"),
        v=v32.astype(np.float32), n=n32.astype(np.float32), eta=eta.astype(np.float32), reflect=refl.astype(np.float32),
        refract=refr.astype(np.float32), schlick_cos=cosv.astype(np.float32), schlick_ior=np.array([1.5, 1.0 / 1.5, 1.33, 2.4], np.float32),
        schlick=schlick.astype(np.float32), tm_sums=sums, tm_spp=10, tm_rgb=rgb,
        sc_d_in=d_in.astype(np.float32), sc_p=p.astype(np.float32), sc_n=n32[:nd].astype(np.float32), sc_front=front,
        die_dirs=die_dirs.astype(np.float32), die_reflect_freq=die_freq, die_trials=trials, met_dir=met_dir.astype(np.float32), met_ok=met_ok,
    )
    print("kat_math written")


def make_renders(src_dir=None):
    for fn, (scene, W, H, spp) in RENDERS.items():
        dst = os.path.join(GOLD, fn)
        if src_dir and os.path.exists(os.path.join(src_dir, fn)):
            shutil.copy(os.path.join(src_dir, fn), dst)
            continue
        if os.path.exists(dst):
            continue
        env = dict(os.environ, OMP_NUM_THREADS="1")
        subprocess.check_call(
            [os.path.join(REF_DIR, "rrto"), "-i", ref_scene_path(scene + ".txt"), "-w", str(W), "-h", str(H), "-s", str(spp), "-d", "50", "-o", dst],
            env=env, stderr=subprocess.DEVNULL,
        )
        print("rendered", fn)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    for name, (W, H, step) in SCENES.items():
        make_scene(name, W, H, step)
    make_kat()
    make_renders(sys.argv[1] if len(sys.argv) > 1 else None)
