"""Tree-quality experiments on the CPU (design aid, not product): record the ray population of a workload with the
oracle (camera rays + scattered rays, wavefront by bounce), then count node visits / box tests / leaf tests of
candidate traversal structures over it (tools/treelab/treelab.c)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_lib import Oracle, philox  # noqa: E402
from rrt_b200.types import SceneArrays  # noqa: E402


class LabCounters(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("rays", "visits", "box_tests", "leaf_tests", "hits", "pushes", "maxstack")]

    def d(self):
        r = max(self.rays, 1)
        return dict(rays=self.rays, visits=self.visits / r, box=self.box_tests / r, leaf=self.leaf_tests / r,
                    hit=self.hits / r, push=self.pushes / r, maxstack=self.maxstack)


def lab():
    so = os.path.join(ROOT, "tools", "treelab", "libtreelab.so")
    lib = C.CDLL(so)
    for f in ("lab_build_sah", "lab_from_arrays", "lab_collapse", "lab_build_ploc", "lab_build_hybrid", "lab_build_hybrid2", "lab_collapse_dp"):
        getattr(lib, f).restype = C.c_void_p
    lib.lab_sah_cost.restype = C.c_double
    lib.lab_sah_cost.argtypes = [C.c_void_p]
    lib.lab_wide_fill.restype = C.c_double
    lib.lab_wide_fill.argtypes = [C.c_void_p]
    lib.lab_set_cleaf.argtypes = [C.c_double]
    return lib


def record_rays(scene, W, H, n_paths, max_depth=50, seed=1984, rng_seed=1):
    """All ray segments of n_paths camera paths (random pixels, sample index = path index)."""
    o = Oracle(scene)
    rng = np.random.default_rng(rng_seed)
    pix = rng.integers(0, W * H, size=n_paths).astype(np.int32)
    rays = np.zeros((n_paths, 7), np.float32)
    for smp in range(1):
        rays = o.camera_rays(W, H, pix, 0, seed)
    all_rays = []
    mats = np.concatenate([scene.spheres["material"], scene.mspheres["material"], scene.triangles["material"],
                           scene.mtriangles["material"]]).astype(np.int32)
    alive_pix = pix.copy()
    for b in range(max_depth):
        if len(rays) == 0:
            break
        all_rays.append(rays.copy())
        ids, t, rec = o.trace(rays, 0.001, "bvh", want_rec=True)
        hit = ids >= 0
        rays, rec, ids, alive_pix = rays[hit], rec[hit], ids[hit], alive_pix[hit]
        n = len(rays)
        if n == 0:
            break
        in16 = np.zeros((n, 16), np.float32)
        in16[:, 0:7] = rays
        in16[:, 7:14] = rec
        in16[:, 14] = mats[ids]
        ctr = np.zeros((n, 4), np.uint32)
        ctr[:, 0] = alive_pix
        ctr[:, 1] = 0
        ctr[:, 2] = 2 + b
        rnd = np.random.default_rng(b + 100).integers(0, 2**32, size=(n, 4), dtype=np.uint64).astype(np.uint32)
        out = o.scatter(in16, rnd)
        ok = out[:, 6] != 0
        nxt = np.zeros((n, 7), np.float32)
        nxt[:, 0:3] = rec[:, 0:3]
        nxt[:, 3:6] = out[:, 0:3]
        nxt[:, 6] = rays[:, 6]
        rays, alive_pix = nxt[ok], alive_pix[ok]
    return np.concatenate(all_rays, axis=0)


def run(scene, rays, label=""):
    L = lab()
    o = Oracle(scene)
    ba = o.bvh_arrays()
    pbox = np.ascontiguousarray(ba["prim_box"], np.float32)
    n = len(pbox)
    vp = lambda a: C.c_void_p(a.ctypes.data)
    rays = np.ascontiguousarray(rays, np.float32)
    m = len(rays)
    ref_ids, _ = o.trace(rays, 0.001, "bvh")
    trees = {}
    trees["lbvh"] = C.c_void_p(L.lab_from_arrays(n, vp(pbox), vp(ba["left"]), vp(ba["right"]), vp(ba["perm"])))
    trees["sah"] = C.c_void_p(L.lab_build_sah(n, vp(pbox)))
    for bins in (8, 16, 32):
        L.lab_set_sah(0, bins)
        trees["hyb512b%d" % bins] = C.c_void_p(L.lab_build_hybrid(trees["lbvh"], 512))
    # a second SAH level over the roots of those subtrees (<= 512 / 4096 roots per upper subtree), and larger subtrees
    L.lab_set_sah(0, 32)
    trees["hyb2_512_512"] = C.c_void_p(L.lab_build_hybrid2(trees["lbvh"], 512, 512))
    trees["hyb2_512_4096"] = C.c_void_p(L.lab_build_hybrid2(trees["lbvh"], 512, 4096))
    trees["hyb2048b32"] = C.c_void_p(L.lab_build_hybrid(trees["lbvh"], 2048))
    L.lab_set_sah(4096, 32)
    trees["hyb512sweep"] = C.c_void_p(L.lab_build_hybrid(trees["lbvh"], 512))
    for r in ():
        trees["ploc%d" % r] = C.c_void_p(L.lab_build_ploc(n, vp(pbox), vp(ba["perm"]), r))
    for name, t in trees.items():
        c = LabCounters()
        ids = np.zeros(m, np.int32)
        L.lab_trace_binary(t, C.byref(o._s), vp(rays), m, C.c_float(0.001), C.byref(c), vp(ids))
        assert np.array_equal(ids, ref_ids), name
        print("%-22s sahcost %7.2f  " % (label + name + " bin", L.lab_sah_cost(t)), c.d())
        for k in (4,):
            for q in (0,):
              for mode in ("greedy", "dp1.0"):
                if mode == "greedy":
                    w = C.c_void_p(L.lab_collapse(t, k))
                else:
                    L.lab_set_cleaf(float(mode[2:]))
                    w = C.c_void_p(L.lab_collapse_dp(t, k))
                L.lab_quantise(w, q)
                L.lab_axis_sort(w)
                for cull, om in ((0, 1),):
                    L.lab_set_order(om)
                    c = LabCounters()
                    L.lab_trace_wide(w, C.byref(o._s), vp(rays), m, C.c_float(0.001), cull, C.byref(c), vp(ids))
                    assert np.array_equal(ids, ref_ids), (name, k)
                    print("%-30s nodes %7d fill %.2f " % ("%s%s w%d %s" % (label, name, k, mode), L.lab_wide_nodes(w),
                                                          L.lab_wide_fill(w)), c.d())


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "final"
    if which == "synthetic":
        import tempfile
        from rrt_b200 import Scene
        from rrt_b200.synthetic import write_synthetic_scene
        kw = dict(n_spheres=int(sys.argv[2]), ico_level=int(sys.argv[3]), grid=7) if len(sys.argv) > 3 else dict(n_spheres=20000, ico_level=4, grid=7)
        pth = os.path.join(tempfile.gettempdir(), "lab_syn.txt")
        write_synthetic_scene(pth, **kw)
        W, H = 3840, 2160
        scene = Scene.from_file(pth, W, H).arrays
    else:
        d = np.load(os.path.join(ROOT, "tests", "golden", "scene_%s.npz" % which))
        scene = SceneArrays.from_npz_dict(d)
        W, H = (1200, 800) if which in ("final", "test1") else (1920, 1080)
    npaths = int(os.environ.get("PATHS", "20000"))
    rays = record_rays(scene, W, H, npaths)
    print("recorded", len(rays), "rays from", npaths, "paths")
    run(scene, rays)
