"""Randomised scenes (mixed spheres / moving spheres / triangles, wide range of sizes and distances).
CPU: the oracle against the LIVE reference (double build) -- the pin is not limited to the four shipped scenes.
GPU: the CUDA path against the oracle, bit for bit (hits, hit records, LBVH)."""
import numpy as np
import pytest

from oracle_lib import Oracle, RefWorld, camera_derive, have_ref
from rrt_b200.types import SceneArrays, material_dtype, msphere_dtype, sphere_dtype, triangle_dtype


def random_scene(seed, n_sph=40, n_msph=6, n_tri=60):
    rng = np.random.default_rng(seed)
    f32 = lambda a: np.asarray(a, dtype=np.float32)
    mats = np.zeros(5, material_dtype)
    mats["type"] = [0, 0, 1, 1, 2]
    mats["albedo"] = f32(rng.uniform(0.1, 0.9, size=(5, 3)))
    mats["param"] = f32([0, 0, 0.0, 0.4, 1.5])
    s = np.zeros(n_sph, sphere_dtype)
    s["center"] = f32(rng.uniform(-6, 6, size=(n_sph, 3)))
    s["radius"] = f32(np.exp(rng.uniform(np.log(0.05), np.log(2.0), size=n_sph)))
    s["material"] = rng.integers(0, 5, size=n_sph)
    if n_sph:
        s["center"][0] = (0, -500.5, 0)  # a huge "ground" sphere: the cancellation-prone case
        s["radius"][0] = 500.0
    ms = np.zeros(n_msph, msphere_dtype)
    ms["center0"] = f32(rng.uniform(-5, 5, size=(n_msph, 3)))
    ms["center1"] = ms["center0"] + f32(rng.uniform(-1, 1, size=(n_msph, 3)))
    ms["time0"] = f32(rng.uniform(0.0, 0.2, size=n_msph))
    ms["time1"] = ms["time0"] + f32(rng.uniform(0.5, 2.0, size=n_msph))
    ms["radius"] = f32(rng.uniform(0.1, 0.8, size=n_msph))
    ms["material"] = rng.integers(0, 5, size=n_msph)
    t = np.zeros(n_tri, triangle_dtype)
    base = rng.uniform(-5, 5, size=(n_tri, 3))
    t["v0"] = f32(base)
    t["v1"] = f32(base + rng.normal(scale=0.8, size=(n_tri, 3)))
    t["v2"] = f32(base + rng.normal(scale=0.8, size=(n_tri, 3)))
    t["material"] = rng.integers(0, 5, size=n_tri)
    cam = camera_derive((0, 2, 12), (0, 0, 0), (0, 1, 0), 40.0, 1.5, 0.1, 12.0, 0.0, 1.0 if n_msph else 0.0)
    return SceneArrays(cam, mats, s, ms, t)


def random_rays(seed, n, scene):
    rng = np.random.default_rng(seed + 1000)
    r = np.zeros((n, 7), np.float32)
    r[:, 0:3] = rng.uniform(-9, 9, size=(n, 3))
    target = rng.uniform(-5, 5, size=(n, 3))
    r[:, 3:6] = (target - r[:, 0:3]) * rng.uniform(0.2, 3.0, size=(n, 1))  # un-normalised directions, like camera rays
    k = min(1, len(scene.spheres) - 1)
    r[: n // 8, 0:3] = scene.spheres["center"][k] + 0.3 * scene.spheres["radius"][k]  # origins inside a sphere
    r[:, 6] = rng.uniform(scene.camera["time0"][0], max(scene.camera["time1"][0], 1e-3), size=n)
    return r


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed", range(6))
def test_oracle_vs_reference_double_on_random_scenes(seed):
    scene = random_scene(seed)
    rays = random_rays(seed, 20000, scene)
    ref_id, ref_t = RefWorld(scene, "d").trace_scan(rays)
    orc = Oracle(scene)
    for mode in ("scan", "bvh"):
        ids, t = orc.trace(rays, 0.001, mode)
        same = ids == ref_id
        m = same & (ref_id >= 0)
        rel = np.abs(t[m].astype(np.float64) - ref_t[m]) / np.abs(ref_t[m])
        ok = same.copy()
        ok[m] &= rel <= 1e-5
        assert ok.mean() >= 0.9999, (seed, mode, ok.mean(), int((~same).sum()), rel.max())
    assert (ref_id >= 0).mean() > 0.3


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(6))
def test_gpu_vs_oracle_on_random_scenes(ctx, seed):
    n = (40, 1, 300, 2, 77, 500)[seed]
    scene = random_scene(100 + seed, n_sph=n, n_msph=seed % 3 * 4, n_tri=(60, 0, 400, 1, 33, 1000)[seed])
    rays = random_rays(seed, 30000, scene)
    ctx.set_scene(scene, use_bvh=True)
    orc = Oracle(scene)
    got, want = ctx.bvh_arrays(), orc.bvh_arrays()
    for k in ("morton", "perm", "left", "right", "parent"):
        assert np.array_equal(got[k], want[k]), (seed, k)
    assert got["node_box"].tobytes() == want["node_box"].tobytes()
    a = ctx.trace(rays, 0.001, "bvh", want_rec=True)
    o = orc.trace(rays, 0.001, "bvh", want_rec=True)
    for x, y in zip(a, o):
        assert x.tobytes() == y.tobytes()
    b = ctx.trace(rays[:4000], 0.001, "scan")
    assert np.array_equal(b[0], a[0][:4000]) and b[1].tobytes() == a[1][:4000].tobytes()
    img, st = ctx.render(64, 40, 4, 50, seed=seed, count_rays=True)
    ref, _, cnt = orc.render(64, 40, 4, 50, seed)
    assert st["paths"] == cnt["paths"] and abs(st["rays"] - cnt["rays"]) <= 0.02 * cnt["rays"] + 8
    assert np.mean(np.abs(np.sqrt(img / 4) - np.sqrt(ref / 4)).max(axis=2) < 2e-3) > 0.95
