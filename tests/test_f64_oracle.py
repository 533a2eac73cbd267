"""-m "not gpu": the DOUBLE-integrator oracle (oracle/rrt_oracle_f64.c) pinned against the live
reference's double build (oracle/_ref/libref_d.so) and against the float oracle.  The reference evaluates its
sphere roots as (-half_b -+ sqrt(disc)) / a (sphere.h:41-48); the oracle uses the cancellation-free pair, so `t`
agrees to a few ulps of double, not bitwise -- the bar below is 1e-9 relative."""
import ctypes as C

import numpy as np
import pytest

from oracle_lib import Oracle, RefScene, RefWorld, have_ref, philox, pinhole_rays, ref_scene_path

CONFIG = {"test1": (1200, 800), "test2": (1920, 1080), "test3": (1920, 1080), "final": (1200, 800)}

needs_ref = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")


@pytest.fixture(scope="module", params=list(CONFIG))
def live(request):
    name = request.param
    W, H = CONFIG[name]
    rs = RefScene(ref_scene_path(name + ".txt"), W, H, "f")
    return name, W, H, rs.arrays()


@needs_ref
def test_f64_primary_hits_match_reference_double(live):
    """Every 3rd pixel-centre ray at the configured size: ids identical to the reference's double list scan, t and
    the hit record within 1e-9."""
    name, W, H, scene = live
    rays = pinhole_rays(scene, W, H, step=3).astype(np.float64)
    if scene.camera["time0"][0] != scene.camera["time1"][0]:
        rays[:, 6] = np.random.default_rng(1).uniform(scene.camera["time0"][0], scene.camera["time1"][0], len(rays))
    w = RefWorld(scene, "d")
    ref_id, ref_t = w.trace_scan(rays)
    orc = Oracle(scene)
    for mode in ("scan", "bvh"):
        ids, t, rec = orc.trace_f64(rays, 0.001, mode, want_rec=True)
        assert np.array_equal(ids, ref_id), (name, mode, (ids != ref_id).sum())
        m = ref_id >= 0
        rel = np.abs(t[m] - ref_t[m]) / np.abs(ref_t[m])
        assert rel.max() <= 1e-9, (name, mode, rel.max())
    # hit records (p, n, front) on a subsample, vs hittable_list::hit's record
    sub = rays[::97]
    r10 = w.trace_world(w.list, sub)  # t, p(3), n(3), front, id, material
    ids, t, rec = orc.trace_f64(sub, 0.001, "bvh", want_rec=True)
    m = ids >= 0
    assert np.allclose(rec[m, 0:3], r10[m, 1:4], rtol=1e-9, atol=1e-9)
    # triangle normals are the float-rounded stored ones: 1e-6; sphere normals are double
    assert np.allclose(rec[m, 3:6], r10[m, 4:7], rtol=0, atol=2e-6)
    assert np.array_equal(rec[m, 6] != 0, r10[m, 7] != 0)


def test_f64_primary_hits_match_reference_fixture(golden):
    """The committed fixtures (tests/golden, produced by the reference's double build, tools/make_golden.py) pin the
    double oracle wherever oracle/_ref is absent: ids identical, t and the hit point within 1e-9, normals of spheres
    within 1e-9 and of triangles within the float rounding of the stored normal."""
    name, scene, d = golden
    orc = Oracle(scene)
    rays = d["rays"].astype(np.float64)
    hit = d["ref_id"] >= 0
    for mode in ("scan", "bvh"):
        ids, t, rec = orc.trace_f64(rays, 0.001, mode, want_rec=True)
        assert np.array_equal(ids, d["ref_id"]), (name, mode, int((ids != d["ref_id"]).sum()))
        rel = np.abs(t[hit] - d["ref_t"][hit]) / np.abs(d["ref_t"][hit])
        assert rel.max() <= 1e-9, (name, mode, rel.max())
        assert np.allclose(rec[hit, 0:3], d["ref_rec"][hit, 1:4], rtol=1e-9, atol=1e-9)
        assert np.allclose(rec[hit, 3:6], d["ref_rec"][hit, 4:7], rtol=0, atol=2e-6)
        assert np.array_equal(rec[hit, 6] != 0, d["ref_rec"][hit, 7] != 0)


def test_f64_agrees_with_float_oracle():
    """The two integrators see the same scene: ids equal, t within the float policy's 2.2e-6."""
    from rrt_b200.synthetic import synthetic_scene_text
    from rrt_b200 import Scene
    import tempfile, os

    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "s.txt")
        open(p, "w").write(synthetic_scene_text(n_spheres=300, ico_level=2, seed=3))
        scene = Scene.from_file(p, 320, 200).arrays
    rays = pinhole_rays(scene, 320, 200)
    orc = Oracle(scene)
    idf, tf = orc.trace(rays, 0.001, "bvh")
    idd, td = orc.trace_f64(rays.astype(np.float64), 0.001, "bvh")
    ids_s, ts = orc.trace_f64(rays.astype(np.float64), 0.001, "scan")
    assert np.array_equal(idd, ids_s) and np.array_equal(td, ts)
    same = idf == idd
    assert same.mean() >= 0.9999
    m = same & (idd >= 0)
    assert (np.abs(tf[m] - td[m]) / np.abs(td[m])).max() <= 4e-6


def test_f64_camera_rays_close_to_float():
    from conftest import load_golden

    scene, _ = load_golden("final")
    orc = Oracle(scene)
    pix = np.random.default_rng(2).integers(0, 1200 * 800, 200)
    a = orc.camera_rays(1200, 800, pix, 3, 77)
    b = orc.camera_rays_f64(1200, 800, pix, 3, 77)
    assert np.allclose(a, b, rtol=0, atol=3e-5)


@needs_ref
def test_f64_estimator_mean_radiance(live):
    """Mean radiance of one explicit primary ray: reference ray_color (double build) vs the f64 oracle's paths."""
    name, W, H, scene = live
    orc = Oracle(scene)
    wd = RefWorld(scene, "d")
    n_samp = 3000
    pix = np.random.default_rng(8).integers(0, W * H, size=4)
    for p in pix:
        ray = orc.camera_rays_f64(W, H, [p], 0, 99)[0]
        ref_s = np.stack([wd.ray_color_mean(wd.list, ray, 50, 1) for _ in range(n_samp)])
        ours = _radiance_samples(orc, ray, int(p), n_samp)
        se = np.sqrt(ref_s.var(axis=0) / n_samp + ours.var(axis=0) / n_samp)
        assert np.all(np.abs(ours.mean(axis=0) - ref_s.mean(axis=0)) < 5 * se + 2e-3), (name, int(p), ours.mean(axis=0), ref_s.mean(axis=0), se)


def _radiance_samples(orc, ray7, pixel, n):
    """ray_color over the f64 oracle's trace + scatter pieces (python loop over bounces, vectorised over samples)."""
    scene = orc.scene
    ns, nms = len(scene.spheres), len(scene.mspheres)
    mats = np.concatenate([scene.spheres["material"], scene.mspheres["material"], scene.triangles["material"]]).astype(np.int64)
    rays = np.tile(ray7, (n, 1))
    thr = np.ones((n, 3))
    out = np.zeros((n, 3))
    alive = np.arange(n)
    for b in range(50):
        if len(alive) == 0:
            break
        ids, t, rec = orc.trace_f64(rays[alive], 0.001, "bvh", want_rec=True)
        miss = ids < 0
        d = rays[alive][miss, 3:6]
        uy = d[:, 1] / np.linalg.norm(d, axis=1)
        tt = 0.5 * (uy + 1.0)
        sky = (1.0 - tt)[:, None] + tt[:, None] * np.array([0.5, 0.7, 1.0])
        out[alive[miss]] = thr[alive[miss]] * sky
        hit = ~miss
        a = alive[hit]
        if len(a) == 0:
            break
        in16 = np.zeros((len(a), 16))
        in16[:, 0:7] = rays[a]
        in16[:, 7:14] = rec[hit]
        in16[:, 14] = mats[ids[hit]]
        rnd = philox(np.stack([[pixel, int(s), 2 + b, 0] for s in a]), 1234, 0)
        o8 = orc.scatter_f64(in16, rnd)
        ok = o8[:, 6] != 0
        thr[a] *= o8[:, 3:6]
        rays[a, 0:3] = rec[hit, 0:3]
        rays[a, 3:6] = o8[:, 0:3]
        alive = a[ok]
    return out
