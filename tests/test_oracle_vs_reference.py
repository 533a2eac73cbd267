"""-m "not gpu": the CPU oracle against the LIVE reference (oracle/_ref/libref_{f,d}.so, built by
oracle/Makefile from /root/reference) at BASELINE.json's full image sizes.  Skipped where oracle/_ref has
not been built (then the committed fixtures of test_oracle_golden.py carry the pin)."""
import numpy as np
import pytest

from oracle_lib import Oracle, RefScene, RefWorld, have_ref, pinhole_rays, ref_scene_path

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")

CONFIG = {"test1": (1200, 800), "test2": (1920, 1080), "test3": (1920, 1080), "final": (1200, 800)}


@pytest.fixture(scope="module", params=list(CONFIG))
def live(request):
    name = request.param
    W, H = CONFIG[name]
    rs = RefScene(ref_scene_path(name + ".txt"), W, H, "f")
    return name, W, H, rs, rs.arrays()


def test_full_size_primary_hits(live):
    """Every pixel-centre ray of the configured image: ids equal to the reference's double-precision
    hittable_list scan on >= 99.99 %, t within 1e-5 relative (the reference's own float build manages
    97.3 % on final.txt -- SURVEY 7 hard part 1)."""
    name, W, H, rs, scene = live
    rays = pinhole_rays(scene, W, H)
    if scene.camera["time0"][0] != scene.camera["time1"][0]:
        rays[:, 6] = np.random.default_rng(1).uniform(scene.camera["time0"][0], scene.camera["time1"][0], len(rays)).astype(np.float32)
    ref_id, ref_t = RefWorld(scene, "d").trace_scan(rays)
    orc = Oracle(scene)
    ids, t = orc.trace(rays, 0.001, "bvh")
    same = ids == ref_id
    m = same & (ref_id >= 0)
    rel = np.abs(t[m].astype(np.float64) - ref_t[m]) / np.abs(ref_t[m])
    ok = same.copy()
    ok[m] &= rel <= 1e-5
    assert ok.mean() >= 0.9999, (name, ok.mean(), rel.max())


def test_bvh_world_hit_equals_list_hit_in_reference(live):
    """Sanity of the oracle's premise: the reference's own bvh_node::hit returns what its list scan returns."""
    name, W, H, rs, scene = live
    rays = pinhole_rays(scene, W, H, step=9)
    w = RefWorld(scene, "d")
    a = w.trace_world(w.list, rays)
    b = w.trace_world(w.bvh(), rays)
    assert np.array_equal(a[:, 8], b[:, 8])
    assert np.allclose(a[:, 0], b[:, 0], rtol=1e-12)


def test_estimator_mean_radiance(live):
    """Mean radiance of ONE explicit primary ray: the reference's ray_color (rrt.cpp:25-52, its own mt19937
    + rejection samplers, double build) vs the oracle's estimator (Philox + direct samplers).  Equal within
    Monte-Carlo error: this is what catches a biased material / sky / depth / t_min rule."""
    import ctypes as C

    name, W, H, rs, scene = live
    orc = Oracle(scene)
    wd = RefWorld(scene, "d")
    rng = np.random.default_rng(5)
    n_rays, n_samp = 8, 4000
    pix = rng.integers(0, W * H, size=n_rays)
    out = np.zeros(3, np.float32)
    for p in pix:
        ray = orc.camera_rays(W, H, [p], 0, 99)[0]
        r64 = ray.astype(np.float64)
        ref_s = np.stack([wd.ray_color_mean(wd.list, r64, 50, 1) for _ in range(n_samp)])
        ours = np.zeros((n_samp, 3))
        for s in range(n_samp):
            orc.lib.orc_radiance(C.byref(orc._s), orc.bvh(), C.c_void_p(ray.ctypes.data), int(p), s, 50, C.c_uint64(1234),
                                 C.c_void_p(out.ctypes.data), None)
            ours[s] = out
        se = np.sqrt(ref_s.var(axis=0) / n_samp + ours.var(axis=0) / n_samp)
        assert np.all(np.abs(ours.mean(axis=0) - ref_s.mean(axis=0)) < 5 * se + 2e-3), (name, int(p), ours.mean(axis=0), ref_s.mean(axis=0), se)


def test_reference_renders_psnr_small(live, tmp_path):
    """Image-level parity on the CPU at a size that runs in seconds: the oracle's estimator vs the
    reference binary rrto (double).  The bar scales with spp (two independent reference
    renders are 26.6 dB apart at 10 spp on final, ~+10 dB per decade); here both sides are far above the
    noise floor only in the mean, so assert mean agreement + a PSNR consistent with pure noise."""
    import os
    import subprocess

    from PIL import Image

    from oracle_lib import REF_DIR, psnr
    from rrt_b200 import Scene

    name, W, H, rs, scene = live
    w, h, spp = 150, 100, 32
    out = tmp_path / "ref.png"
    subprocess.check_call([os.path.join(REF_DIR, "rrto"), "-i", ref_scene_path(name + ".txt"), "-w", str(w), "-h", str(h), "-s", str(spp), "-o", str(out)],
                          stderr=subprocess.DEVNULL, env=dict(os.environ, OMP_NUM_THREADS="1"))
    ref = np.asarray(Image.open(out).convert("RGB")).astype(np.float64)
    sc = Scene.from_file(ref_scene_path(name + ".txt"), w, h).arrays
    orc = Oracle(sc)
    img, _, _ = orc.render(w, h, spp, 50, 1984)
    ours = orc.tonemap(img, spp).astype(np.float64)
    assert abs(ours.mean() - ref.mean()) < 0.8, (name, ours.mean(), ref.mean())
    assert psnr(ours, ref) > 29.0, (name, psnr(ours, ref))
