"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against
(1) the CPU oracle (oracle/liboracle.so) on the same seeded inputs and (2) the committed golden
fixtures produced by the unmodified reference (tests/golden, tools/make_golden.py).

Bars (BASELINE.json north_star):
  * Philox stream, Morton codes, sort order, LBVH topology and boxes: BIT-EXACT vs the oracle
  * primary hits: object id equal to the reference hittable_list scan on >= 99.99 % of rays and
    t within 1e-5 relative of the reference's double-precision result
  * images: PSNR >= 40 dB vs a reference render at the same spp
"""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLD, ROOT, load_golden
from oracle_lib import Oracle, philox as orc_philox, pinhole_rays, psnr

pytestmark = pytest.mark.gpu


def test_philox_bit_exact(ctx):
    rng = np.random.default_rng(7)
    ctr = rng.integers(0, 2**32, size=(4096, 4), dtype=np.uint64).astype(np.uint32)
    ctr[0] = 0
    ctr[1] = 0xFFFFFFFF
    for k0, k1 in ((0, 0), (0xFFFFFFFF, 0xFFFFFFFF), (1984, 0), (0xA4093822, 0x299F31D0)):
        got = ctx.philox(ctr, k0, k1)
        want = orc_philox(ctr, k0, k1)
        assert np.array_equal(got, want)
    # Random123 known-answer vectors (philox4x32-10)
    kat = ctx.philox(np.array([[0, 0, 0, 0]], np.uint32), 0, 0)[0]
    assert [hex(x) for x in kat] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]


def test_lbvh_bit_exact(ctx, golden):
    name, scene, d = golden
    ctx.set_scene(scene, use_bvh=True)
    got = ctx.bvh_arrays()
    want = Oracle(scene).bvh_arrays()
    for k in ("morton", "perm", "left", "right", "parent"):
        assert np.array_equal(got[k], want[k]), (name, k)
    for k in ("prim_box", "node_box"):
        assert got[k].tobytes() == want[k].tobytes(), (name, k)
    # per-primitive boxes equal the reference's bounding_box() (float build)
    assert got["prim_box"].tobytes() == d["ref_boxes"].astype(np.float32).tobytes()


def _check_against_reference(ids, t, d, name):
    ref_id, ref_t = d["ref_id"], d["ref_t"]
    same = ids == ref_id
    hit = ref_id >= 0
    rel = np.zeros(len(ids))
    m = hit & same
    rel[m] = np.abs(t[m].astype(np.float64) - ref_t[m]) / np.abs(ref_t[m])
    ok = same & (rel <= 1e-5)
    assert ok.mean() >= 0.9999, (name, ok.mean(), int((~same).sum()), rel.max())


@pytest.mark.parametrize("mode", ["scan", "bvh"])
def test_primary_hits_vs_reference_and_oracle(ctx, golden, mode):
    name, scene, d = golden
    ctx.set_scene(scene, use_bvh=True)
    rays = d["rays"]
    ids, t, rec = ctx.trace(rays, 0.001, mode, want_rec=True)
    _check_against_reference(ids, t, d, name)
    # hit record vs the reference's double-precision record
    hit = (ids >= 0) & (ids == d["ref_id"])
    ref_rec = d["ref_rec"]
    scale = np.maximum(1.0, np.abs(ref_rec[hit, 1:4]).max(axis=1, keepdims=True))
    assert np.max(np.abs(rec[hit, 0:3] - ref_rec[hit, 1:4]) / scale) < 2e-5
    assert np.max(np.abs(rec[hit, 3:6] - ref_rec[hit, 4:7])) < 2e-3  # normals of the r=100/1000 spheres amplify p error
    assert np.array_equal(rec[hit, 6] != 0, ref_rec[hit, 7] != 0)
    # and BIT-EXACT against the oracle (same op sequence)
    o_ids, o_t, o_rec = Oracle(scene).trace(rays, 0.001, mode, want_rec=True)
    assert np.array_equal(ids, o_ids)
    assert t.tobytes() == o_t.tobytes()
    assert rec.tobytes() == o_rec.tobytes()


def test_full_size_primary_rays_scan_equals_bvh(ctx, golden):
    """BASELINE.json sizes: every pixel-centre ray of the configured image; the LBVH must return exactly
    what the flat scan returns (size-independent property), and both must equal the oracle."""
    name, scene, d = golden
    W, H = int(d["W"]), int(d["H"])
    ctx.set_scene(scene, use_bvh=True)
    rays = pinhole_rays(scene, W, H)
    if scene.camera["time0"][0] != scene.camera["time1"][0]:
        rays[:, 6] = np.random.default_rng(3).uniform(scene.camera["time0"][0], scene.camera["time1"][0], len(rays)).astype(np.float32)
    i_b, t_b = ctx.trace(rays, 0.001, "bvh")
    i_s, t_s = ctx.trace(rays, 0.001, "scan")
    assert np.array_equal(i_b, i_s) and t_b.tobytes() == t_s.tobytes()
    o_i, o_t = Oracle(scene).trace(rays, 0.001, "bvh")
    assert np.array_equal(i_b, o_i) and t_b.tobytes() == o_t.tobytes()


def test_camera_rays_bit_exact(ctx, golden):
    name, scene, d = golden
    W, H = int(d["W"]), int(d["H"])
    ctx.set_scene(scene)
    rng = np.random.default_rng(11)
    pix = rng.integers(0, W * H, size=2000).astype(np.int32)
    pix[:4] = [0, W - 1, W * (H - 1), W * H - 1]
    for sample in (0, 7, 499):
        got = ctx.camera_rays(W, H, pix, sample, seed=1984)
        want = Oracle(scene).camera_rays(W, H, pix, sample, 1984)
        assert got.tobytes() == want.tobytes(), (name, sample)


def test_scatter_matches_oracle(ctx, golden):
    name, scene, d = golden
    ctx.set_scene(scene)
    rng = np.random.default_rng(5)
    n = 4000
    nm = len(scene.materials)
    d_in = rng.normal(size=(n, 3)).astype(np.float32) * rng.uniform(0.2, 8, size=(n, 1)).astype(np.float32)
    nrm = rng.normal(size=(n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    flip = np.sum(d_in * nrm, axis=1) > 0
    nrm[flip] *= -1
    in16 = np.zeros((n, 16), np.float32)
    in16[:, 0:3] = rng.uniform(-3, 3, size=(n, 3))
    in16[:, 3:6] = d_in
    in16[:, 7:10] = rng.uniform(-3, 3, size=(n, 3))
    in16[:, 10:13] = nrm
    in16[:, 13] = rng.integers(0, 2, size=n)
    in16[:, 14] = rng.integers(0, nm, size=n)
    rnd = rng.integers(0, 2**32, size=(n, 4), dtype=np.uint64).astype(np.uint32)
    got = ctx.scatter(in16, rnd)
    want = Oracle(scene).scatter(in16, rnd)
    # the scatter code may use rsqrt / fast division: tolerance, not bitwise
    dn = np.abs(want[:, 0:3]).max(axis=1) + 1e-3
    assert np.max(np.abs(got[:, 0:3] - want[:, 0:3]).max(axis=1) / dn) < 5e-5
    assert np.array_equal(got[:, 3:6], want[:, 3:6])
    # absorbed/continued may flip only when dot(scattered, n) is within rounding of zero
    diff = got[:, 6] != want[:, 6]
    if diff.any():
        assert np.all(np.abs(np.sum(want[diff, 0:3] * in16[diff, 10:13], axis=1)) < 1e-5)


def test_render_matches_oracle_small(ctx, golden):
    """Same Philox streams on both sides: the images agree except where a one-ulp difference in a
    scatter direction sends a path elsewhere."""
    name, scene, d = golden
    W, H, spp = 96, 64, 8
    # the golden camera was derived for the config's aspect ratio; W/H here keeps 3:2 or 16:9 closely enough
    ctx.set_scene(scene, use_bvh=True)
    img, st = ctx.render(W, H, spp, 50, seed=1984, count_rays=True)
    ref, fixed, cnt = Oracle(scene).render(W, H, spp, 50, 1984)
    assert st["paths"] == W * H * spp == cnt["paths"]
    assert abs(st["rays"] - cnt["rays"]) / cnt["rays"] < 0.01
    a = np.sqrt(img / spp).clip(0, 1)
    b = np.sqrt(ref / spp).clip(0, 1)
    close = np.abs(a - b).max(axis=2) < 1e-3
    assert close.mean() > 0.97, (name, close.mean())
    assert abs(a.mean() - b.mean()) < 2e-3
    # flat-scan render is bit-identical to the LBVH render (closest hit is traversal-order independent)
    ctx.set_scene(scene, use_bvh=False)
    img2, _ = ctx.render(W, H, spp, 50, seed=1984)
    assert img2.tobytes() == img.tobytes()


def test_render_deterministic_and_shardable(ctx, golden):
    """Bit-identical image for repeated runs, for tile-sharding and for sample-sharding (any world size):
    the accumulators are integers, the RNG is keyed by (pixel, sample, bounce)."""
    name, scene, d = golden
    W, H, spp = 100, 52, 6  # deliberately not a multiple of the 8x4 tile
    ctx.set_scene(scene, use_bvh=True)
    full, _ = ctx.render(W, H, spp, 50, seed=42)
    again, _ = ctx.render(W, H, spp, 50, seed=42)
    assert full.tobytes() == again.tobytes()
    other, _ = ctx.render(W, H, spp, 50, seed=43)
    assert other.tobytes() != full.tobytes()
    import torch

    for mode in (0, 1):
        for world in (2, 3, 8):
            acc = torch.zeros(H * W * 3, dtype=torch.int64, device="cuda")
            paths = 0
            for rank in range(world):
                p = ctx.params(W, H, spp, 50, 42, rank, world, mode)
                st = ctx.render_device(p, acc.data_ptr())
                paths += st["paths"]
            assert paths == W * H * spp
            out = torch.empty(H * W * 3, dtype=torch.float32, device="cuda")
            ctx.resolve_device(acc.data_ptr(), out.data_ptr(), out.numel())
            assert out.cpu().numpy().tobytes() == full.tobytes(), (name, mode, world)


GOLDEN_RENDERS = {
    "final": ("final_600x400_s500_rrto.png", 600, 400, 500),
    "test1": ("test1_480x320_s256_rrto.png", 480, 320, 256),
    "test2": ("test2_480x270_s256_rrto.png", 480, 270, 256),
    "test3": ("test3_480x270_s256_rrto.png", 480, 270, 256),
}


def _scene_for(name, W, H):
    """Re-derive the camera for this aspect ratio through the product's own parser when the scene text is
    available (oracle/_ref/scenes travels to the GPU box); else fall back to the golden arrays (the golden
    configs share the aspect ratio of these renders)."""
    from oracle_lib import ref_scene_path
    from rrt_b200 import Scene

    p = ref_scene_path(name + ".txt")
    if p:
        return Scene.from_file(p, W, H).arrays
    return load_golden(name)[0]


@pytest.mark.parametrize("name", list(GOLDEN_RENDERS))
def test_psnr_vs_reference_render(ctx, name):
    """PSNR >= 40 dB against the reference's own double-precision render (rrto, deterministic single
    thread) of the same scene at the same spp, on the 8-bit gamma-encoded images (color.h:8-23)."""
    from PIL import Image

    from rrt_b200 import tonemap

    fn, W, H, spp = GOLDEN_RENDERS[name]
    ref = np.asarray(Image.open(os.path.join(GOLD, fn)).convert("RGB"))
    ctx.set_scene(_scene_for(name, W, H), use_bvh=True)
    img, st = ctx.render(W, H, spp, 50, seed=1984)
    ours = tonemap(img, spp)
    val = psnr(ours, ref)
    # two independent unbiased renders at this spp sit at ~43-44 dB; 40 is the bar
    assert val >= 40.0, (name, val)
    assert abs(ours.astype(np.float64).mean() - ref.astype(np.float64).mean()) < 0.5


def test_cli_drop_in(tmp_path, built_lib):
    """The drop-in executable: reference flags, PNG out, stats line on stderr, same pixels as the API."""
    from PIL import Image

    from oracle_lib import ref_scene_path
    from rrt_b200 import Context, Scene, tonemap

    exe = os.path.join(ROOT, "rrt_b200", "bin", "rrt")
    scene_path = ref_scene_path("test1.txt")
    if not (os.path.exists(exe) and scene_path):
        pytest.skip("drop-in executable or scene text not staged")
    out = tmp_path / "t.png"
    r = subprocess.run([exe, "-i", scene_path, "-o", str(out), "-w", "120", "-h", "80", "-s", "4", "-d", "50", "-tx", "16", "-ty", "16"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "stats," in r.stderr and "took " in r.stderr and "sphere count:    4" in r.stderr
    png = np.asarray(Image.open(out).convert("RGB"))
    with Context(0) as c:
        c.set_scene(Scene.from_file(scene_path, 120, 80))
        img, _ = c.render(120, 80, 4, 50, seed=1984)
    assert np.array_equal(png, tonemap(img, 4))
    # PPM on stdout without -o (main.cpp:140-149)
    r = subprocess.run([exe, "-i", scene_path, "-w", "16", "-h", "8", "-s", "1"], capture_output=True, text=True, timeout=300)
    lines = r.stdout.split("\n")
    assert lines[0] == "P3" and lines[1] == "16 8" and lines[2] == "255" and len(lines) >= 3 + 16 * 8


@pytest.mark.slow
@pytest.mark.parametrize("name,W,H,spp", [("final", 1200, 800, 500), ("test2", 1920, 1080, 256), ("test3", 1920, 1080, 256)])
def test_psnr_vs_rrtd_full_size_live(ctx, tmp_path, name, W, H, spp):
    """The north-star bar verbatim, at the sizes of BASELINE.json configs[1..3]: PSNR >= 40 dB against the reference
    `rrtd` (rrt.cu, double, rebuilt for sm_100a by oracle/Makefile) render of the same scene at the same size and spp,
    executed on this very box: scenes/final.txt 1200x800 / 500 spp, the triangle scene and the motion-blur scene at
    1920x1080 / 256 spp."""
    from PIL import Image

    from oracle_lib import REF_DIR, ref_scene_path
    from rrt_b200 import Scene, tonemap

    exe = os.path.join(REF_DIR, "rrtd")
    scene_path = ref_scene_path(name + ".txt")
    if not (os.path.exists(exe) and scene_path):
        pytest.skip("oracle/_ref/rrtd or the scene text is not staged")
    out = tmp_path / "rrtd.png"
    r = subprocess.run([exe, "-i", scene_path, "-w", str(W), "-h", str(H), "-s", str(spp), "-d", "50", "-o", str(out)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-500:]
    ref = np.asarray(Image.open(out).convert("RGB"))
    ctx.set_scene(Scene.from_file(scene_path, W, H), use_bvh=True)
    img, st = ctx.render(W, H, spp, 50, seed=1984)
    ours = tonemap(img, spp)
    val = psnr(ours, ref)
    print("PSNR vs rrtd (%s.txt %dx%d %d spp): %.2f dB; rrtd stats: %s" % (name, W, H, spp, val, [l for l in r.stderr.splitlines() if l.startswith("stats,")]))
    assert val >= 40.0, (name, val)
    assert abs(ours.astype(np.float64).mean() - ref.astype(np.float64).mean()) < 0.5
