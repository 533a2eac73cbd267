"""-m "not gpu": the CPU oracle (oracle/rrt_oracle.c) against the committed fixtures that were produced by
EXECUTING the unmodified reference (tools/make_golden.py).  This is what pins the oracle on machines
without /root/reference."""
import os

import numpy as np
import pytest

from conftest import GOLD
from oracle_lib import Oracle, camera_derive, philox, sincos2pi


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF), (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, want in kat:
        got = philox(np.array([ctr], np.uint32), *key)[0]
        assert tuple(int(x) for x in got) == want


def test_sincos_polynomial_accuracy():
    for u in np.linspace(0, 1, 4097, endpoint=False):
        c, s = sincos2pi(float(np.float32(u)))
        a = 2 * np.pi * (float(np.float32(u)) - 0.5)
        assert abs(c - np.cos(a)) < 3e-7 and abs(s - np.sin(a)) < 3e-7


def test_primary_hits_match_reference_fixture(golden):
    """id equal to the reference hittable_list scan on >= 99.99 % of rays, t within 1e-5 relative of the
    reference's double-precision value (north_star bar), for the flat scan and through the LBVH."""
    name, scene, d = golden
    orc = Oracle(scene)
    for mode in ("scan", "bvh"):
        ids, t, rec = orc.trace(d["rays"], 0.001, mode, want_rec=True)
        same = ids == d["ref_id"]
        hit = d["ref_id"] >= 0
        m = hit & same
        rel = np.abs(t[m].astype(np.float64) - d["ref_t"][m]) / np.abs(d["ref_t"][m])
        assert same.mean() >= 0.9999 and (rel <= 1e-5).mean() >= 0.9999, (name, mode, same.mean(), rel.max())
        assert np.array_equal(rec[m, 6] != 0, d["ref_rec"][m, 7] != 0)


def test_scan_and_bvh_agree_exactly(golden):
    name, scene, d = golden
    orc = Oracle(scene)
    a = orc.trace(d["rays"], 0.001, "scan", want_rec=True)
    b = orc.trace(d["rays"], 0.001, "bvh", want_rec=True)
    for x, y in zip(a, b):
        assert x.tobytes() == y.tobytes()


def test_prim_boxes_equal_reference_bounding_box(golden):
    name, scene, d = golden
    arr = Oracle(scene).bvh_arrays()
    assert arr["prim_box"].tobytes() == d["ref_boxes"].astype(np.float32).tobytes()


def test_lbvh_structure(golden):
    """Topology invariants of the Karras tree: a proper binary tree over the Morton-sorted leaves whose
    boxes enclose their children."""
    name, scene, d = golden
    a = Oracle(scene).bvh_arrays()
    n = len(a["perm"])
    assert sorted(a["perm"].tolist()) == list(range(n))
    keys = (a["morton"][a["perm"]].astype(np.uint64) << np.uint64(32)) | a["perm"].astype(np.uint64)
    assert np.all(keys[1:] > keys[:-1])
    ni = n - 1
    seen_leaf, seen_node = np.zeros(n, int), np.zeros(ni, int)
    for i in range(ni):
        for ch in (a["left"][i], a["right"][i]):
            if ch >= 0:
                seen_node[ch] += 1
                cb = a["node_box"][ch]
                assert a["parent"][ch] == i
            else:
                seen_leaf[~ch] += 1
                cb = a["prim_box"][a["perm"][~ch]]
                assert a["parent"][ni + ~ch] == i
            assert np.all(a["node_box"][i][:3] <= cb[:3]) and np.all(a["node_box"][i][3:] >= cb[3:])
    assert np.all(seen_leaf == 1) and seen_node[0] == 0 and np.all(seen_node[1:] == 1)
    assert a["parent"][0] == -1


def test_math_kats():
    """reflect / refract / Schlick / tonemap / dielectric + metal scatter against reference outputs."""
    from rrt_b200.types import SceneArrays, camera_dtype, material_dtype, sphere_dtype

    k = np.load(os.path.join(GOLD, "kat_math.npz"))
    mats = np.zeros(2, material_dtype)
    mats["type"] = [2, 1]
    mats["param"] = [1.5, 0.0]
    mats["albedo"][1] = (0.8, 0.6, 0.2)
    sph = np.zeros(1, sphere_dtype)
    sph["radius"] = 1
    orc = Oracle(SceneArrays(np.zeros(1, camera_dtype), mats, sph))
    nd = len(k["sc_d_in"])
    in16 = np.zeros((nd, 16), np.float32)
    in16[:, 3:6] = k["sc_d_in"]
    in16[:, 7:10] = k["sc_p"]
    in16[:, 10:13] = k["sc_n"]
    in16[:, 13] = k["sc_front"]
    # dielectric: rnd word 0 = 0 forces "reflect" (reflectance > 0), 0xFFFFFFFF forces "refract unless TIR"
    for word, col in ((0, 0), (0xFFFFFF00, 1)):
        rnd = np.full((nd, 4), word, np.uint32)
        in16[:, 14] = 0
        out = orc.scatter(in16, rnd)
        want = k["die_dirs"][:, col]
        have = ~np.isnan(want[:, 0])
        if col == 1:  # where the reference never refracted (TIR) we reflect as well
            tir = np.isnan(want[:, 0])
            assert np.allclose(out[tir, 0:3], k["die_dirs"][tir, 0], atol=2e-6)
        assert np.allclose(out[have, 0:3], want[have], atol=2e-6)
        assert np.all(out[:, 3:6] == 1.0) and np.all(out[:, 6] == 1.0)
    # Schlick probability: fraction of reflections over many uniform draws ~ the reference's frequency
    rng = np.random.default_rng(3)
    for i in range(0, nd, 6):
        rnd = rng.integers(0, 2**32, size=(3000, 4), dtype=np.uint64).astype(np.uint32)
        rep = np.repeat(in16[i : i + 1], 3000, axis=0)
        out = orc.scatter(rep, rnd)
        refl = np.sum(out[:, 0:3] * in16[i, 10:13], axis=1) > 0
        assert abs(refl.mean() - k["die_reflect_freq"][i]) < 0.04
    # metal with fuzz 0 = mirror reflection of the unit direction (material.h:52)
    in16[:, 14] = 1
    in16[:, 13] = 1
    out = orc.scatter(in16, np.zeros((nd, 4), np.uint32))
    assert np.allclose(out[:, 0:3], k["met_dir"], atol=2e-6)
    assert np.array_equal(out[:, 6] != 0, k["met_ok"] != 0)
    # tonemap (color.h:8-23)
    sums = k["tm_sums"]
    img = orc.tonemap(sums.reshape(1, -1, 3), int(k["tm_spp"]))
    assert np.array_equal(img.reshape(-1, 3).astype(np.int32), k["tm_rgb"])


def test_camera_derivation_matches_reference(golden):
    """camera.h:8-29: our derivation from the camera line equals the reference's stored camera."""
    name, scene, d = golden
    lines = {"test1": ((0, 2, 5), (0, 0, -1), (0, 1, 0), 30.0, 0.1, 6.0, 0.0, 0.0),
             "test2": ((-1, 2, 5), (0, 0.5, -1), (0, 1, 0), 30.0, 0.1, 6.0, 0.0, 0.0),
             "test3": ((0, 2, 5), (0, 0, -1), (0, 1, 0), 30.0, 0.1, 6.0, 0.0, 0.5),
             "final": ((13, 2, 3), (0, 0, 0), (0, 1, 0), 30.0, 0.1, 10.0, 0.0, 0.0)}[name]
    W, H = int(d["W"]), int(d["H"])
    cam = camera_derive(lines[0], lines[1], lines[2], lines[3], float(np.float32(W / H)), lines[4], lines[5], lines[6], lines[7])
    for f in cam.dtype.names:
        assert np.allclose(cam[f], scene.camera[f], rtol=3e-7, atol=1e-7), f


def test_oracle_render_sharding_is_exact():
    """The fixed-point accumulators make tile- and sample-sharded renders sum to the unsharded image bit for
    bit (the property the multi-GPU path relies on)."""
    from conftest import load_golden

    scene, _ = load_golden("test2")
    orc = Oracle(scene)
    W, H, spp = 40, 28, 3
    _, full, cnt = orc.render(W, H, spp, 50, 7)
    assert cnt["paths"] == W * H * spp
    for mode in (0, 1):
        for world in (2, 3):
            acc = np.zeros_like(full)
            for r in range(world):
                acc += orc.render(W, H, spp, 50, 7, rank=r, world=world, shard_mode=mode)[1]
            assert np.array_equal(acc, full)
