"""-m gpu: the boundary proved against the reference's OWN host program.

oracle/_ref/rrtc_dropin, rrt_dropin and rrtd_dropin are the reference's unmodified main.cpp (its CLI, scene.h parser,
color.h tonemap and stb_image_write PNG writer, compiled from /root/reference by oracle/Makefile `dropin`) linked with
the reference-side binding rrt_b200/host/reference_side/rrt_b200.cpp and librrtb200.so.  Their PNGs are compared with
those of this repository's standalone drop-in rrt_b200/bin/rrt{,d} (own CLI, own parser, own tonemap, zlib PNG writer).

The reference has TWO float camera flavours: camera.h:12 calls an unqualified tan(), which is C's double ::tan under g++
(rrtc) and CUDA's float overload under nvcc (rrt); three of the 24 camera floats differ in the last bit.  Our parser
follows the g++ flavour, so rrtc_dropin must match bit for bit and the nvcc-built ones almost everywhere."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

SCENES = [("final.txt", 240, 160, 8), ("test1.txt", 240, 160, 8), ("test2.txt", 256, 144, 8), ("test3.txt", 256, 144, 8)]


def _png(exe, scene, w, h, spp, out, extra=()):
    r = subprocess.run([exe, "-i", scene, "-w", str(w), "-h", str(h), "-s", str(spp), "-o", str(out)] + list(extra),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-800:]
    from PIL import Image

    return np.asarray(Image.open(out)), r.stderr


@pytest.mark.parametrize("name,w,h,spp", SCENES)
def test_reference_main_over_the_binding_writes_the_same_png(tmp_path, name, w, h, spp):
    from oracle_lib import ref_scene_path

    ref_exe = os.path.join(ROOT, "oracle", "_ref", "rrtc_dropin")
    our_exe = os.path.join(ROOT, "rrt_b200", "bin", "rrt")
    scene = ref_scene_path(name)
    if not (os.path.exists(ref_exe) and os.path.exists(our_exe) and scene):
        pytest.skip("oracle/_ref/rrtc_dropin not built (needs /root/reference at build time)")
    a, err_a = _png(ref_exe, scene, w, h, spp, tmp_path / "ref_main.png")
    b, err_b = _png(our_exe, scene, w, h, spp, tmp_path / "own_main.png")
    assert a.shape == (h, w, 3) and np.array_equal(a, b), name
    # the reference's parser summary comes from its own scene.h, ours from the library: same lines (scene.h:443-451)
    head = lambda e: [l for l in e.splitlines() if l.split(":")[0] in ("material count", "sphere count", "msphere count", "obj count", "obj_inst count")]
    assert head(err_a) == head(err_b) and len(head(err_a)) == 5
    # -b (flat scan) through the reference's own flag parsing: same image
    c, _ = _png(ref_exe, scene, w, h, spp, tmp_path / "ref_main_b.png", ["-b"])
    assert np.array_equal(a, c)


def test_reference_main_cuda_flavour(tmp_path):
    """The nvcc-built host (USE_CUDA: -tx/-ty/-D/-q flags, 7-argument Rrt, float tan in the camera): a last-bit camera
    difference redirects a few paths, the rest of the image is identical."""
    from oracle_lib import psnr, ref_scene_path

    ref_exe = os.path.join(ROOT, "oracle", "_ref", "rrt_dropin")
    our_exe = os.path.join(ROOT, "rrt_b200", "bin", "rrt")
    scene = ref_scene_path("final.txt")
    if not (os.path.exists(ref_exe) and os.path.exists(our_exe) and scene):
        pytest.skip("oracle/_ref/rrt_dropin not built")
    a, _ = _png(ref_exe, scene, 240, 160, 8, tmp_path / "a.png", ["-tx", "16", "-ty", "16", "-D", "0"])
    b, _ = _png(our_exe, scene, 240, 160, 8, tmp_path / "b.png")
    assert (a == b).all(axis=2).mean() > 0.99 and psnr(a, b) > 40.0


def test_reference_main_double_build(tmp_path):
    """rrtd: the reference's double parser feeds the binding (camera fields rounded to float there), so a path can
    differ where a camera component rounds differently from the float derivation; the images agree closely."""
    from oracle_lib import psnr, ref_scene_path

    ref_exe = os.path.join(ROOT, "oracle", "_ref", "rrtd_dropin")
    our_exe = os.path.join(ROOT, "rrt_b200", "bin", "rrtd")
    scene = ref_scene_path("test2.txt")
    if not (os.path.exists(ref_exe) and os.path.exists(our_exe) and scene):
        pytest.skip("oracle/_ref/rrtd_dropin not built")
    a, _ = _png(ref_exe, scene, 256, 144, 16, tmp_path / "a.png")
    b, _ = _png(our_exe, scene, 256, 144, 16, tmp_path / "b.png")
    assert (a == b).all(axis=2).mean() > 0.9 and psnr(a, b) > 35.0


def test_reference_main_exit_codes(tmp_path):
    """Failures keep the reference's convention through its own main(): usage -> 1, unreadable scene -> 2."""
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "rrtc_dropin")
    if not os.path.exists(ref_exe):
        pytest.skip("oracle/_ref/rrtc_dropin not built")
    assert subprocess.run([ref_exe, "-z"], capture_output=True).returncode == 1
    assert subprocess.run([ref_exe, "-i", str(tmp_path / "missing.txt")], capture_output=True).returncode == 2
