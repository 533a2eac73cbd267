"""-m "not gpu": the C-ABI library loads and exports every symbol include/rrtb.h declares; the struct
mirrors have the C sizes; the product fails LOUDLY without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "rrtb.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rrtb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(built_lib):
    from rrt_b200 import _lib

    lib = C.CDLL(built_lib)
    syms = header_symbols()
    assert len(syms) >= 28
    for s in syms:
        assert hasattr(lib, s), "librrtb200.so does not export %s" % s
    assert sorted(_lib.SYMBOLS) == syms, "rrt_b200/_lib.py binds a different set than include/rrtb.h declares"
    assert _lib.load().rrtb_abi_version() == 3


def test_struct_layouts_match_header(tmp_path, built_lib):
    """Compile a probe against include/rrtb.h and compare sizeof/offsetof with the numpy/ctypes mirrors."""
    from rrt_b200.types import RenderParams, Stats, camera_dtype, material_dtype, msphere_dtype, sphere_dtype, triangle_dtype

    src = tmp_path / "probe.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "rrtb.h"\n'
        "int main(void){printf(\"%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n\", sizeof(rrtb_camera), sizeof(rrtb_material),"
        " sizeof(rrtb_sphere), sizeof(rrtb_msphere), sizeof(rrtb_triangle), sizeof(rrtb_render_params), sizeof(rrtb_stats),"
        " offsetof(rrtb_render_params, seed), offsetof(rrtb_stats, kernel_launches));return 0;}\n"
    )
    exe = tmp_path / "probe"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [camera_dtype.itemsize, material_dtype.itemsize, sphere_dtype.itemsize, msphere_dtype.itemsize, triangle_dtype.itemsize,
            C.sizeof(RenderParams), C.sizeof(Stats), RenderParams.seed.offset, Stats.kernel_launches.offset]
    assert got == want


def test_no_cpu_fallback(built_lib):
    """Without a CUDA device rrtb_create must fail with RRTB_ERR_NO_DEVICE -- never compute on the CPU."""
    import torch

    from rrt_b200 import Context, RrtbError

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RrtbError) as e:
        Context(0)
    assert e.value.status == -2 and "no CPU fallback" in str(e.value)


def test_product_does_not_touch_the_oracle():
    """The oracle is test infrastructure: nothing under rrt_b200/ may import, link or execute oracle/."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "rrt_b200")):
        if "build" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(base, f), errors="replace").read()
                for ln in txt.splitlines():
                    code = ln.split("//")[0].split("#")[0] if not f.endswith(".py") else ln.split("#")[0]
                    if re.search(r"oracle_lib|liboracle|rrt_oracle|libref_|/oracle/|\"oracle\"", code):
                        bad.append((f, ln.strip()))
    assert not bad, bad
