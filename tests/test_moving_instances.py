"""SURVEY 8f4 -- motion blur for object instances (`mobj` = translating, `kobj` = keyframed between two poses:
rotation / scale / translation over time; both are rrtb_mtriangle).

The reference has no such primitive (it is a README to-do, README.md:62), so there is nothing of the reference to
pin these to directly: PARITY UNPINNED for the moving case itself.  What is anchored instead:
  * a `mobj` with zero displacement is the reference-pinned static `obj` bit for bit (ids, t, records, images);
  * at a frozen time T a `mobj` equals the static `obj` translated to where the instance is at T;
  * a `kobj` whose two poses are equal is the static `obj` bit for bit; at a frozen time T it equals the static
    instance whose vertices are the linear blend of the two poses at T; its boxes cover every pose of the shutter;
  * an open shutter renders the time average of frozen-time renders;
  * the oracle and the CUDA path agree bit for bit on the moving case (GPU tests below).
"""
import numpy as np
import pytest

from oracle_lib import Oracle, pinhole_rays

HEAD = """camera   -1 2 5    0 0.5 -1  0 1 0  30.0  0.1  6.0  {t0} {t1}
material ground lambertian 0.8 0.8 0.0
material pinky  lambertian 0.7 0.3 0.3
material mirror metal      0.8 0.8 0.8   0.1
material glass  dielectric 1.5
obj_beg 4 4
obj_vtx  0.0  0.5774 -0.2041
obj_vtx -0.5 -0.2887 -0.2041
obj_vtx  0.5 -0.2887 -0.2041
obj_vtx  0.0  0.0     0.6124
obj_tri 0 1 3
obj_tri 0 3 2
obj_tri 0 2 1
obj_tri 1 2 3
obj_end
sphere  0.0 -100.5  -1.0   100.0  ground
sphere  1.2    0.3  -1.5     0.4  glass
"""


def scene_text(kind, t0=0.0, t1=1.0, delta=(0.6, 0.35, -0.2), mt=(0.0, 1.0), shift=(0.0, 0.0, 0.0)):
    """kind: 'static' (obj), 'moving' (mobj).  `shift` is added to the static instances' translation."""
    s = HEAD.format(t0=t0, t1=t1)
    places = [("mirror", (0.0, 0.3, -1.0), "r 30 0 1 0"), ("pinky", (-0.8, 0.2, -0.6), "s 1.2 0.8 1.0")]
    for mat, p, xf in places:
        if kind == "static":
            s += "obj 0 %s %s t %r %r %r\n" % (mat, xf, p[0] + shift[0], p[1] + shift[1], p[2] + shift[2])
        else:
            s += "mobj 0 %s %r %r %r %r %r %s t %r %r %r\n" % (mat, delta[0], delta[1], delta[2], mt[0], mt[1], xf, p[0], p[1], p[2])
    return s


def load(tmp_path, text, W=160, H=100, name="s.txt"):
    from rrt_b200 import Scene

    p = tmp_path / name
    p.write_text(text)
    return Scene.from_file(str(p), W, H)


def test_parser_mobj(tmp_path):
    from rrt_b200 import SceneError

    sc = load(tmp_path, scene_text("moving"))
    a = sc.arrays
    assert len(a.triangles) == 0 and len(a.mtriangles) == 8 and sc.n_obj_insts == 2
    assert np.allclose(a.mtriangles["delta"], [0.6, 0.35, -0.2]) and np.all(a.mtriangles["time1"] == 1.0)
    st = load(tmp_path, scene_text("static"), name="st.txt").arrays
    # vertices at time0 are the static instance's vertices, bit for bit
    for k in ("v0", "v1", "v2", "material"):
        assert np.array_equal(a.mtriangles[k], st.triangles[k])
    assert a.n_objects == 2 + 8
    with pytest.raises(SceneError) as e:
        load(tmp_path, HEAD.format(t0=0, t1=1) + "mobj 0 pinky 1 0 0 0.5 0.5\n", name="bad.txt")
    assert e.value.ref_exit_code == 1
    with pytest.raises(SceneError):
        load(tmp_path, HEAD.format(t0=0, t1=1) + "mobj 0 pinky 1 0\n", name="bad2.txt")


def test_zero_displacement_is_the_static_instance(tmp_path):
    """delta = 0: every hit, record and pixel of the moving path equals the reference-pinned static path."""
    W, H = 160, 100
    mv = load(tmp_path, scene_text("moving", delta=(0, 0, 0)), W, H).arrays
    st = load(tmp_path, scene_text("static"), W, H, "st.txt").arrays
    rays = pinhole_rays(st, W, H)
    rays[:, 6] = np.random.default_rng(0).uniform(0, 1, len(rays)).astype(np.float32)
    om, os_ = Oracle(mv), Oracle(st)
    for mode in ("scan", "bvh"):
        a = om.trace(rays, 0.001, mode, want_rec=True)
        b = os_.trace(rays, 0.001, mode, want_rec=True)
        for x, y in zip(a, b):
            assert x.tobytes() == y.tobytes(), mode
        a = om.trace_f64(rays.astype(np.float64), 0.001, mode, want_rec=True)
        b = os_.trace_f64(rays.astype(np.float64), 0.001, mode, want_rec=True)
        for x, y in zip(a, b):
            assert x.tobytes() == y.tobytes(), mode
    ia, fa, _ = om.render(48, 30, 3, 50, 11)
    ib, fb, _ = os_.render(48, 30, 3, 50, 11)
    assert fa.tobytes() == fb.tobytes()
    # boxes: the moving model's vertices are v0(T) + (v1 - v0), one rounding away from the stored v1
    assert np.allclose(om.bvh_arrays()["prim_box"], os_.bvh_arrays()["prim_box"], rtol=0, atol=2e-7)


@pytest.mark.parametrize("T", [0.0, 0.25, 1.0])
def test_frozen_time_equals_translated_static_instance(tmp_path, T):
    """Shutter closed at T: the moving instance is the static one translated by delta * (T - t0)/(t1 - t0)."""
    W, H = 200, 120
    delta, mt = (0.5, 0.25, -0.5), (0.0, 2.0)
    k = (T - mt[0]) / (mt[1] - mt[0])
    mv = load(tmp_path, scene_text("moving", T, T, delta, mt), W, H).arrays
    st = load(tmp_path, scene_text("static", T, T, shift=tuple(k * d for d in delta)), W, H, "st.txt").arrays
    rays = pinhole_rays(st, W, H)
    rays[:, 6] = T
    for mode in ("scan", "bvh"):
        ia, ta = Oracle(mv).trace(rays, 0.001, mode)
        ib, tb = Oracle(st).trace(rays, 0.001, mode)
        same = ia == ib
        assert same.mean() > 0.9995, (mode, same.mean())  # silhouette pixels may flip: the vertices round differently
        m = same & (ia >= 0)
        assert (np.abs(ta[m] - tb[m]) / tb[m]).max() < 2e-6
        ia, ta = Oracle(mv).trace_f64(rays.astype(np.float64), 0.001, mode)
        assert (ia == ib).mean() > 0.9995


def test_boxes_cover_the_shutter_interval(tmp_path):
    """Every position of a moving triangle during the shutter lies inside its LBVH leaf box; bvh == scan."""
    W, H = 200, 120
    mv = load(tmp_path, scene_text("moving", 0.2, 0.9, (0.9, 0.4, -0.7), (0.0, 1.0)), W, H).arrays
    orc = Oracle(mv)
    box = orc.bvh_arrays()["prim_box"][len(mv.spheres):]
    m = mv.mtriangles
    for T in np.linspace(0.2, 0.9, 8):
        kk = (T - m["time0"]) / (m["time1"] - m["time0"])
        for v in ("v0", "v1", "v2"):
            p = m[v] + kk[:, None] * m["delta"]
            assert np.all(p >= box[:, 0:3] - 1e-5) and np.all(p <= box[:, 3:6] + 1e-5)
    rays = pinhole_rays(mv, W, H)
    rays[:, 6] = np.random.default_rng(4).uniform(0.2, 0.9, len(rays)).astype(np.float32)
    a = orc.trace(rays, 0.001, "scan", want_rec=True)
    b = orc.trace(rays, 0.001, "bvh", want_rec=True)
    for x, y in zip(a, b):
        assert x.tobytes() == y.tobytes()
    assert (a[0] >= len(mv.spheres)).sum() > 300  # the instances are actually hit
    # the blur is there: hits on moving triangles at different times land on different points for the same pixel
    r0, r1 = rays.copy(), rays.copy()
    r0[:, 6], r1[:, 6] = 0.2, 0.9
    i0, _ = orc.trace(r0, 0.001, "bvh")
    i1, _ = orc.trace(r1, 0.001, "bvh")
    assert (i0 != i1).sum() > 200


POSE_A = "s 1.0 1.0 1.0 r 0 0 1 0 t 0.0 0.3 -1.0"
POSE_B = "s 1.3 0.8 1.1 r 70 0 1 0 t 0.35 0.45 -1.2"


def kobj_text(t0=0.0, t1=1.0, kt=(0.0, 1.0), pose_a=POSE_A, pose_b=POSE_B, extra=""):
    """One instance keyframed from pose_a (at kt[0]) to pose_b (at kt[1]): it scales, turns 70 degrees and moves."""
    return HEAD.format(t0=t0, t1=t1) + "kobj 0 mirror %r %r %s / %s\n" % (kt[0], kt[1], pose_a, pose_b) + extra


def blended_static(tmp_path, T, kt, W, H, name):
    """The static triangles a keyframed instance IS at time T: vertices blended linearly between the two poses."""
    from rrt_b200.types import SceneArrays, triangle_dtype

    a = load(tmp_path, HEAD.format(t0=T, t1=T) + "obj 0 mirror %s\n" % POSE_A, W, H, name + "a.txt").arrays
    b = load(tmp_path, HEAD.format(t0=T, t1=T) + "obj 0 mirror %s\n" % POSE_B, W, H, name + "b.txt").arrays
    k = np.float64((T - kt[0]) / (kt[1] - kt[0]))
    tri = np.zeros(len(a.triangles), triangle_dtype)
    for v in ("v0", "v1", "v2"):
        tri[v] = (a.triangles[v].astype(np.float64) * (1 - k) + b.triangles[v].astype(np.float64) * k).astype(np.float32)
    tri["material"] = a.triangles["material"]
    return SceneArrays(a.camera, a.materials, a.spheres, a.mspheres, tri)


def test_parser_kobj(tmp_path):
    from rrt_b200 import SceneError

    sc = load(tmp_path, kobj_text())
    m = sc.arrays.mtriangles
    assert len(m) == 4 and len(sc.arrays.triangles) == 0 and sc.n_obj_insts == 1
    a = load(tmp_path, HEAD.format(t0=0, t1=1) + "obj 0 mirror %s\n" % POSE_A, name="a.txt").arrays.triangles
    b = load(tmp_path, HEAD.format(t0=0, t1=1) + "obj 0 mirror %s\n" % POSE_B, name="b.txt").arrays.triangles
    # vertices at time0 are pose A's bit for bit; delta / extras carry each vertex to pose B (to float rounding)
    for v in ("v0", "v1", "v2"):
        assert np.array_equal(m[v], a[v])
    assert np.allclose(m["v0"] + m["delta"], b["v0"], atol=1e-6)
    assert np.allclose(m["v1"] + m["delta"] + m["extra1"], b["v1"], atol=1e-6)
    assert np.allclose(m["v2"] + m["delta"] + m["extra2"], b["v2"], atol=1e-6)
    assert np.abs(m["extra1"]).max() > 0.1  # it really deforms: not a translation
    with pytest.raises(SceneError) as e:  # time0 == time1
        load(tmp_path, HEAD.format(t0=0, t1=1) + "kobj 0 pinky 0.5 0.5 t 0 0 0 / t 1 0 0\n", name="bad.txt")
    assert e.value.ref_exit_code == 1
    with pytest.raises(SceneError):  # three poses
        load(tmp_path, HEAD.format(t0=0, t1=1) + "kobj 0 pinky 0 1 t 0 0 0 / t 1 0 0 / t 2 0 0\n", name="bad2.txt")
    # a translating mobj has zero extras
    assert np.all(load(tmp_path, scene_text("moving"), name="m.txt").arrays.mtriangles["extra1"] == 0)


def test_equal_poses_are_the_static_instance(tmp_path):
    """Both poses equal: every hit, record and pixel of the keyframed path equals the reference-pinned static path."""
    W, H = 160, 100
    mv = load(tmp_path, kobj_text(pose_b=POSE_A), W, H).arrays
    st = load(tmp_path, HEAD.format(t0=0.0, t1=1.0) + "obj 0 mirror %s\n" % POSE_A, W, H, "st.txt").arrays
    assert np.all(mv.mtriangles["delta"] == 0) and np.all(mv.mtriangles["extra1"] == 0) and np.all(mv.mtriangles["extra2"] == 0)
    rays = pinhole_rays(st, W, H)
    rays[:, 6] = np.random.default_rng(0).uniform(0, 1, len(rays)).astype(np.float32)
    om, os_ = Oracle(mv), Oracle(st)
    for mode in ("scan", "bvh"):
        for x, y in zip(om.trace(rays, 0.001, mode, want_rec=True), os_.trace(rays, 0.001, mode, want_rec=True)):
            assert x.tobytes() == y.tobytes(), mode
        for x, y in zip(om.trace_f64(rays.astype(np.float64), 0.001, mode, want_rec=True), os_.trace_f64(rays.astype(np.float64), 0.001, mode, want_rec=True)):
            assert x.tobytes() == y.tobytes(), mode
    assert om.render(48, 30, 3, 50, 11)[1].tobytes() == os_.render(48, 30, 3, 50, 11)[1].tobytes()


@pytest.mark.parametrize("T", [0.0, 0.3, 0.75, 1.5])
def test_frozen_time_equals_the_blended_static_instance(tmp_path, T):
    """Shutter closed at T (inside and outside [time0, time1]): the keyframed instance is the static one whose vertices
    are the linear blend of the two poses -- rotation, scale and translation in one."""
    W, H = 200, 120
    kt = (0.0, 1.5)
    mv = load(tmp_path, kobj_text(T, T, kt), W, H).arrays
    st = blended_static(tmp_path, T, kt, W, H, "st%d" % int(T * 100))
    rays = pinhole_rays(st, W, H)
    rays[:, 6] = T
    for mode in ("scan", "bvh"):
        ia, ta, ra = Oracle(mv).trace(rays, 0.001, mode, want_rec=True)
        ib, tb, rb = Oracle(st).trace(rays, 0.001, mode, want_rec=True)
        same = ia == ib
        assert same.mean() > 0.9995, (mode, same.mean())  # silhouette pixels may flip: the vertices round differently
        m = same & (ia >= 2)  # hits on the instance
        assert m.sum() > 200
        assert (np.abs(ta[m] - tb[m]) / tb[m]).max() < 5e-6
        assert np.abs(ra[m, 3:6] - rb[m, 3:6]).max() < 1e-4  # the normal is the pose's normal at T
        ia, ta = Oracle(mv).trace_f64(rays.astype(np.float64), 0.001, mode)
        assert (ia == ib).mean() > 0.9995


def test_keyframed_boxes_cover_every_pose_of_the_shutter(tmp_path):
    W, H = 200, 120
    t0, t1, kt = 0.2, 0.9, (0.0, 1.0)
    mv = load(tmp_path, kobj_text(t0, t1, kt), W, H).arrays
    orc = Oracle(mv)
    box = orc.bvh_arrays()["prim_box"][len(mv.spheres):]
    m = mv.mtriangles
    for T in np.linspace(t0, t1, 9):
        k = (T - kt[0]) / (kt[1] - kt[0])
        for v, mv_by in (("v0", m["delta"]), ("v1", m["delta"] + m["extra1"]), ("v2", m["delta"] + m["extra2"])):
            p = m[v] + k * mv_by
            assert np.all(p >= box[:, 0:3] - 1e-5) and np.all(p <= box[:, 3:6] + 1e-5)
    rays = pinhole_rays(mv, W, H)
    rays[:, 6] = np.random.default_rng(4).uniform(t0, t1, len(rays)).astype(np.float32)
    a = orc.trace(rays, 0.001, "scan", want_rec=True)
    b = orc.trace(rays, 0.001, "bvh", want_rec=True)
    for x, y in zip(a, b):
        assert x.tobytes() == y.tobytes()
    assert (a[0] >= len(mv.spheres)).sum() > 300


# ---------------------------------------------------------------------------------------------------------------
# GPU: the CUDA path against the oracle, bit for bit
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_moving_instances_bit_exact(ctx, tmp_path):
    W, H = 200, 120
    text = scene_text("moving", 0.1, 0.8, (0.9, 0.4, -0.7), (0.0, 1.0))
    text += "obj 0 glass t 0.9 0.2 -0.3\nobj 0 mirror r 45 1 0 0 t -0.2 0.9 -1.5\n"  # static instances next to moving ones
    sc = load(tmp_path, text, W, H).arrays
    assert len(sc.mtriangles) == 8 and len(sc.triangles) == 8
    orc = Oracle(sc)
    ctx.set_scene(sc, use_bvh=True)
    g, o = ctx.bvh_arrays(), orc.bvh_arrays()
    for k in ("morton", "perm", "left", "right", "parent", "node_box", "prim_box"):
        assert np.array_equal(g[k], o[k]), k
    rays = pinhole_rays(sc, W, H)
    rays[:, 6] = np.random.default_rng(4).uniform(0.1, 0.8, len(rays)).astype(np.float32)
    for mode in ("scan", "bvh"):
        a = ctx.trace(rays, 0.001, mode, want_rec=True)
        b = orc.trace(rays, 0.001, mode, want_rec=True)
        for x, y in zip(a, b):
            assert x.tobytes() == y.tobytes(), mode
        a = ctx.trace_f64(rays.astype(np.float64), 0.001, mode, want_rec=True)
        b = orc.trace_f64(rays.astype(np.float64), 0.001, mode, want_rec=True)
        for x, y in zip(a, b):
            assert x.tobytes() == y.tobytes(), mode
    assert (a[0] >= len(sc.spheres) + len(sc.triangles)).sum() > 300
    # whole framebuffers: the double integrator bit for bit, the float one to its usual tolerance; all schedulers agree
    w, h, spp = 96, 60, 6
    want, _, cnt = orc.render_f64(w, h, spp, 50, 5)
    img, st = ctx.render(w, h, spp, 50, seed=5, count_rays=True, dtype=np.float64, precision="f64")
    assert img.tobytes() == want.tobytes() and st["rays"] == cnt["rays"]
    ref, _, _ = orc.render(w, h, spp, 50, 5)
    f1, _ = ctx.render(w, h, spp, 50, seed=5, scheduler=1)
    f2, _ = ctx.render(w, h, spp, 50, seed=5, scheduler=2)
    assert f1.tobytes() == f2.tobytes()
    close = np.abs(np.sqrt(f1 / spp).clip(0, 1) - np.sqrt(ref / spp).clip(0, 1)).max(axis=2) < 1e-3
    assert close.mean() > 0.97
    ctx.set_scene(sc, use_bvh=False)
    f3, _ = ctx.render(w, h, spp, 50, seed=5)
    assert f3.tobytes() == f1.tobytes()


@pytest.mark.gpu
def test_gpu_motion_blur_is_the_time_average(ctx, tmp_path):
    """An open shutter renders the average of frozen-time renders (same estimator, time is just one more sampled
    dimension): compare against the mean of static frames at stratified times, statistically."""
    W, H, spp = 96, 60, 64
    delta, mt = (0.9, 0.0, 0.0), (0.0, 1.0)
    blur = load(tmp_path, scene_text("moving", 0.0, 1.0, delta, mt), W, H).arrays
    ctx.set_scene(blur)
    img, _ = ctx.render(W, H, spp, 50, seed=3)
    acc = np.zeros_like(img, dtype=np.float64)
    n_t = 16
    for i in range(n_t):
        T = (i + 0.5) / n_t
        fr = load(tmp_path, scene_text("static", T, T, shift=tuple(T * d for d in delta)), W, H, "f%d.txt" % i).arrays
        ctx.set_scene(fr)
        f, _ = ctx.render(W, H, spp // 4, 50, seed=100 + i)
        acc += f.astype(np.float64) / (spp // 4)
    a = np.sqrt(img.astype(np.float64) / spp).clip(0, 1)
    b = np.sqrt(acc / n_t).clip(0, 1)
    assert abs(a.mean() - b.mean()) < 4e-3
    assert np.abs(a - b).mean() < 0.03


@pytest.mark.gpu
def test_gpu_camera_set_rebuilds_for_a_new_shutter(ctx, tmp_path):
    """rrtb_camera_set with moving primitives and a new shutter interval rebuilds the LBVH on the device (the boxes of
    moving primitives span the shutter): everything equals a fresh rrtb_scene_set with that camera."""
    from oracle_lib import camera_derive
    from rrt_b200.types import SceneArrays

    sc = load(tmp_path, scene_text("moving", 0.0, 1.0), 64, 40).arrays
    ctx.set_scene(sc)
    before = ctx.bvh_arrays()
    cam = sc.camera.copy()
    cam["time0"], cam["time1"] = 0.25, 0.5
    ctx.set_camera(cam)
    after = ctx.bvh_arrays()
    assert not np.array_equal(before["prim_box"], after["prim_box"])
    img, _ = ctx.render(64, 40, 4, 50, seed=2)
    fresh = SceneArrays(cam, sc.materials, sc.spheres, sc.mspheres, sc.triangles, sc.mtriangles)
    ctx.set_scene(fresh)
    want = ctx.bvh_arrays()
    for k in ("morton", "perm", "left", "right", "parent", "node_box", "prim_box"):
        assert np.array_equal(after[k], want[k]), k
    img2, _ = ctx.render(64, 40, 4, 50, seed=2)
    assert img.tobytes() == img2.tobytes()
    # a camera moved far away: the traversal boxes are re-padded for its magnitude; tree traversal still equals the flat scan
    far = camera_derive((900.0, 300.0, 700.0), (0, 0.5, -1), (0, 1, 0), 1.0, 1.6, 0.0, 1100.0, 0.25, 0.5)
    ctx.set_camera(far)
    a, _ = ctx.render(64, 40, 2, 50, seed=3)
    ctx.set_scene(SceneArrays(far, sc.materials, sc.spheres, sc.mspheres, sc.triangles, sc.mtriangles), use_bvh=False)
    b, _ = ctx.render(64, 40, 2, 50, seed=3)
    assert a.tobytes() == b.tobytes()
    # staging is consumed: the next scene without moving triangles has none
    st = load(tmp_path, scene_text("static"), 64, 40, "st.txt").arrays
    ctx.set_scene(st)
    assert len(ctx.bvh_arrays()["perm"]) == st.n_objects


@pytest.mark.gpu
def test_gpu_keyframed_instances_bit_exact(ctx, tmp_path):
    """kobj (rotating + scaling + translating instance) next to a translating mobj and static objs: LBVH, hits and
    records (both integrators), the double integrator's framebuffer bit for bit against the oracle; all schedulers and
    the flat scan the same image."""
    W, H = 200, 120
    text = kobj_text(0.1, 0.8, (0.0, 1.0), extra="mobj 0 pinky 0.4 0.2 -0.3 0.0 1.0 s 1.2 0.8 1.0 t -0.8 0.2 -0.6\nobj 0 glass t 0.9 0.2 -0.3\n")
    sc = load(tmp_path, text, W, H).arrays
    assert len(sc.mtriangles) == 8 and len(sc.triangles) == 4
    orc = Oracle(sc)
    ctx.set_scene(sc, use_bvh=True)
    g, o = ctx.bvh_arrays(), orc.bvh_arrays()
    for k in ("morton", "perm", "left", "right", "parent", "node_box", "prim_box"):
        assert np.array_equal(g[k], o[k]), k
    rays = pinhole_rays(sc, W, H)
    rays[:, 6] = np.random.default_rng(4).uniform(0.1, 0.8, len(rays)).astype(np.float32)
    for mode in ("scan", "bvh"):
        for x, y in zip(ctx.trace(rays, 0.001, mode, want_rec=True), orc.trace(rays, 0.001, mode, want_rec=True)):
            assert x.tobytes() == y.tobytes(), mode
        a = ctx.trace_f64(rays.astype(np.float64), 0.001, mode, want_rec=True)
        for x, y in zip(a, orc.trace_f64(rays.astype(np.float64), 0.001, mode, want_rec=True)):
            assert x.tobytes() == y.tobytes(), mode
    assert (a[0] >= len(sc.spheres) + len(sc.triangles)).sum() > 300
    w, h, spp = 96, 60, 6
    want, _, cnt = orc.render_f64(w, h, spp, 50, 5)
    img, st = ctx.render(w, h, spp, 50, seed=5, count_rays=True, dtype=np.float64, precision="f64")
    assert img.tobytes() == want.tobytes() and st["rays"] == cnt["rays"]
    ref, _, _ = orc.render(w, h, spp, 50, 5)
    f1, _ = ctx.render(w, h, spp, 50, seed=5, scheduler=1)
    f2, _ = ctx.render(w, h, spp, 50, seed=5, scheduler=2)
    assert f1.tobytes() == f2.tobytes()
    close = np.abs(np.sqrt(f1 / spp).clip(0, 1) - np.sqrt(ref / spp).clip(0, 1)).max(axis=2) < 1e-3
    assert close.mean() > 0.97
    ctx.set_scene(sc, use_bvh=False)
    f3, _ = ctx.render(w, h, spp, 50, seed=5)
    assert f3.tobytes() == f1.tobytes()


@pytest.mark.gpu
def test_gpu_keyframed_blur_is_the_time_average(ctx, tmp_path):
    """Open shutter over a turning, scaling instance == the mean of frozen-time renders of the blended static instance."""
    W, H, spp = 96, 60, 64
    kt = (0.0, 1.0)
    ctx.set_scene(load(tmp_path, kobj_text(0.0, 1.0, kt), W, H).arrays)
    img, _ = ctx.render(W, H, spp, 50, seed=3)
    acc = np.zeros_like(img, dtype=np.float64)
    n_t = 16
    for i in range(n_t):
        T = (i + 0.5) / n_t
        ctx.set_scene(blended_static(tmp_path, T, kt, W, H, "f%d" % i))
        f, _ = ctx.render(W, H, spp // 4, 50, seed=100 + i)
        acc += f.astype(np.float64) / (spp // 4)
    a = np.sqrt(img.astype(np.float64) / spp).clip(0, 1)
    b = np.sqrt(acc / n_t).clip(0, 1)
    assert abs(a.mean() - b.mean()) < 4e-3
    assert np.abs(a - b).mean() < 0.03


def fast_spheres_scene(n=64, travel=3.0, seed=7, shutter=(0.0, 1.0)):
    """n small spheres that each travel `travel` units (30 radii) during the shutter, over a ground sphere."""
    from oracle_lib import camera_derive
    from rrt_b200.types import SceneArrays, material_dtype, msphere_dtype, sphere_dtype

    rng = np.random.default_rng(seed)
    mats = np.zeros(3, material_dtype)
    mats["type"] = [0, 1, 2]
    mats["albedo"][0] = (0.6, 0.5, 0.4)
    mats["albedo"][1] = (0.8, 0.8, 0.9)
    mats["param"] = [0, 0.1, 1.5]
    ms = np.zeros(n, msphere_dtype)
    c0 = rng.uniform(-4, 4, size=(n, 3)).astype(np.float32)
    c0[:, 1] = rng.uniform(0.2, 2.5, size=n)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True) * travel).astype(np.float32)
    ms["center0"], ms["center1"] = c0, c0 + d
    ms["time0"], ms["time1"] = 0.0, 1.0
    ms["radius"] = 0.1
    ms["material"] = rng.integers(0, 3, size=n)
    sp = np.zeros(1, sphere_dtype)
    sp[0] = ((0, -1000, 0), 1000.0, 0)
    cam = camera_derive((0, 3, 12), (0, 1, 0), (0, 1, 0), 40.0, 1.6, 0.0, 12.0, *shutter)
    return SceneArrays(cam, mats, sp, ms)


@pytest.mark.gpu
def test_gpu_motion_nodes_interpolate_their_boxes(ctx):
    """Moving primitives under an open shutter are traversed through MOTION nodes whose child boxes are interpolated at
    the ray's time (rrtb_device.cuh "Motion node").  (1) Conservative: closest hits (ids, t, records) equal the flat
    scan and the oracle bit for bit at random times of the shutter, both integrators; whole f64 framebuffers equal the
    oracle's.  (2) Tight: far fewer exact sphere tests per ray than the oracle's traversal of boxes that span the
    whole shutter (the reference's moving_sphere::bounding_box)."""
    from test_gpu_edge_cases import check_wide_tree

    sc = fast_spheres_scene()
    orc = Oracle(sc)
    ctx.set_scene(sc, use_bvh=True)
    check_wide_tree(ctx, sc.n_objects)  # downloaded as the union of the two end boxes: encloses the canonical boxes
    W, H = 200, 125
    rays = pinhole_rays(sc, W, H)
    rays[:, 6] = np.random.default_rng(1).uniform(0, 1, len(rays)).astype(np.float32)
    want = orc.trace(rays, 0.001, "bvh", want_rec=True)
    for mode in ("scan", "bvh"):
        for x, y in zip(ctx.trace(rays, 0.001, mode, want_rec=True), want):
            assert x.tobytes() == y.tobytes(), mode
    want64 = orc.trace_f64(rays.astype(np.float64), 0.001, "bvh", want_rec=True)
    for x, y in zip(ctx.trace_f64(rays.astype(np.float64), 0.001, "bvh", want_rec=True), want64):
        assert x.tobytes() == y.tobytes()
    assert (want[0] >= 1).sum() > 150  # moving spheres are hit
    w, h, spp = 96, 60, 8
    ref64, _, oc = orc.render_f64(w, h, spp, 50, 5)
    img64, st64 = ctx.render(w, h, spp, 50, seed=5, count_rays=True, dtype=np.float64, precision="f64")
    assert img64.tobytes() == ref64.tobytes() and st64["rays"] == oc["rays"]
    f1, s1 = ctx.render(w, h, spp, 50, seed=5, scheduler=1, count_rays=True)
    f2, s2 = ctx.render(w, h, spp, 50, seed=5, scheduler=2, count_rays=True)
    assert f1.tobytes() == f2.tobytes() and s1["rays"] == s2["rays"]
    ctx.set_scene(sc, use_bvh=False)
    f3, _ = ctx.render(w, h, spp, 50, seed=5)
    assert f3.tobytes() == f1.tobytes()
    # tightness: exact moving-sphere tests per ray, motion nodes vs the oracle's shutter-spanning boxes
    _, _, ocf = orc.render(w, h, spp, 50, 5)
    ours, theirs = s2["msphere_tests"] / s2["rays"], ocf["msphere_tests"] / ocf["rays"]
    assert ours < 0.5 * theirs, (ours, theirs)
    # a closed shutter has no motion: plain nodes, same answers as the oracle
    frozen = fast_spheres_scene(shutter=(0.4, 0.4))
    ctx.set_scene(frozen, use_bvh=True)
    r2 = pinhole_rays(frozen, W, H)
    r2[:, 6] = 0.4
    for x, y in zip(ctx.trace(r2, 0.001, "bvh", want_rec=True), Oracle(frozen).trace(r2, 0.001, "bvh", want_rec=True)):
        assert x.tobytes() == y.tobytes()
