"""Synthetic scaled-up scenes in the REFERENCE GRAMMAR (scene.h:212-452), so that the reference's own
parser could read them (BASELINE.json configs[4], SURVEY 8d "concrete inputs").

Full size: ground sphere + 100 000 small spheres on a 316x317 grid + 3 hero spheres of the book-1 finale
+ one level-5 icosphere (10 242 vertices, 20 480 triangles) instanced 49x on a 7x7 grid = 1 003 520
triangles.  Scaled-down siblings use fewer spheres / a coarser icosphere / fewer instances.
"""
import numpy as np

SEED = 20221005


def icosphere(level):
    """Unit icosphere: (vertices [n,3] float64, triangles [m,3] int), CCW seen from outside."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    v = [np.array(x, dtype=np.float64) / np.linalg.norm(x) for x in v]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    for _ in range(level):
        cache = {}

        def mid(a, b):
            key = (a, b) if a < b else (b, a)
            if key not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[key] = len(v) - 1
            return cache[key]

        nf = []
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return np.array(v), np.array(f, dtype=np.int64)


def synthetic_scene_text(n_spheres=100_000, ico_level=5, grid=7, seed=SEED):
    """-> scene file text.  Defaults give BASELINE.json configs[4] (1 003 520 triangles + 100 004 spheres)."""
    rng = np.random.default_rng(seed)
    out = ["# synthetic scene: %d small spheres, level-%d icosphere x %d instances (seed %d)" % (n_spheres, ico_level, grid * grid, seed),
           "camera 13 2 3   0 0 0   0 1 0  30.0 0.1 10.0",
           "material ground lambertian 0.5 0.5 0.5",
           "material hero_glass dielectric 1.5",
           "material hero_diffuse lambertian 0.4 0.2 0.1",
           "material hero_metal metal 0.7 0.6 0.5 0.0",
           "sphere 0 -1000 0 1000 ground",
           "sphere 0 1 0 1.0 hero_glass",
           "sphere -4 1 0 1.0 hero_diffuse",
           "sphere 4 1 0 1.0 hero_metal"]
    # small spheres: first n cells (row-major) of a 316 x 317 grid centred at the origin
    nx, nz = 316, 317
    k = np.arange(n_spheres)
    a = (k // nz) - nx // 2
    b = (k % nz) - nz // 2
    u = rng.uniform(size=(n_spheres, 2))
    cx, cz = a + 0.9 * u[:, 0], b + 0.9 * u[:, 1]
    kind = rng.uniform(size=n_spheres)
    alb = rng.uniform(size=(n_spheres, 6))
    fuzz = rng.uniform(0, 0.5, size=n_spheres)
    for i in range(n_spheres):
        name = "s%d" % i
        if kind[i] < 0.80:
            c = alb[i, :3] * alb[i, 3:]
            out.append("material %s lambertian %.6f %.6f %.6f" % (name, c[0], c[1], c[2]))
        elif kind[i] < 0.95:
            c = 0.5 + 0.5 * alb[i, :3]
            out.append("material %s metal %.6f %.6f %.6f %.6f" % (name, c[0], c[1], c[2], fuzz[i]))
        else:
            out.append("material %s dielectric 1.5" % name)
        out.append("sphere %.6f 0.2 %.6f 0.2 %s" % (cx[i], cz[i], name))
    # one icosphere object, grid x grid instances
    v, f = icosphere(ico_level)
    out.append("obj_beg %d %d" % (len(v), len(f)))
    out += ["obj_vtx %.7f %.7f %.7f" % (p[0], p[1], p[2]) for p in v]
    out += ["obj_tri %d %d %d" % (t[0], t[1], t[2]) for t in f]
    out.append("obj_end")
    mats = ["hero_diffuse", "hero_metal", "hero_glass", "ground"]
    for gi in range(grid):
        for gj in range(grid):
            sc = rng.uniform(0.6, 1.2)
            th = rng.uniform(0.0, 360.0)
            x = (gi - (grid - 1) / 2.0) * 6.0
            z = (gj - (grid - 1) / 2.0) * 6.0
            m = mats[(gi * grid + gj) % len(mats)]
            out.append("obj 0 %s s %.6f %.6f %.6f r %.4f 0 1 0 t %.4f 1.0 %.4f" % (m, sc, sc, sc, th, x, z))
    return "\n".join(out) + "\n"


def write_synthetic_scene(path, **kw):
    with open(path, "w") as f:
        f.write(synthetic_scene_text(**kw))
    return path
