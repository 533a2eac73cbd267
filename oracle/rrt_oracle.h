/* rrt_oracle.h -- TEST INFRASTRUCTURE ONLY: the CPU oracle for the rrt path-tracing hot path.
 *
 * A plain-C restatement of (a) the reference's estimator / intersection / material semantics and
 * (b) the canonical LBVH + Philox algorithms the CUDA product implements.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 * It is never linked into, imported by or called from the product (rrt_b200/, librrtb200.so).
 *
 * Parity status: PINNED by execution of the reference itself -- the reference has no golden
 * vectors or tests of its own, so this oracle is checked against
 * oracle/_ref/libref_{f,d}.so (the unmodified reference headers compiled by oracle/Makefile) in
 * tests/test_oracle_vs_reference.py and against fixtures generated from that harness
 * (tests/golden/, tools/make_golden.py).
 *
 * The scene structs are the public C ABI ones (include/rrtb.h) so the same arrays feed both sides.
 */
#ifndef RRT_ORACLE_H
#define RRT_ORACLE_H

#include "../include/rrtb.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_scene {
    rrtb_camera cam;
    const rrtb_material *materials; int n_materials;
    const rrtb_sphere *spheres;     int n_spheres;
    const rrtb_msphere *mspheres;   int n_mspheres;
    const rrtb_triangle *triangles; int n_triangles;
    const rrtb_mtriangle *mtriangles; int n_mtriangles; /* translating instance triangles (ids after triangles) */
} orc_scene;

typedef struct orc_bvh {
    int n;              /* primitives */
    uint32_t *morton;   /* [n]   object-id order */
    uint32_t *perm;     /* [n]   sorted position -> object id */
    int32_t *left;      /* [n-1] >=0 internal, <0 leaf ~k */
    int32_t *right;     /* [n-1] */
    int32_t *parent;    /* [2n-1] */
    float *node_box;    /* [6(n-1)] */
    float *prim_box;    /* [6n]  object-id order */
    float pad;          /* absolute padding applied to boxes during traversal */
} orc_bvh;

typedef struct orc_counters { /* per-ray work counters (V_box, V_sph, V_msph, V_tri, h) */
    uint64_t rays, box_tests, sphere_tests, msphere_tests, triangle_tests, hits, paths;
} orc_counters;

/* Philox4x32-10 (Salmon et al., SC 2011). */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* (x >> 8) * 2^-24, in [0,1) */
float orc_u01(uint32_t x);
/* (cos, sin) of 2*pi*(u - 0.5), the deterministic polynomial both sides use for sampling */
void orc_sincos2pi(float u, float *c, float *s);

/* camera.h:8-29 */
void orc_camera_derive(const float lookfrom[3], const float lookat[3], const float vup[3], float vfov,
                       float aspect, float aperture, float focus, float t0, float t1, rrtb_camera *out);
/* rrt.cu:112-114 + camera.h:31-38 with the Philox streams of the product. ray7 = o d time */
void orc_camera_ray(const rrtb_camera *cam, int W, int H, int pixel, int sample, uint64_t seed, float *ray7);

/* per-primitive tests with the product's precision policy. id = object id. returns 0/1.
 * rec7 (optional) = p(3), face-forwarded normal(3), front_face */
int orc_hit_object(const orc_scene *s, int id, const float *ray7, float t_min, float t_max, float *t, float *rec7);
/* hittable_list.h:95-117 semantics: flat scan, object-id order, shrinking t_max */
void orc_trace_scan(const orc_scene *s, const float *rays7, int n, float t_min, int32_t *id, float *t, float *rec7);

/* canonical LBVH */
orc_bvh *orc_bvh_build(const orc_scene *s);
void orc_bvh_free(orc_bvh *b);
void orc_trace_bvh(const orc_scene *s, const orc_bvh *b, const float *rays7, int n, float t_min, int32_t *id,
                   float *t, float *rec7, orc_counters *cnt);

/* material.h scatter with an explicit Philox block. in16/out8 as rrtb_scatter. */
void orc_scatter(const orc_scene *s, const float *in16, const uint32_t *rnd4, int n, float *out8);

/* the estimator (rrt.cu:42-79,109-121 semantics, product RNG). out_rgb = 3*W*H float sums, bottom-up.
 * bvh may be NULL (flat scan). Pixels with (tile_index % world) != rank are left 0 when world > 1. */
void orc_render(const orc_scene *s, const orc_bvh *bvh, int W, int H, int spp, int max_depth, uint64_t seed,
                int rank, int world, int shard_mode, float *out_rgb, uint64_t *out_fixed, orc_counters *cnt);
/* ray_color for an explicit primary ray ((pixel, sample) only key the bounce RNG) */
void orc_radiance(const orc_scene *s, const orc_bvh *bvh, const float *ray7, int pixel, int sample, int max_depth,
                  uint64_t seed, float *rgb, orc_counters *cnt);
/* one camera path; returns radiance in rgb[3] */
void orc_path(const orc_scene *s, const orc_bvh *bvh, int W, int H, int pixel, int sample, int max_depth,
              uint64_t seed, float *rgb, orc_counters *cnt);

/* triangle.h:9-15: the float unit face normal the product stores in its leaf record */
void orc_triangle_normal(const rrtb_triangle *tr, float n[3]);
/* a moving triangle as the product's leaf record defines it (include/rrtb.h "rrtb_mtriangle"):
 * rate[k] = delta[k] / (time1 - time0), base[k] = fma(-rate[k], time0, v0[k]); v0(time) = fma(rate, time, base) */
void orc_mtriangle_record(const rrtb_mtriangle *m, float base[3], float rate[3], float e1[3], float e2[3]);
/* the full record: (base, rate) of v0, of e1 = v1 - v0 and of e2 = v2 - v0 (extra1 / extra2 make the edges move) */
void orc_mtriangle_record2(const rrtb_mtriangle *m, float base[3], float rate[3], float e1b[3], float e1r[3], float e2b[3], float e2r[3]);
static inline int orc_n_objects(const orc_scene *s) { return s->n_spheres + s->n_mspheres + s->n_triangles + s->n_mtriangles; }

/* ---- the DOUBLE integrator (rrt_oracle_f64.c): FP_T = double semantics over the same float scene ---- */
void orc_d_camera_ray(const rrtb_camera *cam, int W, int H, int pixel, int sample, uint64_t seed, double *ray7);
/* closest hit, leaf tests in double; b may be NULL (flat scan). rec7 optional = p(3) n(3) front */
void orc_d_trace(const orc_scene *s, const orc_bvh *b, const double *rays7, int n, double t_min, int32_t *id, double *t,
                 double *rec7);
void orc_d_scatter(const orc_scene *s, const double *in16, const uint32_t *rnd4, int n, double *out8);
void orc_d_render(const orc_scene *s, const orc_bvh *bvh, int W, int H, int spp, int max_depth, uint64_t seed, double *out_rgb,
                  uint64_t *out_fixed, orc_counters *cnt);

/* color.h:8-23 */
void orc_tonemap_rgb8(const float *rgb_sum, int W, int H, int spp, uint8_t *rgb8_topdown);

#ifdef __cplusplus
}
#endif
#endif
