// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A C-callable window onto the UNMODIFIED reference (rogerallen/rrt), compiled from the
// sources where they lie under /root/reference (-I/root/reference, see oracle/Makefile).
// Nothing of the reference is copied: this file only #includes its headers and rrt.cpp and
// re-exports the functions on the hot path so that tests can (a) validate oracle/rrt_oracle.c
// and (b) generate the golden fixtures under tests/golden/ (tools/make_golden.py).
//
// Built twice: -DUSE_FLOAT -> oracle/_ref/libref_f.so (the `rrt`/`rrtc` precision),
//              (nothing)   -> oracle/_ref/libref_d.so (the `rrtd`/`rrto` precision).
// All arrays crossing this boundary are double (exact for float values) or int32.
//
// Reference entry points used (file:line in /root/reference):
//   scene::scene               scene.h:212-452      parser
//   scene::fill_instance_triangles scene.h:467-472  host-side triangle flattening
//   create_world               rrt.cpp:54-97        object order: spheres, moving spheres, triangles
//   hittable_list::hit         hittable_list.h:95-117  linear closest-hit scan (object-id oracle)
//   bvh_node::hit              bvh.h:167-175
//   sphere/moving_sphere/triangle::hit  sphere.h:33-58, moving_sphere.h:32-58, triangle.h:35-75
//   *::bounding_box            sphere.h:60-64, moving_sphere.h:60-66, triangle.h:77-87
//   material::scatter          material.h:21-32,50-57,76-96
//   ray_color                  rrt.cpp:25-52
//   camera::get_ray            camera.h:31-38
//   convert_color              color.h:8-23

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <random>
#include <sstream>
#include <string>
#include <vector>
#include <omp.h>

// camera keeps its derived fields private (camera.h:40-48); the harness needs to read them.
#define private public
#include "rrt.cpp" // pulls in rrt.h, scene.h and every geometry/material header
#include "color.h"
#undef private

namespace {
inline vec3 v3(const double *p) { return vec3((FP_T)p[0], (FP_T)p[1], (FP_T)p[2]); }
inline void put3(double *o, const vec3 &v)
{
    o[0] = (double)v.e[0];
    o[1] = (double)v.e[1];
    o[2] = (double)v.e[2];
}
inline ray mkray(const double *r7) { return ray(v3(r7), v3(r7 + 3), (FP_T)r7[6]); }
} // namespace

extern "C" {

int ref_fp_bytes() { return (int)sizeof(FP_T); }

// ---- scene (scene.h) -------------------------------------------------------------------
void *ref_scene_load(const char *path, int w, int h) { return new scene(path, w, h); }
void ref_scene_free(void *s) { delete (scene *)s; }

// out[6] = materials, spheres, moving spheres, flattened triangles, objs, obj instances
void ref_scene_counts(void *sp, int *out)
{
    scene *s = (scene *)sp;
    out[0] = (int)s->materials.size();
    out[1] = (int)s->spheres.size();
    out[2] = (int)s->moving_spheres.size();
    out[3] = s->num_triangles();
    out[4] = (int)s->objs.size();
    out[5] = (int)s->obj_insts.size();
}

// out[24] = origin, lower_left_corner, horizontal, vertical, u, v, w (3 each), lens_radius, time0, time1
void ref_scene_camera(void *sp, double *out)
{
    camera *c = ((scene *)sp)->cam;
    put3(out + 0, c->origin);
    put3(out + 3, c->lower_left_corner);
    put3(out + 6, c->horizontal);
    put3(out + 9, c->vertical);
    put3(out + 12, c->u);
    put3(out + 15, c->v);
    put3(out + 18, c->w);
    out[21] = (double)c->lens_radius;
    out[22] = (double)c->time0;
    out[23] = (double)c->time1;
}

// type[i] in {0 lambertian, 1 metal, 2 dielectric}; params[4*i..] = r g b fuzz | ior 0 0 0 (dielectric)
void ref_scene_materials(void *sp, int *type, double *params)
{
    scene *s = (scene *)sp;
    for (size_t i = 0; i < s->materials.size(); ++i) {
        scene_material *m = s->materials[i];
        type[i] = (int)m->type;
        double *p = params + 4 * i;
        p[0] = p[1] = p[2] = p[3] = 0.0;
        if (m->type == LAMBERTIAN) {
            put3(p, m->mat.lambertian.albedo);
        }
        else if (m->type == METAL) {
            put3(p, m->mat.metal.albedo);
            p[3] = (double)m->mat.metal.fuzz;
        }
        else {
            p[0] = (double)m->mat.dielectric.ref_idx;
        }
    }
}

void ref_scene_spheres(void *sp, double *out4, int *mat)
{
    scene *s = (scene *)sp;
    for (size_t i = 0; i < s->spheres.size(); ++i) {
        put3(out4 + 4 * i, s->spheres[i]->center);
        out4[4 * i + 3] = (double)s->spheres[i]->radius;
        mat[i] = s->spheres[i]->material_idx;
    }
}

// out9 = c0(3) c1(3) t0 t1 r
void ref_scene_mspheres(void *sp, double *out9, int *mat)
{
    scene *s = (scene *)sp;
    for (size_t i = 0; i < s->moving_spheres.size(); ++i) {
        scene_moving_sphere *m = s->moving_spheres[i];
        put3(out9 + 9 * i, m->center0);
        put3(out9 + 9 * i + 3, m->center1);
        out9[9 * i + 6] = (double)m->time0;
        out9[9 * i + 7] = (double)m->time1;
        out9[9 * i + 8] = (double)m->radius;
        mat[i] = m->material_idx;
    }
}

void ref_scene_triangles(void *sp, double *out9, int *mat)
{
    scene *s = (scene *)sp;
    int n = s->num_triangles();
    std::vector<scene_instance_triangle> t(n > 0 ? n : 1);
    s->fill_instance_triangles(t.data());
    for (int i = 0; i < n; ++i) {
        for (int k = 0; k < 3; ++k) put3(out9 + 9 * i + 3 * k, t[i].vertices[k]);
        mat[i] = t[i].material_idx;
    }
}

// ---- world (rrt.cpp:54-97) ---------------------------------------------------------------
// use_bvh = 0 -> hittable_list (the `-b` path); 1 -> reference bvh_node tree.
void *ref_world_create(void *sp, int use_bvh) { return (void *)create_world((scene *)sp, use_bvh != 0); }


// The same object list as create_world (rrt.cpp:56-83) but from explicit arrays, so the DOUBLE build
// can be run on the float-rounded scene the product sees ("double arithmetic on identical inputs").
void *ref_world_from_arrays(int nm, const int *mtype, const double *mparams, int ns, const double *sph4,
                            const int *smat, int nms, const double *msph9, const int *msmat, int nt,
                            const double *tri9, const int *tmat)
{
    auto world_list = new hittable_list();
    std::vector<material_ptr_t> materials;
    for (int i = 0; i < nm; ++i) {
        const double *p = mparams + 4 * i;
        if (mtype[i] == LAMBERTIAN)
            materials.push_back(make_shared<lambertian>(color((FP_T)p[0], (FP_T)p[1], (FP_T)p[2])));
        else if (mtype[i] == METAL)
            materials.push_back(make_shared<metal>(color((FP_T)p[0], (FP_T)p[1], (FP_T)p[2]), (FP_T)p[3]));
        else
            materials.push_back(make_shared<dielectric>((FP_T)p[0]));
    }
    for (int i = 0; i < ns; ++i)
        world_list->add(make_shared<sphere>(v3(sph4 + 4 * i), (FP_T)sph4[4 * i + 3], materials[smat[i]]));
    for (int i = 0; i < nms; ++i) {
        const double *m = msph9 + 9 * i;
        world_list->add(make_shared<moving_sphere>(v3(m), v3(m + 3), (FP_T)m[6], (FP_T)m[7], (FP_T)m[8],
                                                   materials[msmat[i]]));
    }
    for (int i = 0; i < nt; ++i) {
        const double *t = tri9 + 9 * i;
        world_list->add(make_shared<triangle>(v3(t), v3(t + 3), v3(t + 6), materials[tmat[i]]));
    }
    return (void *)world_list;
}
// bvh_node over an existing list (bvh.h:15-23); time0/time1 = camera shutter (rrt.cpp:91)
void *ref_bvh_from_list(void *list, double time0, double time1)
{
    return (void *)new bvh_node(*(hittable_list *)list, (FP_T)time0, (FP_T)time1);
}

int ref_world_size(void *list) { return (int)((hittable_list *)list)->objects.size(); }

// The body of hittable_list::hit (hittable_list.h:95-117) with the loop index kept, so the
// winning OBJECT ID is observable.  `list` must come from ref_world_create(s, 0).
void ref_trace_scan(void *list, const double *rays7, int n, double t_min, int *id, double *t)
{
    hittable_list *w = (hittable_list *)list;
    const int nobj = (int)w->objects.size();
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        ray r = mkray(rays7 + 7 * (size_t)i);
        hit_record tmp;
        FP_T closest = infinity;
        int best = -1;
        for (int k = 0; k < nobj; ++k) {
            if (w->objects[k]->hit(r, (FP_T)t_min, closest, tmp, false)) {
                closest = tmp.t;
                best = k;
            }
        }
        id[i] = best;
        t[i] = best >= 0 ? (double)closest : -1.0;
    }
}

// world->hit() through whichever structure `world` is (list or bvh); full hit record out.
// rec10 = t, p(3), normal(3), front_face, hit(0/1), unused
void ref_trace_world(void *world, const double *rays7, int n, double t_min, double *rec10)
{
    hittable *w = (hittable *)world;
    for (int i = 0; i < n; ++i) {
        ray r = mkray(rays7 + 7 * (size_t)i);
        hit_record rec;
        double *o = rec10 + 10 * (size_t)i;
        for (int k = 0; k < 10; ++k) o[k] = 0.0;
        if (w->hit(r, (FP_T)t_min, infinity, rec, false)) {
            o[0] = (double)rec.t;
            put3(o + 1, rec.p);
            put3(o + 4, rec.normal);
            o[7] = rec.front_face ? 1.0 : 0.0;
            o[8] = 1.0;
        }
    }
}

// One object's hit() (sphere.h:33 / moving_sphere.h:32 / triangle.h:35). Returns 0/1.
int ref_hit_one(void *list, int obj, const double *ray7, double t_min, double t_max, double *rec8)
{
    hittable_list *w = (hittable_list *)list;
    ray r = mkray(ray7);
    hit_record rec;
    FP_T tmax = std::isinf(t_max) ? infinity : (FP_T)t_max;
    if (!w->objects[obj]->hit(r, (FP_T)t_min, tmax, rec, false)) return 0;
    rec8[0] = (double)rec.t;
    put3(rec8 + 1, rec.p);
    put3(rec8 + 4, rec.normal);
    rec8[7] = rec.front_face ? 1.0 : 0.0;
    return 1;
}

// One object's bounding_box(time0,time1). out6 = min(3) max(3)
void ref_bounding_box(void *list, int obj, double time0, double time1, double *out6)
{
    hittable_list *w = (hittable_list *)list;
    aabb b;
    w->objects[obj]->bounding_box((FP_T)time0, (FP_T)time1, b);
    put3(out6, b.minimum);
    put3(out6 + 3, b.maximum);
}

// aabb::hit (aabb.h:18-93)
int ref_aabb_hit(const double *box6, const double *ray7, double t_min, double t_max)
{
    aabb b(v3(box6), v3(box6 + 3));
    FP_T tmax = std::isinf(t_max) ? infinity : (FP_T)t_max;
    return b.hit(mkray(ray7), (FP_T)t_min, tmax) ? 1 : 0;
}

// ---- materials (material.h) -------------------------------------------------------------------
void *ref_material_create(int type, double r, double g, double b, double param)
{
    material_ptr_t *m = new material_ptr_t;
    if (type == 0)
        *m = make_shared<lambertian>(color((FP_T)r, (FP_T)g, (FP_T)b));
    else if (type == 1)
        *m = make_shared<metal>(color((FP_T)r, (FP_T)g, (FP_T)b), (FP_T)param);
    else
        *m = make_shared<dielectric>((FP_T)param);
    return m;
}
void ref_material_free(void *m) { delete (material_ptr_t *)m; }

// material::scatter with the reference's own process-global mt19937 (rtweekend.h:64-69).
// in: ray7, p(3), normal(3) (already face-forwarded, as hit_record holds it), front_face
// out9: scattered origin(3), scattered direction(3), attenuation(3). Returns scatter()'s bool.
int ref_scatter(void *mat, const double *ray7, const double *p, const double *nrm, int front_face, double *out9)
{
    material_ptr_t &m = *(material_ptr_t *)mat;
    hit_record rec;
    rec.p = v3(p);
    rec.normal = v3(nrm);
    rec.front_face = front_face != 0;
    rec.t = 0;
    ray scattered;
    color att;
    bool ok = m->scatter(mkray(ray7), rec, att, scattered, false);
    put3(out9, scattered.orig);
    put3(out9 + 3, scattered.dir);
    put3(out9 + 6, att);
    return ok ? 1 : 0;
}

void ref_reflect(const double *v, const double *n, double *out) { put3(out, reflect(v3(v), v3(n))); }
void ref_refract(const double *uv, const double *n, double eta, double *out)
{
    put3(out, refract(v3(uv), v3(n), (FP_T)eta));
}
double ref_reflectance(double cosine, double ref_idx) { return (double)dielectric::reflectance((FP_T)cosine, (FP_T)ref_idx); }

// ---- integrator pieces --------------------------------------------------------------------------
// camera::get_ray (camera.h:31-38); consumes the reference RNG (lens disk + shutter time).
void ref_camera_get_ray(void *sp, double s, double t, double *out7)
{
    ray r = ((scene *)sp)->cam->get_ray((FP_T)s, (FP_T)t);
    put3(out7, r.orig);
    put3(out7 + 3, r.dir);
    out7[6] = (double)r.tm;
}

// ray_color (rrt.cpp:25-52): mean radiance of `nsamples` independent evaluations of one ray.
void ref_ray_color_mean(void *world, const double *ray7, int depth, int nsamples, double *out3)
{
    hittable *w = (hittable *)world;
    ray r = mkray(ray7);
    double acc[3] = {0, 0, 0};
    for (int i = 0; i < nsamples; ++i) {
        color c = ray_color(r, w, depth, false);
        acc[0] += (double)c.e[0];
        acc[1] += (double)c.e[1];
        acc[2] += (double)c.e[2];
    }
    for (int k = 0; k < 3; ++k) out3[k] = acc[k] / nsamples;
}

// convert_color (color.h:8-23)
void ref_convert_color(double r, double g, double b, int spp, int *rgb)
{
    convert_color(color((FP_T)r, (FP_T)g, (FP_T)b), spp, rgb + 0, rgb + 1, rgb + 2);
}

double ref_random_uniform() { return (double)random_uniform(); }

} // extern "C"
