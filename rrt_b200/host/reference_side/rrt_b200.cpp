// rrt_b200.cpp -- the REFERENCE-SIDE binding: a third implementation of the reference's `class Rrt`
// (reference rrt.h:14-48; its own two are rrt.cu and rrt.cpp, chosen at link time, reference Makefile:21-22) that
// forwards to the C ABI of librrtb200.so (include/rrtb.h).
//
// This is the file a maintainer of rogerallen/rrt adds next to main.cpp; nothing else of the reference changes:
//
//   nvcc -O3 -DUSE_FLOAT -DUSE_CUDA main.cpp rrt_b200.cpp -I<rrt-b200>/include -L<rrt-b200>/rrt_b200 -lrrtb200 -o rrt
//   nvcc -O3             -DUSE_CUDA main.cpp rrt_b200.cpp ...                                              -o rrtd
//
// It includes the reference's own headers, so it only compiles where the reference tree is: oracle/Makefile
// (`make dropin`) builds it against the UNMODIFIED /root/reference/main.cpp + scene.h + color.h + stb_image_write.h
// into oracle/_ref/rrt_dropin{,d}, and tests/test_host.py checks that those executables write the same PNGs as this
// repository's standalone rrt_b200/bin/rrt{,d}.
//
// The reference keeps the derived camera fields private (camera.h:40-48) and offers no accessor; a maintainer would add
// `friend class Rrt;` to class camera -- this translation unit reads them without touching the reference source.
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <random>
#include <sstream>
#include <string>
#include <vector>
#ifdef USE_CUDA
#include <curand_kernel.h>
#endif
// (the standard headers come first: the macro below must only reach the reference's own classes)
#define private public
#include "rrt.h" // the reference's header: class Rrt, scene, camera, vec3 (FP_T = float or double)
#undef private

#include "rrtb.h"

static rrtb_ctx *g_ctx = nullptr;

#ifdef USE_CUDA
// main.cpp uses checkCudaErrors for -q / -D (main.cpp:13-30,107-110); rrt.cu:31-40 is where the reference defines it
void check_cuda(cudaError_t result, char const *const func, const char *const file, int const line)
{
    if (result) {
        std::cerr << "CUDA error = " << static_cast<unsigned int>(result) << " at " << file << ":" << line << " '" << func << "' \n";
        exit(99);
    }
}
#endif

static void check(int rc, const char *what) // the reference's failure convention: message, exit(99)
{
    if (rc == RRTB_OK) return;
    std::cerr << "CUDA error = " << rc << " at rrt_b200.cpp '" << what << "' " << rrtb_last_error(g_ctx) << "\n";
    exit(99);
}

static rrtb_camera camera_to_rrtb(const camera &c) // 24 floats, camera.h:40-48
{
    rrtb_camera o;
    for (int k = 0; k < 3; ++k) {
        o.origin[k] = (float)c.origin.e[k];
        o.lower_left_corner[k] = (float)c.lower_left_corner.e[k];
        o.horizontal[k] = (float)c.horizontal.e[k];
        o.vertical[k] = (float)c.vertical.e[k];
        o.u[k] = (float)c.u.e[k];
        o.v[k] = (float)c.v.e[k];
        o.w[k] = (float)c.w.e[k];
    }
    o.lens_radius = (float)c.lens_radius;
    o.time0 = (float)c.time0;
    o.time1 = (float)c.time1;
    return o;
}

vec3 *Rrt::render(scene *s)
{
    int device = 0;
#ifdef USE_CUDA
    cudaGetDevice(&device); // -D <n> was applied by main.cpp with cudaSetDevice (main.cpp:107-110)
#endif
    if (!g_ctx) check(rrtb_create(&g_ctx, device), "rrtb_create");

    // scene -> C structs (public members of class scene, scene.h:474-480), object ids in the reference's insertion order
    std::vector<rrtb_material> mats;
    for (auto m : s->materials) {
        rrtb_material o{};
        o.type = (int)m->type; // LAMBERTIAN = 0, METAL = 1, DIELECTRIC = 2 (scene.h:183)
        if (m->type == DIELECTRIC) {
            o.param = (float)m->mat.dielectric.ref_idx;
        }
        else {
            for (int k = 0; k < 3; ++k) o.albedo[k] = (float)m->mat.metal.albedo.e[k]; // union: lambertian.albedo is at the same offset
            o.param = m->type == METAL ? (float)m->mat.metal.fuzz : 0.f;
        }
        mats.push_back(o);
    }
    std::vector<rrtb_sphere> sph;
    for (auto p : s->spheres) {
        rrtb_sphere o{};
        for (int k = 0; k < 3; ++k) o.center[k] = (float)p->center.e[k];
        o.radius = (float)p->radius;
        o.material = p->material_idx;
        sph.push_back(o);
    }
    std::vector<rrtb_msphere> msph;
    for (auto p : s->moving_spheres) {
        rrtb_msphere o{};
        for (int k = 0; k < 3; ++k) {
            o.center0[k] = (float)p->center0.e[k];
            o.center1[k] = (float)p->center1.e[k];
        }
        o.time0 = (float)p->time0;
        o.time1 = (float)p->time1;
        o.radius = (float)p->radius;
        o.material = p->material_idx;
        msph.push_back(o);
    }
    std::vector<scene_instance_triangle> inst((size_t)s->num_triangles()); // flattened on the host, scene.h:459-472
    if (!inst.empty()) s->fill_instance_triangles(inst.data());
    std::vector<rrtb_triangle> tris(inst.size());
    for (size_t i = 0; i < inst.size(); ++i) {
        for (int k = 0; k < 3; ++k) {
            tris[i].v0[k] = (float)inst[i].vertices[0].e[k];
            tris[i].v1[k] = (float)inst[i].vertices[1].e[k];
            tris[i].v2[k] = (float)inst[i].vertices[2].e[k];
        }
        tris[i].material = inst[i].material_idx;
    }
    const rrtb_camera cam = camera_to_rrtb(*s->cam);
    check(rrtb_scene_set(g_ctx, &cam, mats.data(), (int)mats.size(), sph.data(), (int)sph.size(), msph.data(), (int)msph.size(),
                         tris.data(), (int)tris.size(), bvh ? 1 : 0),
          "rrtb_scene_set");

    fb = new vec3[(size_t)image_width * image_height]; // owned by Rrt and freed in ~Rrt, like rrt.cpp:188-193
    rrtb_render_params p{};
    p.width = image_width;
    p.height = image_height;
    p.spp = samples_per_pixel;
    p.max_depth = max_depth;
    p.seed = 1984; // the reference's curand seed (rrt.cu:88)
    p.rank = 0;
    p.world = 1;
    p.shard_mode = RRTB_SHARD_TILES;
    p.count_rays = 0; // the timed render is the kernel without counters
    rrtb_stats st{};
    std::cerr << "Rendering a " << image_width << "x" << image_height << " image with " << samples_per_pixel
              << " samples per pixel on librrtb200 (sm_100a).\n";
    if (sizeof(FP_T) == 8) { // rrtd: the double integrator and the double framebuffer
        p.precision = RRTB_PRECISION_F64;
        check(rrtb_render_f64(g_ctx, &p, (double *)&fb[0].e[0], &st), "rrtb_render_f64");
    }
    else { // rrt: vec3 is 3 packed floats (vec3.h:81)
        p.precision = RRTB_PRECISION_F32;
        check(rrtb_render(g_ctx, &p, (float *)&fb[0].e[0], &st), "rrtb_render");
    }
    std::cerr << "took " << st.seconds_render << " seconds.\n"; // rrt.cu:302
    return fb;
}

Rrt::~Rrt()
{
    delete[] fb;
    rrtb_destroy(g_ctx);
    g_ctx = nullptr;
}
