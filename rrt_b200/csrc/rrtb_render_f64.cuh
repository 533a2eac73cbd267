// rrtb_render_f64.cuh -- render kernel and test hooks of the DOUBLE integrator (included by
// rrtb_render.cu).  Replaces, for the reference's `rrtd` build, cuda_render + ray_color (rrt.cu:42-122) with
// FP_T = double.  k_render_f64 is the persistent one-path-per-lane form (flat scan, and RRTB_SCHED_SIMPLE); the LBVH
// default is the pool scheduler of rrtb_render_pool.cuh instantiated over the PathF64 policy (final.txt: 1.8x).
#pragma once
#include "rrtb_device_f64.cuh"

namespace rrtb {

template <bool USE_BVH, bool COUNT_RAYS>
__global__ void __launch_bounds__(RENDER_TPB, 2) k_render_f64(const RenderArgs a)
{
    const unsigned lane = threadIdx.x & 31u;
    const DeviceScene &s = a.scene;
    const float4 *__restrict__ leaves = USE_BVH ? s.leaves : s.flat_leaves;
    const LeafAux info = USE_BVH ? LeafAux{s.leaf_info, s.leaf_ext} : LeafAux{s.flat_info, s.flat_ext};

    int pixel = -1, ls = 0, ls_end = 0;
    unsigned long long acc_r = 0, acc_g = 0, acc_b = 0;
    RayD ray;
    double thr_r = 1.0, thr_g = 1.0, thr_b = 1.0;
    int bounce = 0;
    bool in_path = false, done = false;
    unsigned long long rays = 0, hits = 0;
    TravCounters tc = {0ull, 0ull, 0ull, 0ull};

    while (true) {
        // ---- refill: same work items as k_render: (tile, chunk of CHUNK samples, pixel in tile)
        bool need = !done && pixel < 0;
        unsigned need_mask = __ballot_sync(0xffffffffu, need);
        if (need_mask) {
            unsigned long long base = 0;
            const int leader = __ffs(need_mask) - 1;
            if ((int)lane == leader) base = atomicAdd(a.queue, (unsigned long long)__popc(need_mask));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (need) {
                unsigned long long item = base + __popc(need_mask & ((1u << lane) - 1u));
                if (item >= a.n_items) {
                    done = true;
                }
                else {
                    unsigned pit = (unsigned)(item & 31ull);
                    unsigned long long tile_chunk = item >> 5;
                    int chunk = (int)(tile_chunk % (unsigned long long)a.n_chunks);
                    int ltile = (int)(tile_chunk / (unsigned long long)a.n_chunks);
                    int tile = a.shard_mode == RRTB_SHARD_TILES ? ltile * a.world + a.rank : ltile;
                    int ty = tile / a.tiles_x, tx = tile - ty * a.tiles_x;
                    int i = tx * 8 + (int)(pit & 7u), j = ty * 4 + (int)(pit >> 3);
                    if (i < a.W && j < a.H) {
                        pixel = j * a.W + i;
                        ls = chunk * CHUNK;
                        ls_end = min(ls + CHUNK, a.n_local_samples);
                        in_path = false;
                    }
                }
            }
        }
        if (__all_sync(0xffffffffu, done)) break;

        if (pixel >= 0) {
            const int sample = a.shard_mode == RRTB_SHARD_SAMPLES ? ls * a.world + a.rank : ls;
            if (!in_path) {
                ray = camera_ray_d(a.cam, a.W, a.H, pixel % a.W, pixel / a.W, sample, a.key);
                thr_r = thr_g = thr_b = 1.0;
                bounce = 0;
                in_path = true;
            }
            HitD h = USE_BVH ? closest_bvh_d<COUNT_RAYS>(s, ray, 0.001, tc) : closest_scan_d<COUNT_RAYS>(s, ray, 0.001, tc);
            if (COUNT_RAYS) {
                ++rays;
                if (h.ref >= 0) ++hits;
            }
            bool path_end = false;
            double lr = 0.0, lg = 0.0, lb = 0.0;
            if (h.ref < 0) {
                sky_d(ray, thr_r, thr_g, thr_b, lr, lg, lb);
                path_end = true;
            }
            else {
                HitRecordD rec = hit_record_d(leaves, info, ray, h);
                uint4 rnd = philox4x32_10(make_uint4((uint32_t)pixel, (uint32_t)sample, 2u + (uint32_t)bounce, 0u), a.key);
                float4 m = __ldg(&s.materials[rec.mat]);
                int mtype = __ldg(&s.material_type[rec.mat]);
                double dx, dy, dz, ar, ag, ab;
                if (scatter_d(mtype, m, ray, rec, rnd, dx, dy, dz, ar, ag, ab)) {
                    thr_r = __dmul_rn(thr_r, ar);
                    thr_g = __dmul_rn(thr_g, ag);
                    thr_b = __dmul_rn(thr_b, ab);
                    ray.ox = rec.px; ray.oy = rec.py; ray.oz = rec.pz;
                    ray.dx = dx; ray.dy = dy; ray.dz = dz;
                    if (++bounce >= a.max_depth) path_end = true; // exceeded depth: black (rrt.cu:78)
                }
                else {
                    path_end = true; // absorbed: black (rrt.cu:61-63)
                }
            }
            if (path_end) {
                acc_r += to_fixed_d(lr);
                acc_g += to_fixed_d(lg);
                acc_b += to_fixed_d(lb);
                in_path = false;
                if (++ls >= ls_end) {
                    unsigned long long *dst = a.accum + 3ull * (unsigned long long)pixel;
                    atomicAdd(dst + 0, acc_r);
                    atomicAdd(dst + 1, acc_g);
                    atomicAdd(dst + 2, acc_b);
                    acc_r = acc_g = acc_b = 0;
                    pixel = -1;
                }
            }
        }
    }
    if (COUNT_RAYS) {
        unsigned long long v[6] = {rays, tc.box, tc.sph, tc.msph, tc.tri, hits};
        for (int k = 0; k < 6; ++k) {
            for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
            if (lane == 0 && v[k]) atomicAdd(a.queue + 1 + k, v[k]);
        }
    }
}

// ---- test hooks -------------------------------------------------------------------------------------------
__global__ void k_trace_f64(const DeviceScene s, const double *__restrict__ rays7, int n, double t_min, int mode,
                            int *__restrict__ id, double *__restrict__ t, double *__restrict__ rec7)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *q = rays7 + 7 * (size_t)i;
    RayD r;
    r.ox = q[0]; r.oy = q[1]; r.oz = q[2]; r.dx = q[3]; r.dy = q[4]; r.dz = q[5]; r.tm = q[6];
    TravCounters tc;
    HitD h = mode ? closest_bvh_d<false>(s, r, t_min, tc) : closest_scan_d<false>(s, r, t_min, tc);
    if (h.ref < 0) {
        id[i] = -1;
        t[i] = -1.0;
        if (rec7)
            for (int k = 0; k < 7; ++k) rec7[7 * (size_t)i + k] = 0.0;
        return;
    }
    HitRecordD rec = hit_record_d(mode ? s.leaves : s.flat_leaves, mode ? LeafAux{s.leaf_info, s.leaf_ext} : LeafAux{s.flat_info, s.flat_ext}, r, h);
    id[i] = rec.obj;
    t[i] = h.t;
    if (rec7) {
        double *o = rec7 + 7 * (size_t)i;
        o[0] = rec.px; o[1] = rec.py; o[2] = rec.pz; o[3] = rec.nx; o[4] = rec.ny; o[5] = rec.nz;
        o[6] = rec.front ? 1.0 : 0.0;
    }
}

__global__ void k_camera_rays_f64(const DeviceCamera cam, int W, int H, uint2 key, const int *__restrict__ pix, int n,
                                  int sample, double *__restrict__ rays7)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RayD r = camera_ray_d(cam, W, H, pix[i] % W, pix[i] / W, sample, key);
    double *o = rays7 + 7 * (size_t)i;
    o[0] = r.ox; o[1] = r.oy; o[2] = r.oz; o[3] = r.dx; o[4] = r.dy; o[5] = r.dz; o[6] = r.tm;
}

__global__ void k_scatter_test_f64(const DeviceScene s, const double *__restrict__ in16, const uint32_t *__restrict__ rnd4,
                                   int n, double *__restrict__ out8)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *in = in16 + 16 * (size_t)i;
    RayD r;
    r.ox = in[0]; r.oy = in[1]; r.oz = in[2]; r.dx = in[3]; r.dy = in[4]; r.dz = in[5]; r.tm = in[6];
    HitRecordD rec;
    rec.px = in[7]; rec.py = in[8]; rec.pz = in[9];
    rec.nx = in[10]; rec.ny = in[11]; rec.nz = in[12];
    rec.front = in[13] != 0.0;
    rec.mat = (int)in[14];
    rec.obj = 0;
    uint4 rnd = make_uint4(rnd4[4 * i], rnd4[4 * i + 1], rnd4[4 * i + 2], rnd4[4 * i + 3]);
    double dx, dy, dz, ar, ag, ab;
    bool ok = scatter_d(__ldg(&s.material_type[rec.mat]), __ldg(&s.materials[rec.mat]), r, rec, rnd, dx, dy, dz, ar, ag, ab);
    double *o = out8 + 8 * (size_t)i;
    o[0] = dx; o[1] = dy; o[2] = dz; o[3] = ar; o[4] = ag; o[5] = ab; o[6] = ok ? 1.0 : 0.0; o[7] = 0.0;
}

} // namespace rrtb
