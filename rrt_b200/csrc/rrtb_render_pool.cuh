// rrtb_render_pool.cuh -- scheduler RRTB_SCHED_POOL: a per-warp on-chip wavefront (sm_100a).
//
// Why (ncu, profiles/r01_ncu_sched_simple.md): with one path per lane a warp runs at ~7 of 32 lanes --
// every lane waits for the slowest traversal of the round (node visits per ray: mean 12, tail > 60) and
// the shading code runs once per round whatever the number of lanes that need it.
//
// Here each WARP owns a pool of POOL path slots in shared memory (60 B per slot, 228 KB/SM makes room for
// 2048 slots per SM) and runs a small warp-synchronous scheduler over them:
//
//   FETCH  idle lanes pop a slot from the warp's trace stack, load its ray, start a traversal
//   STEP   traversing lanes do ONE step: a BVH node visit, or (by warp vote) one exact leaf test;
//          a lane whose traversal ends writes (t, leaf ref) to its slot, pushes it on the shade stack
//          and becomes idle -- it does NOT wait for the other lanes
//   SHADE  32 slots off the shade stack are shaded by 32 lanes (hit record, scatter / sky, accumulate);
//          a finished path is replaced in place by the next camera path of the global work queue
//          (one warp-aggregated atomicAdd per batch); slots with a new ray go back on the trace stack.
//          Lanes that are in the middle of a traversal keep their traversal registers and resume.
//
// So traversal lanes are refilled as soon as they finish and shading always runs 32 wide; nothing but
// the per-sample radiance (three 64-bit integer atomics into the L2-resident accumulator) leaves the SM.
// Everything is keyed by (pixel, sample, bounce), so the image is bit-identical to the other schedulers.
#pragma once

namespace rrtb {

static constexpr int POOL = 128;               // path slots per warp
static constexpr int POOL_WARPS = RENDER_TPB / 32;
static constexpr int SLOT_FRESH = -2;          // hit_ref marker: slot holds no path yet / path ended
static constexpr int STEP_ITERS = 4;           // node visits per scheduling round

struct WarpPool { // SoA, one per warp, in dynamic shared memory
    float ox[POOL], oy[POOL], oz[POOL], dx[POOL], dy[POOL], dz[POOL], tm[POOL];
    float tr[POOL], tg[POOL], tb[POOL];
    int pixel[POOL], sample[POOL], bounce[POOL];
    float hit_t[POOL];
    int hit_ref[POOL];
    unsigned char tq[POOL]; // trace stack (slots whose ray awaits traversal)
    unsigned char sq[POOL]; // shade stack (slots whose segment is traced, or fresh)
};

template <bool COUNT_RAYS>
__global__ void __launch_bounds__(RENDER_TPB, 3) k_render_pool(const RenderArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpPool &wp = reinterpret_cast<WarpPool *>(smem_raw)[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const DeviceScene &s = a.scene;
    const float4 *__restrict__ nodes = s.nodes;
    const float4 *__restrict__ leaves = s.leaves;
    const int2 *__restrict__ info = s.leaf_info;

    // every slot starts fresh on the shade stack
    for (int k = lane; k < POOL; k += 32) {
        wp.sq[k] = (unsigned char)k;
        wp.hit_ref[k] = SLOT_FRESH;
    }
    __syncwarp();
    // stack heights are warp-uniform and live in registers: every lane derives them from the same ballots
    int tq_n = 0, sq_n = POOL;

    // lane state: the traversal in flight (slot < 0: idle, and then cur == TRAV_DONE)
    int slot = -1;
    Ray ray;
    RayPre pre;
    int cur = TRAV_DONE, sp = 0;
    int stack[RRTB_STACK];
    Hit best;
    best.t = 0.f;
    best.ref = -1;
    best.obj = -1;
    bool queue_empty = false; // the global work queue has run dry (warp-uniform)
    unsigned long long rays = 0, hits = 0;
    TravCounters tc = {0ull, 0ull, 0ull, 0ull};

    while (true) {
        const unsigned idle_mask = __ballot_sync(0xffffffffu, slot < 0);
        const int n_idle = __popc(idle_mask);

        if (n_idle >= a.th_fetch && tq_n > 0) {
            // ---------------- FETCH: idle lanes pop slots off the trace stack
            const int take = min(n_idle, tq_n);
            const int rank = __popc(idle_mask & lt_mask);
            if (slot < 0 && rank < take) {
                slot = wp.tq[tq_n - 1 - rank];
                ray.ox = wp.ox[slot]; ray.oy = wp.oy[slot]; ray.oz = wp.oz[slot];
                ray.dx = wp.dx[slot]; ray.dy = wp.dy[slot]; ray.dz = wp.dz[slot];
                ray.tm = wp.tm[slot];
                pre = ray_pre(ray);
                best.t = __int_as_float(0x7f800000);
                best.ref = -1;
                cur = 0;
                sp = 0;
            }
            tq_n -= take;
        }
        else if (sq_n >= 32 || (sq_n > 0 && tq_n == 0 && n_idle >= a.th_shade)) {
            // ---------------- SHADE: one batch of up to 32 slots, one per lane
            const int take = min(32, sq_n);
            const bool mine = (int)lane < take;
            int sl = mine ? wp.sq[sq_n - 1 - lane] : -1;
            sq_n -= take;
            __syncwarp(); // every lane has read its slot id before the stacks are pushed to below
            bool want_new = false;   // path ended (or slot fresh): needs the next camera path
            bool to_trace = false;   // slot has a ray to trace
            int pixel = 0, sample = 0;
            if (mine) {
                const int href = wp.hit_ref[sl];
                if (href == SLOT_FRESH) {
                    want_new = true;
                }
                else {
                    Ray r;
                    r.ox = wp.ox[sl]; r.oy = wp.oy[sl]; r.oz = wp.oz[sl];
                    r.dx = wp.dx[sl]; r.dy = wp.dy[sl]; r.dz = wp.dz[sl];
                    r.tm = wp.tm[sl];
                    float thr_r = wp.tr[sl], thr_g = wp.tg[sl], thr_b = wp.tb[sl];
                    pixel = wp.pixel[sl];
                    sample = wp.sample[sl];
                    int bounce = wp.bounce[sl];
                    if (COUNT_RAYS) {
                        ++rays;
                        if (href >= 0) ++hits;
                    }
                    float lr = 0.f, lg = 0.f, lb = 0.f;
                    bool path_end = false;
                    if (href < 0) { // sky (rrt.cu:68-75)
                        float cr, cg, cb;
                        sky(r, cr, cg, cb);
                        lr = thr_r * cr;
                        lg = thr_g * cg;
                        lb = thr_b * cb;
                        path_end = true;
                    }
                    else {
                        Hit h;
                        h.t = wp.hit_t[sl];
                        h.ref = href;
                        h.obj = -1;
                        HitRecord rec = hit_record(leaves, info, r, h);
                        uint4 rnd = philox4x32_10(make_uint4((uint32_t)pixel, (uint32_t)sample, 2u + (uint32_t)bounce, 0u), a.key);
                        float4 m = __ldg(&s.materials[rec.mat]);
                        int mtype = __ldg(&s.material_type[rec.mat]);
                        float dx, dy, dz, ar, ag, ab;
                        if (scatter(mtype, m, r, rec, rnd, dx, dy, dz, ar, ag, ab)) {
                            if (++bounce >= a.max_depth) {
                                path_end = true; // exceeded depth: black (rrt.cu:78)
                            }
                            else {
                                wp.ox[sl] = rec.px; wp.oy[sl] = rec.py; wp.oz[sl] = rec.pz;
                                wp.dx[sl] = dx; wp.dy[sl] = dy; wp.dz[sl] = dz;
                                wp.tr[sl] = thr_r * ar; wp.tg[sl] = thr_g * ag; wp.tb[sl] = thr_b * ab;
                                wp.bounce[sl] = bounce;
                                to_trace = true;
                            }
                        }
                        else {
                            path_end = true; // absorbed: black (rrt.cu:61-63)
                        }
                    }
                    if (path_end) {
                        unsigned long long *dst = a.accum + 3ull * (unsigned long long)pixel;
                        const unsigned long long fr = to_fixed(lr), fg = to_fixed(lg), fb = to_fixed(lb);
                        if (fr) atomicAdd(dst + 0, fr);
                        if (fg) atomicAdd(dst + 1, fg);
                        if (fb) atomicAdd(dst + 2, fb);
                        want_new = true;
                    }
                }
            }
            // next camera paths for the slots that ended: one warp-aggregated atomic on the global queue
            unsigned new_mask = __ballot_sync(0xffffffffu, want_new);
            if (new_mask && !queue_empty) {
                unsigned long long base = 0;
                const int leader = __ffs(new_mask) - 1;
                if ((int)lane == leader) base = atomicAdd(a.queue, (unsigned long long)__popc(new_mask));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (base + (unsigned long long)__popc(new_mask) >= a.n_items) queue_empty = true;
                // decode (local tile, local sample) of the first item once per warp (64-bit division), the
                // other lanes are at most one (tile, sample) group further on
                const unsigned long long ts0 = base >> 5;
                int ls0 = 0, ltile0 = 0;
                if ((int)lane == leader) {
                    ltile0 = (int)(ts0 / (unsigned long long)a.n_local_samples);
                    ls0 = (int)(ts0 - (unsigned long long)ltile0 * (unsigned long long)a.n_local_samples);
                }
                ls0 = __shfl_sync(0xffffffffu, ls0, leader);
                ltile0 = __shfl_sync(0xffffffffu, ltile0, leader);
                if (want_new) {
                    const unsigned long long item = base + __popc(new_mask & lt_mask);
                    if (item < a.n_items) {
                        // item = (local_tile * n_local_samples + local_sample) * 32 + pixel_in_tile
                        const unsigned pit = (unsigned)(item & 31ull);
                        int ls = ls0 + (int)((item >> 5) - ts0), ltile = ltile0;
                        if (ls >= a.n_local_samples) {
                            ls -= a.n_local_samples;
                            ++ltile;
                        }
                        const int tile = a.shard_mode == RRTB_SHARD_TILES ? ltile * a.world + a.rank : ltile;
                        int ty = (int)__fdividef((float)tile, (float)a.tiles_x); // float estimate, off by <= 1 ...
                        int tx = tile - ty * a.tiles_x;
                        if (tx < 0) { // ... corrected exactly in integers
                            --ty;
                            tx += a.tiles_x;
                        }
                        else if (tx >= a.tiles_x) {
                            ++ty;
                            tx -= a.tiles_x;
                        }
                        const int i = tx * 8 + (int)(pit & 7u), j = ty * 4 + (int)(pit >> 3);
                        if (i < a.W && j < a.H) { // else: padding pixel of an edge tile -> slot stays fresh
                            pixel = j * a.W + i;
                            sample = a.shard_mode == RRTB_SHARD_SAMPLES ? ls * a.world + a.rank : ls;
                            Ray r = camera_ray(a.cam, a.W, a.H, i, j, sample, a.key);
                            wp.ox[sl] = r.ox; wp.oy[sl] = r.oy; wp.oz[sl] = r.oz;
                            wp.dx[sl] = r.dx; wp.dy[sl] = r.dy; wp.dz[sl] = r.dz;
                            wp.tm[sl] = r.tm;
                            wp.tr[sl] = 1.f; wp.tg[sl] = 1.f; wp.tb[sl] = 1.f;
                            wp.pixel[sl] = pixel;
                            wp.sample[sl] = sample;
                            wp.bounce[sl] = 0;
                            to_trace = true;
                            want_new = false;
                        }
                    }
                }
            }
            // route the slots: traced next / fresh again (retry while the queue has work) / dead
            const bool refresh = want_new && !queue_empty; // padding pixel: ask again
            const bool die = want_new && queue_empty;
            if (want_new) wp.hit_ref[sl] = SLOT_FRESH;
            const unsigned t_mask = __ballot_sync(0xffffffffu, to_trace);
            const unsigned f_mask = __ballot_sync(0xffffffffu, refresh);
            const unsigned d_mask = __ballot_sync(0xffffffffu, die);
            (void)d_mask;
            if (to_trace) wp.tq[tq_n + __popc(t_mask & lt_mask)] = (unsigned char)sl;
            if (refresh) wp.sq[sq_n + __popc(f_mask & lt_mask)] = (unsigned char)sl;
            tq_n += __popc(t_mask);
            sq_n += __popc(f_mask);
            __syncwarp(); // slot contents + stack entries visible to the lanes that will pop them
        }
        else if (n_idle == 32) {
            break; // nothing in flight, nothing to fetch (tq_n == 0), nothing to shade (sq_n == 0): all slots are dead
        }
        else {
            // ---------------- STEP: a few node visits, then (by vote) one exact leaf test
#pragma unroll 1
            for (int it = 0; it < a.step_iters; ++it)
                if (cur >= 0) node_step<COUNT_RAYS>(nodes, pre, 0.001f, best.t, cur, sp, stack, tc);
            const bool at_leaf = cur < 0 && cur != TRAV_DONE;
            const unsigned leaf_mask = __ballot_sync(0xffffffffu, at_leaf);
            if (leaf_mask) {
                const unsigned node_mask = __ballot_sync(0xffffffffu, cur >= 0);
                if (__popc(leaf_mask) >= a.th_leaf || node_mask == 0u) {
                    if (at_leaf) leaf_step<COUNT_RAYS>(leaves, info, ray, pre, 0.001f, best, cur, sp, stack, tc);
                }
            }
            // lanes whose traversal ended publish the hit and go idle
            const bool fin = slot >= 0 && cur == TRAV_DONE;
            const unsigned fin_mask = __ballot_sync(0xffffffffu, fin);
            if (fin_mask) {
                if (fin) {
                    wp.hit_t[slot] = best.t;
                    wp.hit_ref[slot] = best.ref; // -1 = miss
                    wp.sq[sq_n + __popc(fin_mask & lt_mask)] = (unsigned char)slot;
                    slot = -1;
                }
                sq_n += __popc(fin_mask);
                __syncwarp();
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

} // namespace rrtb
