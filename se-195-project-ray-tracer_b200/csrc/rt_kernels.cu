// rt_kernels.cu -- the two sm_100a kernels of librt_b200.so.
//
// Both kernels have the same shape (DESIGN.md "Kernel design"):
//   * persistent CTAs: the grid is (SM count x resident CTAs per SM); every warp pulls work items
//     (pixels, walked in 8x4 screen blocks of the rows this rank owns) from one global counter with a
//     warp-aggregated atomicAdd -- one atomic per refill, not per lane;
//   * one lane = one pixel for its whole life (all sample passes / all nine sub-samples), because the
//     reference's per-pixel RNG stream and float accumulation order are sequential;
//   * the loop body is "one ray query, then advance": every lane that has work tests its current ray
//     (nearest-hit or shadow) against the SAME primitive at the same time, so the primitive record is
//     one broadcast LDS.128 from the structure-of-arrays copy staged in shared memory, and lanes at
//     different bounces / passes / pixels still share the intersection loop.  Shadow lanes drop out
//     at their first blocker; a warp vote ends the loop early once every lane is done;
//   * a lane that finishes its pixel refills on the next iteration, so divergence in path length
//     costs idle lanes only at the very end of the frame;
//   * no tensor cores: the work is branchy FP32 with single-rounded (un-fused) arithmetic, which is
//     what makes hit IDs and RNG streams bit-identical to the reference's CPU path.
#include <cuda_runtime.h>
#include <stdint.h>
#include "rt_kernels.h"

namespace rtb {

#define FULL_MASK 0xffffffffu

// Warp-aggregated work fetch: lanes with `need` set receive consecutive item numbers.
__device__ __forceinline__ uint32_t fetch_items(unsigned *counter, bool need, uint32_t lane) {
    const uint32_t m = __ballot_sync(FULL_MASK, need);
    if (m == 0) return 0xffffffffu;
    const int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(counter, (unsigned)__popc(m));
    base = __shfl_sync(FULL_MASK, base, leader);
    return base + (uint32_t)__popc(m & ((1u << lane) - 1u));
}

__device__ __forceinline__ uint64_t warp_sum(uint64_t v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL_MASK, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------
// smallpt: replaces RadianceGPU (SPT/rendering_kernel.cl:53-97, rendering_kernel_dl.cl) with the pass
// loop inside the kernel.  CHUNKED = the (p, rad^2) array does not fit in shared memory: the CTA walks
// it in chunks, all warps in lock-step (one __syncthreads pair per chunk per query round).
// ALIGNED = the warp runs the shading STEPS of pt_lane.cuh in lock-step: a nearest round for the lanes that need one,
// then light sample -> shadow round -> accumulate per light, then the bounce, each step executed by all the lanes that
// are at it.  With few spheres (Cornell: 9) the shading between two queries is most of the work, and in the plain
// "one query, then advance" loop lanes at different steps serialise it (9 active lanes per instruction, ncu); aligned,
// a shadow lane no longer shares a round with nearest lanes, which costs idle lanes in the sphere loop instead -- the
// right trade only while the loop is short.  Per lane the sequence of operations is the same, so are the results.
template <bool COUNT, bool CHUNKED, bool ALIGNED>
__global__ void __launch_bounds__(PT_THREADS, PT_MIN_BLOCKS)
pt_kernel(PtFrame F, Shard S, uint32_t n_items, float *colors, uint32_t *seeds, uint32_t *pixels,
          unsigned *work_counter, unsigned long long *counters, int chunk) {
    extern __shared__ f4 s_geom[];
    const uint32_t lane = threadIdx.x & 31u;

    if (!CHUNKED) {
        for (int i = threadIdx.x; i < F.n; i += blockDim.x) s_geom[i] = F.geom_global[i];
        __syncthreads();
        F.geom_global = s_geom;     // shading reads of (p, rad^2) also come from the staged copy
    }

    PtLane L;
    L.phase = PH_IDLE;
    L.c_nearest = L.c_shadow = L.c_samples = 0; L.c_tests = 0;
    bool exhausted = false;

    for (;;) {
        const bool need = (L.phase == PH_IDLE) && !exhausted;
        const uint32_t item = fetch_items(work_counter, need, lane);
        if (need) {
            if (item < n_items) {
                int x, y;
                if (item_to_pixel(S, F.w, item, x, y)) pt_begin_pixel(L, F, x, y, colors, seeds);
            } else exhausted = true;
        }
        const bool active = L.phase != PH_IDLE;
        const bool more = active || !exhausted;
        if (CHUNKED) { if (!__syncthreads_or(more)) break; }
        else         { if (!__any_sync(FULL_MASK, more)) break; }

        if (ALIGNED) {
            // (One shared copy of the sphere loop for the nearest and the shadow round -- 30 KB of SASS instead of 35 KB -- was
            // measured 4 % slower, 9.83 against 9.41 ms for 32 spp: the two inlined copies stay.)
            const bool nq = L.phase == PH_NEAREST;
            pt_query_range<COUNT>(L, s_geom, 0, F.n, nq);
            if (nq) pt_hit<COUNT>(L, F);
            while (__any_sync(FULL_MASK, L.phase == PH_LIGHTS)) {
                if (L.phase == PH_LIGHTS) pt_light_step(L, F);
                const bool sq = L.phase == PH_SHADOW;
                if (__any_sync(FULL_MASK, sq)) {
                    pt_query_range<COUNT>(L, s_geom, 0, F.n, sq);
                    if (sq) pt_light_done<COUNT>(L, F);
                }
            }
            if (L.phase == PH_DIFFUSE) pt_diffuse_bounce(L, F);
            if (L.phase == PH_BOUNCE) pt_bounce(L);
            if (L.phase == PH_END && pt_end_sample<COUNT>(L, F)) {
                const size_t i = (size_t)(F.h - L.y - 1) * F.w + L.x;
                colors[3 * i] = L.cr; colors[3 * i + 1] = L.cg; colors[3 * i + 2] = L.cb;
                seeds[2 * i] = L.s0; seeds[2 * i + 1] = L.s1;
                if (!F.sum_mode && !F.defer_pack) pixels[(size_t)L.y * F.w + L.x] = pt_pack_pixel(L.cr, L.cg, L.cb);
            }
            continue;
        }
        if (!CHUNKED) {
            pt_query_range<COUNT>(L, s_geom, 0, F.n, active);
        } else {
            for (int hi = F.n; hi > 0; hi -= chunk) {          // descending index, chunk by chunk
                const int lo = hi > chunk ? hi - chunk : 0;
                __syncthreads();
                RT_CHECK(hi - lo <= chunk && lo >= 0 && hi <= F.n, RT_CHK_STAGING);
                for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) s_geom[i - lo] = F.geom_global[i];
                __syncthreads();
                pt_query_range<COUNT>(L, s_geom, lo, hi, active);
            }
        }

        if (active && pt_advance<COUNT>(L, F)) {
            const size_t i = (size_t)(F.h - L.y - 1) * F.w + L.x;
            colors[3 * i] = L.cr; colors[3 * i + 1] = L.cg; colors[3 * i + 2] = L.cb;
            seeds[2 * i] = L.s0; seeds[2 * i + 1] = L.s1;
            if (!F.sum_mode && !F.defer_pack) pixels[(size_t)L.y * F.w + L.x] = pt_pack_pixel(L.cr, L.cg, L.cb);
        }
    }

    if (COUNT) {
        const uint64_t a = warp_sum(L.c_nearest), b = warp_sum(L.c_shadow), c = warp_sum(L.c_tests), d = warp_sum(L.c_samples);
        if (lane == 0) {
            atomicAdd(&counters[0], (unsigned long long)a); atomicAdd(&counters[1], (unsigned long long)b);
            atomicAdd(&counters[2], (unsigned long long)c); atomicAdd(&counters[4], (unsigned long long)d);
        }
    }
}

// smallpt on a large scene: the same lane state machine, but a query walks the exact hierarchy of pt_bvh.cuh instead of
// all spheres (19 533 spheres: ~4 sphere tests and ~10 node visits per query instead of 19 533 tests; same colours, RNG
// state and pixels).  The traversal is per lane (its own stack in local memory, nodes and spheres read through L1 / L2)
// and its length has a long tail (median 3 node visits, 1 % beyond 50), so a warp neither waits for its slowest query
// nor shades lane by lane.  Every lane is in one of three states -- at an inner node, at a leaf, or holding a finished
// query -- and each iteration the warp runs ONE kind of step, for all the lanes that are in that state:
//   shade     once PT_BVH_SHADE_LANES lanes hold a finished query (or nothing else is left to do): the shading STEPS of
//             pt_lane.cuh in lock-step (hit, light samples, bounce, end of sample) as in pt_kernel<ALIGNED>, then the refill
//             of lanes whose pixel is complete, then the start of the new queries (always-tested spheres, root box);
//   leaf      when more lanes wait at a leaf than at an inner node: the exact sphere tests of the leaf, then a pop;
//   inner     otherwise: two box tests, descend / push / pop.
// Lanes in the other states keep their state and wait.  The reference-order loop of pt_kernel remains the path of small
// scenes, where the whole array sits in shared memory and a test costs less than a node visit, and of counting launches.
__global__ void __launch_bounds__(PT_THREADS, PT_BVH_MIN_BLOCKS)
pt_bvh_kernel(PtFrame F, PtBvh B, Shard S, uint32_t n_items, float *colors, uint32_t *seeds, uint32_t *pixels, unsigned *work_counter) {
    const uint32_t lane = threadIdx.x & 31u;
    int stack[PT_BVH_STACK];
    float stack_t[PT_BVH_STACK];
    PtLane L;
    PtTrav T;
    T.node = PT_BVH_DONE; T.sp = 0;
    L.phase = PH_IDLE;
    L.c_nearest = L.c_shadow = L.c_samples = 0; L.c_tests = 0;
    bool exhausted = false;
    for (;;) {
        const bool inner = pt_bvh_at_inner(T), leaf = pt_bvh_at_leaf(T);
        const bool fin = T.node == PT_BVH_DONE && L.phase != PH_IDLE;
        const int ni = __popc(__ballot_sync(FULL_MASK, inner)), nl = __popc(__ballot_sync(FULL_MASK, leaf));
        const int nf = __popc(__ballot_sync(FULL_MASK, fin));        // (a separate quorum per kind of finished query was 12-20 % slower)
        if (nf >= PT_BVH_SHADE_LANES || ni + nl == 0) {
            if (nf == 0 && !__any_sync(FULL_MASK, !exhausted)) break;              // nothing in flight, nothing left to fetch
            if (fin && L.phase == PH_NEAREST) pt_hit<false>(L, F);
            else if (fin && L.phase == PH_SHADOW) pt_light_done<false>(L, F);
            while (__any_sync(FULL_MASK, fin && L.phase == PH_LIGHTS))
                if (fin && L.phase == PH_LIGHTS) pt_light_step(L, F);             // -> PH_SHADOW (a query), the next light, or past the last one
            if (fin && L.phase == PH_DIFFUSE) pt_diffuse_bounce(L, F);
            if (fin && L.phase == PH_BOUNCE) pt_bounce(L);
            if (fin && L.phase == PH_END && pt_end_sample<false>(L, F)) {
                const size_t i = (size_t)(F.h - L.y - 1) * F.w + L.x;
                colors[3 * i] = L.cr; colors[3 * i + 1] = L.cg; colors[3 * i + 2] = L.cb;
                seeds[2 * i] = L.s0; seeds[2 * i + 1] = L.s1;
                if (!F.sum_mode && !F.defer_pack) pixels[(size_t)L.y * F.w + L.x] = pt_pack_pixel(L.cr, L.cg, L.cb);
            }
            const bool need = (L.phase == PH_IDLE) && !exhausted;
            const uint32_t item = fetch_items(work_counter, need, lane);
            if (need) {
                if (item < n_items) {
                    int x, y;
                    if (item_to_pixel(S, F.w, item, x, y)) pt_begin_pixel(L, F, x, y, colors, seeds);
                } else exhausted = true;
            }
            if ((fin || need) && (L.phase == PH_NEAREST || L.phase == PH_SHADOW)) pt_bvh_begin<false>(L, B, T);
        } else if (nl > ni) {
            if (leaf) pt_bvh_leaf<false>(L, B, T, stack, stack_t);
        } else {
            if (inner) pt_bvh_inner(L, B, T, stack, stack_t);        // (leaf children tested inside this step instead: 8-25 % slower)
        }
    }
}

// The sin / cos table of rt_math.cuh: 2^23 entries made by the function they stand in for.
__global__ void sincos_table_kernel(float *tab) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < (1u << 23); i += gridDim.x * blockDim.x) {
        float sn, cs;
        sincos_glibc(sincos_table_angle(i), &sn, &cs);
        tab[2 * i] = sn; tab[2 * i + 1] = cs;
    }
}

// toInt of every pixel this rank owns (SPT/smallptCPU.cpp:120-122), after the render kernel: one thread per pixel, all lanes busy.
__global__ void pt_pack_kernel(Shard S, int w, int h, uint32_t n_items, const float *colors, uint32_t *pixels) {
    for (uint32_t item = blockIdx.x * blockDim.x + threadIdx.x; item < n_items; item += gridDim.x * blockDim.x) {
        int x, y;
        if (!item_to_pixel(S, w, item, x, y)) continue;
        const size_t i = (size_t)(h - y - 1) * w + x;
        pixels[(size_t)y * w + x] = pt_pack_pixel(colors[3 * i], colors[3 * i + 1], colors[3 * i + 2]);
    }
}

// Sample-sharded mode: colors hold sums; divide and write the 8-bit pixels.
__global__ void pt_resolve_kernel(const float *colors, uint32_t *pixels, int w, int h, float inv_total) {
    const size_t n = (size_t)w * h;
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(p / w), x = (int)(p % w);
        const size_t i = (size_t)(h - y - 1) * w + x;
        pixels[p] = pt_pack_pixel(f_mul(colors[3 * i], inv_total), f_mul(colors[3 * i + 1], inv_total), f_mul(colors[3 * i + 2], inv_total));
    }
}

// Stages the scene tables of a WFrame into shared memory and redirects the frame's pointers to the copies.
// layout: geom[n] | mat_a[n] | mat_b[n] | rbox[2*n_runs] (MODE 2) | pcull[n] (MODE 2) | flags[n] | runs[3*n_runs] | rrad[n] (MODE 2) | lights[n_lights] (MODE 2)
// MODE 2: everything on chip (small scenes; the compiler then knows the material pointers are shared-memory pointers:
// LDS instead of generic loads).  MODE 1: geometry, flags and runs only.  MODE 0: nothing is staged -- a scene larger
// than the per-CTA share of shared memory is read through L1 / L2 (every lane of a warp reads the same record).
// MODE 3: like 2, for scenes of at most W_TAB_CAP primitives in at most W_TAB_RUNS runs (the reference's scenes): the tables sit in a
// STATIC shared-memory struct, so every table address is a compile-time constant -- with the dynamic layout of mode 2 the
// offsets depend on n, and under the register cap the compiler re-derived them at every use (7 % of the executed
// instructions of the 1080p frame were that address arithmetic, plus an S2R / LEA per loop iteration).
struct WTables {
    f4 geom[W_TAB_CAP], ma[W_TAB_CAP], mb[W_TAB_CAP], rbox[2 * W_TAB_RUNS];
    f2 pcull[W_TAB_CAP];
    float rrad[W_TAB_CAP];
    int flags[W_TAB_CAP], lights[W_TAB_CAP], runs[3 * W_TAB_RUNS];
};
template <int MODE>
__device__ __forceinline__ void stage_scene(WFrame &F, f4 *s_raw, const f4 *&s_geom_out, const int *&s_runs_out) {
    if (MODE == 0) { s_geom_out = F.geom; s_runs_out = F.runs; return; }
    if (MODE == 3) {
        __shared__ WTables T;
        RT_CHECK(F.n <= W_TAB_CAP && F.n_runs <= W_TAB_RUNS && F.n_lights <= W_TAB_CAP, RT_CHK_STAGING);
        const bool cull = F.pcull != nullptr;
        for (int i = threadIdx.x; i < F.n; i += blockDim.x) {
            T.geom[i] = F.geom[i]; T.ma[i] = F.mat_a[i]; T.mb[i] = F.mat_b[i]; T.rrad[i] = F.rrad[i]; T.flags[i] = F.flags[i];
            if (cull) T.pcull[i] = F.pcull[i];
        }
        for (int i = threadIdx.x; i < 3 * F.n_runs; i += blockDim.x) T.runs[i] = F.runs[i];
        for (int i = threadIdx.x; i < F.n_lights; i += blockDim.x) T.lights[i] = F.lights[i];
        if (cull) for (int i = threadIdx.x; i < 2 * F.n_runs; i += blockDim.x) T.rbox[i] = F.rbox[i];
        __syncthreads();
        F.geom = T.geom; F.mat_a = T.ma; F.mat_b = T.mb; F.rrad = T.rrad; F.flags = T.flags; F.lights = T.lights;
        if (cull) { F.pcull = T.pcull; F.rbox = T.rbox; }
        s_geom_out = T.geom; s_runs_out = T.runs;
        return;
    }
    const int n = F.n;
    f4 *s_geom = s_raw;
    f4 *s_ma = s_geom + n, *s_mb = s_ma + n, *s_box = s_mb + n;
    f2 *s_pc = (f2 *)(s_box + 2 * F.n_runs);
    int *ibase = MODE == 2 ? (int *)(s_pc + n) : (int *)(s_geom + n);
    int *s_flags = ibase;
    int *s_runs = ibase + n;
    float *s_rr = (float *)(s_runs + 3 * F.n_runs);
    int *s_li = (int *)(s_rr + n);
    const bool cull = MODE == 2 && F.pcull != nullptr;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        s_geom[i] = F.geom[i];
        s_flags[i] = F.flags[i];
        if (MODE == 2) { s_ma[i] = F.mat_a[i]; s_mb[i] = F.mat_b[i]; s_rr[i] = F.rrad[i]; if (cull) s_pc[i] = F.pcull[i]; }
    }
    for (int i = threadIdx.x; i < 3 * F.n_runs; i += blockDim.x) s_runs[i] = F.runs[i];
    if (MODE == 2) {
        for (int i = threadIdx.x; i < F.n_lights; i += blockDim.x) s_li[i] = F.lights[i];
        if (cull) for (int i = threadIdx.x; i < 2 * F.n_runs; i += blockDim.x) s_box[i] = F.rbox[i];
    }
    __syncthreads();
    F.geom = s_geom; F.flags = s_flags;
    if (MODE == 2) { F.mat_a = s_ma; F.mat_b = s_mb; F.rrad = s_rr; F.lights = s_li; if (cull) { F.pcull = s_pc; F.rbox = s_box; } }
    s_geom_out = s_geom; s_runs_out = s_runs;
}

// ------------------------------------------------------------------------------------------------
// Whitted: replaces raytracer_kernel (R323/raytracer_kernel.cl:246-383) with the numerics of its CPU
// twin.  The reference copies the 96-byte AoS primitives to local memory (:254-258) and keeps a
// 64-slot x 80-byte ray queue in private memory; here the primitives are SoA float4 in shared memory
// and the FIFO is 32 slots x 48 bytes (the most a breadth-first walk of a depth-5 binary tree holds).
// NL > 0: the scene has exactly NL lights and all of them are spheres (straight-line shadow set-up, one batch).
// The tree part of a Whitted query round, warp-synchronous: every lane with a query descends through inner nodes until
// all of them hold a leaf (or are done), then the leaves are tested together (same results as w_bvh_nearest /
// w_bvh_blocked, which run one lane to the end).
__device__ __forceinline__ void w_bvh_nearest_round(WLane &L, const PtBvh &B, bool q) {
    if (B.root == PT_BVH_NONE) return;
    int stack[PT_BVH_STACK];
    float stack_t[PT_BVH_STACK];
    PtTrav T;
    T.sp = 0; T.node = PT_BVH_DONE;
    if (q) { T.R = bvh_ray(L.qox, L.qoy, L.qoz, L.qdx, L.qdy, L.qdz, B); T.node = bvh_root(L.qox, L.qoy, L.qoz, L.cumu, B, T.R); }
    while (__any_sync(FULL_MASK, T.node != PT_BVH_DONE)) {
        for (;;) {
            const bool inner = pt_bvh_at_inner(T);
            if (!__any_sync(FULL_MASK, inner)) break;
            if (inner) bvh_inner(L.qox, L.qoy, L.qoz, L.cumu, B, T, stack, stack_t);
        }
        if (pt_bvh_at_leaf(T)) {
            const int code = ~T.node, first = code >> 3, count = (code & 7) + 1;
#pragma unroll 1
            for (int j = 0; j < count; j++) w_bvh_sphere_nearest(L, B.geom[first + j], B.index[first + j]);
            bvh_pop(L.cumu, T, stack, stack_t);
        }
    }
}
__device__ __forceinline__ void w_bvh_shadow_round(WLane &L, const PtBvh &B, bool q) {
    if (B.root == PT_BVH_NONE) return;
    int stack[PT_BVH_STACK];
    float stack_t[PT_BVH_STACK];
#pragma unroll 1
    for (int k = 0; k < W_SHADOW_BATCH; k++) {
        const bool qk = q && k < L.ns && !((L.sblk >> k) & 1);
        if (!__any_sync(FULL_MASK, qk)) continue;
        const float ox = k == 0 ? L.sox[0] : (k == 1 ? L.sox[1] : L.sox[2]), oy = k == 0 ? L.soy[0] : (k == 1 ? L.soy[1] : L.soy[2]);
        const float oz = k == 0 ? L.soz[0] : (k == 1 ? L.soz[1] : L.soz[2]), dx = k == 0 ? L.slx[0] : (k == 1 ? L.slx[1] : L.slx[2]);
        const float dy = k == 0 ? L.sly[0] : (k == 1 ? L.sly[1] : L.sly[2]), dz = k == 0 ? L.slz[0] : (k == 1 ? L.slz[1] : L.slz[2]);
        const float reach = k == 0 ? L.sreach[0] : (k == 1 ? L.sreach[1] : L.sreach[2]);
        PtTrav T;
        T.sp = 0; T.node = PT_BVH_DONE;
        if (qk) { T.R = bvh_ray(ox, oy, oz, dx, dy, dz, B); T.node = bvh_root(ox, oy, oz, reach, B, T.R); }
        while (__any_sync(FULL_MASK, T.node != PT_BVH_DONE)) {
            for (;;) {
                const bool inner = pt_bvh_at_inner(T);
                if (!__any_sync(FULL_MASK, inner)) break;
                if (inner) bvh_inner(ox, oy, oz, reach, B, T, stack, stack_t);
            }
            if (pt_bvh_at_leaf(T)) {
                const int code = ~T.node, first = code >> 3, count = (code & 7) + 1;
                bool blocked = false;
#pragma unroll 1
                for (int j = 0; j < count; j++) blocked = blocked | w_bvh_sphere_blocks(ox, oy, oz, dx, dy, dz, reach, B.geom[first + j]);
                if (blocked) { L.sblk |= 1 << k; T.node = PT_BVH_DONE; }
                else bvh_pop(reach, T, stack, stack_t);
            }
        }
    }
}

// BVH: the run tables hold only what is not in the hierarchy B (planes, lights, odd spheres); after them every query
// continues in the tree, lane by lane (whitted_bvh.cuh) -- scenes of hundreds to thousands of spheres.
// EXACT: the launch that follows the timed kernel (whitted_lane.cuh, "Blocked lights and the redo list"): the pixels of F.redo_list --
// or every item when more were reported than the list holds -- one per lane, blocked lights shaded as the reference shades them.
// GRID: shadow rounds take their candidates from F.grid (the host launches these variants only with a grid).
template <bool COUNT, int STAGED, int NL, bool BVH, bool EXACT = false, bool GRID = false>
__global__ void __launch_bounds__(W_THREADS, W_MIN_BLOCKS)
whitted_kernel(WFrame F, Shard S, uint32_t n_items, const uint32_t *order, const unsigned *class_counts, uint32_t n_stride,
               uint32_t *pixels, unsigned *work_counter, unsigned long long *counters, PtBvh B, const uint8_t *cls, uint32_t filler_items) {
    extern __shared__ f4 s_raw[];
    if (EXACT && *F.redo_count == 0u) return;            // nothing was reported (nearly every frame): not even the tables are staged
    const uint32_t lane = threadIdx.x & 31u;
    const f4 *s_geom; const int *s_runs;
    stage_scene<STAGED>(F, s_raw, s_geom, s_runs);

    f4 queue[3 * W_QUEUE_SLOTS];
    WLane L;
    L.phase = PH_IDLE;
    L.c_nearest = L.c_shadow = L.c_samples = 0; L.c_sphere_tests = L.c_plane_tests = 0; L.c_shadow_lit = 0;
    bool exhausted = false;
    // Two ways of handing out pixels (see whitted_classify_kernel).  Pixels of classes 0 and 1 (a refracting / reflecting
    // surface behind the centre ray: ray trees of very different sizes) go to single lanes, each lane taking the next one
    // when it is done.  Pixels of class 2 (everything else, most of the frame: 9 primary rays + their shadow rays, the same
    // cost for every pixel) are handed out as whole 8x4 SCREEN BLOCKS to whole warps once the lists are empty: a warp whose
    // 32 rays run side by side hits the same few primitives, so the votes in front of the square roots / divisions and
    // the shadow-round culls decide for the warp what they decide for a lane (with single-lane refill a warp soon holds 32
    // unrelated pixels).  `cls` == NULL: lists only.
    // The first `n_items` / 32 blocks ("filler", a launch parameter) are handed out pixel by pixel as well, after the lists: a
    // lane gets fewer than two class-0/1 pixels at 1080p, so without cheap pixels to fill in with, a warp waits for its slowest
    // lane with most lanes idle (25 of 32 lanes per instruction in that phase).
    const uint32_t n_class0 = (order && !F.split0) ? class_counts[0] : 0u;       // split0: class 0 belongs to whitted_split_kernel
    const uint32_t n_listed = (order && cls) ? n_class0 + class_counts[1] : 0u;
    const uint32_t n_filler = (order && cls) ? filler_items : 0u;            // items [0, n_filler) of class 2: by single lanes (a multiple of 32)
    const unsigned n_redo = EXACT ? *F.redo_count : 0u;
    const bool redo_all = EXACT && n_redo > F.redo_cap;
    const uint32_t n_lane_items = EXACT ? (redo_all ? n_items : n_redo) : (order && cls) ? n_listed + n_filler : (order && F.split0) ? n_items - class_counts[0] : n_items;
    const uint32_t n_blocks = cls ? n_stride >> 5 : 0u;

    for (;;) {
        const bool need = (L.phase == PH_IDLE) && !exhausted;
        const uint32_t item = fetch_items(work_counter, need, lane);
        if (need) {
            if (item < n_lane_items) {
                int x, y;
                uint32_t it = item;
                bool take = true;
                if (order) {                     // walk the cost classes in turn (see whitted_classify_kernel)
                    const uint32_t n0 = n_class0, n1 = class_counts[1];
                    RT_CHECK(n0 + n1 <= n_stride && (!cls || item - n_listed < n_stride || item < n_listed), RT_CHK_WORKLIST);
                    if (cls && item >= n_listed) { it = item - n_listed; take = cls[it] == 2; }     // filler: a class-2 pixel of the first blocks
                    else it = item < n0 ? order[item] : (item < n0 + n1 ? order[n_stride + item - n0] : order[2 * (size_t)n_stride + item - n0 - n1]);
                }
                if (EXACT && !redo_all) { RT_CHECK(item < F.redo_cap, RT_CHK_WORKLIST); const uint32_t id = F.redo_list[item]; w_begin_pixel(L, F, (int)(id % (uint32_t)F.w), (int)(id / (uint32_t)F.w)); }
                else if (take && item_to_pixel(S, F.w, it, x, y)) w_begin_pixel(L, F, x, y);
            } else exhausted = true;
        }
        if (!__any_sync(FULL_MASK, L.phase != PH_IDLE || !exhausted)) {
            // every lane is idle and the lists are empty: the next block of class-2 pixels, for the whole warp
            if (!cls) break;
            uint32_t blk = 0;
            if (lane == 0) blk = atomicAdd(work_counter + 1, 1u);
            blk = __shfl_sync(FULL_MASK, blk, 0) + (n_filler >> 5);
            if (blk >= n_blocks) break;
            const uint32_t it = blk * 32u + lane;
            RT_CHECK(it < n_stride, RT_CHK_WORKLIST);
            int x, y;
            if (cls[it] == 2 && item_to_pixel(S, F.w, it, x, y)) w_begin_pixel(L, F, x, y);
            if (!__any_sync(FULL_MASK, L.phase != PH_IDLE)) continue;       // a block without such pixels
        }

        // round 1: every lane with a ray finds its nearest hit; round 2 (repeated while lights remain): every lane
        // that hit a surface tests up to three shadow rays at once; then the finished rays are folded into their pixels.
        const bool nq = L.phase == PH_NEAREST;
#ifdef W_ROUND_STATS      /* debug build: lane participation per round, reported through the counting launch's counters */
        if (COUNT && lane == 0) { atomicAdd(&counters[0], 32ull); atomicAdd(&counters[1], (unsigned long long)__popc(__ballot_sync(FULL_MASK, nq))); }
        else if (COUNT) __ballot_sync(FULL_MASK, nq);
#endif
        if (GRID) w_query_nearest_tiles(L, s_geom, F.flags, s_runs, F.n_runs, nq, F.grid);       // primary rays: candidates from the tile's word
        else w_query_nearest<COUNT>(L, s_geom, s_runs, F.n_runs, nq);
        if (BVH) w_bvh_nearest_round(L, B, nq);
        if (nq) w_after_nearest<COUNT, NL, EXACT>(L, F);
        while (__any_sync(FULL_MASK, L.phase == PH_SHADOW)) {
            const bool sq = L.phase == PH_SHADOW;
#ifdef W_ROUND_STATS
            { const unsigned m = __ballot_sync(FULL_MASK, sq); if (COUNT && lane == 0) { atomicAdd(&counters[2], 32ull); atomicAdd(&counters[3], (unsigned long long)__popc(m)); } }
#endif
            if (GRID) w_query_shadow_grid(L, s_geom, F.flags, sq, F.grid, F.reject_k);     // small scenes: candidates from the grid
            else w_query_shadow<COUNT, (NL > 0 && !COUNT)>(L, s_geom, s_runs, F.n_runs, sq, F.pcull, F.rbox, F.reject_k);     // NL > 0: the host made the cull tables
            if (BVH) w_bvh_shadow_round(L, B, sq);
            if (sq) w_after_shadow<COUNT, NL, EXACT>(L, F);
        }
        if (L.phase == PH_FINAL && w_finalize<COUNT>(L, F, queue)) {
            RT_CHECK(L.x >= 0 && L.x < F.w && L.y >= 0 && L.y < F.h, RT_CHK_PIXEL);
            pixels[(size_t)L.y * F.w + L.x] = w_pack_pixel(L.ar, L.ag, L.ab);
        }
    }

#ifndef W_ROUND_STATS
    if (COUNT) {
        const uint64_t a = warp_sum(L.c_nearest), b = warp_sum(L.c_shadow), c = warp_sum(L.c_sphere_tests),
                       d = warp_sum(L.c_plane_tests), e = warp_sum(L.c_samples);
        if (lane == 0) {
            atomicAdd(&counters[0], (unsigned long long)a); atomicAdd(&counters[1], (unsigned long long)b);
            atomicAdd(&counters[2], (unsigned long long)c); atomicAdd(&counters[3], (unsigned long long)d);
            atomicAdd(&counters[4], (unsigned long long)e);
        }
        const uint64_t f = warp_sum(L.c_shadow_lit);
        if (lane == 0) atomicAdd(&counters[5], (unsigned long long)f);
    }
#endif
}

// The pixels of cost class 0 (their primary-ray tile can see a reflecting or refracting primitive), one lane per SUB-SAMPLE.  A CTA is
// nine warps and takes 32 pixels of the class list at a time: warp s traces sub-sample s of all 32 (neighbouring pixels, the same
// sub-pixel offset: 32 coherent rays, every lane busy), every lane logging what its rays add to its pixel (w_finalize<.., SPLIT>).  When
// the nine warps are done, the nine logs of each pixel are added in the reference's order: warp 0 adds its logs (32 pixels side by
// side) and leaves the running sums in shared memory, warp 1 continues from them, ... and warp 8 stores the pixels.  Same bits as one
// lane tracing the nine sub-samples in turn (whitted_kernel), but the longest chain of rays one lane traces back to back is 63 instead
// of 567: such a pixel alone took 1.7 ms of the 2.4 ms 1080p frame and bounded small frames and strong scaling outright.  (First form:
// three pixels per warp, lanes 9g..9g+8 one pixel, sums handed on by shuffles -- 27 of 32 lanes, and the sums ran at one lane of
// nine: 12 % of the kernel.)  Runs next to the main kernel on a second stream (rtk_launch_whitted); GRID tables as there.
template <int NL>
__global__ void __launch_bounds__(W_SPLIT_THREADS, W_SPLIT_MIN_BLOCKS)
whitted_split_kernel(WFrame F, Shard S, const uint32_t *order, const unsigned *class_counts, uint32_t *pixels, unsigned *work_counter) {
    extern __shared__ f4 s_raw[];
    __shared__ float s_sum[3][32];
    __shared__ uint32_t s_base;
    const uint32_t lane = threadIdx.x & 31u;
    const int sub = (int)(threadIdx.x >> 5);                 // this warp's sub-sample, 0 .. 8
    const f4 *s_geom; const int *s_runs;
    stage_scene<3>(F, s_raw, s_geom, s_runs);
    f4 queue[3 * W_QUEUE_SLOTS];
    float log[3 * 63];
    WLane L;
    L.phase = PH_IDLE;
    L.c_nearest = L.c_shadow = L.c_samples = 0; L.c_sphere_tests = L.c_plane_tests = 0; L.c_shadow_lit = 0;
    const uint32_t n0 = class_counts[0];
    for (;;) {
        if (threadIdx.x == 0) s_base = atomicAdd(work_counter, 32u);
        __syncthreads();
        const uint32_t base = s_base;
        if (base >= n0) break;                               // the same for every thread of the CTA
        bool mine = false;
        if (base + lane < n0) {
            int x, y;
            if (item_to_pixel(S, F.w, order[base + lane], x, y)) {
                L.x = x; L.y = y; L.sub = sub; L.nlog = 0;
                w_start_subsample(L, F);
                mine = true;
            }
        }
        while (__any_sync(FULL_MASK, L.phase != PH_IDLE)) {
            const bool nq = L.phase == PH_NEAREST;
            w_query_nearest_tiles(L, s_geom, F.flags, s_runs, F.n_runs, nq, F.grid);
            if (nq) w_after_nearest<false, NL, false>(L, F);
            while (__any_sync(FULL_MASK, L.phase == PH_SHADOW)) {
                const bool sq = L.phase == PH_SHADOW;
                w_query_shadow_grid(L, s_geom, F.flags, sq, F.grid, F.reject_k);
                if (sq) w_after_shadow<false, NL, false>(L, F);
            }
            if (L.phase == PH_FINAL) w_finalize<false, true>(L, F, queue, log);
        }
        // the nine logs of each pixel, in order: warp s continues the sums warp s-1 left
#pragma unroll 1
        for (int s = 0; s < 9; s++) {
            if (sub == s && mine) {
                float ar = 0.f, ag = 0.f, ab = 0.f;
                if (s > 0) { ar = s_sum[0][lane]; ag = s_sum[1][lane]; ab = s_sum[2][lane]; }
                for (int k = 0; k < L.nlog; k++) { ar = f_add(ar, log[3 * k]); ag = f_add(ag, log[3 * k + 1]); ab = f_add(ab, log[3 * k + 2]); }
                if (s < 8) { s_sum[0][lane] = ar; s_sum[1][lane] = ag; s_sum[2][lane] = ab; }
                else {
                    RT_CHECK(L.x >= 0 && L.x < F.w && L.y >= 0 && L.y < F.h, RT_CHK_PIXEL);
                    pixels[(size_t)L.y * F.w + L.x] = w_pack_pixel(ar, ag, ab);
                }
            }
            __syncthreads();
        }
    }
}

// The pixels of cost class 2 when class 0 goes to whitted_split_kernel: 8x4 blocks whose tile word holds no reflecting or refracting
// primitive, so NO primary ray of the block can have children (the word lists every primitive that can be the accepted hit).  A pixel is
// then nine primary rays, each with its nearest round (the one or two planes of the tile word), its shadow batch and its shading, added
// to the accumulator in turn -- written as that straight line instead of whitted_kernel's general state machine (no ray queue, no
// children, no phases to dispatch on): fewer registers, no local memory, 31.9 of 32 lanes per instruction.  Same lane functions, same
// order of operations, same bits.  A hit that does have children would be a broken table: checked builds test for it.
template <int NL>
__global__ void __launch_bounds__(W_THREADS, W_WALL_MIN_BLOCKS)
whitted_wall_kernel(WFrame F, Shard S, uint32_t n_stride, uint32_t *pixels, unsigned *block_counter, const uint8_t *cls) {
    extern __shared__ f4 s_raw[];
    const uint32_t lane = threadIdx.x & 31u;
    const f4 *s_geom; const int *s_runs;
    stage_scene<3>(F, s_raw, s_geom, s_runs);
    WLane L;
    L.c_nearest = L.c_shadow = L.c_samples = 0; L.c_sphere_tests = L.c_plane_tests = 0; L.c_shadow_lit = 0;
    const uint32_t n_blocks = n_stride >> 5;
    for (;;) {
        uint32_t blk = 0;
        if (lane == 0) blk = atomicAdd(block_counter, 1u);
        blk = __shfl_sync(FULL_MASK, blk, 0);
        if (blk >= n_blocks) break;
        const uint32_t it = blk * 32u + lane;
        int x = 0, y = 0;
        const bool mine = cls[it] == 2 && item_to_pixel(S, F.w, it, x, y);
        if (!__any_sync(FULL_MASK, mine)) continue;
        L.phase = PH_IDLE;
        L.x = x; L.y = y; L.ar = L.ag = L.ab = 0.f;
#pragma unroll 1
        for (int sub = 0; sub < 9; sub++) {
            if (mine) { L.sub = sub; w_start_subsample(L, F); }
            w_query_nearest_tiles(L, s_geom, F.flags, s_runs, F.n_runs, mine, F.grid);
            if (mine) w_after_nearest<false, NL, false>(L, F);
            while (__any_sync(FULL_MASK, L.phase == PH_SHADOW)) {
                const bool sq = L.phase == PH_SHADOW;
                w_query_shadow_grid(L, s_geom, F.flags, sq, F.grid, F.reject_k);
                if (sq) w_after_shadow<false, NL, false>(L, F);
            }
            if (mine) {                                   // the fold of w_finalize for a primary ray (RNO:351-356); no children by construction
                RT_CHECK(L.hit < 0 || !((F.grid.deep >> L.hit) & 1u), RT_CHK_TABLE);
                if (F.hit_ids) F.hit_ids[((size_t)L.y * F.w + L.x) * 9 + sub] = L.hit;
                L.ar = f_add(L.ar, f_mul(L.cr, L.weight));
                L.ag = f_add(L.ag, f_mul(L.cg, L.weight));
                L.ab = f_add(L.ab, f_mul(L.cb, L.weight));
                L.phase = PH_IDLE;
            }
        }
        if (mine) pixels[(size_t)L.y * F.w + L.x] = w_pack_pixel(L.ar, L.ag, L.ab);
    }
}

// ------------------------------------------------------------------------------------------------
// raytracer3.0.06 (BASELINE config 1): the frame of Engine_Render (R306/raytracer.cpp:301-530) on the GPU.  Same shape as
// the Whitted kernel -- persistent warps, one lane per pixel, a nearest round then batched shadow rounds over the
// shared-memory scene -- around the 63-node implicit ray tree of r306_lane.cuh, which lives in local memory.
// SPLIT: the work unit is one SUB-SAMPLE of a pixel (unit u = entry u / 9 of the work list, sub-sample u % 9): the nine ray
// trees of a pixel are independent, and a pixel behind the glass spheres is 9 x 63 rays traced one after the other -- as one
// unit it kept a single lane busy for most of the frame (ncu: SMs busy 56 % of the time).  The sub-sample colours go to
// `subcol`; r306_resolve_kernel adds them in the reference's order and packs the pixel.
template <bool SPLIT, int STAGED = 3>
__global__ void __launch_bounds__(W_THREADS)
r306_kernel(R306Frame F, Shard S, uint32_t n_items, const uint32_t *order, const unsigned *class_counts, uint32_t n_stride, uint32_t *dest, float *subcol, unsigned *work_counter) {
    extern __shared__ f4 s_raw[];
    const uint32_t lane = threadIdx.x & 31u;
    const f4 *s_geom; const int *s_runs;
    stage_scene<STAGED>(F.W, s_raw, s_geom, s_runs);

    R306Tree T;
    R306Lane L;
    L.q.phase = PH_IDLE;
    L.q.c_nearest = L.q.c_shadow = L.q.c_samples = 0; L.q.c_sphere_tests = L.q.c_plane_tests = 0;
    bool exhausted = false;

    for (;;) {
        const bool need = (L.q.phase == PH_IDLE) && !exhausted;
        const uint32_t item = fetch_items(work_counter, need, lane);
        if (need) {
            if (item < n_items) {
                int x, y;
                uint32_t it = SPLIT ? item / 9u : item;
                const int sub = SPLIT ? (int)(item % 9u) : 0;
                const uint32_t item = it;        // position in the work list
                if (order) {                     // expensive pixels first (whitted_classify_kernel): the frame no longer ends on the glass spheres
                    const uint32_t n0 = class_counts[0], n1 = class_counts[1];
                    it = item < n0 ? order[item] : (item < n0 + n1 ? order[n_stride + item - n0] : order[2 * (size_t)n_stride + item - n0 - n1]);
                }
                if (item_to_pixel(S, F.W.w, it, x, y) && y >= F.row0 && y < F.row1) {
                    if (SPLIT) r306_begin_subsample(L, F, x, y, sub);
                    else r306_begin_pixel(L, F, x, y);
                }
            } else exhausted = true;
        }
        const bool active = L.q.phase != PH_IDLE;
        if (!__any_sync(FULL_MASK, active || !exhausted)) break;

        const bool nq = L.q.phase == PH_NEAREST;
        w_query_nearest<false>(L.q, s_geom, s_runs, F.W.n_runs, nq);
        bool node_done = false;
        if (nq) node_done = r306_after_nearest(L, F, T);
        while (__any_sync(FULL_MASK, L.q.phase == PH_SHADOW)) {
            const bool sq = L.q.phase == PH_SHADOW;
            w_query_shadow<false>(L.q, s_geom, s_runs, F.W.n_runs, sq);
            if (sq) r306_after_shadow(L, F);
        }
        if (L.q.phase == PH_FINAL) { r306_finish_hit(L, F, T); node_done = true; }
        if (node_done && r306_next_node(L, F, T, SPLIT)) {
            if (SPLIT) {
                float *c = subcol + (((size_t)L.q.y * F.W.w + L.q.x) * 9 + (size_t)(L.q.sub - 1)) * 3;
                c[0] = L.tr; c[1] = L.tg; c[2] = L.tb;
            } else dest[(size_t)L.q.y * F.W.w + L.q.x] = r306_pack_pixel(L.tr, L.tg, L.tb);
        }
    }
}

// total_acc += the nine sub-sample colours, in order (R306:504-506), then the pixel (R306:512-520).  Rows row0 .. row1-1 of this rank.
__global__ void r306_resolve_kernel(Shard S, int w, int row0, int row1, uint32_t n_items, const float *subcol, uint32_t *dest) {
    for (uint32_t item = blockIdx.x * blockDim.x + threadIdx.x; item < n_items; item += gridDim.x * blockDim.x) {
        int x, y;
        if (!item_to_pixel(S, w, item, x, y) || y < row0 || y >= row1) continue;
        const float *c = subcol + ((size_t)y * w + x) * 27;
        float tr = 0.f, tg = 0.f, tb = 0.f;
#pragma unroll
        for (int s = 0; s < 9; s++) { tr = f_add(tr, c[3 * s]); tg = f_add(tg, c[3 * s + 1]); tb = f_add(tb, c[3 * s + 2]); }
        dest[(size_t)y * w + x] = r306_pack_pixel(tr, tg, tb);
    }
}

// Scheduling pre-pass for the Whitted frame.  A pixel is one indivisible work unit (its float accumulation
// order is fixed), and its cost varies by ~10x: a pixel that looks at a refracting sphere grows a ray tree
// of up to 63 rays per sub-sample, a pixel that looks at a wall traces 9 + 27 rays.  With pixels handed out
// in screen order the expensive ones (the spheres stand on the floor, bottom rows) come last and the frame
// ends in a long tail of warps with one busy lane (ncu: SMs idle 25 % of the frame).  This kernel traces the
// centre primary ray of every pixel (1/60 of the frame's work), classifies the pixel by the material it
// hits -- 0: refracting, 1: reflecting, 2: neither / miss -- and appends it to that class's list.  The
// render kernel then walks list 0, 1, 2: expensive pixels first, and warps that hold pixels of one class.
// It changes WHEN a pixel is rendered, never what is computed for it.
#define W_COST_CLASSES 3
__global__ void __launch_bounds__(W_THREADS)
whitted_classify_kernel(WFrame F, Shard S, uint32_t n_items, uint32_t *lists /* W_COST_CLASSES x n_items */, unsigned *class_counts, int staged, int use_bvh, PtBvh B,
                        uint8_t *cls_out /* NULL, or n_items bytes: the class of every item (255: padding); class 2 then gets no list */) {
    extern __shared__ f4 s_raw[];
    const f4 *s_geom = F.geom;
    const int *s_runs = F.runs;
    if (staged) {
        f4 *g = s_raw;
        int *r = (int *)(g + F.n);
        for (int i = threadIdx.x; i < F.n; i += blockDim.x) g[i] = F.geom[i];
        for (int i = threadIdx.x; i < 3 * F.n_runs; i += blockDim.x) r[i] = F.runs[i];
        __syncthreads();
        s_geom = g; s_runs = r;
    }
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t per_pass = gridDim.x * blockDim.x;
    for (uint32_t base = blockIdx.x * blockDim.x; base < n_items; base += per_pass) {   // warp-uniform trip count
        const uint32_t item = base + threadIdx.x;
        int x = 0, y = 0;
        const bool valid = item < n_items && item_to_pixel(S, F.w, item, x, y);
        int cls = W_COST_CLASSES - 1;
        if (F.split0) {
            // With the primary-ray tiles nothing needs tracing: a pixel is class 0 -- rendered one lane per sub-sample by whitted_split_kernel --
            // iff SOME primary ray of its 8x4 tile can reach a reflecting or refracting primitive (the tile's word), else class 2: nine
            // primary rays that spawn nothing, the same cost for every pixel of the block.  (The centre ray's first hit is a poor guide: a
            // mirror sphere shows the glass ones, a wall pixel on a silhouette has glass behind some of its sub-samples -- pixels "of
            // class 1 or 2" with ray trees of hundreds of rays kept single lanes busy for 0.8 ms of any frame.)
            if (valid && (F.grid.tiles[(y >> 2) * F.grid.tiles_x + (x >> 3)] & F.grid.deep)) cls = 0;
        } else {
            WLane L;
            L.phase = PH_IDLE; L.qhit = -1; L.cumu = 0.f; L.qkind = 0;
            L.qox = L.qoy = L.qoz = 0.f; L.qdx = L.qdy = L.qdz = 0.f;
            if (valid) { L.x = x; L.y = y; L.sub = 4; w_start_subsample(L, F); }
            w_query_nearest<false>(L, s_geom, s_runs, F.n_runs, valid);
            if (use_bvh && valid) w_bvh_nearest(L, B);
            if (valid && L.qhit >= 0) cls = F.mat_b[L.qhit].y > 0.f ? 0 : (F.mat_a[L.qhit].w > 0.f ? 1 : 2);
        }
        const uint32_t below = (1u << lane) - 1u;
        if (cls_out && item < n_items) cls_out[item] = valid ? (uint8_t)cls : (uint8_t)255;
#pragma unroll
        for (int c = 0; c < W_COST_CLASSES; c++) {
            if (cls_out && c == W_COST_CLASSES - 1) break;          // handed out by blocks, not from a list
            const uint32_t m = __ballot_sync(FULL_MASK, valid && cls == c);
            uint32_t b0 = 0;
            if (lane == 0 && m) b0 = atomicAdd(&class_counts[c], (unsigned)__popc(m));
            b0 = __shfl_sync(FULL_MASK, b0, 0);
            if (valid && cls == c) lists[(size_t)c * n_items + b0 + __popc(m & below)] = item;
        }
    }
}

// The shadow-candidate grid of a Whitted scene (whitted_lane.cuh): one thread per cell, in double.
__global__ void whitted_grid_kernel(WGrid G, int gz, uint32_t *cells, const f4 *geom, const int *flags, const f2 *pcull, const float *smargin, const f4 *lcenter, int n_lights) {
    const int n = G.gx * G.gy * gz;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x)
        cells[c] = w_grid_build_cell(G, gz, c, geom, flags, pcull, smargin, lcenter, n_lights);
}

// The primary-ray tile words of a Whitted frame (whitted_lane.cuh): one thread per 8x4-pixel tile, in double.
__global__ void whitted_tiles_kernel(uint32_t *tiles, int tiles_x, int tiles_y, int w, int h, float DX, float DY, const f4 *geom, const int *flags, const float *smargin, uint32_t all) {
    const int n = tiles_x * tiles_y;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x)
        tiles[t] = w_tile_build(t % tiles_x, t / tiles_x, w, h, DX, DY, geom, flags, smargin, all);
}

// Device-side evaluation of the elementary functions of rt_math.cuh on caller-supplied arguments, so
// that tests can compare them with the host libm (tests/test_gpu_parity.py).  Not on the rendering path.
__global__ void selftest_math_kernel(int op, const float *in, void *out, unsigned long long n) {
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const float x = in[i];
        if (op == 0) { float s, c; sincos_glibc(x, &s, &c); ((float *)out)[2 * i] = s; ((float *)out)[2 * i + 1] = c; }
        else if (op == 1) ((float *)out)[i] = expf_glibc(x);
        else if (op == 2) ((int *)out)[i] = to_int_gamma(x);
        else if (op == 3) {
            float xs[1] = { x }, r[1];
            sqrt_group<1>(xs, r);
            ((float *)out)[2 * i] = r[0]; ((float *)out)[2 * i + 1] = __fsqrt_rn(x);
        } else if (op == 4) ((double *)out)[i] = pow20_double(x);
    }
}

}  // namespace rtb

// ------------------------------------------------------------------------------------------------ launchers
using namespace rtb;

// Launch configuration of one kernel at one dynamic shared-memory size on one device: the opt-in attribute is set and the
// occupancy queried ONCE, then remembered -- a 3 ms frame should not pay two driver calls per launch for the same answer.
#include <mutex>
#include <vector>
namespace {
struct KernelCfg { const void *kernel; size_t smem; int threads, device, blocks; };
std::mutex g_cfg_mutex;
std::vector<KernelCfg> g_cfg;
}
template <typename K>
static cudaError_t configure_kernel(K kernel, int threads, size_t smem, int *blocks) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(g_cfg_mutex);
    for (const KernelCfg &c : g_cfg)
        if (c.kernel == (const void *)kernel && c.smem == smem && c.threads == threads && c.device == dev) { *blocks = c.blocks; return cudaSuccess; }
    // the attribute is a per-function MAXIMUM: only ever raised, so that an earlier, larger configuration keeps launching
    size_t have = 0;
    bool seen = false;
    for (const KernelCfg &c : g_cfg)
        if (c.kernel == (const void *)kernel && c.device == dev) { seen = true; if (c.smem > have) have = c.smem; }
    if (!seen || smem > have)
        if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    int nb = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem)) != cudaSuccess) return e;
    if (nb < 1) return cudaErrorLaunchOutOfResources;
    g_cfg.push_back(KernelCfg{ (const void *)kernel, smem, threads, dev, nb });
    *blocks = nb;
    return cudaSuccess;
}

cudaError_t rtk_build_whitted_grid(const WGrid &G, int gz, uint32_t *cells, const f4 *geom, const int *flags, const f2 *pcull, const float *smargin,
                                   const f4 *lcenter, int n_lights, cudaStream_t stream) {
    const long n = (long)G.gx * G.gy * gz;
    whitted_grid_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(G, gz, cells, geom, flags, pcull, smargin, lcenter, n_lights);
    return cudaGetLastError();
}

cudaError_t rtk_build_whitted_tiles(uint32_t *tiles, int tiles_x, int tiles_y, int w, int h, float DX, float DY, const f4 *geom, const int *flags,
                                    const float *smargin, uint32_t all, cudaStream_t stream) {
    const long n = (long)tiles_x * tiles_y;
    whitted_tiles_kernel<<<(unsigned)((n + 63) / 64), 64, 0, stream>>>(tiles, tiles_x, tiles_y, w, h, DX, DY, geom, flags, smargin, all);
    return cudaGetLastError();
}

cudaError_t rtk_fill_sincos_table(float *tab, int sm_count, cudaStream_t stream) {
    sincos_table_kernel<<<sm_count * 8, 256, 0, stream>>>(tab);
    return cudaGetLastError();
}

static cudaError_t rtk_launch_pt_pack(const PtLaunch &p, cudaStream_t stream) {
    if (!p.frame.defer_pack || p.frame.sum_mode) return cudaSuccess;
    long grid = ((long)p.n_items + 255) / 256;
    if (grid > (long)p.sm_count * 8) grid = (long)p.sm_count * 8;
    pt_pack_kernel<<<(unsigned)(grid > 0 ? grid : 1), 256, 0, stream>>>(p.shard, p.frame.w, p.frame.h, p.n_items, p.colors, p.pixels);
    return cudaGetLastError();
}

cudaError_t rtk_launch_pt(const PtLaunch &p, cudaStream_t stream) {
    if (p.use_bvh && !p.count) {
        int nb = 0;
        cudaError_t e = configure_kernel(pt_bvh_kernel, PT_THREADS, 0, &nb);
        if (e != cudaSuccess) return e;
        if (p.max_blocks_per_sm > 0 && nb > p.max_blocks_per_sm) nb = p.max_blocks_per_sm;
        long grid = (long)nb * p.sm_count;
        const long need = ((long)p.n_items + PT_THREADS - 1) / PT_THREADS;
        if (grid > need) grid = need > 0 ? need : 1;
        pt_bvh_kernel<<<(unsigned)grid, PT_THREADS, 0, stream>>>(p.frame, p.bvh, p.shard, p.n_items, p.colors, p.seeds, p.pixels, p.work_counter);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        return rtk_launch_pt_pack(p, stream);
    }
    const size_t geom_bytes = (size_t)p.frame.n * sizeof(f4);
    const bool chunked = geom_bytes > (size_t)p.max_smem_geom;
    int chunk = p.frame.n;
    size_t smem = geom_bytes;
    if (chunked) {
        chunk = p.chunk_spheres;
        smem = (size_t)chunk * sizeof(f4);
    }
    if (smem < 16) smem = 16;
    typedef void (*kern_t)(PtFrame, Shard, uint32_t, float *, uint32_t *, uint32_t *, unsigned *, unsigned long long *, int);
    const bool aligned = !chunked && (p.aligned > 0 || (p.aligned < 0 && p.frame.n <= PT_ALIGNED_MAX_SPHERES));
    kern_t k = chunked ? (p.count ? pt_kernel<true, true, false> : pt_kernel<false, true, false>)
             : aligned ? (p.count ? pt_kernel<true, false, true> : pt_kernel<false, false, true>)
                       : (p.count ? pt_kernel<true, false, false> : pt_kernel<false, false, false>);
    int nb = 0;
    cudaError_t e = configure_kernel(k, PT_THREADS, smem, &nb);
    if (e != cudaSuccess) return e;
    if (p.max_blocks_per_sm > 0 && nb > p.max_blocks_per_sm) nb = p.max_blocks_per_sm;
    long grid = (long)nb * p.sm_count;
    const long need = ((long)p.n_items + PT_THREADS - 1) / PT_THREADS;
    if (grid > need) grid = need > 0 ? need : 1;
    k<<<(unsigned)grid, PT_THREADS, smem, stream>>>(p.frame, p.shard, p.n_items, p.colors, p.seeds, p.pixels,
                                                   p.work_counter, p.counters, chunk);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    return rtk_launch_pt_pack(p, stream);
}

cudaError_t rtk_launch_pt_resolve(const float *colors, uint32_t *pixels, int w, int h, float inv_total, int sm_count, cudaStream_t stream) {
    pt_resolve_kernel<<<sm_count * 8, 256, 0, stream>>>(colors, pixels, w, h, inv_total);
    return cudaGetLastError();
}

cudaError_t rtk_launch_selftest_math(int op, const float *in, void *out, unsigned long long n, int sm_count, cudaStream_t stream) {
    selftest_math_kernel<<<sm_count * 8, 256, 0, stream>>>(op, in, out, n);
    return cudaGetLastError();
}

cudaError_t rtk_launch_r306(const R306Launch &p, cudaStream_t stream) {
    const bool fixed = p.frame.W.n <= W_TAB_CAP && p.frame.W.n_runs <= W_TAB_RUNS;
    const size_t smem = rtk_whitted_smem_bytes(p.frame.W.n, p.frame.W.n_lights, p.frame.W.n_runs, fixed ? 3 : 2);
    const bool split = p.subcol != nullptr;
    auto kernel = fixed ? (split ? r306_kernel<true, 3> : r306_kernel<false, 3>) : (split ? r306_kernel<true, 2> : r306_kernel<false, 2>);
    int nb = 0, cnb = 0;
    cudaError_t e = configure_kernel(kernel, W_THREADS, smem, &nb);
    if (e != cudaSuccess) return e;
    uint32_t n_work = p.n_items;
    if (p.order) {      // the Whitted pre-pass on the same scene tables (its camera rays differ from Engine_Render's running sums by rounding: fine for a schedule)
        const size_t csmem = (size_t)p.frame.W.n * sizeof(f4) + (size_t)p.frame.W.n_runs * 3 * sizeof(int) + 16;
        if ((e = configure_kernel(whitted_classify_kernel, W_THREADS, csmem, &cnb)) != cudaSuccess) return e;
        e = cudaMemsetAsync(p.class_counts, 0, W_COST_CLASSES * sizeof(unsigned), stream);
        if (e != cudaSuccess) return e;
        long cgrid = ((long)p.n_items + W_THREADS - 1) / W_THREADS;
        if (cgrid > (long)p.sm_count * 16) cgrid = (long)p.sm_count * 16;
        PtBvh none;
        memset(&none, 0, sizeof none);
        whitted_classify_kernel<<<(unsigned)cgrid, W_THREADS, csmem, stream>>>(p.frame.W, p.shard, p.n_items, p.order, p.class_counts, 1, 0, none, nullptr);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        n_work = p.n_valid;
    }
    const uint32_t units = split ? n_work * 9u : n_work;
    long grid = (long)nb * p.sm_count;
    const long need = ((long)units + W_THREADS - 1) / W_THREADS;
    if (grid > need) grid = need > 0 ? need : 1;
    kernel<<<(unsigned)grid, W_THREADS, smem, stream>>>(p.frame, p.shard, units, p.order, p.class_counts, p.n_items, p.dest, p.subcol, p.work_counter);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (split) {
        r306_resolve_kernel<<<p.sm_count * 8, 256, 0, stream>>>(p.shard, p.frame.W.w, p.frame.row0, p.frame.row1, p.n_items, p.subcol, p.dest);
        e = cudaGetLastError();
    }
    return e;
}

size_t rtk_whitted_smem_bytes(int n, int n_lights, int n_runs, int stage_mode) {
    if (stage_mode == 0 || stage_mode == 3) return 16;          // 3: static shared memory, nothing dynamic
    size_t b = (size_t)n * (sizeof(f4) + sizeof(int)) + (size_t)n_runs * 3 * sizeof(int);
    if (stage_mode == 2) b += (size_t)n * (2 * sizeof(f4) + sizeof(f2) + sizeof(float)) + (size_t)n_runs * 2 * sizeof(f4) + (size_t)n_lights * sizeof(int);
    return b < 16 ? 16 : b;
}

cudaError_t rtk_launch_whitted(const WLaunch &p, cudaStream_t stream) {
    const size_t smem = rtk_whitted_smem_bytes(p.frame.n, p.frame.n_lights, p.frame.n_runs, p.stage_mode);
    typedef void (*kern_t)(WFrame, Shard, uint32_t, const uint32_t *, const unsigned *, uint32_t, uint32_t *, unsigned *, unsigned long long *, PtBvh, const uint8_t *, uint32_t);
    const bool bvh = p.use_bvh && !p.count;
    kern_t k;
    if (bvh) {          // large scenes: tables through L1 / L2 or geometry-only staging, the light count still compiled in
        const int nl = p.sphere_lights >= 1 && p.sphere_lights <= 3 ? p.sphere_lights : 0;
        k = p.stage_mode == 0 ? (nl == 3 ? whitted_kernel<false, 0, 3, true> : nl == 2 ? whitted_kernel<false, 0, 2, true> : nl == 1 ? whitted_kernel<false, 0, 1, true> : whitted_kernel<false, 0, 0, true>)
                              : (nl == 3 ? whitted_kernel<false, 1, 3, true> : nl == 2 ? whitted_kernel<false, 1, 2, true> : nl == 1 ? whitted_kernel<false, 1, 1, true> : whitted_kernel<false, 1, 0, true>);
    } else
        k = p.stage_mode == 0 ? (p.count ? whitted_kernel<true, 0, 0, false> : whitted_kernel<false, 0, 0, false>)
          : p.stage_mode == 1 ? (p.count ? whitted_kernel<true, 1, 0, false> : whitted_kernel<false, 1, 0, false>)
          : p.stage_mode == 3 ? (p.sphere_lights == 3 ? (p.count ? whitted_kernel<true, 3, 3, false> : whitted_kernel<false, 3, 3, false>)
                               : p.sphere_lights == 2 ? (p.count ? whitted_kernel<true, 3, 2, false> : whitted_kernel<false, 3, 2, false>)
                               : p.sphere_lights == 1 ? (p.count ? whitted_kernel<true, 3, 1, false> : whitted_kernel<false, 3, 1, false>)
                                                      : (p.count ? whitted_kernel<true, 3, 0, false> : whitted_kernel<false, 3, 0, false>))
          : p.sphere_lights == 3 ? (p.count ? whitted_kernel<true, 2, 3, false> : whitted_kernel<false, 2, 3, false>)
          : p.sphere_lights == 2 ? (p.count ? whitted_kernel<true, 2, 2, false> : whitted_kernel<false, 2, 2, false>)
          : p.sphere_lights == 1 ? (p.count ? whitted_kernel<true, 2, 1, false> : whitted_kernel<false, 2, 1, false>)
                                 : (p.count ? whitted_kernel<true, 2, 0, false> : whitted_kernel<false, 2, 0, false>);
    if (p.frame.grid.cells && !bvh && !p.count && p.stage_mode == 3 && p.sphere_lights >= 1 && p.sphere_lights <= 3)
        k = p.sphere_lights == 3 ? whitted_kernel<false, 3, 3, false, false, true> : p.sphere_lights == 2 ? whitted_kernel<false, 3, 2, false, false, true> : whitted_kernel<false, 3, 1, false, false, true>;
    int nb = 0, cnb = 0;
    cudaError_t e = configure_kernel(k, W_THREADS, smem, &nb);
    if (e != cudaSuccess) return e;
    if (p.max_blocks_per_sm > 0 && nb > p.max_blocks_per_sm) nb = p.max_blocks_per_sm;
    long grid = (long)nb * p.sm_count;
    const long need = ((long)p.n_items + W_THREADS - 1) / W_THREADS;
    if (grid > need) grid = need > 0 ? need : 1;
    uint32_t n_work = p.n_items;
    if (p.order && p.classes_ready) n_work = p.n_valid;
    else if (p.order) {
        // scheduling pre-pass: order[] = expensive pixels first; the number of valid entries is known on the host
        size_t csmem = p.stage_mode ? (size_t)p.frame.n * sizeof(f4) + (size_t)p.frame.n_runs * 3 * sizeof(int) : 16;
        if (csmem < 16) csmem = 16;
        if ((e = configure_kernel(whitted_classify_kernel, W_THREADS, csmem, &cnb)) != cudaSuccess) return e;
        e = cudaMemsetAsync(p.class_counts, 0, W_COST_CLASSES * sizeof(unsigned), stream);
        if (e != cudaSuccess) return e;
        long cgrid = ((long)p.n_items + W_THREADS - 1) / W_THREADS;
        if (cgrid > (long)p.sm_count * 16) cgrid = (long)p.sm_count * 16;
        whitted_classify_kernel<<<(unsigned)cgrid, W_THREADS, csmem, stream>>>(p.frame, p.shard, p.n_items, p.order, p.class_counts, p.stage_mode != 0, bvh ? 1 : 0, p.bvh, p.cls);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        n_work = p.n_valid;
    }
    const bool split = p.frame.split0 != 0;
    if (split) {         // class 0, one lane per sub-sample, on the second stream: its blocks take the SMs first, the main kernel's follow as they retire
        auto ks = p.sphere_lights == 3 ? whitted_split_kernel<3> : p.sphere_lights == 2 ? whitted_split_kernel<2> : whitted_split_kernel<1>;
        int nbs = 0;
        if ((e = configure_kernel(ks, W_SPLIT_THREADS, smem, &nbs)) != cudaSuccess) return e;
        if (p.max_blocks_per_sm > 0 && nbs > p.max_blocks_per_sm) nbs = p.max_blocks_per_sm;
        if (p.split_blocks_per_sm > 0 && nbs > p.split_blocks_per_sm) nbs = p.split_blocks_per_sm;
        if ((e = cudaEventRecord(p.ev_fork, stream)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(p.aux_stream, p.ev_fork, 0)) != cudaSuccess) return e;
        ks<<<(unsigned)((long)nbs * p.sm_count), W_SPLIT_THREADS, smem, p.aux_stream>>>(p.frame, p.shard, p.order, p.class_counts, p.pixels, p.split_work_counter);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if ((e = cudaEventRecord(p.ev_join, p.aux_stream)) != cudaSuccess) return e;
    }
    if (split && p.cls && p.wall_kernel) {      // what is left for this stream: pure wall blocks
        auto kw = p.sphere_lights == 3 ? whitted_wall_kernel<3> : p.sphere_lights == 2 ? whitted_wall_kernel<2> : whitted_wall_kernel<1>;
        int nbw = 0;
        if ((e = configure_kernel(kw, W_THREADS, smem, &nbw)) != cudaSuccess) return e;
        if (p.max_blocks_per_sm > 0 && nbw > p.max_blocks_per_sm) nbw = p.max_blocks_per_sm;
        long gridw = (long)nbw * p.sm_count;
        if (gridw > need) gridw = need > 0 ? need : 1;
        kw<<<(unsigned)gridw, W_THREADS, smem, stream>>>(p.frame, p.shard, p.n_items, p.pixels, p.work_counter + 1, p.cls);
    } else
    k<<<(unsigned)grid, W_THREADS, smem, stream>>>(p.frame, p.shard, n_work, p.order, p.class_counts, p.n_items, p.pixels, p.work_counter,
                                                  p.counters, p.bvh, p.order ? p.cls : nullptr, p.filler_items);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (split && (e = cudaStreamWaitEvent(stream, p.ev_join, 0)) != cudaSuccess) return e;
    if (!p.redo_work_counter) return e;
    // The pixels the timed kernel reported (blocked lights it may not skip), again, as the reference computes them.  Launched without
    // looking at the count -- that would be a round trip to the host; a launch that finds the list empty is a few microseconds.
    kern_t kx = bvh ? (p.stage_mode == 0 ? whitted_kernel<false, 0, 0, true, true> : whitted_kernel<false, 1, 0, true, true>)
              : p.stage_mode == 0 ? whitted_kernel<false, 0, 0, false, true> : p.stage_mode == 1 ? whitted_kernel<false, 1, 0, false, true>
              : p.stage_mode == 3 ? whitted_kernel<false, 3, 0, false, true> : whitted_kernel<false, 2, 0, false, true>;
    int nbx = 0;
    if ((e = configure_kernel(kx, W_THREADS, smem, &nbx)) != cudaSuccess) return e;
    long gridx = (long)nbx * p.sm_count;
    if (gridx > need) gridx = need > 0 ? need : 1;
    kx<<<(unsigned)gridx, W_THREADS, smem, stream>>>(p.frame, p.shard, p.n_items, nullptr, nullptr, p.n_items, p.pixels, p.redo_work_counter,
                                                    p.counters, p.bvh, nullptr, 0u);
    return cudaGetLastError();
}

// -DRT_DEVICE_CHECKS builds: the word of failed bounds checks (rt_math.cuh), read and cleared; -1 when the checks are not compiled in.
long long rtk_read_check_flags() {
#ifdef RT_DEVICE_CHECKS
    unsigned v = 0, zero = 0;
    if (cudaMemcpyFromSymbol(&v, rtb::g_rt_check_flags, sizeof v) != cudaSuccess) return -2;
    cudaMemcpyToSymbol(rtb::g_rt_check_flags, &zero, sizeof zero);
    return (long long)v;
#else
    return -1;
#endif
}
