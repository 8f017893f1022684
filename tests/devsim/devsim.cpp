// tests/devsim/devsim.cpp -- TEST-ONLY host build of the device lane code.
//
// Compiles the very headers the CUDA kernels are made of (rt_math.cuh, pt_lane.cuh, whitted_lane.cuh,
// scene_soa.h) with g++ and drives one lane at a time through "query, then advance" -- the loop body
// of the kernels in csrc/rt_kernels.cu -- so that the per-lane state machines, the work-item -> pixel
// mapping and the FP64 replicas of glibc's float functions can be checked against the oracle on a
// machine without a GPU.  It is not part of librt_b200.so, is never loaded by the product, and is not
// a fallback: the product has none (rt_init fails without a CUDA device).
#define PT_BVH_STATS 1
#include <vector>
#include <stdint.h>
#include <string.h>
#include "../../se-195-project-ray-tracer_b200/csrc/scene_soa.h"
#include "../../se-195-project-ray-tracer_b200/csrc/r306_lane.cuh"
#include "../../se-195-project-ray-tracer_b200/csrc/pt_bvh_build.h"
#include "../../se-195-project-ray-tracer_b200/csrc/whitted_bvh.cuh"

using namespace rtb;

extern "C" {

void devsim_sincos(const float *in, float *sin_out, float *cos_out, long n) {
    for (long i = 0; i < n; i++) sincos_glibc(in[i], &sin_out[i], &cos_out[i]);
}
void devsim_expf(const float *in, float *out, long n) { for (long i = 0; i < n; i++) out[i] = expf_glibc(in[i]); }
void devsim_to_int_gamma(const float *in, int *out, long n) { for (long i = 0; i < n; i++) out[i] = to_int_gamma(in[i]); }
double devsim_pow20(float v) { return pow20_double(v); }
float devsim_get_random(uint32_t *s0, uint32_t *s1) { return get_random(*s0, *s1); }

// Item -> pixel coverage: marks every pixel that (rank, world, tile_rows) visits; returns the item count.
uint32_t devsim_cover(int w, int h, int rank, int world, int tile_rows, int32_t *visits /* w*h, incremented */) {
    uint32_t n_items;
    Shard S = make_shard(w, h, rank, world, tile_rows, &n_items);
    for (uint32_t it = 0; it < n_items; it++) {
        int x, y;
        if (item_to_pixel(S, w, it, x, y)) visits[(size_t)y * w + x]++;
    }
    return n_items;
}

// WFrame::tame_reach the host computes for a scene table (scene_soa.h).
void devsim_whitted_tame_reach(const rt_primitive *prims, int n, float *out2) { WSoA soa; build_w_soa(prims, n, soa); out2[0] = soa.tame_reach[0]; out2[1] = soa.tame_reach[1]; }

// Timed-mode devsim_whitted calls use the shadow-candidate grid (default) or only the per-hit-point culls.
static int g_use_grid = 1;
void devsim_whitted_use_grid(int on) { g_use_grid = on; }

// Pixels of the last timed-mode devsim_whitted call that went through the EXACT pass.
static long g_redo_pixels = 0;
long devsim_whitted_redo_pixels() { return g_redo_pixels; }

// Whitted frame through the lane state machine (rows owned by rank/world/tile_rows only).
void devsim_whitted(uint8_t *pixels, int32_t *hit_ids, int w, int h, const rt_primitive *prims, int n,
                    int rank, int world, int tile_rows, uint64_t *counters5, float *acc_out /* NULL or w*h*3 */, int use_runs) {
    WSoA soa;
    build_w_soa(prims, n, soa);
    WFrame F;
    F.geom = soa.geom.data(); F.mat_a = soa.mat_a.data(); F.mat_b = soa.mat_b.data();
    F.flags = soa.flags.data(); F.lights = soa.lights.data(); F.lcenter = soa.lcenter.data(); F.rrad = soa.rrad.data();
    F.runs = soa.runs.data(); F.n_runs = (int)soa.runs.size() / 3;
    if (use_runs == 2) { F.runs = soa.runs_hot.data(); F.n_runs = (int)soa.runs_hot.size() / 3; }    // what timed launches walk
    if (use_runs == 3) { build_w_bvh(prims, n, soa); F.runs = soa.runs_bvh.data(); F.n_runs = (int)soa.runs_bvh.size() / 3; }   // + the hierarchy
    const PtBvh B = soa.bvh.view(soa.bvh.nodes.data(), soa.bvh.geom.data(), soa.bvh.index.data());
    F.n = n; F.n_lights = (int)soa.lights.size(); F.n_spheres = soa.n_spheres; F.n_planes = soa.n_planes;
    F.w = w; F.h = h;
    const float WX1 = -3.0f, WX2 = 3.0f, WY1 = 2.25f, WY2 = -2.25f;
    F.DX = (WX2 - WX1) / w; F.DY = (WY2 - WY1) / h;
    F.hit_ids = hit_ids;
    unsigned redo_count = 0;
    uint32_t redo_one[1];
    F.tame_reach[0] = soa.tame_reach[0]; F.tame_reach[1] = soa.tame_reach[1]; F.redo_count = nullptr; F.redo_list = redo_one; F.redo_cap = 1;
    F.pcull = nullptr; F.rbox = nullptr; F.cull_rp2 = 0.f; F.reject_k = 0.f;
    memset(&F.grid, 0, sizeof F.grid); F.split0 = 0;
    std::vector<uint32_t> grid_cells, tile_words;
    WCull cull;
    const bool split = use_runs == 6;          // 6: like 4, every pixel the way whitted_split_kernel renders class-0 pixels (one lane per sub-sample, logs added in order)
    if (split) use_runs = 4;
    if (use_runs == 4 || use_runs == 5) {      // what a timed launch does: the runs without dead primitives (5: + the hierarchy), the shadow-round culls, no counting
        if (use_runs == 5) { build_w_bvh(prims, n, soa); F.runs = soa.runs_bvh.data(); F.n_runs = (int)soa.runs_bvh.size() / 3; }
        else { F.runs = soa.runs_hot.data(); F.n_runs = (int)soa.runs_hot.size() / 3; }
        build_w_cull(soa, use_runs == 5 ? soa.runs_bvh : soa.runs_hot, cull);
        F.pcull = cull.pcull.data(); F.rbox = cull.rbox.data(); F.cull_rp2 = cull.rp2; F.reject_k = cull.reject_k;
        if (use_runs == 4 && cull.grid_gz > 0 && g_use_grid) {        // the shadow-candidate grid, cell by cell as the device builds it
            F.grid = cull.grid;
            grid_cells.resize((size_t)F.grid.gx * F.grid.gy * cull.grid_gz);
            for (size_t c = 0; c < grid_cells.size(); c++)
                grid_cells[c] = w_grid_build_cell(F.grid, cull.grid_gz, (int)c, F.geom, F.flags, F.pcull, cull.smargin.data(), F.lcenter, F.n_lights);
            F.grid.cells = grid_cells.data();
            F.grid.tiles_x = (w + 7) / 8; F.grid.tiles_y = (h + 3) / 4;
            tile_words.resize((size_t)F.grid.tiles_x * ((h + 3) / 4));
            for (size_t t = 0; t < tile_words.size(); t++)
                tile_words[t] = w_tile_build((int)(t % F.grid.tiles_x), (int)(t / F.grid.tiles_x), w, h, F.DX, F.DY, F.geom, F.flags, cull.smargin.data(), F.grid.all_nearest);
            F.grid.tiles = tile_words.data();
        }
    }
    const PtBvh B5 = soa.bvh.view(soa.bvh.nodes.data(), soa.bvh.geom.data(), soa.bvh.index.data());
    uint32_t n_items;
    Shard S = make_shard(w, h, rank, world, tile_rows, &n_items);
    f4 queue[3 * W_QUEUE_SLOTS];
    uint64_t c[5] = {0, 0, 0, 0, 0};
    g_redo_pixels = 0;
    for (uint32_t it = 0; it < n_items; it++) {
        int x, y;
        if (!item_to_pixel(S, w, it, x, y)) continue;
        WLane L;
        memset(&L, 0, sizeof L);
        w_begin_pixel(L, F, x, y);
        if (split) {
            float ar = 0.f, ag = 0.f, ab = 0.f;
            std::vector<float> log(3 * 63);
            redo_count = 0; F.redo_count = &redo_count;
            for (int sub = 0; sub < 9; sub++) {
                memset(&L, 0, sizeof L);
                L.x = x; L.y = y; L.sub = sub; L.nlog = 0;
                w_start_subsample(L, F);
                for (;;) {
                    w_query_nearest_tiles(L, F.geom, F.flags, F.runs, F.n_runs, true, F.grid);
                    w_after_nearest<false>(L, F);
                    while (L.phase == PH_SHADOW) {
                        if (F.grid.cells) w_query_shadow_grid(L, F.geom, F.flags, true, F.grid, F.reject_k);       // (a table without a grid: the GPU does not split it)
                        else w_query_shadow<false, true>(L, F.geom, F.runs, F.n_runs, true, F.pcull, F.rbox, F.reject_k);
                        w_after_shadow<false>(L, F);
                    }
                    if (w_finalize<false, true>(L, F, queue, log.data())) break;
                }
                for (int k = 0; k < L.nlog; k++) { ar = f_add(ar, log[3 * k]); ag = f_add(ag, log[3 * k + 1]); ab = f_add(ab, log[3 * k + 2]); }
            }
            L.ar = ar; L.ag = ag; L.ab = ab;
            if (redo_count) {           // reported: the pixel again, one lane, as the EXACT launch computes it
                g_redo_pixels++;
                memset(&L, 0, sizeof L);
                w_begin_pixel(L, F, x, y);
                for (;;) {
                    w_query_nearest_tiles(L, F.geom, F.flags, F.runs, F.n_runs, true, F.grid);
                    w_after_nearest<false, 0, true>(L, F);
                    while (L.phase == PH_SHADOW) {
                        if (F.grid.cells) w_query_shadow_grid(L, F.geom, F.flags, true, F.grid, F.reject_k);
                        else w_query_shadow<false, true>(L, F.geom, F.runs, F.n_runs, true, F.pcull, F.rbox, F.reject_k);
                        w_after_shadow<false, 0, true>(L, F);
                    }
                    if (w_finalize<false>(L, F, queue)) break;
                }
            }
        } else if (use_runs >= 4) {
            // the body of the timed kernel's loop, for one lane -- and, when the pixel reported a batch whose blocked lights may not be
            // skipped, the pixel again as the EXACT launch computes it (whitted_lane.cuh, "Blocked lights and the redo list")
            redo_count = 0; F.redo_count = &redo_count;
            for (;;) {
                if (F.grid.cells) w_query_nearest_tiles(L, F.geom, F.flags, F.runs, F.n_runs, true, F.grid);
                else w_query_nearest<false>(L, F.geom, F.runs, F.n_runs, true);
                if (use_runs == 5) w_bvh_nearest(L, B5);
                w_after_nearest<false>(L, F);
                while (L.phase == PH_SHADOW) {
                    if (F.grid.cells) w_query_shadow_grid(L, F.geom, F.flags, true, F.grid, F.reject_k);
                    else w_query_shadow<false, true>(L, F.geom, F.runs, F.n_runs, true, F.pcull, F.rbox, F.reject_k);
                    if (use_runs == 5) w_bvh_shadow(L, B5);
                    w_after_shadow<false>(L, F);
                }
                if (w_finalize<false>(L, F, queue)) break;
            }
            if (redo_count) {
                g_redo_pixels++;
                memset(&L, 0, sizeof L);
                w_begin_pixel(L, F, x, y);
                for (;;) {
                    if (F.grid.cells) w_query_nearest_tiles(L, F.geom, F.flags, F.runs, F.n_runs, true, F.grid);
                    else w_query_nearest<false>(L, F.geom, F.runs, F.n_runs, true);
                    if (use_runs == 5) w_bvh_nearest(L, B5);
                    w_after_nearest<false, 0, true>(L, F);
                    while (L.phase == PH_SHADOW) {
                        if (F.grid.cells) w_query_shadow_grid(L, F.geom, F.flags, true, F.grid, F.reject_k);
                        else w_query_shadow<false, true>(L, F.geom, F.runs, F.n_runs, true, F.pcull, F.rbox, F.reject_k);
                        if (use_runs == 5) w_bvh_shadow(L, B5);
                        w_after_shadow<false, 0, true>(L, F);
                    }
                    if (w_finalize<false>(L, F, queue)) break;
                }
            }
        } else for (;;) {
            w_query_nearest<true>(L, F.geom, F.runs, F.n_runs, true);
            if (use_runs == 3) w_bvh_nearest(L, B);
            w_after_nearest<true>(L, F);
            while (L.phase == PH_SHADOW) {
                w_query_shadow<true>(L, F.geom, F.runs, F.n_runs, true);
                if (use_runs == 3) w_bvh_shadow(L, B);
                w_after_shadow<true>(L, F);
            }
            if (w_finalize<true>(L, F, queue)) break;
        }
        const uint32_t p = w_pack_pixel(L.ar, L.ag, L.ab);
        memcpy(pixels + ((size_t)y * w + x) * 4, &p, 4);
        if (acc_out) { float *a = acc_out + ((size_t)y * w + x) * 3; a[0] = L.ar; a[1] = L.ag; a[2] = L.ab; }
        c[0] += L.c_nearest; c[1] += L.c_shadow; c[2] += L.c_sphere_tests; c[3] += L.c_plane_tests; c[4] += L.c_samples;
    }
    if (counters5) for (int k = 0; k < 5; k++) counters5[k] += c[k];
}

// What build_w_cull makes of a scene table: out4 = (enabled, cullable planes, cullable sphere runs, runs in the hot table).
void devsim_whitted_cull_stats(const rt_primitive *prims, int n, int32_t *out4) {
    WSoA soa;
    build_w_soa(prims, n, soa);
    WCull cull;
    build_w_cull(soa, soa.runs_hot, cull);
    out4[0] = cull.enabled; out4[1] = cull.planes_cullable; out4[2] = cull.runs_cullable; out4[3] = (int)soa.runs_hot.size() / 3;
}

// The two per-scene tables of whitted_lane.cuh checked DIRECTLY against the tests they stand in front of, on the primary rays of a w x h
// frame: out[0] primary rays whose accepted hit (every primitive tested, reference order) is NOT in their tile's word (must be 0),
// out[1] = sum of the bits of the tile words over those rays, out[2] = rays; out[3] shadow batches with a blocker (every primitive tested)
// that is NOT in the word of the hit point's grid cell (must be 0), out[4] = sum of the bits of those words, out[5] = batches;
// out[6] = 1 if the table has a grid at all.
void devsim_whitted_table_check(const rt_primitive *prims, int n, int w, int h, int64_t *out) {
    for (int k = 0; k < 7; k++) out[k] = 0;
    WSoA soa;
    build_w_soa(prims, n, soa);
    WCull cull;
    build_w_cull(soa, soa.runs_hot, cull);
    if (cull.grid_gz <= 0) return;
    out[6] = 1;
    WFrame F;
    memset(&F, 0, sizeof F);
    F.geom = soa.geom.data(); F.mat_a = soa.mat_a.data(); F.mat_b = soa.mat_b.data();
    F.flags = soa.flags.data(); F.lights = soa.lights.data(); F.lcenter = soa.lcenter.data(); F.rrad = soa.rrad.data();
    F.runs = soa.runs_hot.data(); F.n_runs = (int)soa.runs_hot.size() / 3;
    F.n = n; F.n_lights = (int)soa.lights.size(); F.n_spheres = soa.n_spheres; F.n_planes = soa.n_planes;
    F.w = w; F.h = h;
    const float WX1 = -3.0f, WX2 = 3.0f, WY1 = 2.25f, WY2 = -2.25f;
    F.DX = (WX2 - WX1) / w; F.DY = (WY2 - WY1) / h;
    F.cull_rp2 = cull.rp2; F.reject_k = cull.reject_k;
    F.tame_reach[0] = soa.tame_reach[0]; F.tame_reach[1] = soa.tame_reach[1];
    F.grid = cull.grid;
    std::vector<uint32_t> cells((size_t)F.grid.gx * F.grid.gy * cull.grid_gz), tiles((size_t)((w + 7) / 8) * ((h + 3) / 4));
    for (size_t c = 0; c < cells.size(); c++)
        cells[c] = w_grid_build_cell(F.grid, cull.grid_gz, (int)c, F.geom, F.flags, cull.pcull.data(), cull.smargin.data(), F.lcenter, F.n_lights);
    F.grid.cells = cells.data(); F.grid.tiles_x = (w + 7) / 8; F.grid.tiles_y = (h + 3) / 4;
    for (size_t t = 0; t < tiles.size(); t++)
        tiles[t] = w_tile_build((int)(t % F.grid.tiles_x), (int)(t / F.grid.tiles_x), w, h, F.DX, F.DY, F.geom, F.flags, cull.smargin.data(), F.grid.all_nearest);
    F.grid.tiles = tiles.data();
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            for (int sub = 0; sub < 9; sub++) {
                WLane L;
                memset(&L, 0, sizeof L);
                L.x = x; L.y = y; L.sub = sub;
                w_start_subsample(L, F);
                w_query_nearest<false>(L, F.geom, F.runs, F.n_runs, true);          // every primitive, the reference's order
                const uint32_t tw = tiles[(size_t)(y >> 2) * F.grid.tiles_x + (x >> 3)];
                out[2]++; out[1] += __builtin_popcount(tw);
                if (L.qhit >= 0 && !((tw >> L.qhit) & 1u)) out[0]++;
                w_after_nearest<false>(L, F);
                if (L.phase != PH_SHADOW) continue;
                // the batch's blockers one primitive at a time, without any cull
                const uint32_t gw = w_grid_lookup(L, F.grid);
                out[5]++; out[4] += __builtin_popcount(gw);
                for (int i = 0; i < n; i++) {
                    if (!((F.grid.all >> i) & 1u)) continue;
                    const int before = L.sblk;
                    L.sblk = 0;
                    const int alive = w_alive_mask(L, true);
                    if (F.flags[i] & W_FLAG_SPHERE) w_shadow_sphere<false>(L, F.geom[i], alive, true);
                    else w_shadow_plane<false>(L, F.geom[i], alive, true);
                    if (L.sblk && !((gw >> i) & 1u)) out[3]++;
                    L.sblk = before;
                }
            }
}

// raytracer3.0.06 frame through the lane state machine (rows 20 .. h-71, like Engine_Render).
void devsim_r306(uint32_t *dest, int w, int h, const rt_r306_primitive *prims, int n) {
    WSoA soa;
    build_r306_soa(prims, n, soa);
    std::vector<float> sx, sy;
    R306Frame F;
    build_r306_screen(w, h, sx, sy, &F.W.DX, &F.W.DY);
    F.W.geom = soa.geom.data(); F.W.mat_a = soa.mat_a.data(); F.W.mat_b = soa.mat_b.data();
    F.W.flags = soa.flags.data(); F.W.lights = soa.lights.data(); F.W.lcenter = soa.lcenter.data(); F.W.rrad = soa.rrad.data();
    F.W.runs = soa.runs_hot.data(); F.W.n_runs = (int)soa.runs_hot.size() / 3;
    F.W.n = n; F.W.n_lights = (int)soa.lights.size(); F.W.n_spheres = soa.n_spheres; F.W.n_planes = soa.n_planes;
    F.W.w = w; F.W.h = h; F.W.hit_ids = nullptr;
    F.W.pcull = nullptr; F.W.rbox = nullptr; F.W.cull_rp2 = 0.f; F.W.reject_k = 0.f; memset(&F.W.grid, 0, sizeof F.W.grid); F.W.split0 = 0; F.W.tame_reach[0] = F.W.tame_reach[1] = 0.f; F.W.redo_count = nullptr; F.W.redo_list = nullptr; F.W.redo_cap = 0;
    F.sx = sx.data(); F.sy = sy.data(); F.row0 = 20; F.row1 = h - 70;
    R306Tree T;
    for (int y = F.row0; y < F.row1; y++)
        for (int x = 0; x < w; x++) {
            R306Lane L;
            memset(&L, 0, sizeof L);
            r306_begin_pixel(L, F, x, y);
            for (;;) {
                w_query_nearest<false>(L.q, F.W.geom, F.W.runs, F.W.n_runs, true);
                bool node_done = r306_after_nearest(L, F, T);
                while (L.q.phase == PH_SHADOW) {
                    w_query_shadow<false>(L.q, F.W.geom, F.W.runs, F.W.n_runs, true);
                    r306_after_shadow(L, F);
                }
                if (L.q.phase == PH_FINAL) { r306_finish_hit(L, F, T); node_done = true; }
                if (node_done && r306_next_node(L, F, T)) break;
            }
            dest[(size_t)y * w + x] = r306_pack_pixel(L.tr, L.tg, L.tb);
        }
}

// smallpt passes through the lane state machine.  colors/seeds updated in place (CPU-twin indexing).
void devsim_pt(int integrator, const rt_sphere *sph, uint32_t n, const rt_camera *cam, int w, int h,
               int pass0, int n_passes, int sum_mode, float *colors, uint32_t *seeds, uint32_t *pixels,
               int rank, int world, int tile_rows, uint64_t *counters5, int chunk /* -2: hierarchy (pt_bvh.cuh); <=0: plain per-sphere loop */) {
    PtSoA soa;
    build_pt_soa(sph, n, soa);
    PtFrame F;
    F.emis = soa.emis.data(); F.colr = soa.colr.data(); F.geom_global = soa.geom.data(); F.lights = soa.lights.data();
    F.n = (int)n; F.n_lights = (int)soa.lights.size();
    F.cam_ox = cam->orig.x; F.cam_oy = cam->orig.y; F.cam_oz = cam->orig.z;
    F.cam_dx = cam->dir.x; F.cam_dy = cam->dir.y; F.cam_dz = cam->dir.z;
    F.cam_xx = cam->x.x; F.cam_xy = cam->x.y; F.cam_xz = cam->x.z;
    F.cam_yx = cam->y.x; F.cam_yy = cam->y.y; F.cam_yz = cam->y.z;
    F.w = w; F.h = h; F.inv_w = 1.f / w; F.inv_h = 1.f / h;
    F.pass0 = pass0; F.n_passes = n_passes; F.direct_only = integrator; F.sum_mode = sum_mode; F.defer_pack = 0; F.sincos_tab = nullptr;
    uint32_t n_items;
    Shard S = make_shard(w, h, rank, world, tile_rows, &n_items);
    uint64_t c[5] = {0, 0, 0, 0, 0};
    PtBvhHost bh;
    if (chunk == -2) build_pt_bvh(soa.geom, soa.colr, bh);
    const PtBvh B = bh.view(bh.nodes.data(), bh.geom.data(), bh.index.data());
    for (uint32_t it = 0; it < n_items; it++) {
        int x, y;
        if (!item_to_pixel(S, w, it, x, y)) continue;
        PtLane L;
        memset(&L, 0, sizeof L);
        pt_begin_pixel(L, F, x, y, colors, seeds);
        for (;;) {
            if (chunk == -2) pt_query_bvh<true>(L, B);                                                         // the hierarchy of pt_bvh.cuh
            else if (chunk <= 0) for (int i = F.n - 1; i >= 0; --i) pt_test<true>(L, soa.geom[i], i, true);   // plain loop
            else for (int hi = F.n; hi > 0; hi -= chunk) {                                               // the kernel's loops
                const int lo = hi > chunk ? hi - chunk : 0;
                pt_query_range<true>(L, soa.geom.data() + lo, lo, hi, true);
            }
            if (chunk == -2) {      // counters[2] = sphere tests actually executed (pt_advance would add the reference's n per query)
                const uint64_t executed = L.c_tests;
                const bool done = pt_advance<true>(L, F);
                L.c_tests = executed;
                if (done) break;
            }
            else if (pt_advance<true>(L, F)) break;
        }
        const size_t i = (size_t)(h - y - 1) * w + x;
        colors[3 * i] = L.cr; colors[3 * i + 1] = L.cg; colors[3 * i + 2] = L.cb;
        seeds[2 * i] = L.s0; seeds[2 * i + 1] = L.s1;
        if (!sum_mode && pixels) pixels[(size_t)y * w + x] = pt_pack_pixel(L.cr, L.cg, L.cb);
        c[0] += L.c_nearest; c[1] += L.c_shadow; c[2] += L.c_tests; c[4] += L.c_samples;
    }
    if (counters5) for (int k = 0; k < 5; k++) counters5[k] += c[k];
}

// Shape of the hierarchy build_pt_bvh makes for a scene: out5 = (inner nodes, entries, always-tested spheres, depth, leaves).
void devsim_bvh_stats(const rt_sphere *sph, uint32_t n, int32_t *out5) {
    PtSoA soa;
    build_pt_soa(sph, n, soa);
    PtBvhHost bh;
    build_pt_bvh(soa.geom, soa.colr, bh);
    out5[0] = bh.root == PT_BVH_NONE ? 0 : (int)bh.nodes.size() / 4; out5[1] = (int)bh.geom.size(); out5[2] = bh.n_big; out5[3] = bh.depth;
    int leaves = 0;
    if (bh.root != PT_BVH_NONE && bh.root < 0) leaves = 1;
    if (bh.root != PT_BVH_NONE && bh.root >= 0)
        for (size_t i = 0; i < bh.nodes.size() / 4; i++) { int a, b; memcpy(&a, &bh.nodes[4 * i + 3].x, 4); memcpy(&b, &bh.nodes[4 * i + 3].y, 4); leaves += (a < 0) + (b < 0); }
    out5[4] = leaves;
}
// Inner-node and leaf visits of the hierarchy since the last call (out2), then reset.
void devsim_bvh_tail(int64_t *out2, float *ray7) { out2[0] = g_bvh_tail_visits; out2[1] = g_bvh_max_visits; for (int i = 0; i < 7; i++) ray7[i] = g_bvh_max_ray[i]; g_bvh_tail_visits = g_bvh_max_visits = 0; }
void devsim_bvh_hist(int64_t *out64) { for (int i = 0; i < 64; i++) { out64[i] = g_bvh_hist[i]; g_bvh_hist[i] = 0; } }
void devsim_bvh_visits(int64_t *out2) { out2[0] = g_bvh_inner_visits; out2[1] = g_bvh_leaf_visits; g_bvh_inner_visits = g_bvh_leaf_visits = 0; }

}  // extern "C"
