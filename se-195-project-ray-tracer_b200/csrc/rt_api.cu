// rt_api.cu -- the C ABI declared in include/rt_b200.h: context, device buffers, launches, copies.
//
// This layer is what replaces the reference's OpenCL host glue (SPT/smallptGPU.cpp:100-830,
// R323/raytracer.c:84-684).  There is deliberately no CPU path in it: every render entry point
// either launches the sm_100a kernels of rt_kernels.cu or returns an error code.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include "../../include/rt_b200.h"
#include "scene_soa.h"
#include "pt_bvh_build.h"
#include "rt_kernels.h"

using namespace rtb;

struct rt_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t aux_stream = nullptr;             // Whitted: whitted_split_kernel runs here, next to the main kernel
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int w_split = 1;                               // RT_TUNE_WHITTED_SPLIT
    // the cost classes of the last split launch (a function of the table, the frame size and the shard only): kept while those stay the same
    struct { bool valid = false; uint64_t gen = 0; int w = 0, h = 0, rank = 0, world = 0, tile = 0; const void *order = nullptr, *cls = nullptr; } w_classes;
    uint64_t w_table_gen = 0;                      // counts the Whitted tables uploaded so far
    int w_split_blocks = 0;                        // RT_TUNE_WHITTED_SPLIT_BLOCKS (0: as many as fit)
    int sm_count = 0, clock_khz = 0, max_smem_optin = 0;
    char name[256] = {0};
    char err[512] = {0};
    int shard_rank = 0, shard_world = 1, tile_rows = 8;
    int counting = 0;
    unsigned *d_work = nullptr;                    // work counters (items, screen blocks, Whitted redo reports, the redo launch's items), zeroed before each launch
    uint32_t *d_wredo = nullptr;                   // Whitted: pixels reported for the EXACT launch (RT_WHITTED_REDO_CAP entries)
    unsigned long long *d_counters = nullptr;      // 8 x u64
    uint64_t launches = 0;
    // tuning
    int pt_max_resident_bytes = 96 * 1024;
    int pt_chunk_spheres = 3072;                   // 48 KB of (p, rad^2) per chunk
    int max_blocks_per_sm = 0;
    int whitted_sort = 1;                          // cost-sorted work order (scheduling pre-pass)
    int pt_aligned = -1;                           // step-aligned path-tracer warps: -1 by scene size, 0 off, 1 on
    int pt_bvh = -1;                               // exact hierarchy for the sphere queries: -1 by scene size, 0 off, 1 on
    int w_bvh = -1;                                // the same for the Whitted tracer's non-light spheres
    int r306_split = 1;                            // 3.0.06 frame: one sub-sample per work unit (1) or one pixel (0)
    int pt_sincos_table = 1;                       // path tracer: sin / cos of 2*pi*GetRandom() from a 64 MB table (1) or computed (0)
    float *d_sincos = nullptr; bool sincos_failed = false;
    uint64_t setup_launches = 0;                   // one-time set-up kernels (tables); not part of rt_launch_count
    bool w_bvh_ready = false;
    f4 *d_wbnodes = nullptr, *d_wbgeom = nullptr; int *d_wbindex = nullptr, *d_wruns_bvh = nullptr;
    size_t cap_wbnodes = 0, cap_wbgeom = 0, cap_wbindex = 0, cap_wruns_bvh = 0;
    std::vector<rt_primitive> w_prims;             // the last uploaded table (the hierarchy is built from it on first use)
    WCull w_cull, w_cull_bvh;                      // shadow-round cull tables for runs_hot / runs_bvh (scene_soa.h)
    f2 *d_wpcull = nullptr, *d_wpcull_bvh = nullptr; f4 *d_wrbox = nullptr, *d_wrbox_bvh = nullptr;
    size_t cap_wpcull = 0, cap_wpcull_bvh = 0, cap_wrbox = 0, cap_wrbox_bvh = 0;
    float *d_wsmargin = nullptr; size_t cap_wsmargin = 0;      // shadow-candidate grid: per-sphere margins, and the cells
    uint32_t *d_wgrid = nullptr; size_t cap_wgrid = 0;
    uint32_t *d_wtiles = nullptr; size_t cap_wtiles = 0;      // primary-ray tile words of the current (table, frame size)
    bool w_grid_ready = false;
    int w_tiles_w = 0, w_tiles_h = 0;              // the frame size d_wtiles was built for (0: not built)
    int w_grid = 1;                                // RT_TUNE_WHITTED_GRID
    uint32_t *peer_wpixels = nullptr, *peer_ppixels = nullptr;   // rank 0's framebuffers, mapped through CUDA IPC
    size_t peer_wcap = 0, peer_pcap = 0;           // ... and how many pixels rank 0 said they hold
    bool exported_w = false, exported_p = false;   // this context's framebuffers are mapped by other processes: never reallocate them
    uint32_t *d_worder = nullptr; size_t worder_cap = 0;
    uint8_t *d_wcls = nullptr; size_t wcls_cap = 0;
    int whitted_blocks = 1;                        // class-2 pixels as whole screen blocks per warp (needs whitted_sort)
    int w_stage_cap = -1;                          // RT_TUNE_WHITTED_STAGE_CAP
    unsigned w_redo_cap = RT_WHITTED_REDO_CAP;     // RT_TUNE_WHITTED_REDO_CAP
    int whitted_filler_pct = 25;                   // ... except this share of the frame, which fills idle lanes pixel by pixel
    unsigned *d_wclass = nullptr;
    // Whitted
    int w_w = 0, w_h = 0, w_n = 0, w_nl = 0, w_ns = 0, w_np = 0, w_nr = 0, w_want_hits = 0;
    f4 *d_wgeom = nullptr, *d_wma = nullptr, *d_wmb = nullptr, *d_wlcenter = nullptr; size_t cap_wlcenter = 0;
    int *d_wflags = nullptr, *d_wlights = nullptr, *d_wruns = nullptr, *d_wruns_hot = nullptr; size_t cap_wruns_hot = 0; int w_nr_hot = 0;
    float *d_wrrad = nullptr;
    uint32_t *d_wpixels = nullptr;
    int32_t *d_whits = nullptr;
    size_t w_pixels_cap = 0, w_hits_cap = 0;
    size_t cap_wgeom = 0, cap_wma = 0, cap_wmb = 0, cap_wflags = 0, cap_wlights = 0, cap_wrrad = 0, cap_wruns = 0;
    size_t cap_pgeom = 0, cap_pemis = 0, cap_pcolr = 0, cap_plights = 0;
    // raytracer3.0.06 frame (config 1)
    struct {
        f4 *geom = nullptr, *ma = nullptr, *mb = nullptr; int *flags = nullptr, *lights = nullptr, *runs = nullptr; float *rrad = nullptr, *sx = nullptr, *sy = nullptr; f4 *lcenter = nullptr; size_t cap_lcenter = 0;
        size_t cap_geom = 0, cap_ma = 0, cap_mb = 0, cap_flags = 0, cap_lights = 0, cap_runs = 0, cap_rrad = 0, cap_sx = 0, cap_sy = 0;
        uint32_t *dest = nullptr; size_t dest_cap = 0;
        float *subcol = nullptr; size_t subcol_cap = 0;      // sub-sample colours (r306_kernel<SPLIT>)
        int w = 0, h = 0, n = 0, nl = 0, nr = 0, ns = 0, np = 0;
        float DX = 0.f, DY = 0.f;
        WSoA soa; std::vector<float> h_sx, h_sy;
    } r306;
    WSoA w_soa;                                    // host staging of the last uploaded scenes (kept alive
    PtSoA p_soa;                                   //  until the asynchronous copies have been issued and synced)
    bool own_stream = true;
    // path tracer
    int p_w = 0, p_h = 0, p_n = 0, p_nl = 0;
    size_t p_px_cap = 0;                           // pixels the colour / seed / pixel buffers can hold
    float *d_colors = nullptr;
    uint32_t *d_seeds = nullptr, *d_ppixels = nullptr;
    f4 *d_pgeom = nullptr, *d_pemis = nullptr, *d_pcolr = nullptr;
    int *d_plights = nullptr;
    std::vector<rt_sphere> p_scene_copy;           // the table of the last rt_pt_set_scene (an identical one keeps the device tables and the hierarchy)
    PtBvhHost p_bvh;                               // hierarchy over the current scene (csrc/pt_bvh_build.h); built on first use
    bool p_bvh_ready = false;
    f4 *d_bnodes = nullptr, *d_bgeom = nullptr; int *d_bindex = nullptr;
    size_t cap_bnodes = 0, cap_bgeom = 0, cap_bindex = 0;
    rt_camera cam;
    bool have_cam = false, have_scene = false, have_size = false;
    int current_sample = 0, sum_mode = 0;
};

static thread_local char g_init_err[512] = "";

static int fail(rt_ctx *c, int code, const char *fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(c ? c->err : g_init_err, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, RT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// Uploads v into *dptr, growing the allocation only when the capacity (in elements) is too small, so
// that re-uploading a scene of the same size costs copies only (no cudaMalloc / cudaFree).
template <typename T>
static cudaError_t upload_vec(T **dptr, size_t *cap, const std::vector<T> &v, cudaStream_t s) {
    const size_t need = v.size() ? v.size() : 1;
    if (!*dptr || *cap < need) {
        if (*dptr) { cudaFree(*dptr); *dptr = nullptr; *cap = 0; }
        cudaError_t e = cudaMalloc((void **)dptr, need * sizeof(T));
        if (e != cudaSuccess) return e;
        *cap = need;
    }
    if (v.size()) return cudaMemcpyAsync(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
    return cudaSuccess;
}

// The sin / cos table of the path tracer (rt_math.cuh): 2^23 (sin, cos) pairs = 64 MB, filled once per context by the very
// function it stands in for.  Failing to get the memory is not an error: the kernels then compute the same bits.
static void ensure_sincos_table(rt_ctx *ctx) {
    if (!ctx->pt_sincos_table || ctx->d_sincos || ctx->sincos_failed) return;
    if (cudaMalloc((void **)&ctx->d_sincos, (size_t)2 * (1u << 23) * sizeof(float)) != cudaSuccess) {
        cudaGetLastError();                // clear the sticky-less allocation error
        ctx->d_sincos = nullptr; ctx->sincos_failed = true;
        return;
    }
    if (rtk_fill_sincos_table(ctx->d_sincos, ctx->sm_count, ctx->stream) != cudaSuccess) {
        cudaFree(ctx->d_sincos); ctx->d_sincos = nullptr; ctx->sincos_failed = true;
        return;
    }
    ctx->setup_launches++;
}

extern "C" {

int rt_init(rt_ctx **out, int device) {
    rt_ctx *ctx = nullptr;
    if (!out) return fail(nullptr, RT_ERR_ARG, "rt_init: ctx is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, RT_ERR_NO_DEVICE, "rt_init: no CUDA device (%s); this renderer has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= count) return fail(nullptr, RT_ERR_ARG, "rt_init: device %d out of range (0..%d)", device, count - 1);
    ctx = new rt_ctx();
    ctx->device = device;
    memset(&ctx->cam, 0, sizeof ctx->cam);
    auto bail = [&](cudaError_t err, const char *what) {
        fail(nullptr, RT_ERR_CUDA, "rt_init: %s failed: %s", what, cudaGetErrorString(err));
        rt_destroy(ctx);               // frees whatever the partial initialisation already holds
        return (int)RT_ERR_CUDA;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(e, "cudaSetDevice");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail(e, "cudaGetDeviceProperties");
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    cudaDeviceGetAttribute(&ctx->clock_khz, cudaDevAttrClockRate, device);
    snprintf(ctx->name, sizeof ctx->name, "%s", prop.name);
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if ((e = cudaEventCreate(&ctx->ev0)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if ((e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaMalloc((void **)&ctx->d_work, 6 * sizeof(unsigned))) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = cudaMalloc((void **)&ctx->d_wredo, RT_WHITTED_REDO_CAP * sizeof(uint32_t))) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = cudaMalloc((void **)&ctx->d_counters, 8 * sizeof(unsigned long long))) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = cudaMemset(ctx->d_counters, 0, 8 * sizeof(unsigned long long))) != cudaSuccess) return bail(e, "cudaMemset");
    *out = ctx;
    return RT_OK;
}

void rt_destroy(rt_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->peer_wpixels) cudaIpcCloseMemHandle(ctx->peer_wpixels);
    if (ctx->peer_ppixels) cudaIpcCloseMemHandle(ctx->peer_ppixels);
    void *bufs[] = { ctx->d_work, ctx->d_wredo, ctx->d_counters, ctx->d_wgeom, ctx->d_wma, ctx->d_wmb, ctx->d_wflags, ctx->d_wlights, ctx->d_wruns, ctx->d_wruns_hot, ctx->d_wlcenter, ctx->d_worder, ctx->d_wclass,
                     ctx->d_wrrad, ctx->d_wpixels, ctx->d_whits, ctx->d_colors, ctx->d_seeds, ctx->d_ppixels,
                     ctx->d_pgeom, ctx->d_pemis, ctx->d_pcolr, ctx->d_plights, ctx->d_bnodes, ctx->d_bgeom, ctx->d_bindex,
                     ctx->d_wbnodes, ctx->d_wbgeom, ctx->d_wbindex, ctx->d_wruns_bvh, ctx->d_sincos,
                     ctx->d_wpcull, ctx->d_wpcull_bvh, ctx->d_wrbox, ctx->d_wrbox_bvh, ctx->d_wcls, ctx->d_wsmargin, ctx->d_wgrid, ctx->d_wtiles };
    for (void *b : bufs) if (b) cudaFree(b);
    void *rbufs[] = { ctx->r306.geom, ctx->r306.ma, ctx->r306.mb, ctx->r306.flags, ctx->r306.lights, ctx->r306.runs, ctx->r306.rrad, ctx->r306.sx, ctx->r306.sy, ctx->r306.dest, ctx->r306.lcenter, ctx->r306.subcol };
    for (void *b : rbufs) if (b) cudaFree(b);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->aux_stream) { cudaStreamSynchronize(ctx->aux_stream); cudaStreamDestroy(ctx->aux_stream); }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->stream && ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *rt_last_error(const rt_ctx *ctx) { return ctx ? ctx->err : g_init_err; }

int rt_device_info(const rt_ctx *ctx, int *sm_count, int *sm_clock_khz, char *name, int name_cap) {
    if (!ctx) return RT_ERR_ARG;
    if (sm_count) *sm_count = ctx->sm_count;
    if (sm_clock_khz) *sm_clock_khz = ctx->clock_khz;
    if (name && name_cap > 0) snprintf(name, name_cap, "%s", ctx->name);
    return RT_OK;
}

int rt_set_shard(rt_ctx *ctx, int rank, int world, int tile_rows) {
    if (!ctx) return RT_ERR_ARG;
    if (world < 1 || rank < 0 || rank >= world || tile_rows < 1)
        return fail(ctx, RT_ERR_ARG, "rt_set_shard: need 0 <= rank < world and tile_rows >= 1 (got %d, %d, %d)", rank, world, tile_rows);
    ctx->shard_rank = rank; ctx->shard_world = world; ctx->tile_rows = tile_rows;
    return RT_OK;
}

int rt_set_counting(rt_ctx *ctx, int enabled) {
    if (!ctx) return RT_ERR_ARG;
    ctx->counting = enabled ? 1 : 0;
    return RT_OK;
}

int rt_get_counters(rt_ctx *ctx, rt_counters *out) {
    if (!ctx || !out) return RT_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    unsigned long long h[5];
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(h, ctx->d_counters, sizeof h, cudaMemcpyDeviceToHost));
    out->nearest_queries = h[0]; out->shadow_queries = h[1]; out->sphere_tests = h[2]; out->plane_tests = h[3]; out->samples = h[4];
    return RT_OK;
}

int rt_get_counters_ex(rt_ctx *ctx, uint64_t *out8) {
    if (!ctx || !out8) return RT_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(out8, ctx->d_counters, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return RT_OK;
}

int rt_set_tuning(rt_ctx *ctx, int key, int value) {
    if (!ctx) return RT_ERR_ARG;
    switch (key) {
        case RT_TUNE_PT_MAX_RESIDENT_BYTES: if (value < 0) break; ctx->pt_max_resident_bytes = value; return RT_OK;
        case RT_TUNE_PT_CHUNK_SPHERES: if (value < 1) break; ctx->pt_chunk_spheres = value; return RT_OK;
        case RT_TUNE_MAX_BLOCKS_PER_SM: if (value < 0) break; ctx->max_blocks_per_sm = value; return RT_OK;
        case RT_TUNE_WHITTED_COST_ORDER: ctx->whitted_sort = value ? 1 : 0; return RT_OK;
        case RT_TUNE_PT_ALIGNED: if (value < -1 || value > 1) break; ctx->pt_aligned = value; return RT_OK;
        case RT_TUNE_PT_BVH: if (value < -1 || value > 1) break; ctx->pt_bvh = value; return RT_OK;
        case RT_TUNE_WHITTED_BVH: if (value < -1 || value > 1) break; ctx->w_bvh = value; return RT_OK;
        case RT_TUNE_R306_SPLIT: ctx->r306_split = value ? 1 : 0; return RT_OK;
        case RT_TUNE_PT_SINCOS_TABLE: ctx->pt_sincos_table = value ? 1 : 0; return RT_OK;
        case RT_TUNE_WHITTED_BLOCKS: ctx->whitted_blocks = value ? 1 : 0; return RT_OK;
        case RT_TUNE_WHITTED_FILLER_PCT: if (value < 0 || value > 100) break; ctx->whitted_filler_pct = value; return RT_OK;
        case RT_TUNE_WHITTED_REDO_CAP: if (value < 0 || value > (int)RT_WHITTED_REDO_CAP) break; ctx->w_redo_cap = (unsigned)value; return RT_OK;
        case RT_TUNE_WHITTED_SPLIT: if (value < 0 || value > 2) break; ctx->w_split = value; return RT_OK;
        case RT_TUNE_WHITTED_SPLIT_BLOCKS: if (value < 0) break; ctx->w_split_blocks = value; return RT_OK;
        case RT_TUNE_WHITTED_GRID: if (value < 0 || value > 2) break; ctx->w_grid = value; return RT_OK;
        case RT_TUNE_WHITTED_STAGE_CAP: if (value < -1 || value > 3) break; ctx->w_stage_cap = value; return RT_OK;
        default: return fail(ctx, RT_ERR_ARG, "rt_set_tuning: unknown key %d", key);
    }
    return fail(ctx, RT_ERR_ARG, "rt_set_tuning: bad value %d for key %d", value, key);
}

// ------------------------------------------------------------------------------------------------ Whitted

int rt_whitted_upload(rt_ctx *ctx, const rt_primitive *prims, int n, int w, int h, int want_hit_ids) {
    if (!ctx) return RT_ERR_ARG;
    if (!prims || n < 1 || w < 1 || h < 1) return fail(ctx, RT_ERR_ARG, "rt_whitted_upload: need prims, n >= 1, w >= 1, h >= 1");
    CK(cudaSetDevice(ctx->device));
    WSoA &soa = ctx->w_soa;
    const bool same_table = ctx->d_wgeom && (int)ctx->w_prims.size() == n && memcmp(ctx->w_prims.data(), prims, (size_t)n * sizeof(rt_primitive)) == 0;
    if (!same_table) {                     // an identical table keeps the staging copy and the hierarchy built from it
        build_w_soa(prims, n, soa);
        ctx->w_prims.assign(prims, prims + n);
        ctx->w_bvh_ready = false;
        ctx->w_table_gen++;
    }
    CK(upload_vec(&ctx->d_wgeom, &ctx->cap_wgeom, soa.geom, ctx->stream));
    CK(upload_vec(&ctx->d_wma, &ctx->cap_wma, soa.mat_a, ctx->stream));
    CK(upload_vec(&ctx->d_wmb, &ctx->cap_wmb, soa.mat_b, ctx->stream));
    CK(upload_vec(&ctx->d_wflags, &ctx->cap_wflags, soa.flags, ctx->stream));
    CK(upload_vec(&ctx->d_wlights, &ctx->cap_wlights, soa.lights, ctx->stream));
    CK(upload_vec(&ctx->d_wlcenter, &ctx->cap_wlcenter, soa.lcenter, ctx->stream));
    CK(upload_vec(&ctx->d_wrrad, &ctx->cap_wrrad, soa.rrad, ctx->stream));
    CK(upload_vec(&ctx->d_wruns, &ctx->cap_wruns, soa.runs, ctx->stream));
    CK(upload_vec(&ctx->d_wruns_hot, &ctx->cap_wruns_hot, soa.runs_hot, ctx->stream));
    build_w_cull(soa, soa.runs_hot, ctx->w_cull);
    CK(upload_vec(&ctx->d_wpcull, &ctx->cap_wpcull, ctx->w_cull.pcull, ctx->stream));
    CK(upload_vec(&ctx->d_wrbox, &ctx->cap_wrbox, ctx->w_cull.rbox, ctx->stream));
    if (!same_table) ctx->w_grid_ready = false;
    if (ctx->w_cull.grid_gz > 0 && !ctx->w_grid_ready) {          // the shadow-candidate grid of this table, computed on the device, once
        const WGrid &G = ctx->w_cull.grid;
        const size_t cells = (size_t)G.gx * G.gy * ctx->w_cull.grid_gz;
        CK(upload_vec(&ctx->d_wsmargin, &ctx->cap_wsmargin, ctx->w_cull.smargin, ctx->stream));
        if (cells > ctx->cap_wgrid) {
            if (ctx->d_wgrid) cudaFree(ctx->d_wgrid);
            ctx->d_wgrid = nullptr; ctx->cap_wgrid = 0;
            CK(cudaMalloc((void **)&ctx->d_wgrid, cells * sizeof(uint32_t)));
            ctx->cap_wgrid = cells;
        }
        CK(rtk_build_whitted_grid(G, ctx->w_cull.grid_gz, ctx->d_wgrid, ctx->d_wgeom, ctx->d_wflags, ctx->d_wpcull, ctx->d_wsmargin, ctx->d_wlcenter,
                                  (int)soa.lights.size(), ctx->stream));
        ctx->setup_launches++;
        ctx->w_grid_ready = true;
        ctx->w_tiles_w = ctx->w_tiles_h = 0;
    }
    if (ctx->w_cull.grid_gz > 0 && (ctx->w_tiles_w != w || ctx->w_tiles_h != h)) {      // the primary-ray tile words of this table at this frame size
        const int tiles_x = (w + 7) / 8, tiles_y = (h + 3) / 4;
        const size_t tiles = (size_t)tiles_x * tiles_y;
        if (tiles > ctx->cap_wtiles) {
            if (ctx->d_wtiles) cudaFree(ctx->d_wtiles);
            ctx->d_wtiles = nullptr; ctx->cap_wtiles = 0;
            CK(cudaMalloc((void **)&ctx->d_wtiles, tiles * sizeof(uint32_t)));
            ctx->cap_wtiles = tiles;
        }
        const float WX1 = -3.0f, WX2 = 3.0f, WY1 = 2.25f, WY2 = -2.25f;
        CK(rtk_build_whitted_tiles(ctx->d_wtiles, tiles_x, tiles_y, w, h, (WX2 - WX1) / w, (WY2 - WY1) / h, ctx->d_wgeom, ctx->d_wflags, ctx->d_wsmargin,
                                   ctx->w_cull.grid.all_nearest, ctx->stream));
        ctx->setup_launches++;
        ctx->w_tiles_w = w; ctx->w_tiles_h = h;
    }
    const size_t px = (size_t)w * h;
    if (ctx->peer_wpixels && px > ctx->peer_wcap)
        return fail(ctx, RT_ERR_STATE, "rt_whitted_upload: %dx%d exceeds the %zu pixels of the imported rank-0 framebuffer (rt_ipc_close, then share again)", w, h, ctx->peer_wcap);
    if (px > ctx->w_pixels_cap) {
        if (ctx->exported_w)
            return fail(ctx, RT_ERR_STATE, "rt_whitted_upload: the framebuffer (%zu pixels) is mapped by other ranks and cannot grow to %dx%d; rt_ipc_close on every rank first", ctx->w_pixels_cap, w, h);
        if (ctx->d_wpixels) cudaFree(ctx->d_wpixels);
        ctx->d_wpixels = nullptr; ctx->w_pixels_cap = 0;
        CK(cudaMalloc((void **)&ctx->d_wpixels, px * sizeof(uint32_t)));
        ctx->w_pixels_cap = px;
        // rows another rank owns (rt_set_shard) are never written here: they read back as 0, not as stale device memory
        CK(cudaMemsetAsync(ctx->d_wpixels, 0, px * sizeof(uint32_t), ctx->stream));
    }
    if (want_hit_ids && px * 9 > ctx->w_hits_cap) {
        if (ctx->d_whits) cudaFree(ctx->d_whits);
        ctx->d_whits = nullptr; ctx->w_hits_cap = 0;
        CK(cudaMalloc((void **)&ctx->d_whits, px * 9 * sizeof(int32_t)));
        ctx->w_hits_cap = px * 9;
        CK(cudaMemsetAsync(ctx->d_whits, 0xff, px * 9 * sizeof(int32_t), ctx->stream));      // -1 = "no hit" for rows of other ranks
    }
    ctx->w_w = w; ctx->w_h = h; ctx->w_n = n; ctx->w_nl = (int)soa.lights.size();
    ctx->w_ns = soa.n_spheres; ctx->w_np = soa.n_planes; ctx->w_nr = (int)soa.runs.size() / 3; ctx->w_nr_hot = (int)soa.runs_hot.size() / 3; ctx->w_want_hits = want_hit_ids ? 1 : 0;
    return RT_OK;
}

int rt_whitted_launch(rt_ctx *ctx) {
    if (!ctx) return RT_ERR_ARG;
    if (!ctx->d_wgeom || !ctx->d_wpixels) return fail(ctx, RT_ERR_STATE, "rt_whitted_launch: call rt_whitted_upload first");
    CK(cudaSetDevice(ctx->device));
    WLaunch p;
    WFrame &F = p.frame;
    F.geom = ctx->d_wgeom; F.mat_a = ctx->d_wma; F.mat_b = ctx->d_wmb; F.flags = ctx->d_wflags; F.lights = ctx->d_wlights; F.lcenter = ctx->d_wlcenter;
    if (ctx->counting) { F.runs = ctx->d_wruns; F.n_runs = ctx->w_nr; }            // every primitive, so that the test counters equal the oracle's
    else { F.runs = ctx->d_wruns_hot; F.n_runs = ctx->w_nr_hot; }
    F.rrad = ctx->d_wrrad; F.n = ctx->w_n; F.n_lights = ctx->w_nl; F.n_spheres = ctx->w_ns; F.n_planes = ctx->w_np;
    F.w = ctx->w_w; F.h = ctx->w_h;
    // R323/raytracer_non_OpenCL.c:291-296: DX = (WX2 - WX1) / width with float operands.
    const float WX1 = -3.0f, WX2 = 3.0f, WY1 = 2.25f, WY2 = -2.25f;
    F.DX = (WX2 - WX1) / ctx->w_w; F.DY = (WY2 - WY1) / ctx->w_h;
    F.hit_ids = ctx->w_want_hits ? ctx->d_whits : nullptr;
    F.tame_reach[0] = ctx->w_soa.tame_reach[0]; F.tame_reach[1] = ctx->w_soa.tame_reach[1]; F.redo_count = ctx->counting ? nullptr : ctx->d_work + 2; F.redo_list = ctx->d_wredo; F.redo_cap = ctx->w_redo_cap;
    F.reject_k = ctx->w_cull.reject_k;
    if (ctx->counting) { F.pcull = nullptr; F.rbox = nullptr; F.cull_rp2 = 0.f; }         // counting launches execute every test
    else { F.pcull = ctx->d_wpcull; F.rbox = ctx->d_wrbox; F.cull_rp2 = ctx->w_cull.rp2; }   // the tables of runs_hot
    p.shard = make_shard(ctx->w_w, ctx->w_h, ctx->shard_rank, ctx->shard_world, ctx->tile_rows, &p.n_items);
    p.pixels = ctx->peer_wpixels ? ctx->peer_wpixels : ctx->d_wpixels;     // fused gather: store straight into rank 0's frame
    p.work_counter = ctx->d_work; p.counters = ctx->d_counters;
    p.redo_work_counter = (ctx->counting || ctx->w_nl == 0) ? nullptr : ctx->d_work + 3;      // no lights: no shadow batches, nothing to report
    p.count = ctx->counting; p.sm_count = ctx->sm_count; p.max_blocks_per_sm = ctx->max_blocks_per_sm;
    // sized with the LARGEST run table a launch of this scene may stage (all primitives / without the dead ones / without
    // what the hierarchy holds: splitting runs can add runs), so that the choice never overshoots the per-CTA budget
    int nr_max = ctx->w_nr > ctx->w_nr_hot ? ctx->w_nr : ctx->w_nr_hot;
    if ((int)ctx->w_soa.runs_bvh.size() / 3 > nr_max) nr_max = (int)ctx->w_soa.runs_bvh.size() / 3;
    p.stage_mode = rtk_whitted_smem_bytes(ctx->w_n, ctx->w_nl, nr_max, 2) <= RTK_WHITTED_STAGE_LIMIT ? 2
                 : rtk_whitted_smem_bytes(ctx->w_n, ctx->w_nl, nr_max, 1) <= RTK_WHITTED_STAGE_LIMIT ? 1 : 0;
    if (p.stage_mode == 2 && ctx->w_n <= W_TAB_CAP && nr_max <= W_TAB_RUNS) p.stage_mode = 3;
    if (ctx->w_stage_cap >= 0 && p.stage_mode > ctx->w_stage_cap) p.stage_mode = ctx->w_stage_cap;      // diagnostics: a lower mode than the tables would allow
    p.sphere_lights = ctx->w_nl;
    for (int l : ctx->w_soa.lights) if (!(ctx->w_soa.flags[l] & W_FLAG_SPHERE)) p.sphere_lights = 0;
    p.order = nullptr; p.class_counts = nullptr; p.cls = nullptr; p.filler_items = 0;
    p.n_valid = (uint32_t)p.shard.local_rows * (uint32_t)ctx->w_w;
    int tree_candidates = 0;
    for (int i = 0; i < ctx->w_n; i++) tree_candidates += (ctx->w_soa.flags[i] & (W_FLAG_SPHERE | W_FLAG_LIGHT)) == W_FLAG_SPHERE;
    p.use_bvh = !ctx->counting && (ctx->w_bvh > 0 || (ctx->w_bvh < 0 && tree_candidates >= W_BVH_MIN_SPHERES));
    if (p.use_bvh) {
        WSoA &soa = ctx->w_soa;
        if (!ctx->w_bvh_ready) {           // built from the table of the last rt_whitted_upload, once
            build_w_bvh(ctx->w_prims.data(), ctx->w_n, soa);
            CK(upload_vec(&ctx->d_wbnodes, &ctx->cap_wbnodes, soa.bvh.nodes, ctx->stream));
            CK(upload_vec(&ctx->d_wbgeom, &ctx->cap_wbgeom, soa.bvh.geom, ctx->stream));
            CK(upload_vec(&ctx->d_wbindex, &ctx->cap_wbindex, soa.bvh.index, ctx->stream));
            CK(upload_vec(&ctx->d_wruns_bvh, &ctx->cap_wruns_bvh, soa.runs_bvh, ctx->stream));
            build_w_cull(soa, soa.runs_bvh, ctx->w_cull_bvh);
            CK(upload_vec(&ctx->d_wpcull_bvh, &ctx->cap_wpcull_bvh, ctx->w_cull_bvh.pcull, ctx->stream));
            CK(upload_vec(&ctx->d_wrbox_bvh, &ctx->cap_wrbox_bvh, ctx->w_cull_bvh.rbox, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            ctx->w_bvh_ready = true;
        }
        p.bvh = soa.bvh.view(ctx->d_wbnodes, ctx->d_wbgeom, ctx->d_wbindex);
        F.runs = ctx->d_wruns_bvh; F.n_runs = (int)soa.runs_bvh.size() / 3;
        F.pcull = ctx->d_wpcull_bvh; F.rbox = ctx->d_wrbox_bvh; F.cull_rp2 = ctx->w_cull_bvh.rp2; F.reject_k = ctx->w_cull_bvh.reject_k;
        if (p.stage_mode >= 2) p.stage_mode = 1;        // the hierarchy kernels read the materials through L1 / L2
    } else memset(&p.bvh, 0, sizeof p.bvh);
    memset(&F.grid, 0, sizeof F.grid);
    if (p.stage_mode == 3 && !ctx->counting && !p.use_bvh && ctx->w_grid && ctx->w_grid_ready && ctx->w_cull.grid_gz > 0 && p.sphere_lights >= 1 && p.sphere_lights <= 3) { F.grid = ctx->w_cull.grid; F.grid.cells = ctx->d_wgrid; F.grid.tiles = ctx->w_grid == 2 ? nullptr : ctx->d_wtiles; F.grid.tiles_x = (ctx->w_w + 7) / 8; F.grid.tiles_y = (ctx->w_h + 3) / 4; }
    if (ctx->whitted_sort && p.n_items) {
        if (p.n_items > ctx->worder_cap) {
            if (ctx->d_worder) cudaFree(ctx->d_worder);
            ctx->d_worder = nullptr; ctx->worder_cap = 0;
            CK(cudaMalloc((void **)&ctx->d_worder, 3 * (size_t)p.n_items * sizeof(uint32_t)));
            ctx->worder_cap = p.n_items;
        }
        if (!ctx->d_wclass) CK(cudaMalloc((void **)&ctx->d_wclass, 4 * sizeof(unsigned)));
        p.order = ctx->d_worder; p.class_counts = ctx->d_wclass;
        if (ctx->whitted_blocks) {
            if (p.n_items > ctx->wcls_cap) {
                if (ctx->d_wcls) cudaFree(ctx->d_wcls);
                ctx->d_wcls = nullptr; ctx->wcls_cap = 0;
                CK(cudaMalloc((void **)&ctx->d_wcls, p.n_items));
                ctx->wcls_cap = p.n_items;
            }
            p.cls = ctx->d_wcls;
            p.filler_items = (uint32_t)((uint64_t)p.n_items * (uint32_t)ctx->whitted_filler_pct / 100u) & ~31u;
        }
    }
    // class 0 one lane per sub-sample, next to the main kernel: needs the class lists and the tables of the grid variants
    F.split0 = (ctx->w_split && p.order && F.grid.cells && F.grid.tiles) ? 1 : 0;
    if (F.split0) p.filler_items = 0;            // no listed pixels in the main kernel: nothing to fill in behind
    {   // tile-word classes depend on nothing but (table, frame size, shard): an unchanged frame keeps its lists and skips the pre-pass
        auto &K = ctx->w_classes;
        p.classes_ready = F.split0 && K.valid && K.gen == ctx->w_table_gen && K.w == ctx->w_w && K.h == ctx->w_h && K.rank == ctx->shard_rank &&
                          K.world == ctx->shard_world && K.tile == ctx->tile_rows && K.order == (const void *)p.order && K.cls == (const void *)p.cls;
        K.valid = false;                                // until this launch has been issued: its pre-pass writes the list buffers
    }
    p.split_blocks_per_sm = ctx->w_split_blocks;
    p.wall_kernel = ctx->w_split == 1 ? 1 : 0;      // RT_TUNE_WHITTED_SPLIT = 2: the general kernel for the wall blocks
    p.split_work_counter = ctx->d_work + 4; p.aux_stream = ctx->aux_stream; p.ev_fork = ctx->ev_fork; p.ev_join = ctx->ev_join;
    CK(cudaMemsetAsync(ctx->d_work, 0, 6 * sizeof(unsigned), ctx->stream));
    if (ctx->counting) CK(cudaMemsetAsync(ctx->d_counters, 0, 8 * sizeof(unsigned long long), ctx->stream));
    if (p.n_items) { CK(rtk_launch_whitted(p, ctx->stream)); ctx->launches += ((p.order && !p.classes_ready) ? 2 : 1) + (p.redo_work_counter ? 1 : 0) + F.split0; }
    if (p.n_items && F.split0) {
        auto &K = ctx->w_classes;
        K.valid = true; K.gen = ctx->w_table_gen; K.w = ctx->w_w; K.h = ctx->w_h; K.rank = ctx->shard_rank; K.world = ctx->shard_world;
        K.tile = ctx->tile_rows; K.order = p.order; K.cls = p.cls;
    }
    return RT_OK;
}

int rt_whitted_download(rt_ctx *ctx, rt_uchar4 *pixels_out, int32_t *hit_id_out) {
    if (!ctx) return RT_ERR_ARG;
    if (!ctx->d_wpixels) return fail(ctx, RT_ERR_STATE, "rt_whitted_download: nothing rendered yet");
    if (hit_id_out && !ctx->w_want_hits) return fail(ctx, RT_ERR_STATE, "rt_whitted_download: hit IDs were not requested at upload");
    if (pixels_out && ctx->peer_wpixels)
        return fail(ctx, RT_ERR_STATE, "rt_whitted_download: this rank renders into rank 0's framebuffer (rt_ipc_import); read the frame there, or rt_ipc_close first");
    CK(cudaSetDevice(ctx->device));
    const size_t px = (size_t)ctx->w_w * ctx->w_h;
    if (pixels_out) CK(cudaMemcpyAsync(pixels_out, ctx->d_wpixels, px * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (hit_id_out) CK(cudaMemcpyAsync(hit_id_out, ctx->d_whits, px * 9 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}

// Copies the rows this rank owns under rt_set_shard from `dev` (a full w x h frame of 4-byte pixels on the device) into the same
// rows of the host frame `host`: the interleaved tiles are one strided 2-D copy, the ragged last tile (if it is ours) a second one.
static cudaError_t copy_owned_rows(const rt_ctx *ctx, uint32_t *host, const uint32_t *dev, int w, int h, cudaStream_t stream) {
    const size_t tile_bytes = (size_t)ctx->tile_rows * w * 4, pitch = tile_bytes * ctx->shard_world;
    const int n_tiles = h / ctx->tile_rows, ragged = h % ctx->tile_rows;          // full tiles, rows of the last partial tile
    const int mine = n_tiles > ctx->shard_rank ? (n_tiles - ctx->shard_rank + ctx->shard_world - 1) / ctx->shard_world : 0;
    const size_t off = (size_t)ctx->shard_rank * ctx->tile_rows * w;
    cudaError_t e = cudaSuccess;
    if (mine > 0) e = cudaMemcpy2DAsync(host + off, pitch, dev + off, pitch, tile_bytes, (size_t)mine, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess && ragged && n_tiles % ctx->shard_world == ctx->shard_rank) {
        const size_t o2 = (size_t)n_tiles * ctx->tile_rows * w;
        e = cudaMemcpyAsync(host + o2, dev + o2, (size_t)ragged * w * 4, cudaMemcpyDeviceToHost, stream);
    }
    return e;
}

int rt_whitted_download_rows(rt_ctx *ctx, rt_uchar4 *frame) {
    if (!ctx) return RT_ERR_ARG;
    if (!frame) return fail(ctx, RT_ERR_ARG, "rt_whitted_download_rows: frame is NULL");
    if (!ctx->d_wpixels) return fail(ctx, RT_ERR_STATE, "rt_whitted_download_rows: nothing rendered yet");
    if (ctx->peer_wpixels) return fail(ctx, RT_ERR_STATE, "rt_whitted_download_rows: this rank renders into rank 0's framebuffer (rt_ipc_import); rt_ipc_close first");
    CK(cudaSetDevice(ctx->device));
    CK(copy_owned_rows(ctx, (uint32_t *)frame, ctx->d_wpixels, ctx->w_w, ctx->w_h, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}

int rt_host_register(rt_ctx *ctx, void *ptr, uint64_t bytes) {
    if (!ctx) return RT_ERR_ARG;
    if (!ptr || !bytes) return fail(ctx, RT_ERR_ARG, "rt_host_register: need a buffer");
    CK(cudaSetDevice(ctx->device));
    CK(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable));
    return RT_OK;
}

int rt_host_unregister(rt_ctx *ctx, void *ptr) {
    if (!ctx) return RT_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaHostUnregister(ptr));
    return RT_OK;
}

int rt_whitted_render(rt_ctx *ctx, const rt_primitive *prims, int n, int w, int h, rt_uchar4 *pixels_out, int32_t *hit_id_out) {
    if (!ctx) return RT_ERR_ARG;
    if (!pixels_out) return fail(ctx, RT_ERR_ARG, "rt_whitted_render: pixels_out is NULL");
    int rc = rt_whitted_upload(ctx, prims, n, w, h, hit_id_out != nullptr);
    if (rc) return rc;
    if ((rc = rt_whitted_launch(ctx))) return rc;
    return rt_whitted_download(ctx, pixels_out, hit_id_out);
}

// ------------------------------------------------------------------------------------------------ raytracer3.0.06

int rt_r306_upload(rt_ctx *ctx, const rt_r306_primitive *prims, int n, int w, int h) {
    if (!ctx) return RT_ERR_ARG;
    if (!prims || n < 1 || w < 1 || h <= 90) return fail(ctx, RT_ERR_ARG, "rt_r306_upload: need prims, n >= 1, w >= 1 and h > 90 (rows 20 .. h-71 are rendered)");
    CK(cudaSetDevice(ctx->device));
    auto &R = ctx->r306;
    build_r306_soa(prims, n, R.soa);
    build_r306_screen(w, h, R.h_sx, R.h_sy, &R.DX, &R.DY);
    if (rtk_whitted_smem_bytes(n, (int)R.soa.lights.size(), (int)R.soa.runs_hot.size() / 3, 2) > (size_t)ctx->max_smem_optin)   // the table the launch stages
        return fail(ctx, RT_ERR_CAPACITY, "rt_r306_upload: %d primitives exceed the %d-byte shared-memory staging of this build", n, ctx->max_smem_optin);
    CK(upload_vec(&R.geom, &R.cap_geom, R.soa.geom, ctx->stream));
    CK(upload_vec(&R.ma, &R.cap_ma, R.soa.mat_a, ctx->stream));
    CK(upload_vec(&R.mb, &R.cap_mb, R.soa.mat_b, ctx->stream));
    CK(upload_vec(&R.flags, &R.cap_flags, R.soa.flags, ctx->stream));
    CK(upload_vec(&R.lights, &R.cap_lights, R.soa.lights, ctx->stream));
    CK(upload_vec(&R.lcenter, &R.cap_lcenter, R.soa.lcenter, ctx->stream));
    CK(upload_vec(&R.runs, &R.cap_runs, R.soa.runs_hot, ctx->stream));
    CK(upload_vec(&R.rrad, &R.cap_rrad, R.soa.rrad, ctx->stream));
    CK(upload_vec(&R.sx, &R.cap_sx, R.h_sx, ctx->stream));
    CK(upload_vec(&R.sy, &R.cap_sy, R.h_sy, ctx->stream));
    const size_t px = (size_t)w * h;
    if (px > R.dest_cap) {
        if (R.dest) cudaFree(R.dest);
        R.dest = nullptr; R.dest_cap = 0;
        CK(cudaMalloc((void **)&R.dest, px * sizeof(uint32_t)));
        R.dest_cap = px;
        CK(cudaMemsetAsync(R.dest, 0, px * sizeof(uint32_t), ctx->stream));      // rows of other ranks read back as 0
    }
    R.w = w; R.h = h; R.n = n; R.nl = (int)R.soa.lights.size(); R.nr = (int)R.soa.runs_hot.size() / 3;
    R.ns = R.soa.n_spheres; R.np = R.soa.n_planes;
    return RT_OK;
}

int rt_r306_launch(rt_ctx *ctx) {
    if (!ctx) return RT_ERR_ARG;
    auto &R = ctx->r306;
    if (!R.geom || !R.dest) return fail(ctx, RT_ERR_STATE, "rt_r306_launch: call rt_r306_upload first");
    CK(cudaSetDevice(ctx->device));
    R306Launch p;
    WFrame &F = p.frame.W;
    F.geom = R.geom; F.mat_a = R.ma; F.mat_b = R.mb; F.flags = R.flags; F.lights = R.lights; F.lcenter = R.lcenter; F.runs = R.runs; F.n_runs = R.nr;
    F.rrad = R.rrad; F.n = R.n; F.n_lights = R.nl; F.n_spheres = R.ns; F.n_planes = R.np;
    F.w = R.w; F.h = R.h; F.DX = R.DX; F.DY = R.DY; F.hit_ids = nullptr;
    F.pcull = nullptr; F.rbox = nullptr; F.cull_rp2 = 0.f; F.reject_k = 0.f; memset(&F.grid, 0, sizeof F.grid); F.split0 = 0; F.tame_reach[0] = F.tame_reach[1] = 0.f; F.redo_count = nullptr; F.redo_list = nullptr; F.redo_cap = 0;
    p.frame.sx = R.sx; p.frame.sy = R.sy; p.frame.row0 = 20; p.frame.row1 = R.h - 70;
    p.shard = make_shard(R.w, R.h, ctx->shard_rank, ctx->shard_world, ctx->tile_rows, &p.n_items);
    p.dest = R.dest; p.work_counter = ctx->d_work; p.sm_count = ctx->sm_count;
    p.order = nullptr; p.class_counts = nullptr; p.subcol = nullptr;
    p.n_valid = (uint32_t)p.shard.local_rows * (uint32_t)R.w;
    if (ctx->r306_split) {
        const size_t need = (size_t)R.w * R.h * 27;
        if (need > R.subcol_cap) {
            if (R.subcol) cudaFree(R.subcol);
            R.subcol = nullptr; R.subcol_cap = 0;
            CK(cudaMalloc((void **)&R.subcol, need * sizeof(float)));
            R.subcol_cap = need;
        }
        p.subcol = R.subcol;
    }
    if (ctx->whitted_sort && p.n_items) {
        if (p.n_items > ctx->worder_cap) {
            if (ctx->d_worder) cudaFree(ctx->d_worder);
            ctx->d_worder = nullptr; ctx->worder_cap = 0;
            CK(cudaMalloc((void **)&ctx->d_worder, 3 * (size_t)p.n_items * sizeof(uint32_t)));
            ctx->worder_cap = p.n_items;
        }
        if (!ctx->d_wclass) CK(cudaMalloc((void **)&ctx->d_wclass, 4 * sizeof(unsigned)));
        p.order = ctx->d_worder; p.class_counts = ctx->d_wclass;
    }
    ctx->w_classes.valid = false;                 // this launch's pre-pass writes the same list buffers
    CK(cudaMemsetAsync(ctx->d_work, 0, 6 * sizeof(unsigned), ctx->stream));
    if (p.n_items) { CK(rtk_launch_r306(p, ctx->stream)); ctx->launches += (p.order ? 2 : 1) + (p.subcol ? 1 : 0); }
    return RT_OK;
}

int rt_r306_download(rt_ctx *ctx, uint32_t *dest) {
    if (!ctx) return RT_ERR_ARG;
    auto &R = ctx->r306;
    if (!R.dest) return fail(ctx, RT_ERR_STATE, "rt_r306_download: nothing rendered yet");
    if (!dest) return fail(ctx, RT_ERR_ARG, "rt_r306_download: dest is NULL");
    CK(cudaSetDevice(ctx->device));
    // only the rows Engine_Render writes (20 .. h-71): everything else in the caller's buffer stays as it was
    const size_t off = (size_t)20 * R.w, rows = (size_t)(R.h - 70 - 20);
    CK(cudaMemcpyAsync(dest + off, R.dest + off, rows * R.w * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}

int rt_r306_render(rt_ctx *ctx, const rt_r306_primitive *prims, int n, int w, int h, uint32_t *dest) {
    if (!ctx) return RT_ERR_ARG;
    if (!dest) return fail(ctx, RT_ERR_ARG, "rt_r306_render: dest is NULL");
    int rc = rt_r306_upload(ctx, prims, n, w, h);
    if (rc) return rc;
    if ((rc = rt_r306_launch(ctx))) return rc;
    return rt_r306_download(ctx, dest);
}

// ------------------------------------------------------------------------------------------------ smallpt

int rt_pt_resize(rt_ctx *ctx, int w, int h, const uint32_t *seeds) {
    if (!ctx) return RT_ERR_ARG;
    if (w < 1 || h < 1 || !seeds) return fail(ctx, RT_ERR_ARG, "rt_pt_resize: need w >= 1, h >= 1 and a seed array of 2*w*h values");
    CK(cudaSetDevice(ctx->device));
    const size_t px = (size_t)w * h;
    if (ctx->peer_ppixels && px > ctx->peer_pcap)
        return fail(ctx, RT_ERR_STATE, "rt_pt_resize: %dx%d exceeds the %zu pixels of the imported rank-0 framebuffer (rt_ipc_close, then share again)", w, h, ctx->peer_pcap);
    if (px > ctx->p_px_cap && ctx->exported_p)
        return fail(ctx, RT_ERR_STATE, "rt_pt_resize: the framebuffer (%zu pixels) is mapped by other ranks and cannot grow to %dx%d; rt_ipc_close on every rank first", ctx->p_px_cap, w, h);
    ctx->have_size = false;
    if (px > ctx->p_px_cap) {              // the buffers are kept across calls of the same (or a smaller) size
        if (ctx->d_colors) cudaFree(ctx->d_colors);
        if (ctx->d_seeds) cudaFree(ctx->d_seeds);
        if (ctx->d_ppixels) cudaFree(ctx->d_ppixels);
        ctx->d_colors = nullptr; ctx->d_seeds = nullptr; ctx->d_ppixels = nullptr; ctx->p_px_cap = 0;
        CK(cudaMalloc((void **)&ctx->d_colors, px * 3 * sizeof(float)));
        CK(cudaMalloc((void **)&ctx->d_seeds, px * 2 * sizeof(uint32_t)));
        CK(cudaMalloc((void **)&ctx->d_ppixels, px * sizeof(uint32_t)));
        ctx->p_px_cap = px;
    }
    CK(cudaMemsetAsync(ctx->d_colors, 0, px * 3 * sizeof(float), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_ppixels, 0, px * sizeof(uint32_t), ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_seeds, seeds, px * 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->p_w = w; ctx->p_h = h; ctx->have_size = true; ctx->current_sample = 0;
    ensure_sincos_table(ctx);              // one-time set-up belongs here, not inside the first timed launch
    return RT_OK;
}

int rt_pt_restore(rt_ctx *ctx, const float *colors, const uint32_t *seeds, int current_sample) {
    if (!ctx) return RT_ERR_ARG;
    if (!ctx->have_size) return fail(ctx, RT_ERR_STATE, "rt_pt_restore: call rt_pt_resize first");
    if (!colors || !seeds || current_sample < 0) return fail(ctx, RT_ERR_ARG, "rt_pt_restore: need colors, seeds and current_sample >= 0");
    CK(cudaSetDevice(ctx->device));
    const size_t px = (size_t)ctx->p_w * ctx->p_h;
    CK(cudaMemcpyAsync(ctx->d_colors, colors, px * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_seeds, seeds, px * 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->current_sample = current_sample;
    return RT_OK;
}

int rt_pt_set_scene(rt_ctx *ctx, const rt_sphere *spheres, uint32_t n) {
    if (!ctx) return RT_ERR_ARG;
    if (!spheres || n < 1) return fail(ctx, RT_ERR_ARG, "rt_pt_set_scene: need at least one sphere");
    for (uint32_t i = 0; i < n; i++)
        if (spheres[i].refl < 0 || spheres[i].refl > 2)
            return fail(ctx, RT_ERR_ARG, "rt_pt_set_scene: sphere %u has material %d (expected 0, 1 or 2)", i, spheres[i].refl);
    CK(cudaSetDevice(ctx->device));
    // The same table again (the viewer's ReInitSceneGPU after a key that moved nothing, a caller that re-sends its scene every frame):
    // the device tables and the hierarchy built from them are still valid; only the sample counter restarts (SPT/smallptGPU.cpp:784-803).
    if (ctx->have_scene && ctx->p_scene_copy.size() == n && memcmp(ctx->p_scene_copy.data(), spheres, (size_t)n * sizeof(rt_sphere)) == 0) {
        ctx->current_sample = 0;
        return RT_OK;
    }
    ctx->p_scene_copy.assign(spheres, spheres + n);
    PtSoA &soa = ctx->p_soa;
    build_pt_soa(spheres, n, soa);
    CK(upload_vec(&ctx->d_pgeom, &ctx->cap_pgeom, soa.geom, ctx->stream));
    CK(upload_vec(&ctx->d_pemis, &ctx->cap_pemis, soa.emis, ctx->stream));
    CK(upload_vec(&ctx->d_pcolr, &ctx->cap_pcolr, soa.colr, ctx->stream));
    CK(upload_vec(&ctx->d_plights, &ctx->cap_plights, soa.lights, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->p_n = (int)n; ctx->p_nl = (int)soa.lights.size(); ctx->have_scene = true; ctx->current_sample = 0;
    ctx->p_bvh_ready = false;
    return RT_OK;
}

int rt_pt_set_camera(rt_ctx *ctx, const rt_camera *cam) {
    if (!ctx) return RT_ERR_ARG;
    if (!cam) return fail(ctx, RT_ERR_ARG, "rt_pt_set_camera: cam is NULL");
    ctx->cam = *cam; ctx->have_cam = true; ctx->current_sample = 0;
    return RT_OK;
}

int rt_pt_set_accumulate_sums(rt_ctx *ctx, int enabled) {
    if (!ctx) return RT_ERR_ARG;
    ctx->sum_mode = enabled ? 1 : 0;
    return RT_OK;
}

int rt_pt_launch(rt_ctx *ctx, int integrator, int n_passes) {
    if (!ctx) return RT_ERR_ARG;
    if (integrator != 0 && integrator != 1) return fail(ctx, RT_ERR_ARG, "rt_pt_launch: integrator must be 0 (path tracing) or 1 (direct lighting)");
    if (n_passes < 1) return fail(ctx, RT_ERR_ARG, "rt_pt_launch: n_passes must be >= 1");
    if (!ctx->have_size || !ctx->have_scene || !ctx->have_cam)
        return fail(ctx, RT_ERR_STATE, "rt_pt_launch: rt_pt_resize, rt_pt_set_scene and rt_pt_set_camera must be called first");
    CK(cudaSetDevice(ctx->device));
    PtLaunch p;
    PtFrame &F = p.frame;
    F.emis = ctx->d_pemis; F.colr = ctx->d_pcolr; F.geom_global = ctx->d_pgeom; F.lights = ctx->d_plights;
    F.n = ctx->p_n; F.n_lights = ctx->p_nl;
    const rt_camera &c = ctx->cam;
    F.cam_ox = c.orig.x; F.cam_oy = c.orig.y; F.cam_oz = c.orig.z;
    F.cam_dx = c.dir.x; F.cam_dy = c.dir.y; F.cam_dz = c.dir.z;
    F.cam_xx = c.x.x; F.cam_xy = c.x.y; F.cam_xz = c.x.z;
    F.cam_yx = c.y.x; F.cam_yy = c.y.y; F.cam_yz = c.y.z;
    F.w = ctx->p_w; F.h = ctx->p_h;
    F.inv_w = 1.f / ctx->p_w; F.inv_h = 1.f / ctx->p_h;          // SPT/smallptCPU.cpp:80-81
    F.pass0 = ctx->current_sample; F.n_passes = n_passes;
    F.direct_only = integrator; F.sum_mode = ctx->sum_mode; F.defer_pack = ctx->sum_mode ? 0 : 1;
    F.sincos_tab = nullptr;
    if (ctx->pt_sincos_table) {
        ensure_sincos_table(ctx);          // normally already there (rt_pt_resize)
        F.sincos_tab = (const f2 *)ctx->d_sincos;      // NULL if the 64 MB could not be had: the kernel then computes the same bits
    }
    p.shard = make_shard(ctx->p_w, ctx->p_h, ctx->shard_rank, ctx->shard_world, ctx->tile_rows, &p.n_items);
    p.colors = ctx->d_colors; p.seeds = ctx->d_seeds; p.pixels = ctx->peer_ppixels ? ctx->peer_ppixels : ctx->d_ppixels;
    p.work_counter = ctx->d_work; p.counters = ctx->d_counters;
    p.count = ctx->counting; p.sm_count = ctx->sm_count;
    p.max_smem_geom = ctx->pt_max_resident_bytes < ctx->max_smem_optin ? ctx->pt_max_resident_bytes : ctx->max_smem_optin;
    p.chunk_spheres = ctx->pt_chunk_spheres;
    p.max_blocks_per_sm = ctx->max_blocks_per_sm;
    p.aligned = ctx->pt_aligned;
    p.use_bvh = !ctx->counting && (ctx->pt_bvh > 0 || (ctx->pt_bvh < 0 && ctx->p_n >= PT_BVH_MIN_SPHERES));
    if (p.use_bvh) {
        if (!ctx->p_bvh_ready) {           // built from the scene tables of the last rt_pt_set_scene, once
            build_pt_bvh(ctx->p_soa.geom, ctx->p_soa.colr, ctx->p_bvh);
            CK(upload_vec(&ctx->d_bnodes, &ctx->cap_bnodes, ctx->p_bvh.nodes, ctx->stream));
            CK(upload_vec(&ctx->d_bgeom, &ctx->cap_bgeom, ctx->p_bvh.geom, ctx->stream));
            CK(upload_vec(&ctx->d_bindex, &ctx->cap_bindex, ctx->p_bvh.index, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            ctx->p_bvh_ready = true;
        }
        p.bvh = ctx->p_bvh.view(ctx->d_bnodes, ctx->d_bgeom, ctx->d_bindex);
    }
    CK(cudaMemsetAsync(ctx->d_work, 0, 6 * sizeof(unsigned), ctx->stream));
    if (ctx->counting) CK(cudaMemsetAsync(ctx->d_counters, 0, 8 * sizeof(unsigned long long), ctx->stream));
    if (p.n_items) { CK(rtk_launch_pt(p, ctx->stream)); ctx->launches += F.defer_pack ? 2 : 1; }
    ctx->current_sample += n_passes;
    return RT_OK;
}

int rt_pt_resolve_sums(rt_ctx *ctx, int total_samples) {
    if (!ctx) return RT_ERR_ARG;
    if (!ctx->have_size || total_samples < 1) return fail(ctx, RT_ERR_STATE, "rt_pt_resolve_sums: nothing to resolve");
    if (ctx->peer_ppixels) return fail(ctx, RT_ERR_STATE, "rt_pt_resolve_sums: the sample-sharded mode keeps a full frame per rank; rt_ipc_close first");
    CK(cudaSetDevice(ctx->device));
    CK(rtk_launch_pt_resolve(ctx->d_colors, ctx->d_ppixels, ctx->p_w, ctx->p_h, 1.f / (float)total_samples, ctx->sm_count, ctx->stream));
    ctx->launches++;
    return RT_OK;
}

int rt_pt_download(rt_ctx *ctx, uint32_t *pixels_out, float *colors_out, uint32_t *seeds_out) {
    if (!ctx) return RT_ERR_ARG;
    if (!ctx->have_size) return fail(ctx, RT_ERR_STATE, "rt_pt_download: rt_pt_resize has not been called");
    if (pixels_out && ctx->peer_ppixels)
        return fail(ctx, RT_ERR_STATE, "rt_pt_download: this rank renders its 8-bit pixels into rank 0's framebuffer (rt_ipc_import); read them there, or rt_ipc_close first");
    CK(cudaSetDevice(ctx->device));
    const size_t px = (size_t)ctx->p_w * ctx->p_h;
    if (pixels_out) CK(cudaMemcpyAsync(pixels_out, ctx->d_ppixels, px * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (colors_out) CK(cudaMemcpyAsync(colors_out, ctx->d_colors, px * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (seeds_out) CK(cudaMemcpyAsync(seeds_out, ctx->d_seeds, px * 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}

int rt_pt_download_rows(rt_ctx *ctx, uint32_t *frame) {
    if (!ctx) return RT_ERR_ARG;
    if (!frame) return fail(ctx, RT_ERR_ARG, "rt_pt_download_rows: frame is NULL");
    if (!ctx->have_size) return fail(ctx, RT_ERR_STATE, "rt_pt_download_rows: rt_pt_resize has not been called");
    if (ctx->peer_ppixels) return fail(ctx, RT_ERR_STATE, "rt_pt_download_rows: this rank renders into rank 0's framebuffer (rt_ipc_import); rt_ipc_close first");
    CK(cudaSetDevice(ctx->device));
    CK(copy_owned_rows(ctx, frame, ctx->d_ppixels, ctx->p_w, ctx->p_h, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}

int rt_pt_render(rt_ctx *ctx, int integrator, int n_passes, uint32_t *pixels_out, float *colors_out, uint32_t *seeds_out) {
    int rc = rt_pt_launch(ctx, integrator, n_passes);
    if (rc) return rc;
    return rt_pt_download(ctx, pixels_out, colors_out, seeds_out);
}

int rt_pt_current_sample(const rt_ctx *ctx) { return ctx ? ctx->current_sample : RT_ERR_ARG; }

// ------------------------------------------------------------------------------------------------ timing / raw access

int rt_sync(rt_ctx *ctx) {
    if (!ctx) return RT_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}

int rt_timer_begin(rt_ctx *ctx) {
    if (!ctx) return RT_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    return RT_OK;
}

int rt_timer_end(rt_ctx *ctx, float *elapsed_ms) {
    if (!ctx || !elapsed_ms) return RT_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaEventSynchronize(ctx->ev1));
    CK(cudaEventElapsedTime(elapsed_ms, ctx->ev0, ctx->ev1));
    return RT_OK;
}

uint64_t rt_launch_count(const rt_ctx *ctx) { return ctx ? ctx->launches : 0; }

int rt_whitted_redo_reports(rt_ctx *ctx, uint32_t *reports_out) {
    if (!ctx || !reports_out) return RT_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(reports_out, ctx->d_work + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}

void *rt_device_buffer(rt_ctx *ctx, int which, uint64_t *bytes) {
    if (!ctx) return nullptr;
    const uint64_t wpx = (uint64_t)ctx->w_w * ctx->w_h, ppx = (uint64_t)ctx->p_w * ctx->p_h;
    void *p = nullptr; uint64_t b = 0;
    switch (which) {
        case RT_BUF_WHITTED_PIXELS: p = ctx->d_wpixels; b = wpx * 4; break;
        case RT_BUF_WHITTED_HITS: p = ctx->w_want_hits ? ctx->d_whits : nullptr; b = wpx * 36; break;
        case RT_BUF_PT_PIXELS: p = ctx->d_ppixels; b = ppx * 4; break;
        case RT_BUF_PT_COLORS: p = ctx->d_colors; b = ppx * 12; break;
        case RT_BUF_PT_SEEDS: p = ctx->d_seeds; b = ppx * 8; break;
        default: break;
    }
    if (bytes) *bytes = p ? b : 0;
    return p;
}

// Handle blob: the 64-byte CUDA IPC handle, then the exporter's capacity in pixels (u64) and a magic word (u64), so that an
// importer can refuse a frame that does not fit instead of storing past the end of rank 0's allocation.
static const uint64_t RT_IPC_MAGIC = 0x3030326274725f49ull;
int rt_ipc_export(rt_ctx *ctx, int which, unsigned char *handle) {
    if (!ctx || !handle) return RT_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64 && RT_IPC_HANDLE_BYTES == 80, "CUDA IPC handle (64 bytes) + capacity + magic");
    void *p = which == RT_BUF_WHITTED_PIXELS ? (void *)ctx->d_wpixels : which == RT_BUF_PT_PIXELS ? (void *)ctx->d_ppixels : nullptr;
    if (!p) return fail(ctx, RT_ERR_STATE, "rt_ipc_export: buffer %d is not allocated (upload / resize first) or cannot be shared", which);
    CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, p));
    const uint64_t cap = which == RT_BUF_WHITTED_PIXELS ? ctx->w_pixels_cap : ctx->p_px_cap;
    memcpy(handle, &h, 64);
    memcpy(handle + 64, &cap, 8);
    memcpy(handle + 72, &RT_IPC_MAGIC, 8);
    (which == RT_BUF_WHITTED_PIXELS ? ctx->exported_w : ctx->exported_p) = true;    // from now on this allocation must not move
    return RT_OK;
}

int rt_ipc_import(rt_ctx *ctx, int which, const unsigned char *handle) {
    if (!ctx || !handle) return RT_ERR_ARG;
    if (which != RT_BUF_WHITTED_PIXELS && which != RT_BUF_PT_PIXELS) return fail(ctx, RT_ERR_ARG, "rt_ipc_import: only the pixel buffers can be redirected");
    uint64_t cap = 0, magic = 0;
    memcpy(&cap, handle + 64, 8);
    memcpy(&magic, handle + 72, 8);
    if (magic != RT_IPC_MAGIC) return fail(ctx, RT_ERR_ARG, "rt_ipc_import: not a handle made by rt_ipc_export (RT_IPC_HANDLE_BYTES = %d bytes)", RT_IPC_HANDLE_BYTES);
    const size_t mine = which == RT_BUF_WHITTED_PIXELS ? (size_t)ctx->w_w * ctx->w_h : (size_t)ctx->p_w * ctx->p_h;
    if (mine > cap) return fail(ctx, RT_ERR_STATE, "rt_ipc_import: this rank's frame (%zu pixels) does not fit rank 0's buffer (%llu pixels)", mine, (unsigned long long)cap);
    CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void *p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    uint32_t **slot = which == RT_BUF_WHITTED_PIXELS ? &ctx->peer_wpixels : &ctx->peer_ppixels;
    if (*slot) cudaIpcCloseMemHandle(*slot);
    *slot = (uint32_t *)p;
    (which == RT_BUF_WHITTED_PIXELS ? ctx->peer_wcap : ctx->peer_pcap) = (size_t)cap;
    return RT_OK;
}

int rt_ipc_close(rt_ctx *ctx) {
    if (!ctx) return RT_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->peer_wpixels) { cudaIpcCloseMemHandle(ctx->peer_wpixels); ctx->peer_wpixels = nullptr; ctx->peer_wcap = 0; }
    if (ctx->peer_ppixels) { cudaIpcCloseMemHandle(ctx->peer_ppixels); ctx->peer_ppixels = nullptr; ctx->peer_pcap = 0; }
    ctx->exported_w = ctx->exported_p = false;     // the caller closes on every rank (a collective): nobody maps this context's frame any more
    return RT_OK;
}

int rt_selftest_math(rt_ctx *ctx, int op, const float *in, void *out, uint64_t n) {
    if (!ctx) return RT_ERR_ARG;
    if (!in || !out || n == 0 || op < 0 || op > 4) return fail(ctx, RT_ERR_ARG, "rt_selftest_math: bad arguments");
    CK(cudaSetDevice(ctx->device));
    const size_t out_bytes = n * (op == 0 || op == 3 || op == 4 ? 8 : 4);
    float *d_in = nullptr; void *d_out = nullptr;
    CK(cudaMalloc((void **)&d_in, n * sizeof(float)));
    cudaError_t e = cudaMalloc(&d_out, out_bytes);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, in, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = rtk_launch_selftest_math(op, d_in, d_out, n, ctx->sm_count, ctx->stream);
    if (e == cudaSuccess) { ctx->launches++; e = cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream); }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_in); if (d_out) cudaFree(d_out);
    if (e != cudaSuccess) return fail(ctx, RT_ERR_CUDA, "rt_selftest_math: %s", cudaGetErrorString(e));
    return RT_OK;
}

long long rt_debug_check_flags(rt_ctx *ctx) {
    if (!ctx) return RT_ERR_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -2;
    return rtk_read_check_flags();
}

void *rt_stream(rt_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int rt_set_stream(rt_ctx *ctx, void *cuda_stream) {
    if (!ctx) return RT_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return RT_OK;
}

}  // extern "C"
