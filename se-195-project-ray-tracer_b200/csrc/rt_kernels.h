// rt_kernels.h -- launch descriptors shared by rt_kernels.cu (device) and rt_api.cu (C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pt_lane.cuh"
#include "whitted_lane.cuh"
#include "r306_lane.cuh"
#include "pt_bvh.cuh"
#include "whitted_bvh.cuh"

// Launch shape (overridable with -D for the A/B builds of tools/variants.sh).
// smallpt: 128-thread CTAs capped at 64 registers (8 CTAs = 32 warps per SM) measured 6 % faster on Cornell than
// 256 threads at the 65 registers ptxas picks on its own, which costs a whole CTA per SM (profiles/r01_ab_variants.txt).
// Whitted: 64-thread CTAs capped at 85 registers (12 CTAs = 24 warps per SM): with the materials read through shared-memory
// pointers the kernel needs 105 registers uncapped; at 85 the few spills cost less than the extra warps bring (3.50 ms
// against 3.86 ms at 1080p), below that they land in the query loops (profiles/r01_ab_variants.txt).  Parking the cold
// part of the lane state in shared memory by hand was slower than ptxas's own choice.
#ifndef PT_THREADS
#define PT_THREADS 128
#endif
#ifndef PT_MIN_BLOCKS
#define PT_MIN_BLOCKS 8
#endif
// Scenes of up to this many spheres run the step-aligned path tracer (see pt_kernel); larger ones are loop-bound.
#ifndef PT_ALIGNED_MAX_SPHERES
#define PT_ALIGNED_MAX_SPHERES 64
#endif
// Large scenes (pt_bvh_kernel): per-lane traversal; the register cap that measured best (profiles/).
#ifndef PT_BVH_MIN_BLOCKS
#define PT_BVH_MIN_BLOCKS 8
#endif
// The shading steps run once this many lanes of a warp hold a finished query (see pt_bvh_kernel).
#ifndef PT_BVH_SHADE_LANES
#define PT_BVH_SHADE_LANES 16
#endif
// Scenes with at least this many spheres walk the hierarchy (RT_TUNE_PT_BVH = -1).
#ifndef PT_BVH_MIN_SPHERES
#define PT_BVH_MIN_SPHERES 48          /* measured crossover (tools/ab_threshold.py): 33 spheres 1.28 / 1.24 ms, 48: 1.48 / 1.53, 64: 1.53 / 1.86, 128: 2.04 / 2.72 */
#endif
// Whitted scenes with at least this many non-light spheres walk the hierarchy (RT_TUNE_WHITTED_BVH = -1).
#ifndef W_BVH_MIN_SPHERES
#define W_BVH_MIN_SPHERES 24           /* tools/ab_threshold_whitted.py: the reference's scene 1 (59 such spheres) 3.4 / 6.3 ms, its scene 0 (7) 5.0 / 2.9 ms */
#endif
#ifndef W_THREADS
#define W_THREADS 64
#endif
#define W_SPLIT_THREADS 288            /* whitted_split_kernel: nine warps, one per sub-sample */
#ifndef W_SPLIT_MIN_BLOCKS
#define W_SPLIT_MIN_BLOCKS 3     /* 27 warps per SM at 72 registers: 1.49 ms against 1.56 at 2 (94 registers, no spills), 1080p */
#endif
#ifndef W_WALL_MIN_BLOCKS
#define W_WALL_MIN_BLOCKS 16          /* whitted_wall_kernel: 32 warps per SM at 64 registers */
#endif
#ifndef W_MIN_BLOCKS
#define W_MIN_BLOCKS 12
#endif


struct PtLaunch {
    rtb::PtFrame frame;
    rtb::Shard shard;
    uint32_t n_items;
    float *colors; uint32_t *seeds; uint32_t *pixels;
    unsigned *work_counter; unsigned long long *counters;
    int count;                  // 1: counting build of the kernel
    int sm_count;
    int max_smem_geom;          // (p, rad^2) arrays larger than this many bytes are streamed in chunks
    int chunk_spheres;          // spheres per chunk in that case
    int max_blocks_per_sm;      // 0 = whatever fits
    int aligned;                // -1: by scene size, 0: plain query loop, 1: step-aligned warps
    int use_bvh;                // 1: walk the hierarchy `bvh` (device pointers) instead of every sphere; ignored by counting launches
    rtb::PtBvh bvh;
};

struct WLaunch {
    rtb::WFrame frame;
    rtb::Shard shard;
    uint32_t n_items;
    uint32_t *pixels;
    unsigned *work_counter; unsigned long long *counters;
    // frame.split0: class 0 goes to whitted_split_kernel on aux_stream (forked from / joined to the launch stream with the two events)
    int classes_ready;          // 1: order / class_counts / cls still hold this frame's classes (same table, size and shard): no pre-pass
    int wall_kernel;            // with split0 and cls: 1 = the class-2 blocks go to whitted_wall_kernel, 0 = to the general kernel
    int split_blocks_per_sm;    // cap on that kernel's resident CTAs per SM (0: as many as fit); the main kernel's CTAs take the rest at once
    unsigned *split_work_counter; cudaStream_t aux_stream; cudaEvent_t ev_fork, ev_join;
    unsigned *redo_work_counter; // NULL: no EXACT launch after the kernel (counting launches); else its work counter (frame.redo_* name the list)
    int count, sm_count, max_blocks_per_sm;
    int stage_mode;             // 3: all tables in static shared memory (small scenes), 2: all tables in dynamic shared memory, 1: geometry / flags / runs only, 0: read through L1 / L2
    int sphere_lights;          // number of lights when all of them are spheres, else 0
    uint32_t *order;            // NULL: screen order; else scratch of 3 x n_items entries: one work list per cost class
    unsigned *class_counts;     // 3 x u32 scratch: entries in each list
    uint8_t *cls;               // NULL, or n_items bytes of scratch: the class of every item; class-2 pixels are then handed out as 8x4 blocks to whole warps
    uint32_t filler_items;      // with cls: the class-2 pixels among items [0, filler_items) are handed out pixel by pixel after the lists (a multiple of 32)
    uint32_t n_valid;           // pixels owned by this rank (n_items minus the padding of the 8x4 blocks)
    int use_bvh;                // 1: frame.runs lists only what is not in the hierarchy `bvh`; queries continue in the tree (not for counting launches)
    rtb::PtBvh bvh;
};

struct R306Launch {
    rtb::R306Frame frame;
    rtb::Shard shard;
    uint32_t n_items;
    uint32_t *dest;
    unsigned *work_counter;
    int sm_count;
    uint32_t *order;            // as in WLaunch: NULL = screen order, else scratch for the cost-class work lists
    unsigned *class_counts;
    uint32_t n_valid;
    float *subcol;              // NULL: one pixel per work unit; else w*h*9*3 floats of scratch: one sub-sample per work unit + a resolve pass
};

cudaError_t rtk_launch_r306(const R306Launch &p, cudaStream_t stream);
cudaError_t rtk_launch_pt(const PtLaunch &p, cudaStream_t stream);
cudaError_t rtk_build_whitted_grid(const rtb::WGrid &G, int gz, uint32_t *cells, const rtb::f4 *geom, const int *flags, const rtb::f2 *pcull, const float *smargin,
                                   const rtb::f4 *lcenter, int n_lights, cudaStream_t stream);
cudaError_t rtk_build_whitted_tiles(uint32_t *tiles, int tiles_x, int tiles_y, int w, int h, float DX, float DY, const rtb::f4 *geom, const int *flags,
                                    const float *smargin, uint32_t all, cudaStream_t stream);
cudaError_t rtk_fill_sincos_table(float *tab /* 2 x 2^23 floats */, int sm_count, cudaStream_t stream);
cudaError_t rtk_launch_pt_resolve(const float *colors, uint32_t *pixels, int w, int h, float inv_total, int sm_count, cudaStream_t stream);
cudaError_t rtk_launch_whitted(const WLaunch &p, cudaStream_t stream);
size_t rtk_whitted_smem_bytes(int n, int n_lights, int n_runs, int stage_mode);
#define W_TAB_CAP 64                              /* stage mode 3: static shared-memory tables for scenes of at most this many primitives ... */
#define W_TAB_RUNS 24                             /* ... in at most this many runs */
#define RT_WHITTED_REDO_CAP 65536u               /* pixels the timed Whitted kernel can hand to the EXACT launch; more: that launch redoes the frame */
#define RTK_WHITTED_STAGE_LIMIT (18 * 1024)      /* per-CTA share of shared memory with 12 resident CTAs per SM */
cudaError_t rtk_launch_selftest_math(int op, const float *in, void *out, unsigned long long n, int sm_count, cudaStream_t stream);
long long rtk_read_check_flags();            /* -DRT_DEVICE_CHECKS builds: bit mask of failed device-side bounds checks (cleared by the read); -1 otherwise */
