/* include/rt_b200.h -- C ABI of the B200-native renderer (librt_b200.so).
 *
 * This is the drop-in boundary for ONE path of markrosoft/se-195-project-ray-tracer: the per-pixel
 * kernels that its single-threaded host programs launch through OpenCL.  The reference has no
 * plugin registry or FFI; "the interface" is each kernel's argument list plus the buffer
 * lifecycle around clEnqueueNDRangeKernel.  Every entry point below names the reference call
 * site it replaces.  Abbreviations (paths relative to the reference root):
 *     SPT/  = smallptgpu-v1.6/
 *     R323/ = Raytracer3.2.03/raytracer/OpenCL Raytracer/
 *
 * Conventions:
 *   - plain C, plain pointers and sizes; no CUDA, torch or C++ types in any signature;
 *   - every host array is caller-allocated and caller-freed; device memory belongs to the context;
 *   - every call returns RT_OK (0) or a negative rt_status; nothing calls exit() (the reference
 *     prints and exit(-1)s: SPT/smallptGPU.cpp:119-122; or returns 1 up to main: R323/raytracer.c:708-713);
 *     rt_last_error() gives the message;
 *   - calls are synchronous unless named *_launch; a context is not thread-safe;
 *   - there is no CPU fallback: without a CUDA device rt_init() fails with RT_ERR_NO_DEVICE.
 *
 * The POD structs are layout-identical to the reference's own (static_asserts in the library),
 * under rt_-prefixed names so that a reference translation unit can include this header next to
 * its own vec.h / geom.h / camera.h / common.h and simply cast.
 */
#ifndef RT_B200_H
#define RT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ data model */

typedef struct { float x, y, z; } rt_vec;                      /* SPT/vec.h:27-29   Vec    (12 B) */
typedef struct { rt_vec o, d; } rt_ray;                        /* SPT/geom.h:32-34  Ray    (24 B) */
enum { RT_DIFF = 0, RT_SPEC = 1, RT_REFR = 2 };                /* SPT/geom.h:39-41  enum Refl     */
typedef struct {                                               /* SPT/geom.h:43-47  Sphere (44 B) */
    float rad;
    rt_vec p, e, c;
    int32_t refl;
} rt_sphere;
typedef struct { rt_vec orig, target, dir, x, y; } rt_camera;  /* SPT/camera.h:29-34 Camera (60 B) */

typedef struct { float x, y, z, w; } rt_float4;                /* R323/common.h:11-13 float_4 */
typedef struct { unsigned char x, y, z, w; } rt_uchar4;        /* R323/common.h:15-17 uchar_4 */
enum { RT_PLANE = 0, RT_SPHERE = 1 };                          /* R323/common.h:24-27 prim_type */
typedef struct {                                               /* R323/common.h:49-63 Primitive_2 (96 B) */
    rt_float4 m_color;
    float m_refl, m_diff, m_refr, m_refr_index, m_spec, dummy_3;
    int32_t type;
    uint8_t is_light, pad_[3];
    rt_float4 normal, center;
    float depth, radius, sq_radius, r_radius;
} rt_primitive;

enum { RT_R306_SPHERE = 1, RT_R306_PLANE = 2 };                /* R306/raytracer.h:13-17 PRIMTYPE */
typedef struct {                                               /* R306/raytracer.h:19-34 Primitive + Material (96 B) */
    int32_t type, m_light;
    rt_vec centre; float sq_radius, radius, r_radius;
    rt_vec plane_n; float plane_d, plane_cell[4];              /* R306/common.h:47-51 plane */
    rt_vec m_color; float m_refl, m_refr, m_diff, m_spec, m_rindex;
} rt_r306_primitive;

typedef enum {
    RT_OK = 0,
    RT_ERR_NO_DEVICE = -1,   /* no usable CUDA device / driver: the product never falls back to the CPU */
    RT_ERR_CUDA = -2,        /* a CUDA runtime call or kernel failed; see rt_last_error() */
    RT_ERR_ARG = -3,         /* bad argument (null pointer, non-positive size, unknown integrator ...) */
    RT_ERR_STATE = -4,       /* call order: e.g. rt_pt_render before rt_pt_resize / set_scene / set_camera */
    RT_ERR_CAPACITY = -5,    /* scene too large for this build's on-chip staging (rt_r306_* only; the other tracers stream or read through L2) */
    RT_ERR_IO = -6           /* scene file could not be read / parsed */
} rt_status;

typedef struct rt_ctx rt_ctx;

/* Work counters of the most recent counted launch (rt_set_counting).  The same quantities the
 * oracle counts; SURVEY.md 8d derives the algorithmic FLOPs from them
 * (smallpt: 17 per sphere test; Whitted: 16 per sphere test, 12 per plane test). */
typedef struct {
    uint64_t nearest_queries;  /* Whitted raytrace() calls (R323/raytracer_non_OpenCL.c:179) / smallpt Intersect() (SPT/geomfunc.h:71) */
    uint64_t shadow_queries;   /* Whitted shadow rays (:223-241) / smallpt IntersectP() (SPT/geomfunc.h:94) */
    uint64_t sphere_tests;     /* sphere_intersect (:111) / SphereIntersect (SPT/geomfunc.h:32) evaluations */
    uint64_t plane_tests;      /* plane_intersect (:95) evaluations; 0 for smallpt */
    uint64_t samples;          /* smallpt: pixel samples; Whitted: primary rays */
} rt_counters;

/* ------------------------------------------------------------------ context */

/* Replaces SetUpOpenCL (SPT/smallptGPU.cpp:209-615) and initialize_openCL (R323/raytracer.c:84-415):
 * binds one CUDA device, creates the stream and the persistent-kernel work counters.
 * One context per GPU; one process per GPU under torchrun passes LOCAL_RANK as `device`. */
int rt_init(rt_ctx **ctx, int device);
/* Replaces FreeBuffers + the clRelease* sequence (SPT/smallptGPU.cpp:76-98; R323/raytracer.c:642-684). */
void rt_destroy(rt_ctx *ctx);
/* Message of the last failing call on this context (ctx may be NULL for rt_init failures). */
const char *rt_last_error(const rt_ctx *ctx);
/* Number of SMs, SM clock (kHz) and name of the bound device; any pointer may be NULL. */
int rt_device_info(const rt_ctx *ctx, int *sm_count, int *sm_clock_khz, char *name, int name_cap);

/* Multi-GPU image sharding (not in the reference, which drives one OpenCL device).
 * The frame is cut into tiles of `tile_rows` rows; this context renders the tiles t with
 * t % world == rank and leaves every other row of its buffers untouched.  Per-pixel indices and
 * seeds stay global, so each pixel runs exactly the 1-GPU instruction stream.  Default 0/1/8. */
int rt_set_shard(rt_ctx *ctx, int rank, int world, int tile_rows);
/* Enable (1) / disable (0) the work counters; counted launches are slower, never time them. */
int rt_set_counting(rt_ctx *ctx, int enabled);
int rt_get_counters(rt_ctx *ctx, rt_counters *out);
/* The raw counter block of the last counting launch: [0..4] = the fields of rt_counters, [5] = Whitted: the shadow rays a TIMED
 * launch traces (a timed launch traces none for hits on a material with neither a diffuse nor a specular term, RNO:242-276 adds
 * nothing for them; counting launches trace them all so that [1] equals the reference's count), [6..7] reserved. */
int rt_get_counters_ex(rt_ctx *ctx, uint64_t *out8);
/* Launch tuning (no reference counterpart; the reference's only knob is the OpenCL work-group size
 * argument of its command line, SPT/RUN_SCENE_*.bat).  Keys: */
enum {
    RT_TUNE_PT_MAX_RESIDENT_BYTES = 0,  /* sphere (p, rad^2) arrays larger than this are streamed through shared memory in chunks */
    RT_TUNE_PT_CHUNK_SPHERES = 1,       /* spheres per chunk in that mode */
    RT_TUNE_MAX_BLOCKS_PER_SM = 2,      /* cap on resident CTAs per SM (0 = as many as fit) */
    RT_TUNE_WHITTED_COST_ORDER = 3,     /* 1 (default): a pre-pass hands out expensive pixels first; 0: screen order.  Same image either way */
    RT_TUNE_PT_ALIGNED = 4,             /* path tracer: 1 = warps run the shading steps in lock-step, 0 = plain query loop, -1 (default) = by scene size.  Same image either way */
    RT_TUNE_WHITTED_BVH = 6,            /* Whitted tracer: the same hierarchy for the non-light spheres of large scenes (rt_whitted_from_spheres tables); planes,
                                           lights and odd spheres are still tested by every query.  1 / 0 / -1 (default: by scene size).  Same image either way */
    RT_TUNE_R306_SPLIT = 7,             /* rt_r306_*: 1 (default) = each of a pixel's nine sub-samples is a work unit of its own and a second pass adds them in the
                                           reference's order; 0 = one pixel per work unit.  Same image either way */
    RT_TUNE_PT_SINCOS_TABLE = 8,        /* path tracer: 1 (default) = sin / cos of the 2^23 angles 2*pi*GetRandom() can take come from a 64 MB table the device
                                           fills once with the function it replaces; 0 = computed per call.  Same image either way */
    RT_TUNE_WHITTED_BLOCKS = 9,         /* Whitted tracer, with COST_ORDER on: 1 (default) = pixels whose centre ray meets neither a reflecting nor a refracting
                                           surface (same cost each) are handed out as whole 8x4 screen blocks per warp, the others pixel by pixel; 0 = all
                                           pixel by pixel.  Same image either way */
    RT_TUNE_WHITTED_FILLER_PCT = 10,    /* ... with BLOCKS on: the first <value> % of the frame's blocks (default 25) are still handed out pixel by pixel, after
                                           the expensive pixels, so that lanes whose expensive pixel is done do not idle.  Same image for any value */
    RT_TUNE_WHITTED_STAGE_CAP = 11,     /* diagnostics: cap on how much of the Whitted scene tables is staged in shared memory -- 3 static tables, 2 all tables,
                                           1 geometry / flags / runs only, 0 nothing (read through L1 / L2); -1 (default) = whatever fits.  Same image */
    RT_TUNE_WHITTED_GRID = 13,          /* Whitted tracer, scenes of at most 32 primitives with sphere lights: 1 (default) = a shadow round tests only the primitives a
                                           per-scene grid of candidate words lists for the hit point's cell, and a nearest round of primary rays only those a per-frame-size
                                           table lists for their 8x4-pixel tile; 2 = the grid without the tiles; 0 = per-hit-point culls only.  Same image for any value */
    RT_TUNE_WHITTED_SPLIT = 14,         /* Whitted tracer, with COST_ORDER and GRID = 1: 1 (default) = the pixels with a refracting surface behind their centre ray are
                                           rendered one lane per SUB-SAMPLE by a second kernel next to the main one, their accumulators added in the reference's
                                           order afterwards (the longest chain of rays one lane traces is 63 instead of 567), and the remaining pure wall blocks by a
                                           straight-line kernel; 2 = the latter by the general kernel; 0 = everything one lane per pixel.  Same image */
    RT_TUNE_WHITTED_SPLIT_BLOCKS = 15,  /* ... resident CTAs per SM of that second kernel (0 = as many as fit, default); the main kernel's CTAs take the rest at once */
    RT_TUNE_WHITTED_REDO_CAP = 12,      /* diagnostics: how many pixels the timed Whitted kernel can report for the exact launch that follows it (blocked lights
                                           whose shade = 0 product is not provably 0, RNO:250, 270) before that launch redoes the whole frame; 0 .. 65536 (default).
                                           Same image for any value */
    RT_TUNE_PT_BVH = 5                  /* path tracer: 1 = sphere queries walk an exact bounding-volume hierarchy (same hits, distances and tie winners as the
                                           reference's loop over every sphere), 0 = the loop, -1 (default) = by scene size.  Same image either way */
};
int rt_set_tuning(rt_ctx *ctx, int key, int value);

/* ------------------------------------------------------------------ Whitted tracer
 *
 * One call = one frame of R323/raytracer_kernel.cl:246-383 (`raytracer_kernel`), with the numerics
 * of its CPU twin R323/raytracer_non_OpenCL.c:285-450 (`raytracer_non_kernel`), which is what
 * release 3.2.03 actually runs (R323/raytracer.c:752-756) and what the golden test.bmp shows. */

/* Replaces run_openCL_kernel (R323/raytracer.c:417-640: clSetKernelArg x6, clEnqueueNDRangeKernel,
 * clWaitForEvents, clEnqueueReadBuffer) and has the CPU twin's signature
 * raytracer_non_kernel(uchar_4*, int, int, Primitive_2*, int) (R323/raytracer.c:11-16) plus the
 * context and a hit-ID tap.  pixels_out: w*h uchar4, row-major, top row first, (r,g,b,0).
 * hit_id_out: NULL, or int[w*h*9]: the primitive index returned by raytrace() for each of the
 * nine primary rays of a pixel (sub-sample index (tx+1)*3+(ty+1), -1 = miss).
 * Host buffers in, host buffers out; the copies are part of the call. */
int rt_whitted_render(rt_ctx *ctx, const rt_primitive *prims, int n, int w, int h,
                      rt_uchar4 *pixels_out, int32_t *hit_id_out);

/* The same frame split into its three steps, so that a caller (bench.py) can keep the scene and
 * the framebuffer resident in HBM: upload = clCreateBuffer + clEnqueueWriteBuffer
 * (R323/raytracer.c:303-345), launch = clEnqueueNDRangeKernel (:561-570, asynchronous on the
 * context's stream), download = clEnqueueReadBuffer (:600-609, blocking). */
int rt_whitted_upload(rt_ctx *ctx, const rt_primitive *prims, int n, int w, int h, int want_hit_ids);
int rt_whitted_launch(rt_ctx *ctx);
int rt_whitted_download(rt_ctx *ctx, rt_uchar4 *pixels_out, int32_t *hit_id_out);

/* ------------------------------------------------------------------ raytracer3.0.06 (BASELINE config 1)
 *
 * The CPU tracer of raytracer3.0.06.no_rec.samp -- the reference's own baseline program -- as a GPU frame:
 * Engine_SetTarget + Engine_InitRender + Engine_Render (R306/raytracer.cpp:17-23, :278-530) over Engine_Raytrace
 * (:30-271).  Same intersection core as the 3.2.03 tracer, different everything else (63-node implicit ray tree folded
 * bottom-up, running-sum screen coordinates, rows 20 .. h-71 only, 0x00RRGGBB pixels); see csrc/r306_lane.cuh.
 * dest: w*h Pixels (unsigned int); like the reference, rows outside 20 .. h-71 are left untouched, so h must be > 90. */
int rt_r306_render(rt_ctx *ctx, const rt_r306_primitive *prims, int n, int w, int h, uint32_t *dest);
/* The same in three steps (scene + screen tables to HBM, asynchronous kernel, blocking read-back of the rendered rows). */
int rt_r306_upload(rt_ctx *ctx, const rt_r306_primitive *prims, int n, int w, int h);
int rt_r306_launch(rt_ctx *ctx);
int rt_r306_download(rt_ctx *ctx, uint32_t *dest);

/* ------------------------------------------------------------------ smallpt path tracer
 *
 * Replaces the RadianceGPU kernel (SPT/rendering_kernel.cl:53-97 and rendering_kernel_dl.cl) and
 * the host glue around it.  Seed <-> pixel mapping is the CPU twin's: pixel (x,y) uses
 * seeds[2*i], seeds[2*i+1] and colors[i] with i = (h-1-y)*w + x (SPT/smallptCPU.cpp:86-90),
 * and writes pixels[y*w + x]. */

/* Replaces AllocateBuffers (SPT/smallptGPU.cpp:100-167): sizes the colour / seed / pixel buffers
 * and uploads the caller's seeds (2*w*h u32, each >= 2 by the reference's rule at :106-110; seeds
 * are never generated inside the library).  Resets the sample counter. */
int rt_pt_resize(rt_ctx *ctx, int w, int h, const uint32_t *seeds);
/* Replaces the sphere upload of SetUpOpenCL / ReInitSceneGPU (SPT/smallptGPU.cpp:489-498, 784-803).
 * Resets the sample counter. */
int rt_pt_set_scene(rt_ctx *ctx, const rt_sphere *spheres, uint32_t n);
/* Replaces the camera upload of ReInitGPU (SPT/smallptGPU.cpp:805-830); `cam` must already hold
 * dir/x/y as computed by UpdateCamera (SPT/displayfunc.cpp:182-195; rt_update_camera below).
 * Resets the sample counter. */
int rt_pt_set_camera(rt_ctx *ctx, const rt_camera *cam);
/* Replaces UpdateRenderingGPU (SPT/smallptGPU.cpp:642-782): runs n_passes more sample passes
 * (the reference launches one kernel per pass; here the pass loop is inside one kernel) starting
 * at the context's currentSample, then copies back what is asked for:
 * pixels_out w*h u32 (r | g<<8 | b<<16), colors_out 3*w*h float, seeds_out 2*w*h u32; any may be NULL.
 * integrator: 0 = RadiancePathTracing (SPT/geomfunc.h:167), 1 = RadianceDirectLighting (:340). */
int rt_pt_render(rt_ctx *ctx, int integrator, int n_passes,
                 uint32_t *pixels_out, float *colors_out, uint32_t *seeds_out);
/* Asynchronous launch only (state stays in HBM), and the matching blocking read-back. */
int rt_pt_launch(rt_ctx *ctx, int integrator, int n_passes);
int rt_pt_download(rt_ctx *ctx, uint32_t *pixels_out, float *colors_out, uint32_t *seeds_out);
/* Checkpoint / resume of a progressive render (the reference keeps this state only in its
 * buffers and loses it on exit).  The state is exactly (colors, seeds, currentSample): save it with
 * rt_pt_download + rt_pt_current_sample; after rt_pt_resize / set_scene / set_camera on any context,
 * rt_pt_restore puts it back and the next rt_pt_render continues bit-identically. */
int rt_pt_restore(rt_ctx *ctx, const float *colors, const uint32_t *seeds, int current_sample);
/* currentSample of the reference (SPT/smallptGPU.cpp:60): passes accumulated so far. */
int rt_pt_current_sample(const rt_ctx *ctx);
/* Sample-sharded progressive mode: colors hold running SUMS instead of running
 * means, so that per-rank buffers can be added with ncclAllReduce; rt_pt_resolve_sums() then
 * divides by total_samples and writes the 8-bit pixels.  Not bit-identical to the reference's
 * sequential running mean; judged by RMSE only. */
int rt_pt_set_accumulate_sums(rt_ctx *ctx, int enabled);
int rt_pt_resolve_sums(rt_ctx *ctx, int total_samples);

/* ------------------------------------------------------------------ timing and raw access */

int rt_sync(rt_ctx *ctx);
/* CUDA events on the context's stream: begin/end bracket any number of *_launch calls. */
int rt_timer_begin(rt_ctx *ctx);
int rt_timer_end(rt_ctx *ctx, float *elapsed_ms);   /* synchronises on the end event */
/* Render-path kernels launched by this context so far (one-time set-up kernels such as table fills are not counted). */
uint64_t rt_launch_count(const rt_ctx *ctx);
/* Diagnostics: how many shadow batches of the last timed rt_whitted_launch were reported for the exact launch (synchronises). */
int rt_whitted_redo_reports(rt_ctx *ctx, uint32_t *reports_out);
/* Device addresses of the context's buffers, for collectives issued by the caller (NCCL through
 * torch.distributed in bench.py) -- returns NULL if not allocated.  which: */
enum { RT_BUF_WHITTED_PIXELS = 0, RT_BUF_WHITTED_HITS = 1, RT_BUF_PT_PIXELS = 2, RT_BUF_PT_COLORS = 3, RT_BUF_PT_SEEDS = 4 };
void *rt_device_buffer(rt_ctx *ctx, int which, uint64_t *bytes);
/* Fused frame assembly over NVLink (second option): rank 0 exports its pixel buffer as a
 * RT_IPC_HANDLE_BYTES-byte handle (rt_ipc_export, after rt_whitted_upload / rt_pt_resize); every other rank's process
 * imports it (rt_ipc_import) and from then on its render kernel stores the pixels of the rows it owns straight
 * into rank 0's frame through the peer mapping -- the transfer rides inside the kernel, tile by tile, and no
 * gather step exists.  The caller only has to order "all ranks' kernels done" before rank 0 reads the frame
 * (any barrier collective on the render stream).  which: RT_BUF_WHITTED_PIXELS or RT_BUF_PT_PIXELS. */
/* Multi-GPU read-back without the hop through rank 0 (the end-to-end path of bench.py at N > 1): `frame` is ONE full w x h host
 * frame shared by all ranks' processes (POSIX shared memory, optionally page-locked in each process with rt_host_register); every
 * rank copies just the rows it owns under rt_set_shard into their place, over its own PCIe link.  After all ranks have returned
 * (a host barrier is the caller's), the frame is complete -- the same bytes rt_whitted_download / rt_pt_download give on one GPU.
 * Blocking, like the reference's clEnqueueReadBuffer(CL_TRUE) (SPT/smallptGPU.cpp:757-770, R323/raytracer.c:600-640). */
int rt_whitted_download_rows(rt_ctx *ctx, rt_uchar4 *frame);
int rt_pt_download_rows(rt_ctx *ctx, uint32_t *frame /* pixels, w*h */);
int rt_host_register(rt_ctx *ctx, void *ptr, uint64_t bytes);      /* cudaHostRegister: makes copies to `ptr` true DMA */
int rt_host_unregister(rt_ctx *ctx, void *ptr);
#define RT_IPC_HANDLE_BYTES 80   /* CUDA IPC handle (64) + the exporter's capacity in pixels (8) + a magic word (8) */
/* Lifetime rules (errors are RT_ERR_STATE): an exported framebuffer is never reallocated -- an upload / resize that would
 * have to grow it fails until rt_ipc_close has been called (on every rank; it is the caller's collective); an importer
 * refuses a frame larger than the exporter's capacity, at import and at every later upload / resize; while a mapping is
 * open, rt_*_download of the PIXELS on the importing rank fails (they live in rank 0's frame). */
int rt_ipc_export(rt_ctx *ctx, int which, unsigned char *handle /* RT_IPC_HANDLE_BYTES */);
int rt_ipc_import(rt_ctx *ctx, int which, const unsigned char *handle /* RT_IPC_HANDLE_BYTES */);
int rt_ipc_close(rt_ctx *ctx);
/* Diagnostics.  A library built with -DRT_DEVICE_CHECKS (tools/checked_build.sh; not the product build) verifies on the device every
 * index its kernels form -- ray FIFO depth, traversal stack depth, pixel and hit-ID addresses, work-list and class-table positions,
 * staged-table sizes, the 3.0.06 ray tree, table look-ups -- and records a failed check as a bit (RT_CHK_* of csrc/rt_math.cuh)
 * instead of faulting.  Returns that mask since the last call (0 = clean) after synchronising, -1 when the checks are not compiled in. */
long long rt_debug_check_flags(rt_ctx *ctx);
/* The context's cudaStream_t as an opaque pointer (for callers that order their own work after it). */
void *rt_stream(rt_ctx *ctx);
/* Makes the context issue all its work on a caller-owned cudaStream_t (e.g. torch's current stream, so
 * that the caller's collectives and CUDA events are ordered with the render kernels).  The reference
 * analogue is the single in-order cl_command_queue every call shares (SPT/smallptGPU.cpp:463-467). */
int rt_set_stream(rt_ctx *ctx, void *cuda_stream);

/* Evaluates, ON THE DEVICE, the elementary functions the kernels use in place of the reference's libm
 * calls, so that a test can compare them with the host libm the reference's CPU path binds to.
 * op: 0 sinf+cosf (out: float[2n], sin then cos; SPT/geomfunc.h:66-67, 261-262), 1 expf (float[n];
 * R323/raytracer_non_OpenCL.c:424-426), 2 toInt = gamma + 8-bit quantisation (int[n]; SPT/vec.h:62),
 * 3 square root (float[2n]: the loops' grouped fast path, then __fsqrt_rn), 4 x^20 in double (double[n];
 * R323/raytracer_non_OpenCL.c:270).  Diagnostics only; host buffers in and out. */
enum { RT_SELFTEST_SINCOS = 0, RT_SELFTEST_EXPF = 1, RT_SELFTEST_GAMMA = 2, RT_SELFTEST_SQRT = 3, RT_SELFTEST_POW20 = 4 };
int rt_selftest_math(rt_ctx *ctx, int op, const float *in, void *out, uint64_t n);

/* ------------------------------------------------------------------ host-side scene helpers
 * (kept from the reference's host code; pure CPU, no device needed) */

/* UpdateCamera (SPT/displayfunc.cpp:182-195): derives dir, x, y from orig, target and the image size. */
void rt_update_camera(rt_camera *cam, int w, int h);
/* keyFunc / specialFunc of the reference viewer (SPT/displayfunc.cpp:252-420) without GLUT: applies ONE key press
 * to the caller's camera and sphere table exactly as the viewer does (same float / double expression order, including
 * the rotation keys' use of the already-updated component) and returns what the viewer would do next:
 *   RT_KEY_NONE     nothing changed ('h', unknown keys)
 *   RT_KEY_CAMERA   ReInit(0): the camera moved and rt_update_camera has been applied -> rt_pt_set_camera
 *   RT_KEY_SCENE    ReInitScene(): the selected sphere moved or the selection changed -> rt_pt_set_scene
 *   RT_KEY_RESTART  ' ': ReInit(1), the viewer frees and re-allocates its buffers -> rt_pt_resize with fresh seeds
 *   RT_KEY_DUMP     'p': write image.ppm -> rt_write_ppm;    RT_KEY_QUIT  Escape
 * Every one of CAMERA / SCENE / RESTART restarts the progressive image at sample 0, which is what
 * rt_pt_set_camera / rt_pt_set_scene / rt_pt_resize do.  Keys: the viewer's characters; the GLUT special keys are
 * passed as RT_KEY_SPECIAL + GLUT code (UP 101, DOWN 103, LEFT 100, RIGHT 102, PAGE_UP 104, PAGE_DOWN 105).
 * *current_sphere is the viewer's selection ('+' / '-'), 0 at start. */
enum { RT_KEY_NONE = 0, RT_KEY_CAMERA = 1, RT_KEY_SCENE = 2, RT_KEY_RESTART = 3, RT_KEY_DUMP = 4, RT_KEY_QUIT = 5 };
#define RT_KEY_SPECIAL 0x100
int rt_viewer_key(int key, rt_camera *cam, int w, int h, rt_sphere *spheres, uint32_t n, uint32_t *current_sphere);
/* ReadScene (SPT/displayfunc.cpp:120-180): parses a .scn file.  *spheres_out is malloc'd
 * (free with rt_free); cam_out receives orig/target only.  Returns RT_OK or RT_ERR_IO. */
int rt_read_scene(const char *path, rt_camera *cam_out, rt_sphere **spheres_out, uint32_t *n_out);
/* Writes what `perl SPT/scene_build_complex.pl` prints for the given $maxDepth, preceded by the
 * camera / size / light / floor header of SPT/scenes/complex.scn (lines 1-4). */
int rt_write_complex_scene(const char *path, int max_depth);
/* create_scene + the Primitive -> Primitive_2 copy (R323/scene.c:48-128, R323/raytracer.c:721-746).
 * which = CHOOSE_SCENE (0: 17 slots, 1: 64 slots).  Returns the primitive count, or <0. */
int rt_whitted_create_scene(int which, rt_primitive *out, int cap);
/* A smallpt sphere table (.scn scene, e.g. the generated complex scenes) as a Whitted scene (no
 * reference counterpart -- the two programs never shared scenes).  Each sphere becomes a create_sphere record
 * (R323/scene.c:36-46: sq_radius, r_radius derived the same way); an emitter becomes a light whose colour is e scaled
 * to a maximum of 1, DIFF -> m_diff 1 with m_spec 0.5, SPEC -> m_refl 1, REFR -> m_refr 1 at index 1.5 with m_refl 0.1.
 * The Whitted tracer's eye is fixed at (0, 0.25, -7) looking down +z (R323/raytracer_non_OpenCL.c:299-316); with a
 * camera (orig/target as read from the .scn file) the spheres are moved into that frame -- rotated into the camera's
 * axes and scaled so that the target lies 14 units in front of the eye; with cam == NULL coordinates are kept.
 * Returns the primitive count (= n), or RT_ERR_ARG when cap < n. */
int rt_whitted_from_spheres(const rt_sphere *spheres, uint32_t n, const rt_camera *cam, rt_primitive *out, int cap);
/* Scene_InitScene (R306/scene.cpp:217-272): the 17 primitives of the 3.0.06 program.  Returns the count, or <0. */
int rt_r306_create_scene(rt_r306_primitive *out, int cap);
/* write_bmp_file (R323/bitmap.c:8-75): 24-bit BMP, bottom row first, BGR. */
int rt_write_bmp(const char *path, const rt_uchar4 *pixels, int w, int h);
/* The 'p' key of the reference viewer (SPT/displayfunc.cpp:254-271): P3 PPM, bottom row first. */
int rt_write_ppm(const char *path, const uint32_t *pixels, int w, int h);
void rt_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* RT_B200_H */
