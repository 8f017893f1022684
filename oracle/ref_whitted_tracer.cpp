// oracle/_ref harness, part A -- TEST INFRASTRUCTURE ONLY, never linked into the product.
//
// Wraps the reference's own Whitted CPU twin, compiled UNMODIFIED from where it lies under
// /root/reference (Raytracer3.2.03/raytracer/OpenCL Raytracer/raytracer_non_OpenCL.c) by
// including it into this translation unit as C++ (the reference's vcxproj builds it /TP).
// Nothing from the reference is copied into this repository; the include path is supplied
// by oracle/Makefile.
#include <math.h>
#include <string.h>
#include <pthread.h>
#include <stdlib.h>
#include "raytracer_non_OpenCL.c"   // -> Primitive_2, Ray, raytrace(), raytracer_non_kernel()

static_assert(sizeof(Primitive_2) == 96, "reference Primitive_2 must be 96 bytes");
static_assert(sizeof(uchar_4) == 4, "uchar_4");

extern "C" {

// The reference frame function itself (raytracer_non_OpenCL.c:285-450).
void ref_whitted_render(uchar_4 *pixels, int w, int h, const void *prims, int n) {
    raytracer_non_kernel(pixels, w, h, (Primitive_2 *)prims, n);
}

// Primary-hit tap: the return value of the reference raytrace() (:179-281) for every ORIGIN ray.
// The nine primary rays of a pixel are rebuilt here exactly as :299-328 build them; everything
// after that is the reference function.  hit_ids is int[h*w*9], sub-sample index (tx+1)*3+(ty+1).
// dist_out / result_out may be NULL.
void ref_whitted_primary_hits(int *hit_ids, float *dist_out, int *result_out,
                              int w, int h, const void *prims_v, int n) {
    Primitive_2 *prims = (Primitive_2 *)prims_v;
    const float WX1 = -3.0f, WX2 = 3.0f, WY1 = 2.25f, WY2 = -2.25f;
    const float DX = (WX2 - WX1) / w, DY = (WY2 - WY1) / h;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const float SY = WY1 + y * DY, SX = WX1 + x * DX;
            float_4 cam; cam.x = 0; cam.y = 0.25f; cam.z = -7.0f; cam.w = 0;
            for (int tx = -1; tx < 2; tx++)
                for (int ty = -1; ty < 2; ty++) {
                    Ray r;
                    r.direction.x = SX + DX * (tx / 2.0f) - cam.x;
                    r.direction.y = SY + DY * (ty / 2.0f) - cam.y;
                    r.direction.z = 0 - cam.z;
                    r.direction.w = 0;
                    float len = 1.0f / sqrt(r.direction.x * r.direction.x + r.direction.y * r.direction.y +
                                            r.direction.z * r.direction.z);
                    r.direction.x *= len; r.direction.y *= len; r.direction.z *= len;
                    r.origin = cam; r.weight = 1.0f; r.depth = 0; r.origin_primitive = -1;
                    r.type = ORIGIN; r.r_index = 1.0f;
                    r.transparency.x = r.transparency.y = r.transparency.z = 1; r.transparency.w = 0;
                    Color_2 col; col.x = col.y = col.z = col.w = 0;
                    float dist; float_4 pi; int result = 0;
                    int id = raytrace(&r, &col, &dist, &pi, &result, prims, n);
                    size_t k = ((size_t)y * w + x) * 9 + (size_t)((tx + 1) * 3 + (ty + 1));
                    hit_ids[k] = id;
                    if (dist_out) dist_out[k] = dist;
                    if (result_out) result_out[k] = id < 0 ? 0 : result;
                }
        }
}

// Known-answer probe: centre sub-sample primary ray of one pixel.
int ref_whitted_probe(int x, int y, int w, int h, const void *prims_v, int n,
                      float *dist, float *col3, int *result) {
    Primitive_2 *prims = (Primitive_2 *)prims_v;
    const float DX = (3.0f - -3.0f) / w, DY = (-2.25f - 2.25f) / h;
    const float SY = 2.25f + y * DY, SX = -3.0f + x * DX;
    Ray r; memset(&r, 0, sizeof r);
    r.origin.x = 0; r.origin.y = 0.25f; r.origin.z = -7.0f;
    r.direction.x = SX + DX * (0 / 2.0f) - r.origin.x;
    r.direction.y = SY + DY * (0 / 2.0f) - r.origin.y;
    r.direction.z = 0 - r.origin.z;
    float len = 1.0f / sqrt(r.direction.x * r.direction.x + r.direction.y * r.direction.y +
                            r.direction.z * r.direction.z);
    r.direction.x *= len; r.direction.y *= len; r.direction.z *= len;
    r.weight = 1.0f; r.origin_primitive = -1; r.type = ORIGIN; r.r_index = 1.0f;
    r.transparency.x = r.transparency.y = r.transparency.z = 1;
    Color_2 c; c.x = c.y = c.z = c.w = 0; float_4 pi; *result = 0;
    int id = raytrace(&r, &c, dist, &pi, result, prims, n);
    col3[0] = c.x; col3[1] = c.y; col3[2] = c.z;
    return id;
}

// Throughput arm for bench.py --impl reference: `threads` host threads, each rendering its own
// full frame with the unmodified reference function (which cannot be split by rows: its window
// mapping depends on the full height).  Returns nothing; the caller times it.
struct FrameJob { uchar_4 *px; int w, h; const void *prims; int n; };
static void *frame_thread(void *p) {
    FrameJob *j = (FrameJob *)p;
    raytracer_non_kernel(j->px, j->w, j->h, (Primitive_2 *)j->prims, j->n);
    return 0;
}
void ref_whitted_render_mt(uchar_4 *pixels /* threads*w*h */, int w, int h, const void *prims, int n, int threads) {
    pthread_t *t = (pthread_t *)malloc(sizeof(pthread_t) * threads);
    FrameJob *j = (FrameJob *)malloc(sizeof(FrameJob) * threads);
    for (int i = 0; i < threads; i++) {
        j[i].px = pixels + (size_t)i * w * h; j[i].w = w; j[i].h = h; j[i].prims = prims; j[i].n = n;
        pthread_create(&t[i], 0, frame_thread, &j[i]);
    }
    for (int i = 0; i < threads; i++) pthread_join(t[i], 0);
    free(t); free(j);
}

}  // extern "C"
