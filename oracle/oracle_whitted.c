/* oracle/oracle_whitted.c -- CPU restatement of the reference's Whitted tracer.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Restates, in this repository's own words, the
 * algorithm of  Raytracer3.2.03/raytracer/OpenCL Raytracer/raytracer_non_OpenCL.c  (cited
 * below as RNO:line), adding what the reference lacks: a primary-hit-ID tap, ray/test counters,
 * row ranges (for host threads) and a defined behaviour on a miss (RNO:373 reads prims[-1]).
 * Pinned: tests/test_oracle_whitted.py checks it byte-for-byte against oracle/_ref (the
 * reference compiled unmodified) and against the golden test.bmp fixture.
 *
 * Every float expression keeps the reference's C evaluation order; compile without contraction.
 */
#include "oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define TRACE_DEPTH 5          /* RNO:4 */
#define RAY_EPS 0.001f         /* RNO:26 */
#define FAR_AWAY 10000000.0f   /* RNO:181 */
#define KIND_HIT 1             /* RNO:22-24 */
#define KIND_INSIDE (-1)

enum { RAY_PRIMARY = 0, RAY_REFLECTED = 1, RAY_REFRACTED = 2 };   /* RNO:59-63 */

typedef struct { float x, y, z; } v3;

typedef struct {
    v3 o, d;
    float weight, depth, r_index;
    int from_prim, kind;
    v3 transp;
} wray;                         /* RNO:83-92 minus the unused w lanes */

static inline float dot3(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }   /* RNO:50 */

/* RNO:95-109.  Plane normals are not unit length in the reference scenes; nothing normalises them. */
static int test_plane(const ow_prim *p, v3 o, v3 d, float *nearest, ow_counters *c) {
    if (c) c->plane_tests++;
    v3 nrm = { p->normal.x, p->normal.y, p->normal.z };
    float denom = dot3(nrm, d);
    if (denom != 0) {
        float t = -(dot3(nrm, o) + p->depth) / denom;
        if (t > 0 && t < *nearest) { *nearest = t; return KIND_HIT; }
    }
    return 0;
}

/* RNO:111-148 */
static int test_sphere(const ow_prim *p, v3 o, v3 d, float *nearest, ow_counters *c) {
    if (c) c->sphere_tests++;
    v3 v = { o.x - p->center.x, o.y - p->center.y, o.z - p->center.z };
    float b = -dot3(v, d);
    float disc = (b * b) - dot3(v, v) + p->sq_radius;
    if (disc > 0) {
        float root = sqrtf(disc);
        float t_near = b - root, t_far = b + root;
        if (t_far > 0) {
            if (t_near < 0) {
                if (t_far < *nearest) { *nearest = t_far; return KIND_INSIDE; }
            } else if (t_near < *nearest) { *nearest = t_near; return KIND_HIT; }
        }
    }
    return 0;
}

static inline int test_prim(const ow_prim *p, v3 o, v3 d, float *nearest, ow_counters *c) {   /* RNO:150-160 */
    if (p->type == 0) return test_plane(p, o, d, nearest, c);
    if (p->type == 1) return test_sphere(p, o, d, nearest, c);
    return 0;
}

static inline v3 surface_normal(const ow_prim *p, v3 at) {   /* RNO:162-177 */
    v3 n = { 0, 0, 0 };
    if (p->type == 0) { n.x = p->normal.x; n.y = p->normal.y; n.z = p->normal.z; }
    else if (p->type == 1) {
        n.x = (at.x - p->center.x) * p->r_radius;
        n.y = (at.y - p->center.y) * p->r_radius;
        n.z = (at.z - p->center.z) * p->r_radius;
    }
    return n;
}

/* One ray: nearest hit (ascending index, strict '<' so the lower index keeps an exact tie),
 * then local shading with one hard-shadow ray per sphere light.  RNO:179-281.
 * Returns the hit primitive or -1. */
static int trace_one(const wray *ray, const ow_prim *prims, int n, v3 *col, float *dist, v3 *at,
                     int *kind, ow_counters *c) {
    if (c) c->traced_rays++;
    float nearest = FAR_AWAY;
    int hit = -1;
    for (int s = 0; s < n; s++) {
        int k = test_prim(&prims[s], ray->o, ray->d, &nearest, c);
        if (k) { hit = s; *kind = k; }
    }
    *dist = nearest;
    if (hit < 0) return -1;
    const ow_prim *hp = &prims[hit];
    if (hp->is_light) { col->x = hp->color.x; col->y = hp->color.y; col->z = hp->color.z; return hit; }

    at->x = ray->o.x + ray->d.x * nearest;
    at->y = ray->o.y + ray->d.y * nearest;
    at->z = ray->o.z + ray->d.z * nearest;
    for (int l = 0; l < n; l++) {
        const ow_prim *lp = &prims[l];
        if (!lp->is_light) continue;
        float lit = 1.0f;
        v3 to = { lp->center.x - at->x, lp->center.y - at->y, lp->center.z - at->z };
        float reach = sqrtf(to.x * to.x + to.y * to.y + to.z * to.z);
        v3 L = { (1.0f / reach) * to.x, (1.0f / reach) * to.y, (1.0f / reach) * to.z };
        if (lp->type == 1) {   /* only sphere lights cast shadows, RNO:223 */
            if (c) c->shadow_rays++;
            v3 so = { at->x + L.x * RAY_EPS, at->y + L.y * RAY_EPS, at->z + L.z * RAY_EPS };
            for (int s = 0; s < n; s++)
                if (!prims[s].is_light && test_prim(&prims[s], so, L, &reach, c)) { lit = 0; break; }
        }
        v3 N = surface_normal(hp, *at);
        if (hp->diff > 0) {   /* RNO:244-254 */
            float nl = dot3(N, L);
            if (nl > 0) {
                float k = nl * hp->diff * lit;
                col->x += k * hp->color.x * lp->color.x;
                col->y += k * hp->color.y * lp->color.y;
                col->z += k * hp->color.z * lp->color.z;
            }
        }
        if (hp->spec > 0) {   /* RNO:256-275 */
            float ln = dot3(L, N);
            v3 R = { L.x - 2.0f * ln * N.x, L.y - 2.0f * ln * N.y, L.z - 2.0f * ln * N.z };
            float vr = dot3(ray->d, R);
            if (vr > 0) {
                /* pow(float,int) is the double overload in the reference's C++ build; the product
                 * with m_spec and shade is therefore formed in double and rounded to float once. */
                float k = (float)(pow((double)vr, 20.0) * (double)hp->spec * (double)lit);
                col->x += k * lp->color.x;
                col->y += k * lp->color.y;
                col->z += k * lp->color.z;
            }
        }
    }
    return hit;
}

void oracle_whitted_rows(uint8_t *pixels, int32_t *hit_ids, int w, int h, int y0, int y1,
                         const ow_prim *prims, int n, ow_counters *ctr) {
    const float WX1 = -3.0f, WX2 = 3.0f, WY1 = 2.25f, WY2 = -2.25f;   /* RNO:291-294 */
    const float DX = (WX2 - WX1) / w, DY = (WY2 - WY1) / h;
    for (int y = y0; y < y1; y++)
        for (int x = 0; x < w; x++) {
            const float SY = WY1 + y * DY, SX = WX1 + x * DX;
            const v3 eye = { 0, 0.25f, -7.0f };
            v3 acc = { 0, 0, 0 };
            /* FIFO => breadth-first.  The reference keeps a 64-slot ring whose cursors persist across
             * sub-samples (RNO:30-40, 309-312); the queue drains completely per sub-sample, so a
             * linear queue restarted per sub-sample pops in the same order. 63 = 1+2+..+32 nodes. */
            wray fifo[64];
            for (int tx = -1; tx < 2; tx++)
                for (int ty = -1; ty < 2; ty++) {
                    int head = 0, tail = 0;
                    wray pr;
                    pr.d.x = SX + DX * (tx / 2.0f) - eye.x;
                    pr.d.y = SY + DY * (ty / 2.0f) - eye.y;
                    pr.d.z = 0 - eye.z;
                    float inv = 1.0f / sqrtf(pr.d.x * pr.d.x + pr.d.y * pr.d.y + pr.d.z * pr.d.z);
                    pr.d.x *= inv; pr.d.y *= inv; pr.d.z *= inv;
                    pr.o = eye; pr.weight = 1.0f; pr.depth = 0; pr.from_prim = -1; pr.kind = RAY_PRIMARY;
                    pr.r_index = 1.0f; pr.transp.x = pr.transp.y = pr.transp.z = 1;
                    fifo[tail++] = pr;
                    while (head < tail) {
                        if (ctr && (uint32_t)(tail - head) > ctr->queue_high_water) ctr->queue_high_water = tail - head;
                        const wray cur = fifo[head++];
                        /* `at` stays (0,0,0) when the ray hits a light: the reference leaves point_intersect
                         * uninitialised there (RNO:197-200) yet spawns children from it if the light's material
                         * reflects or refracts (RNO:370-431).  Defined here as the origin; no shipped scene has
                         * such a light. */
                        v3 col = { 0, 0, 0 }, at = { 0, 0, 0 };
                        float dist; int kind = 0;
                        int hit = trace_one(&cur, prims, n, &col, &dist, &at, &kind, ctr);
                        if (cur.kind == RAY_PRIMARY) {   /* RNO:351-368 */
                            if (hit_ids) hit_ids[((size_t)y * w + x) * 9 + (size_t)((tx + 1) * 3 + (ty + 1))] = hit;
                            acc.x += col.x * cur.weight; acc.y += col.y * cur.weight; acc.z += col.z * cur.weight;
                        } else if (cur.kind == RAY_REFLECTED) {
                            const ow_prim *fp = &prims[cur.from_prim];
                            acc.x += col.x * cur.weight * fp->color.x * cur.transp.x;
                            acc.y += col.y * cur.weight * fp->color.y * cur.transp.y;
                            acc.z += col.z * cur.weight * fp->color.z * cur.transp.z;
                        } else {
                            acc.x += col.x * cur.weight * cur.transp.x;
                            acc.y += col.y * cur.weight * cur.transp.y;
                            acc.z += col.z * cur.weight * cur.transp.z;
                        }
                        if (hit < 0) continue;          /* defined here: a miss spawns nothing */
                        if (!(cur.depth < TRACE_DEPTH)) continue;
                        const ow_prim *hp = &prims[hit];
                        if (hp->refl > 0.0f) {          /* RNO:374-394 */
                            v3 N = surface_normal(hp, at);
                            float dn = dot3(cur.d, N);
                            wray nr;
                            nr.d.x = cur.d.x - 2.0f * dn * N.x;
                            nr.d.y = cur.d.y - 2.0f * dn * N.y;
                            nr.d.z = cur.d.z - 2.0f * dn * N.z;
                            nr.o.x = at.x + nr.d.x * RAY_EPS; nr.o.y = at.y + nr.d.y * RAY_EPS; nr.o.z = at.z + nr.d.z * RAY_EPS;
                            nr.depth = cur.depth + 1; nr.weight = hp->refl * cur.weight; nr.kind = RAY_REFLECTED;
                            nr.from_prim = hit; nr.r_index = cur.r_index; nr.transp = cur.transp;
                            fifo[tail++] = nr;
                        }
                        if (hp->refr > 0.0f) {          /* RNO:397-431 */
                            float into = hp->refr_index;
                            float ratio = cur.r_index / into;
                            v3 g = surface_normal(hp, at);
                            v3 N = { g.x * (float)kind, g.y * (float)kind, g.z * (float)kind };
                            float cosI = -dot3(N, cur.d);
                            float cosT2 = 1.0f - ratio * ratio * (1.0f - cosI * cosI);
                            if (cosT2 > 0.0f) {
                                wray nr;
                                nr.d.x = (ratio * cur.d.x) + (ratio * cosI - sqrtf(cosT2)) * N.x;
                                nr.d.y = (ratio * cur.d.y) + (ratio * cosI - sqrtf(cosT2)) * N.y;
                                nr.d.z = (ratio * cur.d.z) + (ratio * cosI - sqrtf(cosT2)) * N.z;
                                nr.o.x = at.x + nr.d.x * RAY_EPS; nr.o.y = at.y + nr.d.y * RAY_EPS; nr.o.z = at.z + nr.d.z * RAY_EPS;
                                nr.depth = cur.depth + 1; nr.weight = cur.weight; nr.kind = RAY_REFRACTED;
                                nr.from_prim = hit; nr.r_index = into;
                                nr.transp.x = cur.transp.x * expf(hp->color.x * 0.15f * (-dist));   /* Beer's law */
                                nr.transp.y = cur.transp.y * expf(hp->color.y * 0.15f * (-dist));
                                nr.transp.z = cur.transp.z * expf(hp->color.z * 0.15f * (-dist));
                                fifo[tail++] = nr;
                            }
                        }
                    }
                }
            int r = (int)(acc.x * (256 / 9)), g = (int)(acc.y * (256 / 9)), b = (int)(acc.z * (256 / 9));   /* RNO:436-447 */
            if (r > 255) r = 255;
            if (g > 255) g = 255;
            if (b > 255) b = 255;
            uint8_t *px = pixels + ((size_t)y * w + x) * 4;
            px[0] = (uint8_t)r; px[1] = (uint8_t)g; px[2] = (uint8_t)b; px[3] = 0;
        }
}

typedef struct {
    uint8_t *pixels; int32_t *hit_ids; int w, h, bands; const ow_prim *prims; int n;
    int *next_band; ow_counters ctr;
} wjob;

/* Each host thread pulls row bands off a shared counter until none are left. */
static void *wjob_run(void *p) {
    wjob *j = (wjob *)p;
    for (;;) {
        int b = __sync_fetch_and_add(j->next_band, 1);
        if (b >= j->bands) break;
        int y0 = (int)((long)j->h * b / j->bands), y1 = (int)((long)j->h * (b + 1) / j->bands);
        oracle_whitted_rows(j->pixels, j->hit_ids, j->w, j->h, y0, y1, j->prims, j->n, &j->ctr);
    }
    return 0;
}

void oracle_whitted_render(uint8_t *pixels, int32_t *hit_ids, int w, int h,
                           const ow_prim *prims, int n, int threads, ow_counters *ctr) {
    if (threads < 1) threads = 1;
    int bands = threads * 16; if (bands > h) bands = h; if (bands < 1) bands = 1;
    int next = 0;
    wjob *jobs = (wjob *)calloc(threads, sizeof(wjob));
    pthread_t *t = (pthread_t *)malloc(sizeof(pthread_t) * threads);
    for (int k = 0; k < threads; k++) {
        jobs[k].pixels = pixels; jobs[k].hit_ids = hit_ids; jobs[k].w = w; jobs[k].h = h; jobs[k].bands = bands;
        jobs[k].prims = prims; jobs[k].n = n; jobs[k].next_band = &next;
        pthread_create(&t[k], 0, wjob_run, &jobs[k]);
    }
    for (int k = 0; k < threads; k++) pthread_join(t[k], 0);
    if (ctr)
        for (int k = 0; k < threads; k++) {
            ctr->traced_rays += jobs[k].ctr.traced_rays; ctr->shadow_rays += jobs[k].ctr.shadow_rays;
            ctr->sphere_tests += jobs[k].ctr.sphere_tests; ctr->plane_tests += jobs[k].ctr.plane_tests;
            if (jobs[k].ctr.queue_high_water > ctr->queue_high_water) ctr->queue_high_water = jobs[k].ctr.queue_high_water;
        }
    free(t); free(jobs);
}

/* The scene table of create_scene() with CHOOSE_SCENE 0 (R323/scene.c:48-96) as flat Primitive_2 records (the copy of
 * R323/raytracer.c:721-746): n_primitives is 17 although only 16 slots are filled (scene.c:55-57 memsets the list), so slot 16
 * is an all-zero plane.  Used by bench.py's --impl reference arm when oracle/_ref (the reference's own scene.c) was not shipped,
 * so that the arm never touches the product library.  Row: type, r, g, b, refl, refr, refr_index, diff, spec, is_light, x, y, z, depth|radius. */
int oracle_whitted_scene0(ow_prim *out, int cap) {
    static const float rows[16][14] = {
        { 0, 0.6f, 0.6f, 0.6f, 0.0f, 0.0f, 0.0f, 0.4f, 1.8f, 0, 0.0f, 0.75f, 0.0f, 4.4f },
        { 1, 0.08f, 0.08f, 0.08f, 0.2f, 1.0f, 1.4f, 0.0f, 0.0f, 0, 3.4f, -3.4f, 23.0f, 2.5f },
        { 1, 0.07f, 0.17f, 0.07f, 0.1f, 1.0f, 1.2f, 0.0f, 0.0f, 0, -0.7f, -4.90f, 27.0f, 1.0f },
        { 1, 1.0f, 1.0f, 1.0f, 0.8f, 0.0f, 0.0f, 0.0f, 0.0f, 0, -3.4f, -3.4f, 29.0f, 2.5f },
        { 1, 1.5f, 0.7f, 0.7f, 0.1f, 0.0f, 0.0f, 0.2f, 0.2f, 0, 0.5f, -4.1f, 29.0f, 1.5f },
        { 1, 0.7f, 0.7f, 1.7f, 0.2f, 0.0f, 0.0f, 0.2f, 0.2f, 0, -6.0f, -4.1f, 32.0f, 1.5f },
        { 1, 0.07f, 0.17f, 0.07f, 0.3f, 1.0f, 1.2f, 0.2f, 0.8f, 0, -6.7f, -4.90f, 29.0f, 1.0f },
        { 1, 0.08f, 0.08f, 0.08f, 0.7f, 1.0f, 1.3f, 0.8f, 0.0f, 0, 6.4f, -4.9f, 18.0f, 1.0f },
        { 0, 1.0f, 0.6f, 0.6f, 0.0f, 0.0f, 0.0f, 0.8f, 1.5f, 0, 0.7f, 0.0f, 0.0f, 5.4f },
        { 0, 0.7f, 0.6f, 1.0f, 0.0f, 0.0f, 0.0f, 0.8f, 0.8f, 0, -0.7f, 0.0f, 0.0f, 5.4f },
        { 0, 1.0f, 1.0f, 1.0f, 0.0f, 0.0f, 0.0f, 1.2f, 0.8f, 0, 0.0f, -0.8f, 0.0f, 5.4f },
        { 0, 1.5f, 1.5f, 1.5f, 0.0f, 0.0f, 0.0f, 1.2f, 0.8f, 0, 0.0f, 0.0f, -0.14f, 5.4f },
        { 0, 0.1f, 0.1f, 0.1f, 0.0f, 0.0f, 0.0f, 1.0f, 1.0f, 0, 0.0f, 0.0f, 0.72f, 5.4f },
        { 1, 0.85f, 0.85f, 0.85f, 0.0f, 0.0f, 0.0f, 0.0f, 1.8f, 1, 0.0f, 6.5f, 22.0f, 0.35f },
        { 1, 0.85f, 0.85f, 0.85f, 0.0f, 0.0f, 0.0f, 0.0f, 1.8f, 1, -3.0f, 6.5f, 22.0f, 0.35f },
        { 1, 0.85f, 0.85f, 0.85f, 0.0f, 0.0f, 0.0f, 0.0f, 1.8f, 1, 3.0f, 6.5f, 22.0f, 0.35f },
    };
    if (!out || cap < 17) return -1;
    memset(out, 0, sizeof(ow_prim) * 17);
    for (int i = 0; i < 16; i++) {
        const float *r = rows[i];
        ow_prim *p = &out[i];
        p->type = (int32_t)r[0];
        p->color.x = r[1]; p->color.y = r[2]; p->color.z = r[3];
        p->refl = r[4]; p->refr = r[5]; p->refr_index = r[6]; p->diff = r[7]; p->spec = r[8];
        p->is_light = (uint8_t)r[9];
        if (p->type == 1) {        /* create_sphere, scene.c:34-46 */
            p->center.x = r[10]; p->center.y = r[11]; p->center.z = r[12];
            p->radius = r[13]; p->sq_radius = r[13] * r[13]; p->r_radius = 1.0f / r[13];
        } else {                   /* create_plane, scene.c:20-32 */
            p->normal.x = r[10]; p->normal.y = r[11]; p->normal.z = r[12]; p->depth = r[13];
        }
    }
    return 17;
}
