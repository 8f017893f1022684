/* oracle/oracle_smallpt.c -- CPU restatement of the reference's smallpt path-trace loop.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Restates, in this repository's own words, the CPU
 * twin of the smallptGPU kernel: smallptgpu-v1.6/smallptCPU.cpp:84-124 (pixel loop, "SCPU:"),
 * geomfunc.h (intersection + integrators, "GF:"), simplernd.h:34-48 (RNG), vec.h (macros,
 * "VEC:") and displayfunc.cpp:182-195 (UpdateCamera), with the float-overload numerics of the
 * reference's C++ (/TP) build.  Adds query/test counters, row ranges, multi-pass rendering in one
 * call and a driver for the direct-lighting integrator (the reference has none on the CPU).
 * Pinned against oracle/_ref by tests/test_oracle_smallpt.py (colors, pixels, RNG state bit-exact).
 *
 * Behaviours kept on purpose: the "is it an emitter" test looks at e.x and e.z
 * only (VEC:44); sign(0) is -1 (VEC:59); the two GetRandom() arguments of the light-sample call
 * (GF:131) are evaluated right-to-left, as g++ does on x86-64, so u2 is drawn BEFORE u1.
 */
#include "oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>

#define HIT_EPS 0.01f                      /* geom.h:29 */
#define PI_F 3.14159265358979323846f       /* geom.h:30 */

float oracle_pt_get_random(uint32_t *s0, uint32_t *s1) {     /* simplernd.h:34-48 */
    *s0 = 36969u * (*s0 & 65535u) + (*s0 >> 16);
    *s1 = 18000u * (*s1 & 65535u) + (*s1 >> 16);
    union { uint32_t u; float f; } cvt;
    cvt.u = (((*s0) << 16) + (*s1)) & 0x007fffffu;
    cvt.u |= 0x40000000u;
    return (cvt.f - 2.f) / 2.f;
}
#define RND(s) oracle_pt_get_random(&(s)[0], &(s)[1])

static inline float dot(op_vec a, op_vec b) { return a.x * b.x + a.y * b.y + a.z * b.z; }      /* VEC:40 */
static inline op_vec unit(op_vec v) {                                                           /* VEC:41 */
    float l = 1.f / sqrtf(dot(v, v));
    op_vec r = { l * v.x, l * v.y, l * v.z };
    return r;
}
static inline op_vec cross(op_vec a, op_vec b) {                                                /* VEC:42 */
    op_vec r = { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x };
    return r;
}
static inline int emits(const op_sphere *s) { return !(s->e.x == 0.f && s->e.z == 0.f); }       /* VEC:44 */

static inline float hit_distance(const op_sphere *s, op_vec o, op_vec d) {                      /* GF:32-59 */
    op_vec op = { s->p.x - o.x, s->p.y - o.y, s->p.z - o.z };
    float b = dot(op, d);
    float disc = b * b - dot(op, op) + s->rad * s->rad;
    if (disc < 0.f) return 0.f;
    disc = sqrtf(disc);
    float t = b - disc;
    if (t > HIT_EPS) return t;
    t = b + disc;
    return t > HIT_EPS ? t : 0.f;
}

float oracle_pt_sphere_intersect(const op_sphere *s, const float *o3, const float *d3) {
    op_vec o = { o3[0], o3[1], o3[2] }, d = { d3[0], d3[1], d3[2] };
    return hit_distance(s, o, d);
}

/* GF:71-92: descending index, strict '<' => on an exact tie the higher index is kept. */
static int nearest_hit(const op_sphere *sph, uint32_t n, op_vec o, op_vec d, float *t, uint32_t *id, op_counters *c) {
    const float inf = 1e20f;
    *t = inf;
    if (c) { c->nearest_queries++; c->sphere_tests += n; }
    for (uint32_t i = n; i--;) {
        float k = hit_distance(&sph[i], o, d);
        if (k != 0.f && k < *t) { *t = k; *id = i; }
    }
    return *t < inf;
}

/* GF:94-110: any hit closer than maxt, descending index, stops at the first one. */
static int occluded(const op_sphere *sph, uint32_t n, op_vec o, op_vec d, float maxt, op_counters *c) {
    if (c) c->shadow_queries++;
    for (uint32_t i = n; i--;) {
        if (c) c->sphere_tests++;
        float k = hit_distance(&sph[i], o, d);
        if (k != 0.f && k < maxt) return 1;
    }
    return 0;
}

/* GF:112-165: next-event estimation towards every emitter, ascending index. */
static op_vec direct_light(const op_sphere *sph, uint32_t n, uint32_t *seed, op_vec at, op_vec nl, op_counters *c) {
    op_vec sum = { 0.f, 0.f, 0.f };
    for (uint32_t i = 0; i < n; i++) {
        const op_sphere *lt = &sph[i];
        if (!emits(lt)) continue;
        /* GF:131 with g++'s right-to-left argument evaluation: second argument drawn first. */
        const float u2 = RND(seed);
        const float u1 = RND(seed);
        const float zz = 1.f - 2.f * u1;                       /* GF:61-69 */
        const float inside = 1.f - zz * zz;
        const float rr = sqrtf(0.f > inside ? 0.f : inside);
        const float phi = 2.f * PI_F * u2;
        op_vec on_unit = { rr * cosf(phi), rr * sinf(phi), zz };
        op_vec on_light = { lt->rad * on_unit.x + lt->p.x, lt->rad * on_unit.y + lt->p.y, lt->rad * on_unit.z + lt->p.z };
        op_vec sd = { on_light.x - at.x, on_light.y - at.y, on_light.z - at.z };
        const float len = sqrtf(dot(sd, sd));
        const float inv = 1.f / len;
        sd.x = inv * sd.x; sd.y = inv * sd.y; sd.z = inv * sd.z;
        float wo = dot(sd, on_unit);
        if (wo > 0.f) continue;                                /* far half of the light */
        wo = -wo;
        const float wi = dot(sd, nl);
        if (wi > 0.f && !occluded(sph, n, at, sd, len - HIT_EPS, c)) {
            const float s = (4.f * PI_F * lt->rad * lt->rad) * wi * wo / (len * len);
            sum.x = sum.x + s * lt->e.x; sum.y = sum.y + s * lt->e.y; sum.z = sum.z + s * lt->e.z;
        }
    }
    return sum;
}

/* GF:167-338 (direct_only == 0) and GF:340-483 (direct_only == 1). */
static op_vec radiance(int direct_only, const op_sphere *sph, uint32_t n, op_vec o, op_vec d, uint32_t *seed, op_counters *c) {
    op_vec rad = { 0.f, 0.f, 0.f }, thr = { 1.f, 1.f, 1.f };
    int after_specular = 1;
    for (unsigned depth = 0; ; ++depth) {
        if (depth > 6) return rad;
        float t; uint32_t id = 0;
        if (!nearest_hit(sph, n, o, d, &t, &id, c)) return rad;
        const op_sphere *obj = &sph[id];
        op_vec at = { o.x + t * d.x, o.y + t * d.y, o.z + t * d.z };
        op_vec nrm = { at.x - obj->p.x, at.y - obj->p.y, at.z - obj->p.z };
        nrm = unit(nrm);
        const float dp = dot(nrm, d);
        const float flip = -1.f * (dp > 0 ? 1 : -1);
        op_vec nl = { flip * nrm.x, flip * nrm.y, flip * nrm.z };
        if (emits(obj)) {
            if (after_specular) {
                const float a = fabsf(dp);
                rad.x = rad.x + thr.x * (a * obj->e.x);
                rad.y = rad.y + thr.y * (a * obj->e.y);
                rad.z = rad.z + thr.z * (a * obj->e.z);
            }
            return rad;
        }
        if (obj->refl == 0) {                                  /* diffuse, GF:228-276 */
            after_specular = 0;
            thr.x = thr.x * obj->c.x; thr.y = thr.y * obj->c.y; thr.z = thr.z * obj->c.z;
            op_vec ld = direct_light(sph, n, seed, at, nl, c);
            rad.x = rad.x + thr.x * ld.x; rad.y = rad.y + thr.y * ld.y; rad.z = rad.z + thr.z * ld.z;
            if (direct_only) return rad;
            const float r1 = 2.f * PI_F * RND(seed);
            const float r2 = RND(seed);
            const float r2s = sqrtf(r2);
            op_vec w = nl, a;
            if (fabsf(w.x) > .1f) { a.x = 0.f; a.y = 1.f; a.z = 0.f; } else { a.x = 1.f; a.y = 0.f; a.z = 0.f; }
            op_vec u = unit(cross(a, w));
            op_vec v = cross(w, u);
            const float ku = cosf(r1) * r2s, kv = sinf(r1) * r2s, kw = sqrtf(1 - r2);
            op_vec nd = { (ku * u.x + kv * v.x) + kw * w.x, (ku * u.y + kv * v.y) + kw * w.y, (ku * u.z + kv * v.z) + kw * w.z };
            o = at; d = nd;
            continue;
        }
        after_specular = 1;
        const float two_dn = 2.f * dot(nrm, d);
        op_vec mirror = { d.x - two_dn * nrm.x, d.y - two_dn * nrm.y, d.z - two_dn * nrm.z };
        if (obj->refl == 1) {                                  /* mirror, GF:277-288 */
            thr.x = thr.x * obj->c.x; thr.y = thr.y * obj->c.y; thr.z = thr.z * obj->c.z;
            o = at; d = mirror;
            continue;
        }
        /* glass, GF:289-336 */
        const int into = dot(nrm, nl) > 0;
        const float nc = 1.f, nt = 1.5f;
        const float nnt = into ? nc / nt : nt / nc;
        const float ddn = dot(d, nl);
        const float cos2t = 1.f - nnt * nnt * (1.f - ddn * ddn);
        if (cos2t < 0.f) {                                     /* total internal reflection */
            thr.x = thr.x * obj->c.x; thr.y = thr.y * obj->c.y; thr.z = thr.z * obj->c.z;
            o = at; d = mirror;
            continue;
        }
        const float kk = (into ? 1 : -1) * (ddn * nnt + sqrtf(cos2t));
        op_vec td = { nnt * d.x - kk * nrm.x, nnt * d.y - kk * nrm.y, nnt * d.z - kk * nrm.z };
        td = unit(td);
        const float ea = nt - nc, eb = nt + nc;
        const float R0 = ea * ea / (eb * eb);
        const float cc = 1 - (into ? -ddn : dot(td, nrm));
        const float Re = R0 + (1 - R0) * cc * cc * cc * cc * cc;
        const float Tr = 1.f - Re;
        const float P = .25f + .5f * Re;
        const float RP = Re / P, TP = Tr / (1.f - P);
        if (RND(seed) < P) {
            thr.x = (RP * thr.x) * obj->c.x; thr.y = (RP * thr.y) * obj->c.y; thr.z = (RP * thr.z) * obj->c.z;
            o = at; d = mirror;
        } else {
            thr.x = (TP * thr.x) * obj->c.x; thr.y = (TP * thr.y) * obj->c.y; thr.z = (TP * thr.z) * obj->c.z;
            o = at; d = td;
        }
    }
}

void oracle_pt_update_camera(op_camera *cam, int w, int h) {   /* displayfunc.cpp:182-195 */
    op_vec dir = { cam->target.x - cam->orig.x, cam->target.y - cam->orig.y, cam->target.z - cam->orig.z };
    cam->dir = unit(dir);
    const op_vec up = { 0.f, 1.f, 0.f };
    const float fov = (M_PI / 180.f) * 45.f;      /* double product, rounded to float once */
    op_vec cx = unit(cross(cam->dir, up));
    const float kx = w * fov / h;
    cam->x.x = kx * cx.x; cam->x.y = kx * cx.y; cam->x.z = kx * cx.z;
    op_vec cy = unit(cross(cam->x, cam->dir));
    cam->y.x = fov * cy.x; cam->y.y = fov * cy.y; cam->y.z = fov * cy.z;
}

static inline int to_byte(float v) {                            /* VEC:47, 62 */
    float cl = v < 0.f ? 0.f : (v > 1.f ? 1.f : v);
    return (int)(powf(cl, 1.f / 2.2f) * 255.f + .5f);
}

void oracle_pt_rows(int integrator, const op_sphere *sph, uint32_t n, const op_camera *cam,
                    int w, int h, int y0, int y1, int pass0, int n_passes,
                    float *colors, uint32_t *seeds, uint32_t *pixels, op_counters *ctr) {
    const float inv_w = 1.f / w, inv_h = 1.f / h;
    for (int y = y0; y < y1; y++)
        for (int x = 0; x < w; x++) {
            const size_t i = (size_t)(h - y - 1) * w + x;        /* SCPU:86: flipped index */
            uint32_t *seed = &seeds[2 * i];
            op_vec c = { colors[3 * i], colors[3 * i + 1], colors[3 * i + 2] };
            for (int s = pass0; s < pass0 + n_passes; s++) {
                const float r1 = RND(seed) - .5f;
                const float r2 = RND(seed) - .5f;
                const float kcx = (x + r1) * inv_w - .5f;
                const float kcy = (y + r2) * inv_h - .5f;
                op_vec rd = { cam->x.x * kcx + cam->y.x * kcy + cam->dir.x,
                              cam->x.y * kcx + cam->y.y * kcy + cam->dir.y,
                              cam->x.z * kcx + cam->y.z * kcy + cam->dir.z };
                op_vec ro = { 0.1f * rd.x + cam->orig.x, 0.1f * rd.y + cam->orig.y, 0.1f * rd.z + cam->orig.z };
                rd = unit(rd);
                if (ctr) ctr->samples++;
                op_vec r = radiance(integrator, sph, n, ro, rd, seed, ctr);
                if (s == 0) c = r;
                else {                                          /* SCPU:110-118 */
                    const float k1 = s, k2 = 1.f / (k1 + 1.f);
                    c.x = (c.x * k1 + r.x) * k2; c.y = (c.y * k1 + r.y) * k2; c.z = (c.z * k1 + r.z) * k2;
                }
            }
            colors[3 * i] = c.x; colors[3 * i + 1] = c.y; colors[3 * i + 2] = c.z;
            if (pixels) pixels[(size_t)y * w + x] = to_byte(c.x) | (to_byte(c.y) << 8) | (to_byte(c.z) << 16);
        }
}

typedef struct {
    int integrator; const op_sphere *sph; uint32_t n; const op_camera *cam; int w, h, bands, pass0, n_passes;
    float *colors; uint32_t *seeds, *pixels; int *next_band; op_counters ctr;
} pjob;

static void *pjob_run(void *pv) {
    pjob *j = (pjob *)pv;
    for (;;) {
        int b = __sync_fetch_and_add(j->next_band, 1);
        if (b >= j->bands) break;
        int y0 = (int)((long)j->h * b / j->bands), y1 = (int)((long)j->h * (b + 1) / j->bands);
        oracle_pt_rows(j->integrator, j->sph, j->n, j->cam, j->w, j->h, y0, y1, j->pass0, j->n_passes,
                       j->colors, j->seeds, j->pixels, &j->ctr);
    }
    return 0;
}

void oracle_pt_render(int integrator, const op_sphere *sph, uint32_t n, const op_camera *cam,
                      int w, int h, int pass0, int n_passes,
                      float *colors, uint32_t *seeds, uint32_t *pixels, int threads, op_counters *ctr) {
    if (threads < 1) threads = 1;
    int bands = threads * 16; if (bands > h) bands = h; if (bands < 1) bands = 1;
    int next = 0;
    pjob *jobs = (pjob *)calloc(threads, sizeof(pjob));
    pthread_t *t = (pthread_t *)malloc(sizeof(pthread_t) * threads);
    for (int k = 0; k < threads; k++) {
        pjob *j = &jobs[k];
        j->integrator = integrator; j->sph = sph; j->n = n; j->cam = cam; j->w = w; j->h = h; j->bands = bands;
        j->pass0 = pass0; j->n_passes = n_passes; j->colors = colors; j->seeds = seeds; j->pixels = pixels;
        j->next_band = &next;
        pthread_create(&t[k], 0, pjob_run, j);
    }
    for (int k = 0; k < threads; k++) pthread_join(t[k], 0);
    if (ctr)
        for (int k = 0; k < threads; k++) {
            ctr->samples += jobs[k].ctr.samples; ctr->nearest_queries += jobs[k].ctr.nearest_queries;
            ctr->shadow_queries += jobs[k].ctr.shadow_queries; ctr->sphere_tests += jobs[k].ctr.sphere_tests;
        }
    free(t); free(jobs);
}

/* Host libm taps for tests/test_math_parity.py: the float functions the reference's C++ build binds to. */
void oracle_libm_sincosf(const float *in, float *sin_out, float *cos_out, long n) {
    for (long i = 0; i < n; i++) { sin_out[i] = sinf(in[i]); cos_out[i] = cosf(in[i]); }
}
void oracle_libm_expf(const float *in, float *out, long n) { for (long i = 0; i < n; i++) out[i] = expf(in[i]); }
void oracle_libm_to_int_gamma(const float *in, int *out, long n) { for (long i = 0; i < n; i++) out[i] = to_byte(in[i]); }
double oracle_libm_pow20(float v) { return pow((double)v, 20.0); }
