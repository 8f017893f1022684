// pt_lane.cuh -- per-lane state machine of the smallpt path tracer (host/device).
//
// One lane owns one pixel for all of its sample passes (the per-pixel RNG stream and the running
// mean are sequential, SPT/smallptCPU.cpp:84-124), and walks the reference's control flow
// (SPT/geomfunc.h:167-338 RadiancePathTracing, :340-483 RadianceDirectLighting, :112-165
// SampleLights) as a sequence of RAY QUERIES: a nearest-hit query (Intersect, :71-92) or an any-hit
// shadow query (IntersectP, :94-110).  Between two queries pt_advance() runs the shading code.
// The kernel keeps 32 such lanes per warp in one loop whose body is "one query, then advance", so
// lanes that are at different bounces, passes or pixels still execute the sphere loop together.
//
// Expression order follows SURVEY.md 9.1 / the macros of SPT/vec.h; see rt_math.cuh for the rules.
#pragma once
#include "rt_math.cuh"

namespace rtb {

struct alignas(16) f4 { float x, y, z, w; };
struct alignas(8) f2 { float x, y; };

enum { PH_IDLE = 0, PH_NEAREST = 1, PH_SHADOW = 2 };

#define PT_EPS 0.01f                              /* SPT/geom.h:29 */
#define PT_PI 3.14159265358979323846f             /* SPT/geom.h:30 */
#define PT_INF 1e20f                              /* SPT/geomfunc.h:79 */

// Scene in structure-of-arrays form (built once per rt_pt_set_scene):
//   geom[i] = (p.x, p.y, p.z, rad*rad)   -- everything the sphere test reads (16 B, one LDS.128)
//   emis[i] = (e.x, e.y, e.z, bits(refl))
//   colr[i] = (c.x, c.y, c.z, rad)
//   lights[k] = indices i with !viszero(e), ascending (SPT/geomfunc.h:128-131; viszero tests x,x,z: SPT/vec.h:44)
struct PtFrame {
    const f4 *emis, *colr;
    const f4 *geom_global;          // full geometry array in global memory (the kernel stages it to smem)
    const int *lights;
    int n, n_lights;
    float cam_ox, cam_oy, cam_oz, cam_dx, cam_dy, cam_dz, cam_xx, cam_xy, cam_xz, cam_yx, cam_yy, cam_yz;
    int w, h;
    float inv_w, inv_h;
    int pass0, n_passes;
    int direct_only;                // 0: RadiancePathTracing, 1: RadianceDirectLighting
    const f2 *sincos_tab;           // NULL, or 2^23 (sin, cos) pairs: every angle 2*pi*GetRandom() can be (rt_math.cuh)
    int sum_mode;                   // 1: accumulate sums instead of the running mean (sample-sharded mode)
    int defer_pack;                 // 1: the render kernel leaves `pixels` alone; pt_pack_kernel converts the colours afterwards with all lanes busy
                                    //  (inside the render kernel the FP64 gamma ran for the 2 lanes of a warp whose pixel had just finished)
};

struct PtLane {
    int x, y, pass;
    uint32_t s0, s1;
    float cr, cg, cb;               // colors[i]
    float ox, oy, oz, dx, dy, dz;   // ray of the query in flight
    float tr, tg, tb;               // throughput
    float rr, rg, rb;               // radiance of this sample so far
    float nlx, nly, nlz;            // oriented normal at the current diffuse hit
    float lr, lg, lb;               // direct light gathered at the current diffuse hit
    float lscale;                   // geometric term of the light sample whose shadow query is in flight
    float cumu;                     // nearest distance so far / shadow max-t
    int hit;                        // sphere index of the accepted hit, -1 none (shadow query: >= 0 means occluded)
    int li;                         // cursor into lights[]
    int depth, after_spec, phase;
    // work counters (only maintained by counting builds)
    uint32_t c_nearest, c_shadow, c_samples;
    uint64_t c_tests;
};

// One sphere against the lane's query: SphereIntersect (SPT/geomfunc.h:32-59) on the staged (p, rad^2)
// record, followed by the acceptance test shared by Intersect (:80-88) and IntersectP (:102-109):
// `d != 0 && d < limit`.  Written without per-lane branches -- every lane of the warp evaluates the same
// instruction stream and the update is a predicated select -- because the loop index is warp-uniform; the
// square root (only needed when some lane has det >= 0) sits behind a warp vote.  `live` masks lanes that
// have no query in flight.  A nearest query keeps the closest hit (scanning i = n-1 .. 0 with a strict '<'
// leaves the HIGHER index on an exact tie, as the reference does); a shadow query only needs "any hit", so
// recording the index there too (hit >= 0 means occluded) lets both kinds share the update.
template <bool COUNT>
RT_HD void pt_test(PtLane &L, const f4 g, int i, bool live) {
    if (COUNT && live && L.phase == PH_SHADOW && L.hit < 0) L.c_tests++;   // IntersectP stops at its first hit
    const float opx = f_sub(g.x, L.ox), opy = f_sub(g.y, L.oy), opz = f_sub(g.z, L.oz);
    const float b = dot3(opx, opy, opz, L.dx, L.dy, L.dz);
    const float det = f_add(f_sub(f_mul(b, b), dot3(opx, opy, opz, opx, opy, opz)), g.w);
    const bool cand = live & !(det < 0.f);
    if (warp_any(cand)) {
        const float dv[1] = { det };
        const bool need[1] = { cand };
        float sqv[1];
        sqrt_group<1>(dv, need, sqv);
        const float sq = sqv[0];
        const float t1 = f_sub(b, sq), t2 = f_add(b, sq);
        const float t = t1 > PT_EPS ? t1 : t2;                 // first root if beyond EPSILON, else the second
        if (cand & (t > PT_EPS) & (t < L.cumu)) { L.cumu = t; L.hit = i; }
    }
}

// Four consecutive spheres (indices i, i-1, i-2, i-3; g points at sphere i-3) in one go: the four
// discriminants are independent instruction chains (ILP), one warp vote covers the four square roots, and
// the four acceptance tests are applied in the reference's descending index order.
template <bool COUNT>
RT_HD void pt_test4(PtLane &L, const f4 *g, int i, bool live) {
    float b[4], det[4];
    bool cand[4];
    bool any = false;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const f4 s = g[3 - k];
        const float opx = f_sub(s.x, L.ox), opy = f_sub(s.y, L.oy), opz = f_sub(s.z, L.oz);
        b[k] = dot3(opx, opy, opz, L.dx, L.dy, L.dz);
        det[k] = f_add(f_sub(f_mul(b[k], b[k]), dot3(opx, opy, opz, opx, opy, opz)), s.w);
        cand[k] = live & !(det[k] < 0.f);
        any = any | cand[k];
    }
    if (warp_any(any)) {
        float t[4], sq[4];
        sqrt_group<4>(det, cand, sq);              // a negative discriminant of a lane that is not a candidate must not cost the slow path
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float t1 = f_sub(b[k], sq[k]), t2 = f_add(b[k], sq[k]);
            t[k] = t1 > PT_EPS ? t1 : t2;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (COUNT && live && L.phase == PH_SHADOW && L.hit < 0) L.c_tests++;
            if (cand[k] & (t[k] > PT_EPS) & (t[k] < L.cumu)) { L.cumu = t[k]; L.hit = i - k; }
        }
    } else if (COUNT && live && L.phase == PH_SHADOW && L.hit < 0) L.c_tests += 4;
}

// The sphere loop of one query round over spheres [lo, hi), descending (geom[0] is sphere `lo`): groups of
// four, then the ragged rest.  Every 16 spheres a warp vote ends the loop once no lane can change any more
// (all idle, or all shadow queries already occluded) -- the any-hit early-out of IntersectP at warp level.
template <bool COUNT>
RT_HD void pt_query_range(PtLane &L, const f4 *geom, int lo, int hi, bool active) {
    int i = hi - 1;
    while (i >= lo) {
        const bool live = active && !(L.phase == PH_SHADOW && L.hit >= 0);
        if (!warp_any(live)) return;
        int stop = i - 15 > lo ? i - 15 : lo;           // this vote covers spheres i .. stop
        for (; i - 3 >= stop; i -= 4) pt_test4<COUNT>(L, geom + (i - 3 - lo), i, live);
        for (; i >= stop; --i) pt_test<COUNT>(L, geom[i - lo], i, live);
    }
}

RT_HD void pt_unit(float &x, float &y, float &z) {     // vnorm, SPT/vec.h:41
    const float l = f_rcp(f_sqrt(dot3(x, y, z, x, y, z)));
    x = f_mul(l, x); y = f_mul(l, y); z = f_mul(l, z);
}

// Camera ray of SPT/smallptCPU.cpp:89-105 (= GenerateCameraRay, SPT/rendering_kernel.cl:29-51).
RT_HD void pt_start_sample(PtLane &L, const PtFrame &F) {
    const float r1 = f_sub(get_random(L.s0, L.s1), .5f);
    const float r2 = f_sub(get_random(L.s0, L.s1), .5f);
    const float kcx = f_sub(f_mul(f_add((float)L.x, r1), F.inv_w), .5f);
    const float kcy = f_sub(f_mul(f_add((float)L.y, r2), F.inv_h), .5f);
    float dx = f_add(f_add(f_mul(F.cam_xx, kcx), f_mul(F.cam_yx, kcy)), F.cam_dx);
    float dy = f_add(f_add(f_mul(F.cam_xy, kcx), f_mul(F.cam_yy, kcy)), F.cam_dy);
    float dz = f_add(f_add(f_mul(F.cam_xz, kcx), f_mul(F.cam_yz, kcy)), F.cam_dz);
    L.ox = f_add(f_mul(0.1f, dx), F.cam_ox);
    L.oy = f_add(f_mul(0.1f, dy), F.cam_oy);
    L.oz = f_add(f_mul(0.1f, dz), F.cam_oz);
    pt_unit(dx, dy, dz);
    L.dx = dx; L.dy = dy; L.dz = dz;
    L.tr = L.tg = L.tb = 1.f;
    L.rr = L.rg = L.rb = 0.f;
    L.depth = 0; L.after_spec = 1;
    L.phase = PH_NEAREST; L.cumu = PT_INF; L.hit = -1;
}

RT_HD void pt_begin_pixel(PtLane &L, const PtFrame &F, int x, int y, const float *colors, const uint32_t *seeds) {
    const size_t i = (size_t)(F.h - y - 1) * F.w + x;        // SPT/smallptCPU.cpp:86
    RT_CHECK(x >= 0 && x < F.w && y >= 0 && y < F.h, RT_CHK_PIXEL);
    L.x = x; L.y = y; L.pass = F.pass0;
    L.s0 = seeds[2 * i]; L.s1 = seeds[2 * i + 1];
    L.cr = colors[3 * i]; L.cg = colors[3 * i + 1]; L.cb = colors[3 * i + 2];
    pt_start_sample(L, F);
}

// The shading code between two queries, cut into STEPS so that a warp can run each step with all the lanes that are at
// it (pt_kernel, ALIGNED) -- or one lane can run them back to back (pt_advance).  Either way a lane performs the same
// operations in the same order (in particular the same RNG draws), so the result does not depend on the schedule.
//   PH_NEAREST  a nearest-hit query is pending         PH_SHADOW   a shadow query is pending
//   PH_LIGHTS   SampleLights is at light L.li          PH_DIFFUSE  lights done: cosine-weighted bounce next
//   PH_BOUNCE   a new direction is set: depth check    PH_END      the sample is complete
enum { PH_LIGHTS = 4, PH_DIFFUSE = 5, PH_BOUNCE = 6, PH_END = 7 };

// After a nearest-hit query (SPT/geomfunc.h:190-288).
template <bool COUNT>
RT_HD void pt_hit(PtLane &L, const PtFrame &F) {
    if (COUNT) { L.c_nearest++; L.c_tests += (uint32_t)F.n; }
    if (!(L.cumu < PT_INF)) { L.phase = PH_END; return; }            // miss: SPT/geomfunc.h:190-193
    const int id = L.hit;
    RT_CHECK(id >= 0 && id < F.n, RT_CHK_SCENE_INDEX);
    const f4 g = F.geom_global[id];
    const f4 em = F.emis[id];
    const float t = L.cumu;
    const float ax = f_add(L.ox, f_mul(t, L.dx)), ay = f_add(L.oy, f_mul(t, L.dy)), az = f_add(L.oz, f_mul(t, L.dz));
    float nx = f_sub(ax, g.x), ny = f_sub(ay, g.y), nz = f_sub(az, g.z);
    pt_unit(nx, ny, nz);
    const float dp = dot3(nx, ny, nz, L.dx, L.dy, L.dz);
    const float flip = dp > 0.f ? -1.f : 1.f;                        // -1.f * sign(dp), sign(0) = -1 (SPT/vec.h:59)
    const float nlx = f_mul(flip, nx), nly = f_mul(flip, ny), nlz = f_mul(flip, nz);
    if (!(em.x == 0.f && em.z == 0.f)) {                             // emitter: SPT/geomfunc.h:216-227
        if (L.after_spec) {
            const float a = fabsf(dp);
            L.rr = f_add(L.rr, f_mul(L.tr, f_mul(a, em.x)));
            L.rg = f_add(L.rg, f_mul(L.tg, f_mul(a, em.y)));
            L.rb = f_add(L.rb, f_mul(L.tb, f_mul(a, em.z)));
        }
        L.phase = PH_END;
        return;
    }
    const f4 cl = F.colr[id];
    const int refl = (int)f_bits(em.w);
    if (refl == 0) {                                                  // DIFF: SPT/geomfunc.h:228-239
        L.after_spec = 0;
        L.tr = f_mul(L.tr, cl.x); L.tg = f_mul(L.tg, cl.y); L.tb = f_mul(L.tb, cl.z);
        L.ox = ax; L.oy = ay; L.oz = az;
        L.nlx = nlx; L.nly = nly; L.nlz = nlz;
        L.lr = L.lg = L.lb = 0.f;
        L.li = 0;
        L.phase = PH_LIGHTS;
        return;
    }
    L.after_spec = 1;
    const float k2 = f_mul(2.f, dot3(nx, ny, nz, L.dx, L.dy, L.dz));
    const float mx = f_sub(L.dx, f_mul(k2, nx)), my = f_sub(L.dy, f_mul(k2, ny)), mz = f_sub(L.dz, f_mul(k2, nz));
    if (refl == 1) {                                                  // SPEC: SPT/geomfunc.h:277-288
        L.tr = f_mul(L.tr, cl.x); L.tg = f_mul(L.tg, cl.y); L.tb = f_mul(L.tb, cl.z);
        L.dx = mx; L.dy = my; L.dz = mz;
    } else {                                                          // REFR: SPT/geomfunc.h:289-336
        const bool into = dot3(nx, ny, nz, nlx, nly, nlz) > 0.f;
        const float nnt = into ? 0x1.555556p-1f : 1.5f;               // nc / nt = fl(1.f / 1.5f), nt / nc = 1.5f
        const float ddn = dot3(L.dx, L.dy, L.dz, nlx, nly, nlz);
        const float cos2t = f_sub(1.f, f_mul(f_mul(nnt, nnt), f_sub(1.f, f_mul(ddn, ddn))));
        if (cos2t < 0.f) {                                            // total internal reflection
            L.tr = f_mul(L.tr, cl.x); L.tg = f_mul(L.tg, cl.y); L.tb = f_mul(L.tb, cl.z);
            L.dx = mx; L.dy = my; L.dz = mz;
        } else {
            const float kk = f_mul(into ? 1.f : -1.f, f_add(f_mul(ddn, nnt), f_sqrt(cos2t)));
            float tx = f_sub(f_mul(nnt, L.dx), f_mul(kk, nx));
            float ty = f_sub(f_mul(nnt, L.dy), f_mul(kk, ny));
            float tz = f_sub(f_mul(nnt, L.dz), f_mul(kk, nz));
            pt_unit(tx, ty, tz);
            const float R0 = 0.04f;                                   // a*a/(b*b) = fl(0.25f / 6.25f), a = nt-nc, b = nt+nc
            const float c = f_sub(1.f, into ? -ddn : dot3(tx, ty, tz, nx, ny, nz));
            const float Re = f_add(R0, f_mul(f_mul(f_mul(f_mul(f_mul(f_sub(1.f, R0), c), c), c), c), c));
            const float Tr = f_sub(1.f, Re);
            const float P = f_add(.25f, f_mul(.5f, Re));
            const float RP = f_div(Re, P), TP = f_div(Tr, f_sub(1.f, P));
            if (get_random(L.s0, L.s1) < P) {
                L.tr = f_mul(f_mul(RP, L.tr), cl.x); L.tg = f_mul(f_mul(RP, L.tg), cl.y); L.tb = f_mul(f_mul(RP, L.tb), cl.z);
                L.dx = mx; L.dy = my; L.dz = mz;
            } else {
                L.tr = f_mul(f_mul(TP, L.tr), cl.x); L.tg = f_mul(f_mul(TP, L.tg), cl.y); L.tb = f_mul(f_mul(TP, L.tb), cl.z);
                L.dx = tx; L.dy = ty; L.dz = tz;
            }
        }
    }
    L.ox = ax; L.oy = ay; L.oz = az;
    L.phase = PH_BOUNCE;
}

// After a shadow query: SPT/geomfunc.h:157-162.
template <bool COUNT>
RT_HD void pt_light_done(PtLane &L, const PtFrame &F) {
    if (COUNT) L.c_shadow++;
    if (L.hit < 0) {
        const f4 le = F.emis[F.lights[L.li]];
        L.lr = f_add(L.lr, f_mul(L.lscale, le.x));
        L.lg = f_add(L.lg, f_mul(L.lscale, le.y));
        L.lb = f_add(L.lb, f_mul(L.lscale, le.z));
    }
    L.li++;
    L.phase = PH_LIGHTS;
}

// One light of SampleLights (SPT/geomfunc.h:112-165): a shadow query (PH_SHADOW), the next light (PH_LIGHTS), or, past
// the last light, the direct light folded into the radiance (PH_DIFFUSE, or PH_END for the direct-lighting integrator).
RT_HD void pt_light_step(PtLane &L, const PtFrame &F) {
    if (L.li >= F.n_lights) {
        L.rr = f_add(L.rr, f_mul(L.tr, L.lr));
        L.rg = f_add(L.rg, f_mul(L.tg, L.lg));
        L.rb = f_add(L.rb, f_mul(L.tb, L.lb));
        L.phase = F.direct_only ? PH_END : PH_DIFFUSE;               // SPT/geomfunc.h:412-413
        return;
    }
    const int lid = F.lights[L.li];
    const f4 lg = F.geom_global[lid];
    const float lrad = F.colr[lid].w;
    // UniformSampleSphere(GetRandom(), GetRandom(), ..): the reference's compiler evaluates the
    // arguments right to left, so the SECOND argument (u2) takes the first draw.
    uint32_t u2_bits;
    const float u2 = get_random_bits(L.s0, L.s1, u2_bits);
    const float u1 = get_random(L.s0, L.s1);
    const float zz = f_sub(1.f, f_mul(2.f, u1));
    const float inside = f_sub(1.f, f_mul(zz, zz));
    const float r = f_sqrt(0.f > inside ? 0.f : inside);
    const float phi = f_mul(f_mul(2.f, PT_PI), u2);
    float sn, cs;
    RT_CHECK(u2_bits < (1u << 23), RT_CHK_TABLE);
    if (F.sincos_tab) { const f2 t = F.sincos_tab[u2_bits]; sn = t.x; cs = t.y; }     // phi == sincos_table_angle(u2_bits)
    else sincos_glibc(phi, &sn, &cs);
    const float ux = f_mul(r, cs), uy = f_mul(r, sn), uz = zz;
    const float spx = f_add(f_mul(lrad, ux), lg.x), spy = f_add(f_mul(lrad, uy), lg.y), spz = f_add(f_mul(lrad, uz), lg.z);
    float sx = f_sub(spx, L.ox), sy = f_sub(spy, L.oy), sz = f_sub(spz, L.oz);
    const float len = f_sqrt(dot3(sx, sy, sz, sx, sy, sz));
    const float inv = f_rcp(len);
    sx = f_mul(inv, sx); sy = f_mul(inv, sy); sz = f_mul(inv, sz);
    float wo = dot3(sx, sy, sz, ux, uy, uz);
    const float wi = dot3(sx, sy, sz, L.nlx, L.nly, L.nlz);
    if (wo > 0.f || !(wi > 0.f)) { L.li++; return; }                // sample on the far half of the light / light below the horizon
    wo = -wo;
    L.lscale = f_div(f_mul(f_mul(f_mul(f_mul(f_mul(4.f, PT_PI), lrad), lrad), wi), wo), f_mul(len, len));
    L.dx = sx; L.dy = sy; L.dz = sz;
    L.cumu = f_sub(len, PT_EPS);
    L.hit = -1;
    L.phase = PH_SHADOW;
}

// Cosine-weighted bounce, SPT/geomfunc.h:243-275.
RT_HD void pt_diffuse_bounce(PtLane &L, const PtFrame &F) {
    uint32_t r1_bits;
    const float r1 = f_mul(f_mul(2.f, PT_PI), get_random_bits(L.s0, L.s1, r1_bits));
    const float r2 = get_random(L.s0, L.s1);
    const float r2s = f_sqrt(r2);
    const float wx = L.nlx, wy = L.nly, wz = L.nlz;
    float ux, uy, uz;
    if (fabsf(wx) > .1f) {        // a = (0,1,0): u = a x w
        ux = f_sub(f_mul(1.f, wz), f_mul(0.f, wy)); uy = f_sub(f_mul(0.f, wx), f_mul(0.f, wz)); uz = f_sub(f_mul(0.f, wy), f_mul(1.f, wx));
    } else {                      // a = (1,0,0)
        ux = f_sub(f_mul(0.f, wz), f_mul(0.f, wy)); uy = f_sub(f_mul(0.f, wx), f_mul(1.f, wz)); uz = f_sub(f_mul(1.f, wy), f_mul(0.f, wx));
    }
    pt_unit(ux, uy, uz);
    const float vx = f_sub(f_mul(wy, uz), f_mul(wz, uy)), vy = f_sub(f_mul(wz, ux), f_mul(wx, uz)), vz = f_sub(f_mul(wx, uy), f_mul(wy, ux));
    float sn, cs;
    if (F.sincos_tab) { const f2 t = F.sincos_tab[r1_bits]; sn = t.x; cs = t.y; }
    else sincos_glibc(r1, &sn, &cs);
    const float ku = f_mul(cs, r2s), kv = f_mul(sn, r2s), kw = f_sqrt(f_sub(1.f, r2));
    L.dx = f_add(f_add(f_mul(ku, ux), f_mul(kv, vx)), f_mul(kw, wx));
    L.dy = f_add(f_add(f_mul(ku, uy), f_mul(kv, vy)), f_mul(kw, wy));
    L.dz = f_add(f_add(f_mul(ku, uz), f_mul(kv, vz)), f_mul(kw, wz));
    L.phase = PH_BOUNCE;
}

// A new direction is set: SPT/geomfunc.h:182-185.
RT_HD void pt_bounce(PtLane &L) {
    L.depth++;
    if (L.depth > 6) L.phase = PH_END;
    else { L.phase = PH_NEAREST; L.cumu = PT_INF; L.hit = -1; }
}

// Folds the finished sample into the pixel (SPT/smallptCPU.cpp:110-118) and starts the next pass; returns true when the
// last pass of the pixel is done.
template <bool COUNT>
RT_HD bool pt_end_sample(PtLane &L, const PtFrame &F) {
    if (COUNT) L.c_samples++;
    if (F.sum_mode) {
        L.cr = f_add(L.cr, L.rr); L.cg = f_add(L.cg, L.rg); L.cb = f_add(L.cb, L.rb);
    } else if (L.pass == 0) {
        L.cr = L.rr; L.cg = L.rg; L.cb = L.rb;
    } else {
        const float k1 = (float)L.pass;
        const float k2 = f_rcp(f_add(k1, 1.f));
        L.cr = f_mul(f_add(f_mul(L.cr, k1), L.rr), k2);
        L.cg = f_mul(f_add(f_mul(L.cg, k1), L.rg), k2);
        L.cb = f_mul(f_add(f_mul(L.cb, k1), L.rb), k2);
    }
    L.pass++;
    if (L.pass >= F.pass0 + F.n_passes) { L.phase = PH_IDLE; return true; }
    pt_start_sample(L, F);
    return false;
}

// One lane, steps back to back: runs the shading code that follows a finished query until the lane needs its next
// query (returns false) or has completed the last pass of its pixel (returns true).
template <bool COUNT>
RT_HD bool pt_advance(PtLane &L, const PtFrame &F) {
    if (L.phase == PH_NEAREST) pt_hit<COUNT>(L, F);
    else pt_light_done<COUNT>(L, F);
    while (L.phase == PH_LIGHTS) pt_light_step(L, F);
    if (L.phase == PH_SHADOW) return false;
    if (L.phase == PH_DIFFUSE) pt_diffuse_bounce(L, F);
    if (L.phase == PH_BOUNCE) pt_bounce(L);
    if (L.phase == PH_NEAREST) return false;
    return pt_end_sample<COUNT>(L, F);
}

// pixels[y*w + x] of SPT/smallptCPU.cpp:120-122.
RT_HD uint32_t pt_pack_pixel(float r, float g, float b) {
    return (uint32_t)to_int_gamma(r) | ((uint32_t)to_int_gamma(g) << 8) | ((uint32_t)to_int_gamma(b) << 16);
}

// Work-item -> pixel mapping shared by both kernels.  The frame is cut into tiles of tile_rows rows;
// rank r owns tiles t with t % world == r.  The owned rows are walked in 8x4-pixel blocks so that the
// 32 lanes of a warp start on a compact screen patch (coherent rays, 32-byte store segments).
struct Shard { int rank, world, tile_rows, local_rows, blocks_x; };

RT_HD bool item_to_pixel(const Shard &S, int w, uint32_t item, int &x, int &y) {
    const uint32_t blk = item >> 5, j = item & 31u;
    const int bx = (int)(blk % (uint32_t)S.blocks_x), by = (int)(blk / (uint32_t)S.blocks_x);
    x = bx * 8 + (int)(j & 7u);
    const int lr = by * 4 + (int)(j >> 3);
    if (x >= w || lr >= S.local_rows) return false;
    const int lt = lr / S.tile_rows;
    y = (lt * S.world + S.rank) * S.tile_rows + lr % S.tile_rows;
    return true;
}

}  // namespace rtb
