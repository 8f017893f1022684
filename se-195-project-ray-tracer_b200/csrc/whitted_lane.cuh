// whitted_lane.cuh -- per-lane state machine of the Whitted tracer (host/device).
//
// Follows R323/raytracer_non_OpenCL.c ("RNO:") -- the CPU twin of R323/raytracer_kernel.cl that
// release 3.2.03 actually runs and whose output is the golden test.bmp.  One lane owns one pixel:
// nine sub-samples in the reference's order (tx outer, ty inner), each expanded breadth-first
// through a FIFO of secondary rays (RNO:30-40, 330-433), because the float accumulation order
// inside a pixel is part of the result.  As in pt_lane.cuh the control flow is cut into RAY QUERIES:
// a nearest-hit query over all primitives (RNO:185-193), then ONE batched hard-shadow query in which the
// lane tests up to three shadow rays (one per light, RNO:206-241) against the non-light primitives at
// once.  Every warp alternates the two rounds, so that all of its lanes run the same kind of query (a
// shadow lane never rides along a nearest loop, and the primitive record / loop overhead is shared by three
// rays).  Between the rounds: w_after_nearest (hit point, shadow-ray set-up), w_after_shadow (shading in
// light order) and w_finalize (accumulation, reflected / refracted children, next ray / sub-sample).
#pragma once
#include "rt_math.cuh"
#include "pt_lane.cuh"   // f4, PH_*, Shard

namespace rtb {

#define W_EPS 0.001f             /* RNO:26 */
#define W_FAR 10000000.0f        /* RNO:181 */
#define W_TRACEDEPTH 5           /* RNO:4   */
#ifndef W_QUEUE_SLOTS
#define W_QUEUE_SLOTS 32         /* a breadth-first queue over a depth-5 binary ray tree never holds more */
#endif
// Planes are tested one at a time in the nearest round: the paired variant (w_plane2, one vote for two divisions) made
// the loop a third longer in code for no fewer issue slots -- 3.2 ms against 3.0 ms at 1080p (profiles/r01_ab_variants.txt).
#ifndef W_PLANE_PAIRS
#define W_PLANE_PAIRS 0
#endif
#ifndef W_QUEUE_CHECK_SLOTS
#define W_QUEUE_CHECK_SLOTS W_QUEUE_SLOTS    /* tools/checked_build.sh builds a second library with 4 here: its FIFO check MUST fire */
#endif
#define W_SHADOW_BATCH 3          /* shadow rays per lane per shadow round */
static_assert(W_SHADOW_BATCH == 3, "w_after_shadow picks the batch's rays with three-way selects");
#define PH_FINAL 3                /* the current ray is complete: w_finalize() folds it into the pixel */
#define W_FLAG_SPHERE 1
#define W_FLAG_LIGHT 2
enum { W_PRIMARY = 0, W_REFLECTED = 1, W_REFRACTED = 2 };   /* RNO:59-63 */

// Scene in structure-of-arrays form (built once per frame from the 96-byte Primitive_2 records):
//   geom[i]  sphere: (center.xyz, sq_radius)   plane: (normal.xyz, depth)    <- all a test reads
//   flags[i] bit0 = sphere, bit1 = is_light
//   mat_a[i] (color.xyz, refl)     mat_b[i] (diff, refr, refr_index, spec)    rrad[i] = r_radius
// Shadow-candidate grid (see "Shadow-candidate grid" below): one word per cell of a box around the scene.
struct WGrid {
    const uint32_t *cells;      // NULL: no grid.  Bit i of a cell: primitive i (< 32) may block a shadow ray from a hit point in the cell
    float x0, y0, z0;           // low corner
    float ix, iy, iz;           // cells per unit length
    float fgx, fgy, fgz;        // cells per axis, as floats
    int gx, gy, gz;
    uint32_t all;               // the word of a hit point outside the box: every primitive a shadow query tests
    // Primary-ray candidates ("Primary-ray tiles" below): one word per 8x4-pixel tile of the frame
    const uint32_t *tiles;      // NULL: none
    int tiles_x, tiles_y;
    uint32_t all_nearest;       // the word of any other ray: every primitive a nearest query tests
    uint32_t deep;              // the primitives that reflect or refract (m_refl > 0 or m_refr > 0): a ray that hits one has children
};

struct WFrame {
    const f4 *geom, *mat_a, *mat_b;
    const int *flags, *lights;
    const f4 *lcenter;          // lcenter[k] = the `center` field of light k (in lights[] order), whatever its type: a light that
                                //  is not a sphere is shaded towards that point without a shadow ray (RNO:214-223); global memory
    const int *runs;            // maximal index runs of equal (type, is_light): triples (start, count, flags)
    const float *rrad;
    int n, n_lights, n_spheres, n_planes, n_runs;
    int w, h;
    float DX, DY;               // (WX2-WX1)/w, (WY2-WY1)/h computed on the host exactly as RNO:295-296
    int32_t *hit_ids;           // NULL or int[w*h*9]
    // Shadow-round culls (see "Exact culls of the shadow round" below); NULL = none.  Built on the host by build_w_cull.
    const f2 *pcull;            // per primitive, planes only: (sgn, T) -- no light-bound ray from a point P with sgn*(N.P + depth) > T can hit the plane
    const f4 *rbox;             // per run of `runs`: (lo.xyz, -) and (hi.xyz, -) of the run's spheres, grown by the proven margin; a face
                                //  that does not separate the run from every light is at -+inf
    float cull_rp2;             // the culls' margins hold for hit points with |P|^2 < cull_rp2
    float reject_k;             // K of w_shadow_sphere_keep for this scene
    WGrid grid;
    int split0;                 // 1: the pixels of cost class 0 are rendered by whitted_split_kernel; this launch starts at class 1
    // Blocked lights (see "Blocked lights and the redo list" below).
    float tame_reach[2];        // a hit point on a plane [0] / a sphere [1] closer than this to every light of its batch may skip its blocked lights;
                                //  0: none may (build_w_soa)
    unsigned *redo_count;       // NULL (EXACT / counting launches), or the number of pixels reported so far
    uint32_t *redo_list;        // pixel numbers y * w + x of the first redo_cap reports
    unsigned redo_cap;
};

struct WLane {
    int x, y, sub;
    float ar, ag, ab;                               // pixel accumulator
    int nlog;                                       // SPLIT: triples in the lane's log
    int head, tail;                                 // FIFO cursors of the current sub-sample
    float dx, dy, dz;                               // direction of the ray being processed
    float weight, r_index, tr, tg, tb;              // its weight, medium index, transparency
    int depth, kind, from;
    float qox, qoy, qoz, qdx, qdy, qdz;             // ray of the query in flight
    float cumu;                                     // nearest distance so far / distance to the light
    int qhit, qkind;                                // primitive of the accepted hit (-1 none) + HIT(1)/INPRIM(-1); shadow: qhit >= 0 = blocked
    float dist; int hit, hkind;                     // result of the nearest query while lights are processed
    float px, py, pz;                               // intersection point
    float cr, cg, cb;                               // colour gathered for this ray
    float sox[W_SHADOW_BATCH], soy[W_SHADOW_BATCH], soz[W_SHADOW_BATCH];   // shadow rays of the current batch: origins,
    float slx[W_SHADOW_BATCH], sly[W_SHADOW_BATCH], slz[W_SHADOW_BATCH];   //  unit directions to lights li .. li+ns-1
    float sreach[W_SHADOW_BATCH];                   //  and distances to them
    int ns, sblk;                                   // rays in the batch; bit k set = ray k is blocked
    bool pnear;                                     // |P|^2 < F.cull_rp2: the shadow culls' margins cover this hit point
    int li, phase;
    uint32_t c_nearest, c_shadow, c_samples;
    uint32_t c_shadow_lit;                          // shadow rays a TIMED launch traces: those of hits on a material with a diffuse or specular term
    uint64_t c_sphere_tests, c_plane_tests;
};

// "Sphere entirely behind the ray": b < 0 and det < b*b*(1-2^-22) imply sqrt_rn(det) <= |b|, so the far root
// b + sqrt(det) is <= 0 and the reference's `i2 > 0` fails (bb = fl(b*b) <= b*b*(1+2^-24), and the rounded
// product with 1-2^-22 stays below b*b; sqrt_rn is monotonic and |b| is representable).  Skips most roots.
RT_HD bool w_sphere_behind(float b, float bb, float det) { return (b < 0.f) & (det < f_mul(bb, 0.999999761581420898437500f)); }

// plane_intersect (RNO:95-109) and sphere_intersect (RNO:111-148) against the lane's query ray, written
// without per-lane branches: the primitive index is warp-uniform, so all 32 lanes run one instruction
// stream and a hit is a predicated update of (cumu, qhit, qkind).  `live` masks lanes that take no part
// (no query in flight, or a shadow query looking at a light).  The two expensive IEEE operations sit
// behind warp votes: the square root is only evaluated when some lane has det > 0, the division only
// when some lane can still be hit.  For a shadow query qhit >= 0 simply means "blocked"; the reference
// stops at the first blocker, which changes nothing but the test count (kept exact in counting builds).
//
// Plane pre-filters (exact: they only discard tests the reference's own comparison would fail):
//   A. dist = num/d with num = -(N.o + depth).  dist > 0 needs num and d non-zero and of equal sign.
//   B. |num| > (cumu*|d|)*(1+2^-21), all factors rounded, implies |num|/|d| > cumu*(1+2^-22), hence the
//      correctly rounded quotient is >= cumu + ulp and `dist < cumu` fails.  Skipped when the product is
//      not comfortably normal (the error bound would not hold for subnormals; inf never rejects).
//   Both at once: with s = N.o + depth (num = -s) and the signed limit w = (d*cumu)*(1+2^-21), A and B say that the
//   division is only worth taking if num lies between 0 and w, i.e. num*(w - num) >= 0, i.e. s*(w + s) <= 0: one
//   addition, one multiplication and one comparison (w_plane_candidate).  The sign of a rounded sum is the exact one,
//   an underflowing product keeps its sign, and a zero counts as "take the division".  The candidate set is a superset
//   of A-and-B (it keeps d == 0 / num == 0); it only gates the exact stage, which re-checks 0 < dist < cumu and
//   rejects the inf / NaN a zero d produces, so the accepted hits are the same.
RT_HD bool w_plane_candidate(float s, float d, float limit) {
    const float w = f_mul(f_mul(d, limit), 1.000000476837158203125f);
    const float m = f_mul(s, f_add(w, s));
    return (m <= 0.f) | !(fabsf(w) > 1e-30f);
}
template <bool COUNT>
RT_HD void w_plane(WLane &L, const f4 g, int i, bool live) {
    if (COUNT && live && L.phase == PH_SHADOW && L.qhit < 0) L.c_plane_tests++;
    const float d = dot3(g.x, g.y, g.z, L.qdx, L.qdy, L.qdz);
    const float s = f_add(dot3(g.x, g.y, g.z, L.qox, L.qoy, L.qoz), g.w);
    const bool cand = w_plane_candidate(s, d, L.cumu);
    if (warp_any(cand & live)) {
        const float dist = f_div(-s, d);
        if (live & cand & (dist > 0.f) & (dist < L.cumu)) { L.cumu = dist; L.qhit = i; L.qkind = 1; }
    }
}
template <bool COUNT>
RT_HD void w_sphere(WLane &L, const f4 g, int i, bool live) {
    if (COUNT && live && L.phase == PH_SHADOW && L.qhit < 0) L.c_sphere_tests++;
    const float vx = f_sub(L.qox, g.x), vy = f_sub(L.qoy, g.y), vz = f_sub(L.qoz, g.z);
    const float b = -dot3(vx, vy, vz, L.qdx, L.qdy, L.qdz);
    const float bb = f_mul(b, b);
    const float det = f_add(f_sub(bb, dot3(vx, vy, vz, vx, vy, vz)), g.w);
    const bool cand = (det > 0.f) & !w_sphere_behind(b, bb, det);
    if (warp_any(cand & live)) {
        const float dv[1] = { det };
        const bool need[1] = { cand & live };
        float sqv[1];
        sqrt_group<1>(dv, need, sqv);
        const float sq = sqv[0];
        const float i1 = f_sub(b, sq), i2 = f_add(b, sq);
        const bool inside = i1 < 0.f;                           // ray starts inside: take the far root, INPRIM
        const float t = inside ? i2 : i1;
        if (live & cand & (i2 > 0.f) & (t < L.cumu)) { L.cumu = t; L.qhit = i; L.qkind = inside ? -1 : 1; }
    }
}

// Two consecutive planes (i, i+1) at once: two independent chains, one vote for both divisions.
template <bool COUNT>
RT_HD void w_plane2(WLane &L, const f4 *g, int i, bool live) {
    float d[2], sv[2];
    bool cand[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const f4 s = g[k];
        d[k] = dot3(s.x, s.y, s.z, L.qdx, L.qdy, L.qdz);
        sv[k] = f_add(dot3(s.x, s.y, s.z, L.qox, L.qoy, L.qoz), s.w);
        // pre-filter B uses the limit before either plane of the pair is applied: cumu only shrinks, so a
        // plane that is beyond this limit is beyond the later one too
        cand[k] = w_plane_candidate(sv[k], d[k], L.cumu);
    }
    if (warp_any((cand[0] | cand[1]) & live)) {
        const float q0 = f_div(-sv[0], d[0]), q1 = f_div(-sv[1], d[1]);
        if (COUNT && live && L.phase == PH_SHADOW && L.qhit < 0) L.c_plane_tests++;
        if (live & cand[0] & (q0 > 0.f) & (q0 < L.cumu)) { L.cumu = q0; L.qhit = i; L.qkind = 1; }
        if (COUNT && live && L.phase == PH_SHADOW && L.qhit < 0) L.c_plane_tests++;
        if (live & cand[1] & (q1 > 0.f) & (q1 < L.cumu)) { L.cumu = q1; L.qhit = i + 1; L.qkind = 1; }
    } else if (COUNT && live && L.phase == PH_SHADOW && L.qhit < 0) L.c_plane_tests += 2;
}

// Two consecutive spheres (i, i+1) at once: two independent discriminant chains, one vote for both roots,
// acceptance in ascending index order.
template <bool COUNT>
RT_HD void w_sphere2(WLane &L, const f4 *g, int i, bool live) {
    float b[2], det[2];
    bool cand[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const f4 s = g[k];
        const float vx = f_sub(L.qox, s.x), vy = f_sub(L.qoy, s.y), vz = f_sub(L.qoz, s.z);
        b[k] = -dot3(vx, vy, vz, L.qdx, L.qdy, L.qdz);
        const float bb = f_mul(b[k], b[k]);
        det[k] = f_add(f_sub(bb, dot3(vx, vy, vz, vx, vy, vz)), s.w);
        cand[k] = live & (det[k] > 0.f) & !w_sphere_behind(b[k], bb, det[k]);
    }
    if (warp_any(cand[0] | cand[1])) {
        float sqv[2];
        sqrt_group<2>(det, cand, sqv);
#pragma unroll
        for (int k = 0; k < 2; k++) {
            if (COUNT && live && L.phase == PH_SHADOW && L.qhit < 0) L.c_sphere_tests++;
            const float sq = sqv[k];
            const float i1 = f_sub(b[k], sq), i2 = f_add(b[k], sq);
            const bool inside = i1 < 0.f;
            const float t = inside ? i2 : i1;
            if (cand[k] & (i2 > 0.f) & (t < L.cumu)) { L.cumu = t; L.qhit = i + k; L.qkind = inside ? -1 : 1; }
        }
    } else if (COUNT && live && L.phase == PH_SHADOW && L.qhit < 0) L.c_sphere_tests += 2;
}

// Nearest-hit round: runs of equal (type, is_light) in ascending index order (ties: the lowest index keeps
// an exact tie because the acceptance test is a strict '<', RNO:185-193).  `has` = this lane has a ray.
template <bool COUNT>
RT_HD void w_query_nearest(WLane &L, const f4 *geom, const int *runs, int n_runs, bool has) {
    if (!warp_any(has)) return;
    for (int r = 0; r < n_runs; ++r) {
        const int start = runs[3 * r], count = runs[3 * r + 1], fl = runs[3 * r + 2];
        int i = start;
        const int end = start + count;
        if (fl & W_FLAG_SPHERE) {
            for (; i + 1 < end; i += 2) w_sphere2<COUNT>(L, geom + i, i, has);
            if (i < end) w_sphere<COUNT>(L, geom[i], i, has);
        } else {
#if W_PLANE_PAIRS
            for (; i + 1 < end; i += 2) w_plane2<COUNT>(L, geom + i, i, has);
#endif
            for (; i < end; ++i) w_plane<COUNT>(L, geom[i], i, has);
        }
    }
}

// Shadow round, one primitive against the lane's (up to) three shadow rays.  Only the boolean matters
// ("is there a non-light primitive with 0 < dist < distance to the light", RNO:232-240), so:
//   sphere: det > 0, far root > 0, and the entry root (or the far root when the origin is inside) < reach;
//           the three square roots share one warp vote;
//   plane:  the candidate test of the nearest round with the distance to the light as the limit (pre-filters A
//           and B): dist = num/d is only formed for a plane that may really lie between the point and the light.
// `alive` = per-ray participation mask (bit k: ray k exists and was not blocked when the run started; counting
// builds refresh it per primitive so that the test counters stop exactly at the first blocker).
RT_HD int w_alive_mask(const WLane &L, bool has) { return has ? (((1 << L.ns) - 1) & ~L.sblk) : 0; }

template <bool COUNT>
RT_HD void w_shadow_sphere(WLane &L, const f4 g, int alive_in, bool has) {
    float b[W_SHADOW_BATCH], det[W_SHADOW_BATCH];
    bool cand[W_SHADOW_BATCH];
    bool any = false;
    const int alive = COUNT ? w_alive_mask(L, has) : alive_in;
#pragma unroll
    for (int k = 0; k < W_SHADOW_BATCH; k++) {
        if (COUNT && ((alive >> k) & 1)) L.c_sphere_tests++;
        const float vx = f_sub(L.sox[k], g.x), vy = f_sub(L.soy[k], g.y), vz = f_sub(L.soz[k], g.z);
        b[k] = -dot3(vx, vy, vz, L.slx[k], L.sly[k], L.slz[k]);
        const float bb = f_mul(b[k], b[k]);
        det[k] = f_add(f_sub(bb, dot3(vx, vy, vz, vx, vy, vz)), g.w);
        cand[k] = (((alive >> k) & 1) != 0) & (det[k] > 0.f) & !w_sphere_behind(b[k], bb, det[k]);
        any = any | cand[k];
    }
    if (warp_any(any)) {
        float sq[W_SHADOW_BATCH];
        sqrt_group<W_SHADOW_BATCH>(det, cand, sq);
#pragma unroll
        for (int k = 0; k < W_SHADOW_BATCH; k++) {
            const float i1 = f_sub(b[k], sq[k]), i2 = f_add(b[k], sq[k]);
            const float t = i1 < 0.f ? i2 : i1;
            if (cand[k] & (i2 > 0.f) & (t < L.sreach[k])) L.sblk |= 1 << k;
        }
    }
}
template <bool COUNT>
RT_HD void w_shadow_plane(WLane &L, const f4 g, int alive_in, bool has) {
    float d[W_SHADOW_BATCH], sv[W_SHADOW_BATCH];
    bool cand[W_SHADOW_BATCH];
    bool any = false;
    const int alive = COUNT ? w_alive_mask(L, has) : alive_in;
#pragma unroll
    for (int k = 0; k < W_SHADOW_BATCH; k++) {
        if (COUNT && ((alive >> k) & 1)) L.c_plane_tests++;
        d[k] = dot3(g.x, g.y, g.z, L.slx[k], L.sly[k], L.slz[k]);
        sv[k] = f_add(dot3(g.x, g.y, g.z, L.sox[k], L.soy[k], L.soz[k]), g.w);
        // pre-filters A and B with the distance to the light as the limit: what survives -- a plane that may
        // really block -- takes the division.
        cand[k] = (((alive >> k) & 1) != 0) & w_plane_candidate(sv[k], d[k], L.sreach[k]);
        any = any | cand[k];
    }
    if (warp_any(any)) {
#pragma unroll
        for (int k = 0; k < W_SHADOW_BATCH; k++) {
            const float dist = f_div(-sv[k], d[k]);
            if (cand[k] & (dist > 0.f) & (dist < L.sreach[k])) L.sblk |= 1 << k;
        }
    }
}
// Exact culls of the shadow round (timed launches of scenes whose lights are all spheres; tables from build_w_cull,
// scene_soa.h, which also states the margins).  A shadow ray runs from o = fl(P + L*EPS) along L = fl((1/len) * (c_light - P))
// and is blocked by a primitive iff the reference's float test returns a distance in (0, len) (RNO:232-240).  Both culls use
// one fact: a linear function of the ray parameter that has the same sign at both ends of the segment has no zero on it.
//   plane    g(t) = N.(o + t L) + depth.  g at the light's centre is a constant of the scene (sign sgn, magnitude >= 2T);
//            if sgn * (N.P + depth) > T -- evaluated ONCE per hit point for the whole batch, fused, any rounding is inside
//            T -- then the plane's zero lies before o or beyond the light by more than the reference's rounding can move it,
//            so its test returns a distance <= 0, >= len, or inf / NaN for every ray of the batch: not a blocker.  The plane a
//            point lies on (|N.P + depth| ~ 0) and planes that pass between the lights always take the full test.
//   spheres  a run of spheres lies inside its box; if P and every light's centre lie beyond the same face of the box grown by
//            m (the inflated-sphere bound of pt_bvh.cuh: det >= 0 means the ray's line passes within R' of the centre, and
//            the accepted distance is the parameter of a point of that inflated sphere up to 18u|op|), every point of the
//            segment does, and no sphere of the run can return an accepted distance: the whole run is skipped for the lane.
// A lane that is culled keeps its rays out of the run (alive = 0); the tests themselves are unchanged.
RT_HD float w_fused_plane_side(const f4 g, float px, float py, float pz) {
#ifdef __CUDA_ARCH__
    return __fmaf_rn(g.x, px, __fmaf_rn(g.y, py, __fmaf_rn(g.z, pz, g.w)));
#else
    return fmaf(g.x, px, fmaf(g.y, py, fmaf(g.z, pz, g.w)));
#endif
}
// Fused conservative reject of a sphere for the (up to three) shadow rays of a hit point P (SURVEY 7 "hard parts" 1).  The rays start at
// o_k = fl(P + L_k EPS); with v0 = P - c the discriminant of ray k does not depend on the EPS offset (sliding the origin along the ray
// keeps its line), so |v0|^2 - r^2 is shared by the batch and a ray costs one 3-term FMA dot product and two FMAs instead of 16 single
// roundings.  A ray is rejected iff its fused discriminant is below -E, or the sphere lies behind it (q = v0.L > E) with the origin
// outside (|v_k|^2 - r^2 > E): then the reference's det is negative, resp. its far root is <= 0 (w_sphere_behind's argument), whatever
// the roundings, because E = K (|v0|^2 + r^2 + 1) with K = (64 + 8 RP) u exceeds every difference between the two evaluations: 20u
// (|v|^2 + r^2) from the arithmetic itself (pt_bvh.cuh) and 7u |v| |P| from the rounding of o_k, |P| < RP.  NaN compares false: kept.
// Returns the rays of `alive` that the exact test still has to look at.
RT_HD float w_fma(float a, float b, float c) {
#ifdef __CUDA_ARCH__
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}
RT_HD int w_shadow_sphere_keep(const WLane &L, const f4 g, int alive, float K) {
    const float vx = L.px - g.x, vy = L.py - g.y, vz = L.pz - g.z;
    const float vv = w_fma(vx, vx, w_fma(vy, vy, vz * vz));
    const float c0 = vv - g.w, E = w_fma(K, vv + g.w, K), c1 = c0 + W_EPS * W_EPS;
    int keep = 0;
#pragma unroll
    for (int k = 0; k < W_SHADOW_BATCH; k++) {
        const float q = w_fma(vx, L.slx[k], w_fma(vy, L.sly[k], vz * L.slz[k]));
        const float det = w_fma(q, q, -c0), ck = w_fma(2.0f * W_EPS, q, c1);
        const bool rej = (det < -E) | ((q > E) & (ck > E));
        keep |= (rej ? 0 : 1) << k;
    }
    return keep & alive;
}
// Shadow-candidate grid (timed launches of scenes of at most 32 primitives whose lights are all spheres).  The two culls above decide per
// hit point, with a handful of instructions per plane and per run of spheres.  Nearly all of that work has the same outcome for all hit
// points of a neighbourhood, so it is done once per scene instead: a box around the scene is cut into cells, and for every cell the device
// computes (w_grid_cell_mask, in double, margins rounded towards "keep") which primitives could block a shadow ray from ANY hit point in
// the cell to ANY light:
//   plane i    unless the plane cull's condition sgn (N.P + depth) > T holds at every point of the cell (the minimum over the cell is at a
//              corner: the value at the centre minus sum |N_a| h_a);
//   sphere i   iff for some light the segment from the cell's centre Pc to the light's centre passes within rad + grow + h of the sphere's
//              centre (h = half the cell's diagonal).  build_w_cull's argument for a run's box holds for any convex region: the points the
//              reference's test can return a distance for lie within `grow` - (R' - rad) of the real segment [P, c_l], which lies within
//              |P - Pc| <= h of [Pc, c_l]; det >= 0 needs one of them within R' of the centre.
// A shadow round then loads ONE word for the lane's hit point (cell index from three multiplications; the float rounding of the index is
// covered by cells taken 0.1 % larger), ORs the words of the warp, and runs the tests of the primitives whose bit is set somewhere in the
// warp -- typically the wall the points lie on and no sphere, or one.  A hit point outside the box or beyond the culls' radius (L.pnear)
// takes G.all.  The tests themselves, and the fused reject in front of the sphere test, are unchanged.
RT_HD uint32_t w_grid_cell_mask(double cx, double cy, double cz, double hx, double hy, double hz, const f4 *geom, const int *flags,
                                const f2 *pcull, const float *smargin, uint32_t all, const f4 *lcenter, int n_lights) {
    uint32_t m = 0;
    const double h = sqrt(hx * hx + hy * hy + hz * hz);
    for (int i = 0; i < 32; i++) {
        if (!((all >> i) & 1u)) continue;
        const f4 g = geom[i];
        bool keep = true;
        if (flags[i] & W_FLAG_SPHERE) {
            keep = false;
            const double reach = (double)smargin[i] + h;
            for (int l = 0; l < n_lights && !keep; l++) {
                const f4 c = lcenter[l];
                const double vx = (double)c.x - cx, vy = (double)c.y - cy, vz = (double)c.z - cz;
                const double wx = (double)g.x - cx, wy = (double)g.y - cy, wz = (double)g.z - cz;
                const double vv = vx * vx + vy * vy + vz * vz;
                double t = vv > 0.0 ? (wx * vx + wy * vy + wz * vz) / vv : 0.0;
                t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
                const double ex = wx - t * vx, ey = wy - t * vy, ez = wz - t * vz;
                if (!(ex * ex + ey * ey + ez * ez > reach * reach * 1.000001)) keep = true;         // NaN: keep
            }
        } else {
            const f2 cu = pcull[i];
            const double side = (double)cu.x * ((double)g.x * cx + (double)g.y * cy + (double)g.z * cz + (double)g.w);
            const double slack = fabs((double)g.x) * hx + fabs((double)g.y) * hy + fabs((double)g.z) * hz;
            if (side - slack > (double)cu.y * 1.000001 + 1e-30) keep = false;                       // T = inf (no cull for this plane), NaN: keep
        }
        if (keep) m |= 1u << i;
    }
    return m;
}
// Cell c of the grid (cell sizes from 1 / G.i*; the cells are taken 0.1 % larger than they are).
RT_HD uint32_t w_grid_build_cell(const WGrid &G, int gz, int c, const f4 *geom, const int *flags, const f2 *pcull, const float *smargin, const f4 *lcenter, int n_lights) {
    const int x = c % G.gx, y = (c / G.gx) % G.gy, z = c / (G.gx * G.gy);
    (void)gz;
    const double sx = 1.0 / (double)G.ix, sy = 1.0 / (double)G.iy, sz = 1.0 / (double)G.iz;
    return w_grid_cell_mask((double)G.x0 + (x + 0.5) * sx, (double)G.y0 + (y + 0.5) * sy, (double)G.z0 + (z + 0.5) * sz,
                            0.5005 * sx, 0.5005 * sy, 0.5005 * sz, geom, flags, pcull, smargin, G.all, lcenter, n_lights);
}
RT_HD uint32_t w_grid_lookup(const WLane &L, const WGrid &G) {
    const float fx = (L.px - G.x0) * G.ix, fy = (L.py - G.y0) * G.iy, fz = (L.pz - G.z0) * G.iz;
    const bool in = L.pnear & (fx >= 0.f) & (fx < G.fgx) & (fy >= 0.f) & (fy < G.fgy) & (fz >= 0.f) & (fz < G.fgz);       // NaN: outside
    RT_CHECK(!in || ((int)fx < G.gx && (int)fy < G.gy && (int)fz < G.gz), RT_CHK_TABLE);
    return in ? G.cells[((int)fz * G.gy + (int)fy) * G.gx + (int)fx] : G.all;
}
// The shadow round over the grid's candidates (same tests as w_query_shadow<false, true>, fewer of them).
RT_HD void w_query_shadow_grid(WLane &L, const f4 *geom, const int *flags, bool has, const WGrid &G, float reject_k) {
    int alive = w_alive_mask(L, has);
    const uint32_t m = alive ? w_grid_lookup(L, G) : 0u;
    for (uint32_t todo = warp_or(m); todo; todo &= todo - 1u) {
        const int i = w_lowest_bit(todo);
        alive = ((m >> i) & 1u) ? w_alive_mask(L, has) : 0;
        if (!warp_any(alive != 0)) continue;
        const f4 g = geom[i];
        if (flags[i] & W_FLAG_SPHERE) {
            const int keep = L.pnear ? w_shadow_sphere_keep(L, g, alive, reject_k) : alive;
            if (warp_any(keep != 0)) w_shadow_sphere<false>(L, g, keep, has);
        } else w_shadow_plane<false>(L, g, alive, has);
    }
}

// Primary-ray tiles (same launches as the grid).  Every primary ray starts at the eye, so which primitives it can reach is a function of
// the pixel; for each 8x4-pixel tile the device computes once per (scene, frame size) which primitives can be the ACCEPTED hit of some
// primary ray of the tile (w_tile_mask, in double).  For the rectangle of un-normalised directions D = (u, v, 7) the tile's nine
// sub-samples per pixel span (grown by a thousandth of a pixel), each primitive gets: can it be hit at all, is it certainly hit by every
// ray, and between which distances:
//   plane   t = -s |D| / (N.D), s = N.O + depth: N.D is linear in (u, v), so its range is taken at the corners; a range that (nearly)
//           contains 0, or s ~ 0, leaves the plane "possible" from the smallest distance any ray can see; otherwise the sign of -s / N.D
//           decides between "missed by all" (the reference's dist <= 0) and "hit by all" with t in |s| [min |D|, max |D|] / |N.D|;
//   sphere  against the cone around the tile's centre ray that holds its corners: outside the sphere inflated to rad + grow (scene_soa.h:
//           where det can still be positive) = missed; cone inside the sphere deflated by the same = hit by all, entry distance at most that
//           of the deflated sphere along the cone's rim; never closer than |C - O| - rad - grow.
// The float distances of the reference differ from these by parts in 10^6; the intervals carry 10^-4.  A primitive is dropped from the
// tile's word iff it is missed by all rays or lies farther than the farthest possible distance of some primitive that all rays hit:
// then it is never the accepted hit, and leaving out its test changes no accepted (cumu, qhit, qkind) -- the order of the remaining tests
// (ascending index, strict '<') is unchanged.  A nearest round ORs the words of the warp's lanes (any other ray: G.all_nearest) and
// tests the primitives whose bit is set; a warp of 8x4 wall pixels tests one or two planes instead of every primitive.
struct WTileCam { double ox, oy, oz, u0, u1, v0, v1, dz; };
RT_HD uint32_t w_tile_mask(const WTileCam &T, const f4 *geom, const int *flags, const float *smargin, uint32_t all) {
    const double cu[4] = { T.u0, T.u1, T.u0, T.u1 }, cv[4] = { T.v0, T.v0, T.v1, T.v1 };
    // |D| over the rectangle: smallest at the point closest to (0, 0), largest at a corner
    const double un = T.u0 > 0.0 ? T.u0 : (T.u1 < 0.0 ? T.u1 : 0.0), vn = T.v0 > 0.0 ? T.v0 : (T.v1 < 0.0 ? T.v1 : 0.0);
    const double la = sqrt(un * un + vn * vn + T.dz * T.dz);
    double lb = 0.0;
    for (int c = 0; c < 4; c++) { const double l = sqrt(cu[c] * cu[c] + cv[c] * cv[c] + T.dz * T.dz); lb = l > lb ? l : lb; }
    // the cone: axis through the rectangle's centre, half-angle to the farthest corner
    double ax = 0.5 * (T.u0 + T.u1), ay = 0.5 * (T.v0 + T.v1), az = T.dz;
    const double al = sqrt(ax * ax + ay * ay + az * az);
    ax /= al; ay /= al; az /= al;
    double cmin = 1.0;
    for (int c = 0; c < 4; c++) {
        const double l = sqrt(cu[c] * cu[c] + cv[c] * cv[c] + T.dz * T.dz);
        const double cs = (ax * cu[c] + ay * cv[c] + az * T.dz) / l;
        cmin = cs < cmin ? cs : cmin;
    }
    const double theta = acos(cmin < -1.0 ? -1.0 : (cmin > 1.0 ? 1.0 : cmin)) + 1e-7;
    const double olen = sqrt(T.ox * T.ox + T.oy * T.oy + T.oz * T.oz);
    uint32_t possible = 0;
    double tlo[32];
    double tstar = INFINITY;                      // no accepted distance exceeds this: the farthest hit of some primitive all rays hit
    for (int i = 0; i < 32; i++) {
        tlo[i] = 0.0;
        if (!((all >> i) & 1u)) continue;
        const f4 g = geom[i];
        bool poss = true, cert = false;
        double lo = 0.0, hi = INFINITY;
        if (flags[i] & W_FLAG_SPHERE) {
            const double wx = (double)g.x - T.ox, wy = (double)g.y - T.oy, wz = (double)g.z - T.oz;
            const double dist = sqrt(wx * wx + wy * wy + wz * wz);
            const double rad = sqrt((double)g.w), rp = (double)smargin[i], rm = 2.0 * rad - rp;
            if (dist > rp * 1.0001 && rp >= rad) {            // the eye is outside (else: possible, from distance 0)
                double ca = (wx * ax + wy * ay + wz * az) / dist;
                ca = ca < -1.0 ? -1.0 : (ca > 1.0 ? 1.0 : ca);
                const double alpha = acos(ca), bp = asin(rp / dist);
                if (alpha > theta + bp + 1e-6) poss = false;
                else {
                    lo = dist - rp;
                    if (rm > 0.0) {
                        const double am = alpha + theta + 1e-6, bm = asin(rm / dist);
                        if (am < bm) {
                            const double sn = dist * sin(am), q = rm * rm - sn * sn;
                            if (q > 0.0) { cert = true; hi = dist * cos(am) - sqrt(q); }
                        }
                    }
                }
            }
        } else {
            // float side: s_f = fl(N.o + depth) within es of sv, d_f = fl(N.q) within ed of c = N.D / |D| (q the float unit direction),
            // dist = fl(-s_f / d_f), accepted iff 0 < dist < cumu (RNO:97-108)
            const double nn = sqrt((double)g.x * g.x + (double)g.y * g.y + (double)g.z * g.z);
            const double sv = (double)g.x * T.ox + (double)g.y * T.oy + (double)g.z * T.oz + (double)g.w;
            const double es = 1e-6 * (nn * olen + fabs((double)g.w)), ed = 4e-6 * nn;
            double a = INFINITY, b = -INFINITY;
            for (int c = 0; c < 4; c++) {
                const double nd = (double)g.x * cu[c] + (double)g.y * cv[c] + (double)g.z * T.dz;
                a = nd < a ? nd : a; b = nd > b ? nd : b;
            }
            const double amax = fabs(a) > fabs(b) ? fabs(a) : fabs(b), amin = fabs(a) < fabs(b) ? fabs(a) : fabs(b);
            const double cmax = amax / la, cmn = amin / lb;              // |c| <= cmax; |c| >= cmn when N.D keeps its sign
            const bool signed_nd = (a > 0.0) == (b > 0.0) && a != 0.0 && b != 0.0 && cmn > 100.0 * ed;
            if (!(nn > 0.0) || !(fabs(sv) <= 1e300) || !(cmax <= 1e300)) { /* degenerate: possible, from 0 */ }
            else if (signed_nd && fabs(sv) > 100.0 * es) {
                if ((sv > 0.0) == (a > 0.0)) poss = false;             // -s_f / d_f < 0 for every ray
                else { cert = true; lo = (fabs(sv) - es) / (cmax + ed); hi = (fabs(sv) + es) / (cmn - ed); }
            } else lo = (fabs(sv) > es ? fabs(sv) - es : 0.0) / (cmax + ed);
        }
        if (!poss) continue;
        possible |= 1u << i;
        tlo[i] = lo * (1.0 - 1e-4) - 1e-4;
        if (cert) { const double h = hi * (1.0 + 1e-4) + 1e-4; if (h < tstar) tstar = h; }
    }
    uint32_t m = 0;
    for (int i = 0; i < 32; i++) if (((possible >> i) & 1u) && !(tlo[i] > tstar)) m |= 1u << i;       // NaN: keep
    return m;
}
// Tile (tx, ty) of a w x h frame with the camera of w_start_subsample (RNO:299-328).
RT_HD uint32_t w_tile_build(int tx, int ty, int w, int h, float DX, float DY, const f4 *geom, const int *flags, const float *smargin, uint32_t all) {
    const int x0 = tx * 8, y0 = ty * 4, x1 = (x0 + 7 < w - 1) ? x0 + 7 : w - 1, y1 = (y0 + 3 < h - 1) ? y0 + 3 : h - 1;
    WTileCam T;
    T.ox = 0.0; T.oy = 0.25; T.oz = -7.0; T.dz = 7.0;
    const double dx = (double)DX, dy = (double)DY;
    const double ua = -3.0 + (x0 - 0.5) * dx, ub = -3.0 + (x1 + 0.5) * dx, va = 2.0 + (y0 - 0.5) * dy, vb = 2.0 + (y1 + 0.5) * dy;
    const double pu = 1e-3 * fabs(dx) + 1e-5, pv = 1e-3 * fabs(dy) + 1e-5;
    T.u0 = (ua < ub ? ua : ub) - pu; T.u1 = (ua < ub ? ub : ua) + pu;
    T.v0 = (va < vb ? va : vb) - pv; T.v1 = (va < vb ? vb : va) + pv;
    return w_tile_mask(T, geom, flags, smargin, all);
}
RT_HD void w_query_nearest_tiles(WLane &L, const f4 *geom, const int *flags, const int *runs, int n_runs, bool has, const WGrid &G) {
    if (!G.tiles) { w_query_nearest<false>(L, geom, runs, n_runs, has); return; }       // RT_TUNE_WHITTED_GRID = 2: the grid without the tiles
    uint32_t m = 0;
    RT_CHECK(!has || L.kind != W_PRIMARY || ((L.x >> 3) < G.tiles_x && (L.y >> 2) < G.tiles_y && L.x >= 0 && L.y >= 0), RT_CHK_TABLE);
    if (has) m = (L.kind == W_PRIMARY) ? G.tiles[(L.y >> 2) * G.tiles_x + (L.x >> 3)] : G.all_nearest;
    uint32_t todo = warp_or(m);
    if (todo == G.all_nearest) { w_query_nearest<false>(L, geom, runs, n_runs, has); return; }       // nothing to leave out: the paired loops
    for (; todo; todo &= todo - 1u) {
        const int i = w_lowest_bit(todo);
        const bool live = (m >> i) & 1u;
        const f4 g = geom[i];
        if (flags[i] & W_FLAG_SPHERE) w_sphere<false>(L, g, i, live);
        else w_plane<false>(L, g, i, live);
    }
}

template <bool COUNT, bool CULL = false>
RT_HD void w_query_shadow(WLane &L, const f4 *geom, const int *runs, int n_runs, bool has, const f2 *pcull = nullptr, const f4 *rbox = nullptr, float reject_k = 0.f) {
    for (int r = 0; r < n_runs; ++r) {
        const int start = runs[3 * r], count = runs[3 * r + 1], fl = runs[3 * r + 2];
        if (fl & W_FLAG_LIGHT) continue;                                   // RNO:234: lights cast no shadow
        int alive = w_alive_mask(L, has);                                  // rays of this lane that are still unblocked
        if (!warp_any(alive != 0)) return;                                 // the `break` of RNO:237, for the whole warp
        const int end = start + count;
        // (loading the next primitive's record before testing the current one changed nothing: 2.97 against 2.98 ms)
        if (fl & W_FLAG_SPHERE) {
#ifndef W_NO_RUN_CULL
            if (CULL) {
                const f4 lo = rbox[2 * r], hi = rbox[2 * r + 1];
                const bool outside = L.pnear & ((L.px > hi.x) | (L.px < lo.x) | (L.py > hi.y) | (L.py < lo.y) | (L.pz > hi.z) | (L.pz < lo.z));
                if (outside) alive = 0;
                if (!warp_any(alive != 0)) continue;
            }
#endif
#ifndef W_NO_FUSED_REJECT
            if (CULL) {                                                    // fused reject first, the exact test for what it keeps
                for (int i = start; i < end; ++i) {
                    const f4 g = geom[i];
                    const int m = L.pnear ? w_shadow_sphere_keep(L, g, alive, reject_k) : alive;
                    if (warp_any(m != 0)) w_shadow_sphere<COUNT>(L, g, m, has);
                }
            } else
#endif
            for (int i = start; i < end; ++i) w_shadow_sphere<COUNT>(L, geom[i], alive, has);
#ifndef W_NO_PLANE_CULL
        } else if (CULL) {
            for (int i = start; i < end; ++i) {
                const f4 g = geom[i];
                const f2 cu = pcull[i];
                const bool clear = L.pnear & (f_mul(cu.x, w_fused_plane_side(g, L.px, L.py, L.pz)) > cu.y);
                const int a = clear ? 0 : alive;
                if (warp_any(a != 0)) w_shadow_plane<COUNT>(L, g, a, has);
            }
#endif
        } else for (int i = start; i < end; ++i) w_shadow_plane<COUNT>(L, geom[i], alive, has);
    }
}

RT_HD void w_normal(const WFrame &F, int prim, float px, float py, float pz, float &nx, float &ny, float &nz) {   // RNO:162-177
    const f4 g = F.geom[prim];
    if (F.flags[prim] & W_FLAG_SPHERE) {
        const float rr = F.rrad[prim];
        nx = f_mul(f_sub(px, g.x), rr); ny = f_mul(f_sub(py, g.y), rr); nz = f_mul(f_sub(pz, g.z), rr);
    } else { nx = g.x; ny = g.y; nz = g.z; }
}

RT_HD void w_set_nearest_query(WLane &L, float ox, float oy, float oz) {
    L.qox = ox; L.qoy = oy; L.qoz = oz; L.qdx = L.dx; L.qdy = L.dy; L.qdz = L.dz;
    L.cumu = W_FAR; L.qhit = -1; L.qkind = 0; L.phase = PH_NEAREST;
}

// Primary ray of sub-sample L.sub: RNO:299-328.
RT_HD void w_start_subsample(WLane &L, const WFrame &F) {
    const int tx = L.sub / 3 - 1, ty = L.sub % 3 - 1;
    const float SY = f_add(2.25f, f_mul((float)L.y, F.DY));
    const float SX = f_add(-3.0f, f_mul((float)L.x, F.DX));
    float dx = f_sub(f_add(SX, f_mul(F.DX, f_mul((float)tx, 0.5f))), 0.f);
    float dy = f_sub(f_add(SY, f_mul(F.DY, f_mul((float)ty, 0.5f))), 0.25f);
    float dz = f_sub(0.f, -7.0f);
    const float len = f_rcp(f_sqrt(f_add(f_add(f_mul(dx, dx), f_mul(dy, dy)), f_mul(dz, dz))));
    L.dx = f_mul(dx, len); L.dy = f_mul(dy, len); L.dz = f_mul(dz, len);
    L.weight = 1.0f; L.depth = 0; L.from = -1; L.kind = W_PRIMARY; L.r_index = 1.0f;
    L.tr = L.tg = L.tb = 1.0f;
    L.head = L.tail = 0;
    w_set_nearest_query(L, 0.f, 0.25f, -7.0f);
}

RT_HD void w_begin_pixel(WLane &L, const WFrame &F, int x, int y) {
    L.x = x; L.y = y; L.sub = 0; L.ar = L.ag = L.ab = 0.f;
    w_start_subsample(L, F);
}

// FIFO record: 12 words = three f4.
RT_HD void w_push(f4 *q, WLane &L, float ox, float oy, float oz, float dx, float dy, float dz,
                  float weight, float r_index, float tr, float tg, float tb, int depth, int kind, int from) {
    RT_CHECK(L.tail - L.head < W_QUEUE_CHECK_SLOTS, RT_CHK_FIFO);      // a depth-5 binary ray tree never queues more
    f4 *slot = q + 3 * (L.tail & (W_QUEUE_SLOTS - 1));
    L.tail++;
    f4 a = { ox, oy, oz, dx }, b = { dy, dz, weight, r_index }, c = { tr, tg, tb, bits_f((uint32_t)(depth | (kind << 4) | ((from + 1) << 8))) };
    slot[0] = a; slot[1] = b; slot[2] = c;
}
RT_HD void w_pop(const f4 *q, WLane &L) {
    const f4 *slot = q + 3 * (L.head & (W_QUEUE_SLOTS - 1));
    L.head++;
    // Empty again: start over at slot 0.  94 % of the pushes find at most two rays waiting, but cursors that only count up touch a new
    // 48-byte slot per push -- ten slots per lane on a glass pixel, 68 MB over the 113 000 resident lanes, more than the L2 keeps:
    // 460 MB of DRAM writes per 1080p frame (ncu).  Reusing the first slots keeps the queue of all lanes in about 20 MB.
    if (L.head == L.tail) L.head = L.tail = 0;
    const f4 a = slot[0], b = slot[1], c = slot[2];
    L.dx = a.w; L.dy = b.x; L.dz = b.y; L.weight = b.z; L.r_index = b.w;
    L.tr = c.x; L.tg = c.y; L.tb = c.z;
    const uint32_t pk = f_bits(c.w);
    L.depth = (int)(pk & 15u); L.kind = (int)((pk >> 4) & 15u); L.from = (int)(pk >> 8) - 1;
    w_set_nearest_query(L, a.x, a.y, a.z);
}

// Diffuse + specular contribution of light `l` (RNO:242-276); Lx.. is the unit vector to the light.
// (ma, mb) = material of the hit primitive, (nx, ny, nz) its normal at the hit point, lc = the light's colour.
RT_HD void w_shade_with(WLane &L, const f4 ma, const f4 mb, const f4 lc, float nx, float ny, float nz, float Lx, float Ly, float Lz, float lit) {
    if (mb.x > 0.f) {
        const float nl = dot3(nx, ny, nz, Lx, Ly, Lz);
        if (nl > 0.f) {
            const float k = f_mul(f_mul(nl, mb.x), lit);
            L.cr = f_add(L.cr, f_mul(f_mul(k, ma.x), lc.x));
            L.cg = f_add(L.cg, f_mul(f_mul(k, ma.y), lc.y));
            L.cb = f_add(L.cb, f_mul(f_mul(k, ma.z), lc.z));
        }
    }
    if (mb.w > 0.f) {
        const float ln = dot3(Lx, Ly, Lz, nx, ny, nz);
        const float k2 = f_mul(2.0f, ln);
        const float rx = f_sub(Lx, f_mul(k2, nx)), ry = f_sub(Ly, f_mul(k2, ny)), rz = f_sub(Lz, f_mul(k2, nz));
        const float vr = dot3(L.dx, L.dy, L.dz, rx, ry, rz);
        if (vr > 0.f) {
            // pow(float,int) binds to the double overload in the reference's C++ build; the product with
            // m_spec and shade stays in double and is rounded to float once (RNO:270).
            const float k = (float)d_mul(d_mul(pow20_double(vr), (double)mb.w), (double)lit);
            L.cr = f_add(L.cr, f_mul(k, lc.x));
            L.cg = f_add(L.cg, f_mul(k, lc.y));
            L.cb = f_add(L.cb, f_mul(k, lc.z));
        }
    }
}
RT_HD void w_shade(WLane &L, const WFrame &F, int l, float Lx, float Ly, float Lz, float lit) {
    float nx, ny, nz;
    w_normal(F, L.hit, L.px, L.py, L.pz, nx, ny, nz);
    w_shade_with(L, F.mat_a[L.hit], F.mat_b[L.hit], F.mat_a[l], nx, ny, nz, Lx, Ly, Lz, lit);
}

// Blocked lights and the redo list.  A blocked light does not add "nothing": the reference multiplies both terms by shade = 0 (RNO:250,
// 270), and 0 times a factor that is not finite is NaN.  That happens -- a refraction direction is not re-normalised, a few bounces later
// |d| is 10^3 .. 10^7, pow(V.R, 20) overflows the double and the pixel's accumulator turns NaN, which x86's (int) makes a black pixel (found
// by tools/cull_fuzz.py on a random room; round 1 skipped blocked lights outright).  EXACT kernels and counting launches shade every
// blocked light with shade = 0 as the reference does.  The timed kernel skips a blocked light -- which is only right when every factor is
// PROVABLY finite, so that the products with 0 are +-0 and change no sum that started at +0: the scene's tables are bounded (host,
// w_scene_tame_reach), |d|^2 < W_TAME_D2 and every distance to a light of the batch lies inside (1e-18, F.tame_reach[kind of the hit]).  This is decided
// when a batch is set up, from values that are in registers there; a batch that fails REPORTS ITS PIXEL (w_report_untame) and carries on,
// and an EXACT launch over the reported pixels follows the timed kernel and overwrites them (a handful per frame in scenes that have
// any; more reports than the list holds: the EXACT launch redoes the frame).  Anything inline in the timed kernel -- keeping the distances
// alive into the shading code, a call to an out-of-line exact shader -- cost 5-12 % of the 1080p frame through register pressure alone.
#define W_TAME_D2 1e6f
RT_HD bool w_reach_is_tame(float reach, float cap) { return (reach > 1e-18f) & (reach < cap); }                     // NaN: false
RT_HD bool w_dir_is_tame(const WLane &L) { return dot3(L.dx, L.dy, L.dz, L.dx, L.dy, L.dz) < W_TAME_D2; }           // NaN: false
RT_HD void w_report_untame(const WLane &L, const WFrame &F) {
    if (!F.redo_count) return;
#ifdef __CUDA_ARCH__
    const unsigned at = atomicAdd(F.redo_count, 1u);
#else
    const unsigned at = (*F.redo_count)++;
#endif
    if (at < F.redo_cap) F.redo_list[at] = (uint32_t)L.y * (uint32_t)F.w + (uint32_t)L.x;
}

// Sets up the next batch of shadow rays (lights li, li+1, ... in index order, RNO:206-241), or marks the
// ray complete when no light is left.  A light that is not a sphere casts no shadow ray (RNO:223) and is
// shaded on the spot -- but only when no batch is pending before it, so that the float accumulation keeps
// the reference's light order.
RT_HD void w_light_vector(const WFrame &F, const WLane &L, int l, float &Lx, float &Ly, float &Lz, float &reach) {
    const f4 lg = F.geom[l];
    const float ex = f_sub(lg.x, L.px), ey = f_sub(lg.y, L.py), ez = f_sub(lg.z, L.pz);
    reach = f_sqrt(f_add(f_add(f_mul(ex, ex), f_mul(ey, ey)), f_mul(ez, ez)));
    const float inv = f_rcp(reach);
    Lx = f_mul(inv, ex); Ly = f_mul(inv, ey); Lz = f_mul(inv, ez);
}
// A light that is not a sphere (none in the reference's scenes): shaded without a shadow ray.
RT_HD void w_shade_unshadowed(WLane &L, const WFrame &F, int l, int li) {
    const f4 lg = F.lcenter[li];
    const float ex = f_sub(lg.x, L.px), ey = f_sub(lg.y, L.py), ez = f_sub(lg.z, L.pz);
    const float inv = f_rcp(f_sqrt(f_add(f_add(f_mul(ex, ex), f_mul(ey, ey)), f_mul(ez, ez))));
    w_shade(L, F, l, f_mul(inv, ex), f_mul(inv, ey), f_mul(inv, ez), 1.0f);
}
template <bool EXACT>
RT_HD void w_next_shadow_batch(WLane &L, const WFrame &F) {
    for (;;) {
        if (L.li >= F.n_lights) { L.phase = PH_FINAL; return; }
        const int l = F.lights[L.li];
        if (F.flags[l] & W_FLAG_SPHERE) break;
        w_shade_unshadowed(L, F, l, L.li);
        L.li++;
    }
    L.ns = 0; L.sblk = 0;
    bool tame = w_dir_is_tame(L);
    const float cap = (F.flags[L.hit] & W_FLAG_SPHERE) ? F.tame_reach[1] : F.tame_reach[0];
#pragma unroll
    for (int k = 0; k < W_SHADOW_BATCH; k++) {
        if (L.ns == k && L.li + k < F.n_lights && (F.flags[F.lights[L.li + k]] & W_FLAG_SPHERE)) {
            float Lx, Ly, Lz, reach;
            w_light_vector(F, L, F.lights[L.li + k], Lx, Ly, Lz, reach);
            L.sox[k] = f_add(L.px, f_mul(Lx, W_EPS)); L.soy[k] = f_add(L.py, f_mul(Ly, W_EPS)); L.soz[k] = f_add(L.pz, f_mul(Lz, W_EPS));
            L.slx[k] = Lx; L.sly[k] = Ly; L.slz[k] = Lz; L.sreach[k] = reach;
            tame &= w_reach_is_tame(reach, cap);
            L.ns = k + 1;
        }
    }
    if (!EXACT && !tame) w_report_untame(L, F);
    L.phase = PH_SHADOW;
}

// The common case as straight-line code: exactly NL (<= W_SHADOW_BATCH) lights, all of them spheres (the reference's
// scenes: 3), so the one batch holds lights[0 .. NL-1] and no selects, cursors or type checks are needed.  Same
// operations in the same order as the general functions (NL = 0 selects those).
template <int NL, bool EXACT>
RT_HD void w_shadow_batch_fixed(WLane &L, const WFrame &F) {
    L.ns = NL; L.sblk = 0;
    bool tame = w_dir_is_tame(L);
    const float cap = (F.flags[L.hit] & W_FLAG_SPHERE) ? F.tame_reach[1] : F.tame_reach[0];
#pragma unroll
    for (int k = 0; k < NL; k++) {
        float Lx, Ly, Lz, reach;
        w_light_vector(F, L, F.lights[k], Lx, Ly, Lz, reach);
        L.sox[k] = f_add(L.px, f_mul(Lx, W_EPS)); L.soy[k] = f_add(L.py, f_mul(Ly, W_EPS)); L.soz[k] = f_add(L.pz, f_mul(Lz, W_EPS));
        L.slx[k] = Lx; L.sly[k] = Ly; L.slz[k] = Lz; L.sreach[k] = reach;
        tame &= w_reach_is_tame(reach, cap);
    }
    if (!EXACT && !tame) w_report_untame(L, F);
    L.phase = PH_SHADOW;
}

// After the nearest-hit round (RNO:194-205).
template <bool COUNT, int NL = 0, bool EXACT = false>
RT_HD void w_after_nearest(WLane &L, const WFrame &F) {
    if (COUNT) { L.c_nearest++; L.c_sphere_tests += (uint32_t)F.n_spheres; L.c_plane_tests += (uint32_t)F.n_planes; }
    L.dist = L.cumu; L.hit = L.qhit; L.hkind = L.qkind;
    RT_CHECK(L.hit >= -1 && L.hit < F.n, RT_CHK_SCENE_INDEX);
    L.cr = L.cg = L.cb = 0.f;
    L.phase = PH_FINAL;
    if (L.hit >= 0) {
        if (F.flags[L.hit] & W_FLAG_LIGHT) {                           // RNO:197-200
            const f4 ma = F.mat_a[L.hit];
            L.cr = ma.x; L.cg = ma.y; L.cb = ma.z;
            // The reference leaves point_intersect unwritten here, and its caller still spawns children from it when
            // the light's material reflects or refracts (RNO:370-431) -- an uninitialised stack variable.  None of its
            // scenes has such a light; the oracle and this kernel define the point as (0,0,0) in that case.
            L.px = L.py = L.pz = 0.f;
        } else {
            L.px = f_add(L.qox, f_mul(L.qdx, L.dist));
            L.py = f_add(L.qoy, f_mul(L.qdy, L.dist));
            L.pz = f_add(L.qoz, f_mul(L.qdz, L.dist));
            L.li = 0;
            L.pnear = dot3(L.px, L.py, L.pz, L.px, L.py, L.pz) < F.cull_rp2;      // false for NaN / inf and when there are no cull tables (rp2 = 0)
#ifndef W_NO_UNLIT_SKIP
            // A material with neither a diffuse nor a specular term (m_diff <= 0 and m_spec <= 0: glass, mirrors) gathers
            // exactly nothing from any light whatever the shadow rays say (RNO:242-276 skips both terms): no shadow rays.
            // Counting launches still trace them, so that the ray and test counters equal the reference's.
#ifdef W_ROUND_STATS
            { const f4 mb = F.mat_b[L.hit]; if (!(mb.x > 0.f) & !(mb.w > 0.f)) return; }
#else
            if (!COUNT) { const f4 mb = F.mat_b[L.hit]; if (!(mb.x > 0.f) & !(mb.w > 0.f)) return; }
#endif
#endif
            if (NL > 0) w_shadow_batch_fixed<NL, (EXACT || COUNT)>(L, F);
            else w_next_shadow_batch<(EXACT || COUNT)>(L, F);
        }
    }
}

// After a shadow round: shade the batch's lights in index order, then set up the next batch or complete the ray.  A blocked light:
// shaded with shade = 0 in EXACT kernels and counting launches, skipped otherwise (see "Blocked lights and the redo list").
template <bool COUNT, int NL = 0, bool EXACT = false>
RT_HD void w_after_shadow(WLane &L, const WFrame &F) {
    constexpr bool AS_REFERENCE = EXACT || COUNT;
    if (COUNT) {
        L.c_shadow += (uint32_t)L.ns;
        const f4 mbc = F.mat_b[L.hit];
        if ((mbc.x > 0.f) | (mbc.w > 0.f)) L.c_shadow_lit += (uint32_t)L.ns;
    }
    if (NL > 0) {
        const f4 ma = F.mat_a[L.hit], mb = F.mat_b[L.hit];                // once per hit, not per light
        float nx, ny, nz;
        w_normal(F, L.hit, L.px, L.py, L.pz, nx, ny, nz);
#ifdef W_UNROLL_SHADE
#pragma unroll
#else
#pragma unroll 1
#endif
        for (int k = 0; k < NL; k++) {        // one copy of the shading code; k-th ray picked with selects
            const float Lx = k == 0 ? L.slx[0] : (k == 1 ? L.slx[NL > 1 ? 1 : 0] : L.slx[NL > 2 ? 2 : 0]);
            const float Ly = k == 0 ? L.sly[0] : (k == 1 ? L.sly[NL > 1 ? 1 : 0] : L.sly[NL > 2 ? 2 : 0]);
            const float Lz = k == 0 ? L.slz[0] : (k == 1 ? L.slz[NL > 1 ? 1 : 0] : L.slz[NL > 2 ? 2 : 0]);
            const bool blocked = (L.sblk >> k) & 1;
            if (AS_REFERENCE) w_shade_with(L, ma, mb, F.mat_a[F.lights[k]], nx, ny, nz, Lx, Ly, Lz, blocked ? 0.0f : 1.0f);
            else if (!blocked) w_shade_with(L, ma, mb, F.mat_a[F.lights[k]], nx, ny, nz, Lx, Ly, Lz, 1.0f);
        }
        L.phase = PH_FINAL;
        return;
    }
#pragma unroll 1
    for (int k = 0; k < L.ns; k++) {          // a real loop (one copy of the shading code); the rays are picked with selects
        const float Lx = k == 0 ? L.slx[0] : (k == 1 ? L.slx[1] : L.slx[2]);
        const float Ly = k == 0 ? L.sly[0] : (k == 1 ? L.sly[1] : L.sly[2]);
        const float Lz = k == 0 ? L.slz[0] : (k == 1 ? L.slz[1] : L.slz[2]);
        const bool blocked = (L.sblk >> k) & 1;
        if (AS_REFERENCE) w_shade(L, F, F.lights[L.li + k], Lx, Ly, Lz, blocked ? 0.0f : 1.0f);
        else if (!blocked) w_shade(L, F, F.lights[L.li + k], Lx, Ly, Lz, 1.0f);
    }
    L.li += L.ns;
    w_next_shadow_batch<AS_REFERENCE>(L, F);
}

// The ray is complete: fold its colour into the pixel (RNO:351-368), spawn its children (RNO:370-432) and move
// on to the next ray of the FIFO, the next sub-sample, or the end of the pixel (returns true).
// SPLIT (whitted_split_kernel: one lane per SUB-SAMPLE of a pixel): the pixel's accumulator runs through all rays of all nine sub-samples
// in turn (RNO:351-368), so a lane that traces sub-sample s cannot add to it -- it appends what each of its rays would have added to
// `log` (at most 63 triples: a depth-5 binary ray tree), L.tail of the NEXT free slot kept in L.nlog, and the nine logs are added in the
// reference's order afterwards.  A triple of zeros is not logged: the accumulator starts at +0 and x + (+-0) = x for every x that is
// not -0, which a sum that started at +0 never is.  The sub-sample ends the lane's work (returns true).
template <bool COUNT, bool SPLIT = false>
RT_HD bool w_finalize(WLane &L, const WFrame &F, f4 *q, float *log = nullptr) {
    // The ray is finished: fold its colour into the pixel (RNO:351-368).
    float ar, ag, ab;
    if (L.kind == W_PRIMARY) {
        if (COUNT) L.c_samples++;
        RT_CHECK(L.x >= 0 && L.x < F.w && L.y >= 0 && L.y < F.h && L.sub >= 0 && L.sub < 9, RT_CHK_PIXEL);
        if (F.hit_ids) F.hit_ids[((size_t)L.y * F.w + L.x) * 9 + L.sub] = L.hit;
        ar = f_mul(L.cr, L.weight); ag = f_mul(L.cg, L.weight); ab = f_mul(L.cb, L.weight);
    } else if (L.kind == W_REFLECTED) {
        const f4 fa = F.mat_a[L.from];
        ar = f_mul(f_mul(f_mul(L.cr, L.weight), fa.x), L.tr);
        ag = f_mul(f_mul(f_mul(L.cg, L.weight), fa.y), L.tg);
        ab = f_mul(f_mul(f_mul(L.cb, L.weight), fa.z), L.tb);
    } else {
        ar = f_mul(f_mul(L.cr, L.weight), L.tr);
        ag = f_mul(f_mul(L.cg, L.weight), L.tg);
        ab = f_mul(f_mul(L.cb, L.weight), L.tb);
    }
    if (SPLIT) {
        if (!((ar == 0.f) & (ag == 0.f) & (ab == 0.f))) {          // NaN: logged
            RT_CHECK(L.nlog < 63, RT_CHK_FIFO);
            log[3 * L.nlog] = ar; log[3 * L.nlog + 1] = ag; log[3 * L.nlog + 2] = ab;
            L.nlog++;
        }
    } else { L.ar = f_add(L.ar, ar); L.ag = f_add(L.ag, ag); L.ab = f_add(L.ab, ab); }
    // Children (RNO:370-432).  A miss spawns nothing (the reference reads prims[-1] there; its shipped
    // scenes are closed boxes, so it never happens -- SURVEY.md 2.3).
    if (L.hit >= 0 && L.depth < W_TRACEDEPTH) {
        const f4 ma = F.mat_a[L.hit], mb = F.mat_b[L.hit];
        if (ma.w > 0.0f) {                                             // reflection
            float nx, ny, nz;
            w_normal(F, L.hit, L.px, L.py, L.pz, nx, ny, nz);
            const float k2 = f_mul(2.0f, dot3(L.dx, L.dy, L.dz, nx, ny, nz));
            const float rx = f_sub(L.dx, f_mul(k2, nx)), ry = f_sub(L.dy, f_mul(k2, ny)), rz = f_sub(L.dz, f_mul(k2, nz));
            w_push(q, L, f_add(L.px, f_mul(rx, W_EPS)), f_add(L.py, f_mul(ry, W_EPS)), f_add(L.pz, f_mul(rz, W_EPS)), rx, ry, rz,
                   f_mul(ma.w, L.weight), L.r_index, L.tr, L.tg, L.tb, L.depth + 1, W_REFLECTED, L.hit);
        }
        if (mb.y > 0.0f) {                                             // refraction
            const float m_rindex = mb.z;
            const float nn = f_div(L.r_index, m_rindex);
            float gx, gy, gz;
            w_normal(F, L.hit, L.px, L.py, L.pz, gx, gy, gz);
            const float sgn = (float)L.hkind;
            const float nx = f_mul(gx, sgn), ny = f_mul(gy, sgn), nz = f_mul(gz, sgn);
            const float cosI = -dot3(nx, ny, nz, L.dx, L.dy, L.dz);
            const float cosT2 = f_sub(1.0f, f_mul(f_mul(nn, nn), f_sub(1.0f, f_mul(cosI, cosI))));
            if (cosT2 > 0.0f) {
                const float kk = f_sub(f_mul(nn, cosI), f_sqrt(cosT2));
                const float tx = f_add(f_mul(nn, L.dx), f_mul(kk, nx));
                const float ty = f_add(f_mul(nn, L.dy), f_mul(kk, ny));
                const float tz = f_add(f_mul(nn, L.dz), f_mul(kk, nz));
                const float nd = -L.dist;
                w_push(q, L, f_add(L.px, f_mul(tx, W_EPS)), f_add(L.py, f_mul(ty, W_EPS)), f_add(L.pz, f_mul(tz, W_EPS)), tx, ty, tz,
                       L.weight, m_rindex,
                       f_mul(L.tr, expf_glibc(f_mul(f_mul(ma.x, 0.15f), nd))),     // Beer's law, RNO:424-426
                       f_mul(L.tg, expf_glibc(f_mul(f_mul(ma.y, 0.15f), nd))),
                       f_mul(L.tb, expf_glibc(f_mul(f_mul(ma.z, 0.15f), nd))),
                       L.depth + 1, W_REFRACTED, L.hit);
            }
        }
    }
    if (L.head < L.tail) { w_pop(q, L); return false; }
    if (!SPLIT) {
        L.sub++;
        if (L.sub < 9) { w_start_subsample(L, F); return false; }
    }
    L.phase = PH_IDLE;
    return true;
}

// RNO:436-447: min(255, (int)(acc * (256/9))), alpha 0.
RT_HD uint32_t w_pack_pixel(float r, float g, float b) {
    // The reference is x86 code: (int) of a float outside the int range (its tracer does blow up to ~1e13 on a
    // few pixels inside the front glass sphere) or of a NaN is cvttss2si's "integer indefinite" 0x80000000,
    // which then fails `> 255` and truncates to byte 0.  CUDA's cvt would saturate to 255 instead.
    int ir = x86_float_to_int(f_mul(r, 28.0f)), ig = x86_float_to_int(f_mul(g, 28.0f)), ib = x86_float_to_int(f_mul(b, 28.0f));
    if (ir > 255) ir = 255;
    if (ig > 255) ig = 255;
    if (ib > 255) ib = 255;
    return (uint32_t)(ir & 255) | ((uint32_t)(ig & 255) << 8) | ((uint32_t)(ib & 255) << 16);
}

}  // namespace rtb
