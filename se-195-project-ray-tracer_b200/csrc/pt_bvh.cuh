// pt_bvh.cuh -- exact culling for the sphere queries of the path tracer on large scenes (host/device).
//
// The reference tests every sphere against every ray (Intersect / IntersectP, SPT/geomfunc.h:71-110).  For the generated
// many-sphere scenes (SPT/scene_build_complex.pl: 783 / 3 908 / 19 533 spheres) almost all of those tests are misses that a
// bounding-volume hierarchy can rule out -- provided the hierarchy never rules out a test whose FLOAT result the
// reference would have accepted.  This file is that hierarchy's traversal.  It changes which tests are executed, never
// the arithmetic of a test (pt_bvh_sphere is SphereIntersect, operation for operation) nor the winner:
//
//   nearest query   result = the smallest non-zero d_i; among equal d_i the HIGHEST index (the reference scans
//                   i = n-1 .. 0 with a strict '<', SPT/geomfunc.h:80-88).  Any visiting order gives that result with
//                   the update rule "d < t, or d == t and i > id", as long as every sphere with d_i <= final t is visited.
//   shadow query    result = "some d_i != 0 is < maxt" (SPT/geomfunc.h:102-109): order-free.
//
// Conservative node test.  u = 2^-24.  For a sphere (centre c, float rad2 = fl(rad*rad)) the reference computes
//   op = fl(c - o), b = fl(op.d), det = fl(fl(fl(b*b) - fl(op.op)) + rad2), and returns 0 when det < 0.
// With B = op.d, OO = op.op as real numbers and |d|^2 = 1 + eps, standard rounding bounds give
//   |det - (B*B - OO + rad2)| <= 12u * OO * max(1, |d|^2) + u * rad2,      B*B - OO = eps*OO - (1+eps)*rho^2,
// rho = distance from c to the line through o' = c - op (|o' - o| <= u*|op| per component) along d.  Hence det >= 0 needs
//   rho^2 <= rad2 + eta,   eta = (|eps| + 13u)(1 + 2|eps|) * OO + (2|eps| + 2u) * rad2:
// the line passes through the sphere inflated to R' = sqrt(rad2 + eta), and R' - rad <= min(eta / (2*rad), sqrt(eta)).
// (The second bound matters far from the scene, where eta exceeds rad2 -- the reference's own det is rounding noise
// there, which the hierarchy has to reproduce -- and the first one would swallow the whole tree.)  The accepted distance
// is fl(b -+ sqrt(det)), which lies within (|eps| + 10u) * |op| of the parametric entry / exit of that inflated sphere.
//
// A node stores the box of its spheres and hinv = 0.5 / (smallest radius below it); the scene stores the largest radius
// r_max of the tree.  Per ray, D_k = the largest |root box corner - o| per axis bounds |op_k| of every sphere of the tree
// (every node's box lies inside the root's), so with
//   K1 = 2*e + 34u, K2 = 4*e + 24u  (e = |fl(d.d) - 1| >= |eps| - 4u: TWICE the bounds above),
//   eta = K1 * (Dx^2+Dy^2+Dz^2) + K2 * r_max^2,
//   m  = min(eta * hinv, 1.001 * sqrt(eta)) + 1e-6 * (1 + Dx+Dy+Dz)     (the last term covers o' - o and the slab roundings)
// every sphere below a node that could return d != 0 has its inflated sphere inside the node's box grown by m; the slab
// test on the grown box yields [te, tx], and no accepted distance below the node is smaller than te - kT*(Dx+Dy+Dz) or
// larger than tx + kT*(Dx+Dy+Dz), kT = 2*e + 40u.  A node is skipped only if the grown box is missed, lies behind the
// origin or lies beyond the current limit by those margins; every comparison is written so that a NaN (0 * inf on a slab
// face) means "visit".  The bounds hold for any |eps|: a direction that is not exactly of unit length -- the Whitted
// tracer's reflections off computed sphere normals are off by ~1e-4 -- just sees slightly larger spheres, which is what
// the reference's unit-length formula does with it; beyond |eps| = 2^-7 the ray gets K1 = K2 = kT = inf and visits
// everything.  Spheres that are much larger than the rest (the 10 000-unit floor) or not finite are not in the tree at
// all: they form a short list that every query tests first.
//
// tests/: the traversal against the plain loop on the lane simulator (CPU), against the brute-force kernel on the GPU at
// full size, and tools/bvh_fuzz.py (thousands of random scenes; a build without the margin is caught) -- colours, RNG
// state and pixels bit-identical.
#pragma once
#include "pt_lane.cuh"

namespace rtb {

#define PT_BVH_STACK 64
#ifndef PT_BVH_LEAF_MAX
#define PT_BVH_LEAF_MAX 8                        /* 8 against 4: 24 ms against 29 ms on the 19 533-sphere scene at 4K (a shallower tree), equal on the smaller ones */
#endif
#define PT_BVH_NONE 0x7fffffff
#define PT_BVH_U 5.9604644775390625e-8f          /* 2^-24 */
#define PT_BVH_MAX_EPS 0.0078125f                 /* 2^-7: directions further from unit length than this get no culling */

// Inner node i = nodes[4i .. 4i+3]:
//   [0] = (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y)   [1] = (c1.lo.x, c1.hi.x, c1.lo.y, c1.hi.y)
//   [2] = (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z)   [3] = (bits child0, bits child1, hinv0, hinv1)
// child >= 0: inner node index; child < 0: leaf ~((first << 3) | (count - 1)) over geom[] / index[].
struct PtBvh {
    const f4 *nodes;
    const f4 *geom;          // (p, rad^2) in leaf order; entries [0, n_big) are the always-tested spheres
    const int *index;        // original sphere index of each entry
    int n_big;
    int root;                // child code of the root, PT_BVH_NONE when every sphere is in the always-tested list
    float root_hinv;
    float root_lo[3], root_hi[3];
    float rmax2;             // (largest radius in the tree)^2, rounded up
};

RT_HD int pt_bvh_leaf_code(int first, int count) { return ~((first << 3) | (count - 1)); }

// SphereIntersect (SPT/geomfunc.h:32-59) + the acceptance of Intersect / IntersectP, for ONE lane (no warp votes: the
// traversal is per lane).  Same operations in the same order as pt_test.
template <bool COUNT>
RT_HD void pt_bvh_sphere(PtLane &L, const f4 g, int idx) {
    if (COUNT) L.c_tests++;                      // tests actually executed (the reference's count is n per query)
    const float opx = f_sub(g.x, L.ox), opy = f_sub(g.y, L.oy), opz = f_sub(g.z, L.oz);
    const float b = dot3(opx, opy, opz, L.dx, L.dy, L.dz);
    const float det = f_add(f_sub(f_mul(b, b), dot3(opx, opy, opz, opx, opy, opz)), g.w);
    if (det < 0.f) return;
    const float sq = f_sqrt(det);
    const float t1 = f_sub(b, sq), t2 = f_add(b, sq);
    const float t = t1 > PT_EPS ? t1 : t2;
    if (!(t > PT_EPS)) return;
    if (L.phase == PH_SHADOW) { if (t < L.cumu) L.hit = idx; }
    else if (t < L.cumu || (t == L.cumu && idx > L.hit)) { L.cumu = t; L.hit = idx; }
}

// Per-ray constants of the node test: reciprocal direction, eta, 1.001*sqrt(eta), the additive term of m, the slack.
struct PtBvhRay { float ix, iy, iz, eta, seta, c, slack; };

// 1/x for the slab test: two units in the last place are enough (the slacks above allow for it), so the device uses
// the fast reciprocal; a denormal component flushes to "parallel to the slab" like a zero one.
RT_HD float bvh_rcp(float x) {
#ifdef __CUDA_ARCH__
    return __fdividef(1.f, x);
#else
    return 1.f / x;
#endif
}

RT_HD PtBvhRay bvh_ray(float ox, float oy, float oz, float dx, float dy, float dz, const PtBvh &B) {
    PtBvhRay R;
    const float dd = dot3(dx, dy, dz, dx, dy, dz);
    const float e = fabsf(f_sub(dd, 1.f));
    float K1, K2, kT;
    if (e <= PT_BVH_MAX_EPS) { K1 = 2.f * e + 34.f * PT_BVH_U; K2 = 4.f * e + 24.f * PT_BVH_U; kT = 2.f * e + 40.f * PT_BVH_U; }
    else K1 = K2 = kT = INFINITY;                  // far from a unit direction (or NaN): no culling
    R.ix = bvh_rcp(dx); R.iy = bvh_rcp(dy); R.iz = bvh_rcp(dz);
    const float Dx = fmaxf(fabsf(B.root_lo[0] - ox), fabsf(B.root_hi[0] - ox));
    const float Dy = fmaxf(fabsf(B.root_lo[1] - oy), fabsf(B.root_hi[1] - oy));
    const float Dz = fmaxf(fabsf(B.root_lo[2] - oz), fabsf(B.root_hi[2] - oz));
    const float D1 = Dx + Dy + Dz;
    R.eta = K1 * (Dx * Dx + Dy * Dy + Dz * Dz) + K2 * B.rmax2;
    R.seta = 1.001f * sqrtf(R.eta);
    R.c = 1e-6f * (1.f + D1);
    R.slack = kT * D1;
    return R;
}

// Conservative "may some sphere below this box return an accepted distance below `limit`" + a lower bound of such distances.
RT_HD bool bvh_box(float ox, float oy, float oz, float limit, const PtBvhRay &R, float lox, float hix, float loy, float hiy, float loz, float hiz,
                   float hinv, float &lb_out) {
    const float a0x = lox - ox, a1x = hix - ox, a0y = loy - oy, a1y = hiy - oy, a0z = loz - oz, a1z = hiz - oz;
#ifdef PT_BVH_TEST_NO_MARGIN          /* tools/bvh_fuzz.py self-check: a hierarchy WITHOUT the rounding margins must be caught */
    const float m = 0.f;
#else
    const float m = fminf(R.eta * hinv, R.seta) + R.c;
#endif
    const float t0x = (a0x - m) * R.ix, t1x = (a1x + m) * R.ix;
    const float t0y = (a0y - m) * R.iy, t1y = (a1y + m) * R.iy;
    const float t0z = (a0z - m) * R.iz, t1z = (a1z + m) * R.iz;
    const float te = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fminf(t0z, t1z));
    const float tx = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
    lb_out = te - R.slack;                 // no accepted distance below this box is smaller (NaN: compares false = keep)
    const bool missed = (te - tx) > 1e-6f * (fabsf(te) + fabsf(tx));
    const bool behind = (tx + R.slack) < 0.f;
    const bool beyond = (te - R.slack) > limit;
    return !(missed | behind | beyond);
}

// A query in flight: the ray's constants, the node to visit next and the stack depth.  The stacks themselves are two
// caller-owned arrays of PT_BVH_STACK entries (local memory on the device).
#define PT_BVH_DONE 0x7ffffffe
struct PtTrav { PtBvhRay R; int node, sp; };

RT_HD int bvh_root(float ox, float oy, float oz, float limit, const PtBvh &B, const PtBvhRay &R);
// Starts the lane's query (nearest or shadow, by L.phase): the always-tested spheres, then the root box.
template <bool COUNT>
RT_HD void pt_bvh_begin(PtLane &L, const PtBvh &B, PtTrav &T) {
    T.node = PT_BVH_DONE; T.sp = 0;
    const bool shadow = L.phase == PH_SHADOW;
    for (int j = 0; j < B.n_big; j++) {
        pt_bvh_sphere<COUNT>(L, B.geom[j], B.index[j]);
        if (shadow && L.hit >= 0) return;
    }
    if (B.root == PT_BVH_NONE) return;
    T.R = bvh_ray(L.ox, L.oy, L.oz, L.dx, L.dy, L.dz, B);
    T.node = bvh_root(L.ox, L.oy, L.oz, L.cumu, B, T.R);
}

#ifdef PT_BVH_STATS          /* test-only visit counters of the host build (tests/devsim) */
static long g_bvh_inner_visits = 0, g_bvh_leaf_visits = 0, g_bvh_hist[64] = {0}, g_bvh_tail_visits = 0, g_bvh_max_visits = 0;
static float g_bvh_max_ray[7] = {0};
#define PT_BVH_STAT(x) ((x)++)
#else
#define PT_BVH_STAT(x) ((void)0)
#endif
RT_HD bool pt_bvh_at_inner(const PtTrav &T) { return (unsigned)T.node < (unsigned)PT_BVH_DONE; }
RT_HD bool pt_bvh_at_leaf(const PtTrav &T) { return T.node < 0; }

// Next node from the stack; a pushed child that is now beyond the limit is dropped.
RT_HD void bvh_pop(float limit, PtTrav &T, const int *stack, const float *stack_t) {
    for (;;) {
        if (T.sp == 0) { T.node = PT_BVH_DONE; return; }
        --T.sp;
        if (!(stack_t[T.sp] > limit)) break;
    }
    T.node = stack[T.sp];
}

// The root box: the node to start at, or PT_BVH_DONE.
RT_HD int bvh_root(float ox, float oy, float oz, float limit, const PtBvh &B, const PtBvhRay &R) {
    float lb;
    return bvh_box(ox, oy, oz, limit, R, B.root_lo[0], B.root_hi[0], B.root_lo[1], B.root_hi[1], B.root_lo[2], B.root_hi[2], B.root_hinv, lb) ? B.root : PT_BVH_DONE;
}

// One inner node: two box tests, descend into the nearer child that may matter, push the other.
RT_HD void bvh_inner(float ox, float oy, float oz, float limit, const PtBvh &B, PtTrav &T, int *stack, float *stack_t) {
    PT_BVH_STAT(g_bvh_inner_visits);
    const int node = T.node;
    const f4 n0 = B.nodes[4 * node], n1 = B.nodes[4 * node + 1], n2 = B.nodes[4 * node + 2], n3 = B.nodes[4 * node + 3];
    float lb0, lb1;
    const bool h0 = bvh_box(ox, oy, oz, limit, T.R, n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, n3.z, lb0);
    const bool h1 = bvh_box(ox, oy, oz, limit, T.R, n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, n3.w, lb1);
    const int c0 = (int)f_bits(n3.x), c1 = (int)f_bits(n3.y);
    if (h0 & h1) {
        const bool swap = lb1 < lb0;                     // nearer child first (any order is correct)
        RT_CHECK(T.sp >= 0 && T.sp < PT_BVH_STACK, RT_CHK_STACK);
        stack[T.sp] = swap ? c0 : c1;
        stack_t[T.sp++] = swap ? lb0 : lb1;
        T.node = swap ? c1 : c0;
    } else if (h0 | h1) T.node = h0 ? c0 : c1;
    else bvh_pop(limit, T, stack, stack_t);
}
RT_HD void pt_bvh_inner(const PtLane &L, const PtBvh &B, PtTrav &T, int *stack, float *stack_t) { bvh_inner(L.ox, L.oy, L.oz, L.cumu, B, T, stack, stack_t); }

// One leaf: its spheres (exact tests), then the next node from the stack.
template <bool COUNT>
RT_HD void pt_bvh_leaf(PtLane &L, const PtBvh &B, PtTrav &T, const int *stack, const float *stack_t) {
    PT_BVH_STAT(g_bvh_leaf_visits);
    const int code = ~T.node;
    const int first = code >> 3, count = (code & 7) + 1;
#pragma unroll 1                                         /* rolled: 8 % faster than eight inlined copies of the test (8 KB less code) */
    for (int j = 0; j < count; j++) pt_bvh_sphere<COUNT>(L, B.geom[first + j], B.index[first + j]);
    if (L.phase == PH_SHADOW && L.hit >= 0) { T.node = PT_BVH_DONE; return; }
    bvh_pop(L.cumu, T, stack, stack_t);
}

// One whole query of one lane.  Replaces pt_query_range (tests/devsim; the kernel interleaves the steps of 32 lanes).
template <bool COUNT>
RT_HD void pt_query_bvh(PtLane &L, const PtBvh &B) {
    int stack[PT_BVH_STACK];
    float stack_t[PT_BVH_STACK];
    PtTrav T;
    pt_bvh_begin<COUNT>(L, B, T);
#ifdef PT_BVH_STATS
    const long v0 = g_bvh_inner_visits;
    struct Tally { long v0; const PtLane &L; ~Tally() { long n = g_bvh_inner_visits - v0; g_bvh_hist[n > 63 ? 63 : n]++; if (n >= 63) g_bvh_tail_visits += n;
        if (n > g_bvh_max_visits) { g_bvh_max_visits = n; g_bvh_max_ray[0] = L.ox; g_bvh_max_ray[1] = L.oy; g_bvh_max_ray[2] = L.oz; g_bvh_max_ray[3] = L.dx; g_bvh_max_ray[4] = L.dy; g_bvh_max_ray[5] = L.dz; g_bvh_max_ray[6] = (float)L.phase; } } } tally{v0, L};
#endif
    while (T.node != PT_BVH_DONE) {
        if (pt_bvh_at_inner(T)) pt_bvh_inner(L, B, T, stack, stack_t);
        else pt_bvh_leaf<COUNT>(L, B, T, stack, stack_t);
    }
}

}  // namespace rtb
