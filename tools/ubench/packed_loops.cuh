// packed_loops.cuh -- EXPERIMENT (not part of the product): the Whitted query loops with Blackwell's packed FP32
// instructions (FADD2 / FMUL2, PTX add.rn.f32x2 / mul.rn.ftz.f32x2), two primitives of a run per step.
//
// Outcome (profiles/r01_f32x2_experiment.md): bit-identical frames, but no win.  Every packed instruction holds
// the issue port for two cycles (FADD2, FMUL2 and FFMA2 alike; tools/ubench/f32x2_forms.cu), so un-fused packed
// arithmetic has the throughput of the scalar instructions; only an ALU instruction can slip in behind an FMUL2.  In
// isolation the packed loops are 20-30 % faster per test (tools/ubench/loop_bench.cu), inside the kernel the fillers of
// odd runs, the per-ray range votes and the larger code made the frame slower (6.2 ms against 4.8 ms).  Kept for the
// record, together with the ptxas 12.9 finding that mul.rn.f32x2 + add.rn.f32x2 is contracted into FFMA2 even with
// --fmad=false (a .ftz multiply next to a non-.ftz add is not).
#pragma once
#include "whitted_lane.cuh"

namespace rtb {
struct alignas(8) f2 { float x, y; };
struct alignas(16) f2x2 { f2 a, b; };             // one LDS.128: two packed pairs
#define W_RUN_STRIDE 4
#ifdef __CUDA_ARCH__
#define RT_PK(u, v) asm("mov.b64 %0, {%1, %2};" : "=l"(u) : "f"((v).x), "f"((v).y))
#define RT_UNPK(v, u) asm("mov.b64 {%0, %1}, %2;" : "=f"((v).x), "=f"((v).y) : "l"(u))
#endif

RT_HD f2 f2_make(float x, float y) { f2 r; r.x = x; r.y = y; return r; }
RT_HD f2 f2_bcast(float s) { f2 r; r.x = s; r.y = s; return r; }

RT_HD f2 f2_add(f2 a, f2 b) {
    f2 r;
#ifdef __CUDA_ARCH__
    unsigned long long A, B, C;
    RT_PK(A, a); RT_PK(B, b);
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(C) : "l"(A), "l"(B));
    RT_UNPK(r, C);
#else
    r.x = a.x + b.x; r.y = a.y + b.y;
#endif
    return r;
}
RT_HD f2 f2_sub(f2 a, f2 b) {
    f2 r;
#ifdef __CUDA_ARCH__
    unsigned long long A, B, C;
    RT_PK(A, a); RT_PK(B, b);
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(C) : "l"(A), "l"(B));
    RT_UNPK(r, C);
#else
    r.x = a.x - b.x; r.y = a.y - b.y;
#endif
    return r;
}
// Packed multiply, flush-to-zero form (rule 1 above).  The host build (tests/devsim) is plain IEEE.
RT_HD f2 f2_mul(f2 a, f2 b) {
    f2 r;
#ifdef __CUDA_ARCH__
    unsigned long long A, B, C;
    RT_PK(A, a); RT_PK(B, b);
    asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(C) : "l"(A), "l"(B));
    RT_UNPK(r, C);
#else
    r.x = a.x * b.x; r.y = a.y * b.y;
#endif
    return r;
}
// Pair-with-scalar forms.  The scalar is packed inside the same asm block as its use, which is what makes ptxas
// encode it as a broadcast operand (FMUL2 R, R.F32x2, R.F32) instead of keeping a duplicated register pair alive.
RT_HD f2 f2_mul_s(f2 a, float s) {
    f2 r;
#ifdef __CUDA_ARCH__
    unsigned long long A, C;
    RT_PK(A, a);
    asm("{\n\t.reg .b64 t;\n\tmov.b64 t, {%2, %2};\n\tmul.rn.ftz.f32x2 %0, %1, t;\n\t}" : "=l"(C) : "l"(A), "f"(s));
    RT_UNPK(r, C);
#else
    r.x = a.x * s; r.y = a.y * s;
#endif
    return r;
}
RT_HD f2 f2_rsub_s(float s, f2 a) {            // (s - a.x, s - a.y)
    f2 r;
#ifdef __CUDA_ARCH__
    unsigned long long A, C;
    RT_PK(A, a);
    asm("{\n\t.reg .b64 t;\n\tmov.b64 t, {%2, %2};\n\tsub.rn.f32x2 %0, t, %1;\n\t}" : "=l"(C) : "l"(A), "f"(s));
    RT_UNPK(r, C);
#else
    r.x = s - a.x; r.y = s - a.y;
#endif
    return r;
}
// (a.x*s.x + a.y*s.y) + a.z*s.z of two vectors a against one scalar vector s.
RT_HD f2 f2_dot3_s(f2 ax, f2 ay, f2 az, float sx, float sy, float sz) {
    return f2_add(f2_add(f2_mul_s(ax, sx), f2_mul_s(ay, sy)), f2_mul_s(az, sz));
}
// (a.x*b.x + a.y*b.y) + a.z*b.z for two vectors at once: 3 FMUL2 + 2 FADD2.
RT_HD f2 f2_dot3(f2 ax, f2 ay, f2 az, f2 bx, f2 by, f2 bz) {
    return f2_add(f2_add(f2_mul(ax, bx), f2_mul(ay, by)), f2_mul(az, bz));
}

// Rule 2: a coordinate may enter the packed loops when it is zero or at least 2^-40 in magnitude (NaN fails).
#define RT_PACKED_MIN 9.094947017729282379150390625e-13f      /* 2^-40 */
RT_HD bool packed_range_ok(float c) { return fabsf(c) >= RT_PACKED_MIN || c == 0.f; }
RT_HD bool packed_range_ok3(float x, float y, float z) { return packed_range_ok(x) && packed_range_ok(y) && packed_range_ok(z); }


// ---------------------------------------------------------------------------------------------------------
// Packed loops (f32x2.cuh): the same tests, two primitives of a run per step, every quantity a pair
// (prim i, prim i+1).  Per sphere pair 8 FMUL2 + 9 FADD2 replace 32 scalar operations; the comparisons
// and the rare exact stage (roots, division) stay scalar and are the very code of the scalar loops.
//
// Bit-identity with the scalar loops, given rule 2 of f32x2.cuh (geometry and ray coordinates are 0 or
// >= 2^-40 in magnitude, sq_radius >= 2^-40): FMUL2.FTZ differs from an IEEE multiply only if an operand is
// subnormal or the exact product is below 2^-126.
//   * v = o - c is an IEEE subtraction of two such numbers: 0, or at least one ulp of the larger = 2^-63.
//   * v_i*d_i >= 2^-103, v_i*v_i >= 2^-126, N_i*d_i and N_i*o_i >= 2^-80: normal, never flushed.  All sums are IEEE.
//   * bb = (v.d)^2 may be flushed (|v.d| can be as small as 2^-126).  It only feeds det = (bb - v.v) + sq_radius
//     and the filter bound bb*(1-2^-22).  A flushed bb is below 2^-126; if v.v >= 2^-102 it vanishes in the first
//     addition either way, otherwise |bb - v.v| < 2^-101 vanishes against sq_radius >= 2^-40: det is the same
//     float.  The bound is then 0 or below 2^-126 and `det < bound` is false either way.
//   * plane bounds (cumu*|d|)*(1+2^-21) may be flushed; both forms then fail the `> 1e-30` guard, so the test
//     falls through to the exact division.
// The filler of an odd run can never be hit: sq_radius = -inf makes det = -inf, a zero plane has d = 0.
RT_HD void w_sphere_pair(WLane &L, const f2x2 A, const f2x2 B, int i, int live_i) {
    const bool live = live_i != 0;
    const f2 vx = f2_rsub_s(L.qox, A.a), vy = f2_rsub_s(L.qoy, A.b), vz = f2_rsub_s(L.qoz, B.a);
    const f2 vd = f2_dot3_s(vx, vy, vz, L.qdx, L.qdy, L.qdz);                                    // b = -vd
    const f2 bb = f2_mul(vd, vd);
    const f2 det = f2_add(f2_sub(bb, f2_dot3(vx, vy, vz, vx, vy, vz)), B.b);
    const f2 lim = f2_mul_s(bb, 0.999999761581420898437500f);                                   // w_sphere_behind
    const bool c0 = (det.x > 0.f) & !((vd.x > 0.f) & (det.x < lim.x));
    const bool c1 = (det.y > 0.f) & !((vd.y > 0.f) & (det.y < lim.y));
    if (warp_any((c0 | c1) & live)) {
        const float dv[2] = { det.x, det.y };
        const bool need[2] = { c0, c1 };
        float sq[2];
        sqrt_group<2>(dv, need, sq);
        {
            const float b = -vd.x, i1 = f_sub(b, sq[0]), i2 = f_add(b, sq[0]);
            const bool inside = i1 < 0.f;
            const float t = inside ? i2 : i1;
            if (live & c0 & (i2 > 0.f) & (t < L.cumu)) { L.cumu = t; L.qhit = i; L.qkind = inside ? -1 : 1; }
        }
        {
            const float b = -vd.y, i1 = f_sub(b, sq[1]), i2 = f_add(b, sq[1]);
            const bool inside = i1 < 0.f;
            const float t = inside ? i2 : i1;
            if (live & c1 & (i2 > 0.f) & (t < L.cumu)) { L.cumu = t; L.qhit = i + 1; L.qkind = inside ? -1 : 1; }
        }
    }
}
// Candidate test of the plane pairs.  With num = -s and the signed limit w = (d*cumu)*(1+2^-21), pre-filters A and B
// together say: the division is only worth taking if num lies between 0 and w, i.e. num*(w - num) >= 0, which is
// s*(w + s) <= 0 -- one FADD2, one FMUL2 and one comparison per plane instead of a sign test and two magnitude tests.  The sign of the rounded
// difference and of the (possibly flushed) product are the exact ones, a zero counts as "take the division", and so
// does a limit that is not comfortably normal (pre-filter B's error bound needs that).  The test is a superset of the
// scalar loops' candidate test and only gates the exact stage, which re-checks 0 < dist < cumu (a zero d gives inf or
// NaN there and is rejected) -- the accepted hits are the same.
RT_HD bool w_plane_candidate(float w, float m) { return (m <= 0.f) | !(fabsf(w) > 1e-30f); }
RT_HD void w_plane_pair(WLane &L, const f2x2 A, const f2x2 B, int i, int live_i) {
    const bool live = live_i != 0;
    const f2 d = f2_dot3_s(A.a, A.b, B.a, L.qdx, L.qdy, L.qdz);
    const f2 s = f2_add(f2_dot3_s(A.a, A.b, B.a, L.qox, L.qoy, L.qoz), B.b);
    const f2 w = f2_mul_s(f2_mul_s(d, L.cumu), 1.000000476837158203125f);
    const f2 m = f2_mul(s, f2_add(w, s));                                                          // -num * (w - num)
    const bool c0 = w_plane_candidate(w.x, m.x);
    const bool c1 = w_plane_candidate(w.y, m.y);
    if (warp_any((c0 | c1) & live)) {
        const float q0 = f_div(-s.x, d.x), q1 = f_div(-s.y, d.y);
        if (live & c0 & (q0 > 0.f) & (q0 < L.cumu)) { L.cumu = q0; L.qhit = i; L.qkind = 1; }
        if (live & c1 & (q1 > 0.f) & (q1 < L.cumu)) { L.cumu = q1; L.qhit = i + 1; L.qkind = 1; }
    }
}
RT_HD void w_query_nearest_x2(WLane &L, const f2x2 *pairs, const int *runs, int n_runs, bool has_query) {
    if (!warp_any(has_query)) return;
    const int has = has_query ? 1 : 0;        // an integer: re-tested per pair, so that it does not pin a predicate register across the loops
    for (int r = 0; r < n_runs; ++r) {
        const int start = runs[W_RUN_STRIDE * r], count = runs[W_RUN_STRIDE * r + 1], fl = runs[W_RUN_STRIDE * r + 2];
        const f2x2 *g = pairs + 2 * runs[W_RUN_STRIDE * r + 3];
        const int end = start + count;
        if (fl & W_FLAG_SPHERE) {
#pragma unroll 1
            for (int i = start; i < end; i += 2, g += 2) w_sphere_pair(L, g[0], g[1], i, has);
        } else {
#pragma unroll 1
            for (int i = start; i < end; i += 2, g += 2) w_plane_pair(L, g[0], g[1], i, has);
        }
    }
}

// Shadow round, one primitive pair against the lane's (up to) three shadow rays, ray by ray (holding the
// discriminants of all three rays until one common vote costs ~60 registers more than the scalar loops use).
RT_HD void w_shadow_sphere_pair(WLane &L, const f2x2 A, const f2x2 B, int alive) {
#pragma unroll
    for (int k = 0; k < W_SHADOW_BATCH; k++) {
        const f2 vx = f2_rsub_s(L.sox[k], A.a), vy = f2_rsub_s(L.soy[k], A.b), vz = f2_rsub_s(L.soz[k], B.a);
        const f2 vd = f2_dot3_s(vx, vy, vz, L.slx[k], L.sly[k], L.slz[k]);
        const f2 bb = f2_mul(vd, vd);
        const f2 det = f2_add(f2_sub(bb, f2_dot3(vx, vy, vz, vx, vy, vz)), B.b);
        const f2 lim = f2_mul_s(bb, 0.999999761581420898437500f);
        const bool on = (alive >> k) & 1;
        const bool c0 = (det.x > 0.f) & !((vd.x > 0.f) & (det.x < lim.x));
        const bool c1 = (det.y > 0.f) & !((vd.y > 0.f) & (det.y < lim.y));
        if (warp_any((c0 | c1) & on)) {
            const float dv[2] = { det.x, det.y };
            const bool need[2] = { c0, c1 };
            float sq[2];
            sqrt_group<2>(dv, need, sq);
            const float b0 = -vd.x, b1 = -vd.y;
            const float i1a = f_sub(b0, sq[0]), i2a = f_add(b0, sq[0]), i1b = f_sub(b1, sq[1]), i2b = f_add(b1, sq[1]);
            const float ta = i1a < 0.f ? i2a : i1a, tb = i1b < 0.f ? i2b : i1b;
            if (on & ((c0 & (i2a > 0.f) & (ta < L.sreach[k])) | (c1 & (i2b > 0.f) & (tb < L.sreach[k])))) L.sblk |= 1 << k;
        }
    }
}
RT_HD void w_shadow_plane_pair(WLane &L, const f2x2 A, const f2x2 B, int alive) {
#pragma unroll
    for (int k = 0; k < W_SHADOW_BATCH; k++) {
        const f2 d = f2_dot3_s(A.a, A.b, B.a, L.slx[k], L.sly[k], L.slz[k]);
        const f2 s = f2_add(f2_dot3_s(A.a, A.b, B.a, L.sox[k], L.soy[k], L.soz[k]), B.b);
        const f2 w = f2_mul_s(f2_mul_s(d, L.sreach[k]), 1.000000476837158203125f);
        const f2 m = f2_mul(s, f2_add(w, s));
        const bool on = (alive >> k) & 1;
        const bool c0 = w_plane_candidate(w.x, m.x);
        const bool c1 = w_plane_candidate(w.y, m.y);
        if (warp_any((c0 | c1) & on)) {
            const float q0 = f_div(-s.x, d.x), q1 = f_div(-s.y, d.y);
            if (on & ((c0 & (q0 > 0.f) & (q0 < L.sreach[k])) | (c1 & (q1 > 0.f) & (q1 < L.sreach[k])))) L.sblk |= 1 << k;
        }
    }
}
RT_HD void w_query_shadow_x2(WLane &L, const f2x2 *pairs, const int *runs, int n_runs, bool has) {
    for (int r = 0; r < n_runs; ++r) {
        const int count = runs[W_RUN_STRIDE * r + 1], fl = runs[W_RUN_STRIDE * r + 2];
        if (fl & W_FLAG_LIGHT) continue;                                   // RNO:234: lights cast no shadow
        const int alive = w_alive_mask(L, has);
        if (!warp_any(alive != 0)) return;                                 // the `break` of RNO:237, for the whole warp
        const f2x2 *g = pairs + 2 * runs[W_RUN_STRIDE * r + 3];
        if (fl & W_FLAG_SPHERE) {
#pragma unroll 1
            for (int i = 0; i < count; i += 2, g += 2) w_shadow_sphere_pair(L, g[0], g[1], alive);
        } else {
#pragma unroll 1
            for (int i = 0; i < count; i += 2, g += 2) w_shadow_plane_pair(L, g[0], g[1], alive);
        }
    }
}
// Rule 2 of f32x2.cuh for the rays of the query in flight.
RT_HD bool w_nearest_ray_packed_ok(const WLane &L) {
    return packed_range_ok3(L.qox, L.qoy, L.qoz) && packed_range_ok3(L.qdx, L.qdy, L.qdz);
}
RT_HD bool w_shadow_rays_packed_ok(const WLane &L) {
    bool ok = true;
#pragma unroll
    for (int k = 0; k < W_SHADOW_BATCH; k++)
        ok = ok && (k >= L.ns || (packed_range_ok3(L.sox[k], L.soy[k], L.soz[k]) && packed_range_ok3(L.slx[k], L.sly[k], L.slz[k])));
    return ok;
}


}  // namespace rtb
