// scene_soa.h -- host-side conversion of the reference's array-of-structs scene tables into the
// structure-of-arrays float4 layout the kernels stage into shared memory.  Pure host C++.
#pragma once
#include <vector>
#include <stdint.h>
#include <string.h>
#include <cmath>
#include "../../include/rt_b200.h"
#include "pt_lane.cuh"
#include "whitted_lane.cuh"
#include "pt_bvh_build.h"

namespace rtb {

static_assert(sizeof(rt_sphere) == 44, "rt_sphere must match the reference Sphere (SPT/geom.h:43-47)");
static_assert(sizeof(rt_camera) == 60, "rt_camera must match the reference Camera (SPT/camera.h:29-34)");
static_assert(sizeof(rt_primitive) == 96, "rt_primitive must match the reference Primitive_2 (R323/common.h:49-63)");
static_assert(sizeof(rt_uchar4) == 4 && sizeof(rt_float4) == 16 && sizeof(f4) == 16, "vector PODs");

struct PtSoA {
    std::vector<f4> geom, emis, colr;
    std::vector<int> lights;
};

inline void build_pt_soa(const rt_sphere *s, uint32_t n, PtSoA &out) {
    out.geom.resize(n); out.emis.resize(n); out.colr.resize(n); out.lights.clear();
    for (uint32_t i = 0; i < n; i++) {
        const float rad2 = s[i].rad * s[i].rad;                 // `s->rad * s->rad`, SPT/geomfunc.h:42 (one rounding)
        f4 g = { s[i].p.x, s[i].p.y, s[i].p.z, rad2 };
        uint32_t refl_bits = (uint32_t)s[i].refl; float refl_as_float; memcpy(&refl_as_float, &refl_bits, 4);
        f4 e = { s[i].e.x, s[i].e.y, s[i].e.z, refl_as_float };
        f4 c = { s[i].c.x, s[i].c.y, s[i].c.z, s[i].rad };
        out.geom[i] = g; out.emis[i] = e; out.colr[i] = c;
        // viszero() of SPT/vec.h:44 tests x, x and z -- never y.  Reproduced.
        if (!(s[i].e.x == 0.f && s[i].e.z == 0.f)) out.lights.push_back((int)i);
    }
}

struct WSoA {
    std::vector<f4> geom, mat_a, mat_b;
    std::vector<f4> lcenter;            // `center` of every light, in lights[] order
    std::vector<int> flags, lights, runs;
    std::vector<int> runs_hot;          // the same runs without primitives that can never be hit (timed launches use these)
    std::vector<float> rrad;
    int n_spheres = 0, n_planes = 0;
    float tame_reach[2] = { 0.f, 0.f }; // WFrame::tame_reach of this table (w_scene_tame_reach below)
    // large scenes (build_w_bvh): the non-light spheres in the exact hierarchy of pt_bvh.cuh, and the run table of
    // everything else that can be hit (planes, lights, spheres the tree does not take) -- what every query still walks
    PtBvhHost bvh;
    std::vector<int> runs_bvh;
    int n_tree = 0;
};


// WFrame::tame_reach: how close to every light of its batch a hit point must be for the timed kernels to skip a BLOCKED light instead of
// shading it with shade = 0 as the reference does (RNO:250, 270).  Skipping is right only when 0 times every factor is +-0, that is when
// nothing in the two terms can be inf or NaN.  The kernel checks the per-ray part (|d| < 1e3, distance to each light of the batch inside
// (1e-18, tame_reach): whitted_lane.cuh); this is the scene's part, in double with room for the float roundings: colours, m_diff and m_spec
// finite; |N| <= nmax at any hit point that close to a light (plane: |normal|; sphere: (|P - c_l| + |c_l - c|) / r), so that
// |N.L| m_diff <= nmax m_diff fits a float and |V.R| <= |d| |L| (1 + 2 |N|^2) to the 20th power times m_spec fits a double.
// out[0]: the largest such distance for hit points on planes (N is the plane's own normal: any distance whose square is a float, 1e18, if the
// normals allow it), out[1]: on spheres; 0 when there is none (every such batch is then reported, and the EXACT kernel renders those pixels).
inline void w_scene_tame_reach(const struct WSoA &s, int n, float out[2]);

inline void build_w_soa(const rt_primitive *p, int n, WSoA &out) {
    out.geom.resize(n); out.mat_a.resize(n); out.mat_b.resize(n); out.flags.resize(n); out.rrad.resize(n);
    out.lights.clear(); out.lcenter.clear(); out.n_spheres = out.n_planes = 0;
    for (int i = 0; i < n; i++) {
        int fl = 0;
        f4 g;
        if (p[i].type == RT_SPHERE) {
            fl |= W_FLAG_SPHERE; out.n_spheres++;
            g.x = p[i].center.x; g.y = p[i].center.y; g.z = p[i].center.z; g.w = p[i].sq_radius;
        } else if (p[i].type == RT_PLANE) {
            out.n_planes++;
            g.x = p[i].normal.x; g.y = p[i].normal.y; g.z = p[i].normal.z; g.w = p[i].depth;
        } else {                       // intersect() returns MISS for any other type (RNO:150-160): a plane that is never hit
            g.x = g.y = g.z = g.w = 0.f;
        }
        if (p[i].is_light) {
            fl |= W_FLAG_LIGHT; out.lights.push_back(i);
            f4 c = { p[i].center.x, p[i].center.y, p[i].center.z, 0.f };
            out.lcenter.push_back(c);
        }
        f4 a = { p[i].m_color.x, p[i].m_color.y, p[i].m_color.z, p[i].m_refl };
        f4 b = { p[i].m_diff, p[i].m_refr, p[i].m_refr_index, p[i].m_spec };
        out.geom[i] = g; out.mat_a[i] = a; out.mat_b[i] = b; out.flags[i] = fl; out.rrad[i] = p[i].r_radius;
    }
    w_scene_tame_reach(out, n, out.tame_reach);
    // Index order is part of the result (ties), so the kernel walks the primitives in order -- but as runs
    // of equal (type, is_light), which takes the type dispatch out of the inner loop.
    out.runs.clear();
    for (int i = 0; i < n;) {
        int j = i;
        while (j < n && out.flags[j] == out.flags[i] && j - i < 32) j++;   // <= 32 per run: the kernel re-votes per run
        out.runs.push_back(i); out.runs.push_back(j - i); out.runs.push_back(out.flags[i]);
        i = j;
    }
    // A plane whose normal is (0,0,0) -- the reference's zeroed slot (R323/scene.c:55-57), or a primitive of unknown
    // type -- has d = N.dir = 0 (or NaN) for every ray and fails `d != 0` / `dist > 0` (RNO:97-108): it is never hit and
    // never shadows.  The work counters still include it (counting launches walk `runs`); timed launches skip it.
    auto dead = [&](int i) { return !(out.flags[i] & W_FLAG_SPHERE) && out.geom[i].x == 0.f && out.geom[i].y == 0.f && out.geom[i].z == 0.f; };
    out.runs_hot.clear();
    for (int i = 0; i < n;) {
        if (dead(i)) { i++; continue; }
        int j = i;
        while (j < n && !dead(j) && out.flags[j] == out.flags[i] && j - i < 32) j++;
        out.runs_hot.push_back(i); out.runs_hot.push_back(j - i); out.runs_hot.push_back(out.flags[i]);
        i = j;
    }
}

inline void w_scene_tame_reach(const WSoA &s, int n, float out[2]) {
    out[0] = out[1] = 0.f;
    const double dmax = 1e3;                        // sqrt(W_TAME_D2)
    double spec = 0.0, diff = 0.0, plane_n = 0.0, far = 0.0, rrad = 0.0;
    for (int i = 0; i < n; i++) {
        const f4 a = s.mat_a[i], b = s.mat_b[i], g = s.geom[i];
        for (float v : { a.x, a.y, a.z, b.x, b.w }) if (!(std::fabs(v) < 1e30f)) return;         // NaN lands here too
        spec = std::max(spec, (double)std::fabs(b.w)); diff = std::max(diff, (double)std::fabs(b.x));
        if (s.flags[i] & W_FLAG_SPHERE) {
            for (const f4 &c : s.lcenter) {
                const double ex = (double)c.x - g.x, ey = (double)c.y - g.y, ez = (double)c.z - g.z;
                const double e = std::sqrt(ex * ex + ey * ey + ez * ez);
                if (!(e < 1e30)) return;
                far = std::max(far, e);
            }
            if (!(std::fabs(s.rrad[i]) < 1e30f)) return;
            rrad = std::max(rrad, (double)std::fabs(s.rrad[i]));
        } else {
            const double e = std::sqrt((double)g.x * g.x + (double)g.y * g.y + (double)g.z * g.z);
            if (!(e < 1e30)) return;
            plane_n = std::max(plane_n, e * 1.001);
        }
    }
    // the largest |N| the two products allow
    const double vmax = std::pow(10.0, (300.0 - std::log10(std::max(spec, 1e-300))) / 20.0);      // |V.R| below this: pow(V.R, 20) m_spec < 1e300
    const double q = (vmax / (3.0 * 1.001 * dmax) - 1.001) / 2.004;                                // 1 + 2 |N|^2 (with roundings) below vmax / (3 |d|)
    if (!(q > 0.0)) return;
    double nmax = std::sqrt(q);
    if (diff > 0.0) nmax = std::min(nmax, 1e37 / (1.01 * diff));
    if (plane_n < nmax) out[0] = 1e18f;
    double reach = 1e18;
    if (rrad > 0.0) reach = std::min(reach, (nmax / (rrad * 1.001) - far) / 1.001);
    if (!(reach > 0.0)) return;
    float r = (float)reach;
    if ((double)r > reach) r = nextafterf(r, 0.f);
    out[1] = r;
}

// The hierarchy over the non-light spheres of a Whitted scene table, after build_w_soa.  Lights stay in the run table
// (a shadow ray skips them, RNO:234, and there are few); so do planes and whatever the builder itself keeps out of the tree.
inline void build_w_bvh(const rt_primitive *p, int n, WSoA &out) {
    std::vector<f4> g((size_t)n), c((size_t)n);
    std::vector<char> skip((size_t)n, 1);
    for (int i = 0; i < n; i++) {
        g[i] = out.geom[i]; c[i] = f4{0, 0, 0, 0};
        if ((out.flags[i] & W_FLAG_SPHERE) && !(out.flags[i] & W_FLAG_LIGHT)) { c[i].w = p[i].radius; skip[i] = 0; }
    }
    build_pt_bvh(g, c, out.bvh, &skip);
    std::vector<char> in_tree((size_t)n, 0);
    for (size_t j = (size_t)out.bvh.n_big; j < out.bvh.index.size(); j++) in_tree[out.bvh.index[j]] = 1;
    out.n_tree = (int)out.bvh.index.size() - out.bvh.n_big;
    out.bvh.n_big = 0;                                 // the builder's always-tested spheres join the run table below
    auto dead = [&](int i) { return !(out.flags[i] & W_FLAG_SPHERE) && out.geom[i].x == 0.f && out.geom[i].y == 0.f && out.geom[i].z == 0.f; };
    out.runs_bvh.clear();
    for (int i = 0; i < n;) {
        if (dead(i) || in_tree[i]) { i++; continue; }
        int j = i;
        while (j < n && !dead(j) && !in_tree[j] && out.flags[j] == out.flags[i] && j - i < 32) j++;
        out.runs_bvh.push_back(i); out.runs_bvh.push_back(j - i); out.runs_bvh.push_back(out.flags[i]);
        i = j;
    }
    if (out.runs_bvh.empty()) { out.runs_bvh.push_back(0); out.runs_bvh.push_back(0); out.runs_bvh.push_back(0); }
}

// Tables of the shadow-round culls of whitted_lane.cuh ("Exact culls of the shadow round"), for one run table of one scene.
// Everything here is evaluated in double and rounded AWAY from "cull", so the float comparisons of the kernel are safe.
//
// Setting: every light is a sphere (else `enabled` stays false).  Rs = largest |coordinate| + radius over all spheres, RL =
// sqrt(3) Rs bounds |c_light|; the culls only apply to hit points with |P| < RP = 2 Rs + 16, so a shadow ray (o = fl(P +
// fl(L EPS)), L = fl(fl(1 / len) (c_l - P)), RNO:214-231) has |o - P| <= 1.01 EPS + 1.8u RP, length len < RR = RP + RL + 0.002,
// |op| < V = RR for every sphere, | |L|^2 - 1 | <= 12u, and the real point o + (len - EPS) L lies within 6u (RP + RL) of the
// light's centre.  u = 2^-24, EPS = 0.001 (RNO:26); scenes with RR > 8 000 get no culls.
//
// Plane (N, depth), nn = |N|, g(t) = N.(o + t L) + depth, D = N.L (reals).  The reference computes d = fl(N.L) and, iff d != 0,
// s = fl(fl(N.o) + depth) and dist = fl(-s / d); the plane blocks iff 0 < dist < len (RNO:97-108, 232-240).  Bounds:
// |s - g(0)| <= e2 = 4u (nn |o| + |depth|), |d - D| <= 3.1u nn; the kernel's fused N.P + depth is within 4u (nn RP + |depth|)
// of the real value.  T = nn (3 EPS + 40u (RP + RR)) + 40u |depth|.  The kernel culls when sgn (N.P + depth) > T (float, T
// rounded up); the table holds an entry only if sgn (N.c_l + depth) >= 2T for EVERY light (double, here).  Then
//   sgn g(0) >= 1.99 nn EPS + 34u nn (RP + RR) and sgn s >= 1.99 nn EPS: s is non-zero and has the sign sgn;
//   sgn g(len - EPS) >= 2T - 6u nn (RP + RL) >= 1.8 T: g has no zero on [0, len - EPS].
//   d == 0: no test (RNO:98).   sign d == sgn: dist <= 0.
//   sign d == -sgn but sign D != -sgn (rounding flipped it): |d| <= 3.1u nn, dist >= 1.99 nn EPS / (3.1u nn) > 10 000 > RR > len.
//   sign d == sign D == -sgn: |g(0)| = |g(len - EPS)| + (len - EPS)|D|, so |s| / |d| >= (1.7 T + (len - EPS)|D|) / (|D| + 3.1u nn),
//     which is >= len (1 + 2u) because 1.7 T >= nn (5.1 EPS + 68u (RP + RR)) > (EPS + 2u RR)|D| + 3.2u nn RR: dist >= len.
//   In no case a blocker, for any ray of the batch (T does not depend on the light).
// Sphere run with box [lo, hi] (centres -+ radii), r_min / r_max its radii.  pt_bvh.cuh (the Whitted test has the same
// roundings): det >= 0 implies the line through o' (|o' - o| <= 1.8u V) along L passes within R' of the centre, R' - rad <=
// min(eta_S / (2 rad), sqrt(eta_S)), eta_S = (|e| + 13u)(1 + 2|e|) OO + (2|e| + 2u) rad^2 <= eta = 28u V^2 + 28u r_max^2 for
// |e| <= 12u, and every accepted distance is the parameter of a point of that inflated sphere up to (|e| + 10u)|op| <= 22u V.
// grow = min(eta / (2 r_min), 1.001 sqrt(eta)) + 3 EPS + 96u (V + RP).  If P (kernel, float, box rounded outwards) and every
// light centre (double, here) lie beyond one face of the box grown by `grow`, then every point o' + t L with -22u V <= t <=
// len + 22u V lies beyond that face grown by R' - rad (the coordinate is linear in t; the ends are within 1.01 EPS + 1.8u RP
// + 1.8u V, resp. 6u (RP + RL) + 1.8u V + EPS, of P and of the light), where no point of an inflated sphere of the run is:
// no sphere of the run returns an accepted distance in [0, len).
#ifndef W_GRID_CELLS
#define W_GRID_CELLS 49152          /* about this many cells (4 bytes each) */
#endif
struct WCull {
    std::vector<f2> pcull;      // per primitive
    std::vector<f4> rbox;       // two per run
    float rp2 = 0.f;
    float reject_k = 0.f;       // K of w_shadow_sphere_keep: (64 + 8 RP) u
    bool enabled = false;
    int planes_cullable = 0, runs_cullable = 0;     // for the record (tests, bench)
    // shadow-candidate grid (whitted_lane.cuh): per-sphere rad + grow (rounded up), and the box / cell counts; grid.cells stays NULL here
    std::vector<float> smargin;
    WGrid grid;
    int grid_gz = 0;            // cells along z (0: no grid for this table)
};

inline float w_cull_up(double v) { float f = (float)v; if ((double)f < v) f = nextafterf(f, INFINITY); return f; }
inline float w_cull_down(double v) { float f = (float)v; if ((double)f > v) f = nextafterf(f, -INFINITY); return f; }

inline void build_w_cull(const WSoA &soa, const std::vector<int> &runs, WCull &out) {
    const int n = (int)soa.geom.size(), n_runs = (int)runs.size() / 3;
    const f2 never = { 0.f, INFINITY };
    const f4 open_lo = { -INFINITY, -INFINITY, -INFINITY, 0.f }, open_hi = { INFINITY, INFINITY, INFINITY, 0.f };
    out.pcull.assign((size_t)(n > 0 ? n : 1), never);
    out.rbox.assign((size_t)(n_runs > 0 ? 2 * n_runs : 2), open_lo);
    for (int r = 0; r < n_runs; r++) out.rbox[2 * r + 1] = open_hi;
    out.rp2 = 0.f; out.enabled = false; out.planes_cullable = out.runs_cullable = 0;
    out.smargin.assign((size_t)(n > 0 ? n : 1), 0.f); memset(&out.grid, 0, sizeof out.grid); out.grid_gz = 0;
    if (soa.lights.empty()) return;
    const double u = 1.0 / 16777216.0, EPS = 0.001;
    double Rs = 0.0;
    for (int i = 0; i < n; i++) {
        if (!(soa.flags[i] & W_FLAG_SPHERE)) continue;
        const f4 g = soa.geom[i];
        if (!(std::isfinite(g.x) && std::isfinite(g.y) && std::isfinite(g.z) && std::isfinite(g.w)) || g.w < 0.f) return;
        const double ext = std::fmax(std::fabs((double)g.x), std::fmax(std::fabs((double)g.y), std::fabs((double)g.z))) + std::sqrt((double)g.w);
        if (ext > Rs) Rs = ext;
    }
    for (int l : soa.lights) if (!(soa.flags[l] & W_FLAG_SPHERE)) return;       // a light that is not a sphere: the general path, no culls
    const double RL = std::sqrt(3.0) * Rs, RP = 2.0 * Rs + 16.0, RR = RP + RL + 0.002, V = RR;
    if (!(RR <= 8000.0)) return;
    // planes
    for (int i = 0; i < n; i++) {
        if (soa.flags[i] & (W_FLAG_SPHERE | W_FLAG_LIGHT)) continue;
        const f4 g = soa.geom[i];
        if (!(std::isfinite(g.x) && std::isfinite(g.y) && std::isfinite(g.z) && std::isfinite(g.w))) continue;
        const double nn = std::sqrt((double)g.x * g.x + (double)g.y * g.y + (double)g.z * g.z);
        if (!(nn > 0.0)) continue;
#ifdef W_CULL_TEST_NO_MARGIN      /* tools/cull_fuzz.py --self-check: tables WITHOUT the margins must be caught by the fuzz */
        const double T = 0.0;
#else
        const double T = nn * (3.0 * EPS + 40.0 * u * (RP + RR)) + 40.0 * u * std::fabs((double)g.w);
#endif
        int side = 0;
        bool ok = true;
        for (int l : soa.lights) {
            const f4 c = soa.geom[l];
            const double sl = (double)g.x * c.x + (double)g.y * c.y + (double)g.z * c.z + (double)g.w;
            const int sg = sl > 0.0 ? 1 : -1;
            if (!(std::fabs(sl) >= 2.0 * T) || (side != 0 && sg != side)) { ok = false; break; }
            side = sg;
        }
        if (!ok || side == 0) continue;
        out.pcull[i] = f2{ (float)side, w_cull_up(T) };
        out.planes_cullable++;
    }
    // runs of non-light spheres
    for (int r = 0; r < n_runs; r++) {
        const int start = runs[3 * r], count = runs[3 * r + 1], fl = runs[3 * r + 2];
        if (!(fl & W_FLAG_SPHERE) || (fl & W_FLAG_LIGHT) || count < 1) continue;
        double lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY }, rmin = INFINITY, rmax = 0.0;
        for (int i = start; i < start + count; i++) {
            const f4 g = soa.geom[i];
            const double rad = std::sqrt((double)g.w), c[3] = { g.x, g.y, g.z };
            for (int a = 0; a < 3; a++) { lo[a] = std::fmin(lo[a], c[a] - rad); hi[a] = std::fmax(hi[a], c[a] + rad); }
            rmin = std::fmin(rmin, rad); rmax = std::fmax(rmax, rad);
        }
        const double eta = 28.0 * u * V * V + 28.0 * u * rmax * rmax;
#ifdef W_CULL_TEST_NO_MARGIN
        const double grow = -0.02 * rmax;
#else
        const double grow = (rmin > 0.0 ? std::fmin(eta / (2.0 * rmin), 1.001 * std::sqrt(eta)) : 1.001 * std::sqrt(eta)) + 3.0 * EPS + 96.0 * u * (V + RP);
#endif
        float flo[3], fhi[3];
        bool any = false;
        for (int a = 0; a < 3; a++) {
            bool above = true, below = true;             // every light centre beyond the grown face
            for (int l : soa.lights) {
                const f4 cg = soa.geom[l];
                const double c = a == 0 ? cg.x : (a == 1 ? cg.y : cg.z);
                if (!(c > hi[a] + grow)) above = false;
                if (!(c < lo[a] - grow)) below = false;
            }
            fhi[a] = above ? w_cull_up(hi[a] + grow) : INFINITY;
            flo[a] = below ? w_cull_down(lo[a] - grow) : -INFINITY;
            any = any || above || below;
        }
        out.rbox[2 * r] = f4{ flo[0], flo[1], flo[2], 0.f };
        out.rbox[2 * r + 1] = f4{ fhi[0], fhi[1], fhi[2], 0.f };
        if (any) out.runs_cullable++;
    }
    out.rp2 = w_cull_down(RP * RP * (1.0 - 8.0 * u));
    out.reject_k = w_cull_up((64.0 + 8.0 * RP) * u);
    out.enabled = out.planes_cullable > 0 || out.runs_cullable > 0;
    if (!out.enabled) out.rp2 = 0.f;
    // Shadow-candidate grid: scenes of at most 32 primitives (one bit each).  The box: the spheres and the lights, stretched to the planes
    // that are perpendicular to an axis (the walls of a room) but not beyond the culls' radius; about W_GRID_CELLS cells of equal size.
    if (!out.enabled || n > 32) return;
    uint32_t all = 0;
    for (int r = 0; r < n_runs; r++) {
        const int start = runs[3 * r], count = runs[3 * r + 1], fl = runs[3 * r + 2];
        if (fl & W_FLAG_LIGHT) continue;
        for (int i = start; i < start + count; i++) all |= 1u << i;
    }
    double lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
    for (int i = 0; i < n; i++) {
        const f4 g = soa.geom[i];
        if (soa.flags[i] & W_FLAG_SPHERE) {
            const double rad = (soa.flags[i] & W_FLAG_LIGHT) ? 0.0 : std::sqrt((double)g.w), c[3] = { g.x, g.y, g.z };
            for (int a = 0; a < 3; a++) { lo[a] = std::fmin(lo[a], c[a] - rad); hi[a] = std::fmax(hi[a], c[a] + rad); }
            {
                const double rad = std::sqrt((double)g.w);           // lights too: a nearest query can hit them (the primary-ray tiles)
                const double eta = 28.0 * u * V * V + 28.0 * u * rad * rad;
#ifdef W_CULL_TEST_NO_MARGIN
                const double grow = -0.02 * rad;
#else
                const double grow = (rad > 0.0 ? std::fmin(eta / (2.0 * rad), 1.001 * std::sqrt(eta)) : 1.001 * std::sqrt(eta)) + 3.0 * EPS + 96.0 * u * (V + RP);
#endif
                out.smargin[i] = w_cull_up(rad + grow);
            }
        }
    }
    for (int i = 0; i < n; i++) {
        if (soa.flags[i] & W_FLAG_SPHERE) continue;
        const f4 g = soa.geom[i];
        const int nz = (g.x != 0.f) + (g.y != 0.f) + (g.z != 0.f);
        if (nz != 1 || !std::isfinite(g.w)) continue;
        const int a = g.x != 0.f ? 0 : (g.y != 0.f ? 1 : 2);
        const double wpos = -(double)g.w / (double)(a == 0 ? g.x : (a == 1 ? g.y : g.z));
        if (!std::isfinite(wpos)) continue;
        lo[a] = std::fmin(lo[a], wpos); hi[a] = std::fmax(hi[a], wpos);
    }
    double ext[3], vol = 1.0;
    for (int a = 0; a < 3; a++) {
        if (!(lo[a] <= hi[a])) return;
        lo[a] = std::fmax(lo[a], -RP); hi[a] = std::fmin(hi[a], RP);
        const double pad = 1e-3 * (hi[a] - lo[a]) + 1e-3;
        lo[a] -= pad; hi[a] += pad;
        ext[a] = hi[a] - lo[a];
        vol *= ext[a];
    }
    const double edge = std::cbrt(vol / (double)W_GRID_CELLS);
    int gdim[3];
    for (int a = 0; a < 3; a++) { double c = std::ceil(ext[a] / edge); gdim[a] = c < 1.0 ? 1 : (c > 256.0 ? 256 : (int)c); }
    WGrid &G = out.grid;
    G.cells = nullptr; G.all = all;
    G.tiles = nullptr; G.tiles_x = 0; G.tiles_y = 0; G.all_nearest = 0; G.deep = 0;
    for (int i = 0; i < n; i++) if (soa.mat_a[i].w > 0.f || soa.mat_b[i].y > 0.f) G.deep |= 1u << i;
    for (int r = 0; r < n_runs; r++) for (int i = runs[3 * r]; i < runs[3 * r] + runs[3 * r + 1]; i++) G.all_nearest |= 1u << i;
    G.x0 = (float)lo[0]; G.y0 = (float)lo[1]; G.z0 = (float)lo[2];
    G.ix = (float)(gdim[0] / ext[0]); G.iy = (float)(gdim[1] / ext[1]); G.iz = (float)(gdim[2] / ext[2]);
    G.fgx = (float)gdim[0]; G.fgy = (float)gdim[1]; G.fgz = (float)gdim[2];
    G.gx = gdim[0]; G.gy = gdim[1]; G.gz = gdim[2];
    out.grid_gz = gdim[2];
}

static_assert(sizeof(rt_r306_primitive) == 96, "rt_r306_primitive must match the reference Primitive (R306/raytracer.h:24-34)");

// The 3.0.06 scene table in the same structure-of-arrays form (its intersection core is the 3.2.03 one):
// mat_a = (m_Color, m_Refl), mat_b = (m_Diff, m_Refr, m_RIndex, m_Spec).  A type that is neither sphere nor plane is
// never hit (R306/scene.cpp:185-187): an all-zero plane.
inline void build_r306_soa(const rt_r306_primitive *p, int n, WSoA &out) {
    std::vector<rt_primitive> q((size_t)n);
    for (int i = 0; i < n; i++) {
        rt_primitive &r = q[i];
        memset(&r, 0, sizeof r);
        r.m_color.x = p[i].m_color.x; r.m_color.y = p[i].m_color.y; r.m_color.z = p[i].m_color.z;
        r.m_refl = p[i].m_refl; r.m_diff = p[i].m_diff; r.m_refr = p[i].m_refr; r.m_refr_index = p[i].m_rindex; r.m_spec = p[i].m_spec;
        r.is_light = p[i].m_light > 0 ? 1 : 0;
        if (p[i].type == RT_R306_SPHERE) {
            r.type = RT_SPHERE;
            r.center.x = p[i].centre.x; r.center.y = p[i].centre.y; r.center.z = p[i].centre.z;
            r.radius = p[i].radius; r.sq_radius = p[i].sq_radius; r.r_radius = p[i].r_radius;
        } else if (p[i].type == RT_R306_PLANE) {
            r.type = RT_PLANE;
            r.normal.x = p[i].plane_n.x; r.normal.y = p[i].plane_n.y; r.normal.z = p[i].plane_n.z; r.depth = p[i].plane_d;
        } else r.type = -1;
    }
    build_w_soa(q.data(), n, out);
}

// Engine_InitRender / Engine_Render screen coordinates (R306/raytracer.cpp:278-296, :313, :523-525): running float sums.
inline void build_r306_screen(int w, int h, std::vector<float> &sx, std::vector<float> &sy, float *DX, float *DY) {
    const float WX1 = -3, WX2 = 3, WY1 = 2.25f, WY2 = -2.25f;
    const float dx = (WX2 - WX1) / w, dy = (WY2 - WY1) / h;
    sx.assign((size_t)w, 0.f); sy.assign((size_t)h, 0.f);
    float v = WX1;
    for (int x = 0; x < w; x++) { sx[x] = v; v += dx; }
    float u = WY1;
    u += 20 * dy;
    for (int y = 20; y < h; y++) { sy[y] = u; u += dy; }
    *DX = dx; *DY = dy;
}

// Rows of an h-row frame owned by `rank` when tiles of tile_rows rows are dealt round-robin.
inline int shard_local_rows(int h, int rank, int world, int tile_rows) {
    int rows = 0;
    for (int t = rank; t * tile_rows < h; t += world) {
        const int r = h - t * tile_rows;
        rows += r < tile_rows ? r : tile_rows;
    }
    return rows;
}

inline Shard make_shard(int w, int h, int rank, int world, int tile_rows, uint32_t *n_items) {
    Shard s;
    s.rank = rank; s.world = world; s.tile_rows = tile_rows;
    s.local_rows = shard_local_rows(h, rank, world, tile_rows);
    s.blocks_x = (w + 7) / 8;
    const int blocks_y = (s.local_rows + 3) / 4;
    *n_items = (uint32_t)s.blocks_x * (uint32_t)blocks_y * 32u;
    return s;
}

}  // namespace rtb
