// scene_soa.h -- host-side conversion of the reference's array-of-structs scene tables into the
// structure-of-arrays float4 layout the kernels stage into shared memory.  Pure host C++.
#pragma once
#include <vector>
#include <stdint.h>
#include <string.h>
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
    // large scenes (build_w_bvh): the non-light spheres in the exact hierarchy of pt_bvh.cuh, and the run table of
    // everything else that can be hit (planes, lights, spheres the tree does not take) -- what every query still walks
    PtBvhHost bvh;
    std::vector<int> runs_bvh;
    int n_tree = 0;
};

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
