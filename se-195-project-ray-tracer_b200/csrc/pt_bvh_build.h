// pt_bvh_build.h -- host-side construction of the hierarchy pt_bvh.cuh walks.  Pure host C++.
//
// Binned surface-area splits over the sphere boxes, leaves of <= PT_BVH_LEAF_MAX spheres, median splits below depth 24
// so that the depth stays under the traversal stack.  Boxes are formed in double and rounded outwards to float.  Spheres
// that are not finite, or much larger than the typical sphere (radius > 64 x median: the floor of the
// generated scenes), stay out of the tree: every query tests them directly.
#pragma once
#include <vector>
#include <algorithm>
#include <cmath>
#include <cfloat>
#include <stdint.h>
#include <string.h>
#include "pt_bvh.cuh"

namespace rtb {

struct PtBvhHost {
    std::vector<f4> nodes;
    std::vector<f4> geom;
    std::vector<int> index;
    int n_big = 0, root = PT_BVH_NONE, depth = 0;
    float root_hinv = 0.f, root_lo[3] = {0, 0, 0}, root_hi[3] = {0, 0, 0}, rmax2 = 0.f;
    PtBvh view(const f4 *nodes_p, const f4 *geom_p, const int *index_p) const {
        PtBvh B;
        B.nodes = nodes_p; B.geom = geom_p; B.index = index_p;
        B.n_big = n_big; B.root = root; B.root_hinv = root_hinv; B.rmax2 = rmax2;
        for (int k = 0; k < 3; k++) { B.root_lo[k] = root_lo[k]; B.root_hi[k] = root_hi[k]; }
        return B;
    }
};

namespace bvh_detail {
struct Box { double lo[3], hi[3]; };
inline void grow(Box &b, const Box &o) { for (int k = 0; k < 3; k++) { b.lo[k] = std::min(b.lo[k], o.lo[k]); b.hi[k] = std::max(b.hi[k], o.hi[k]); } }
inline Box empty() { Box b; for (int k = 0; k < 3; k++) { b.lo[k] = DBL_MAX; b.hi[k] = -DBL_MAX; } return b; }
inline double area(const Box &b) { const double x = b.hi[0] - b.lo[0], y = b.hi[1] - b.lo[1], z = b.hi[2] - b.lo[2]; return x * y + y * z + z * x; }
inline float down(double v) { float f = (float)v; if ((double)f > v) f = std::nextafterf(f, -INFINITY); return f; }
inline float up(double v) { float f = (float)v; if ((double)f < v) f = std::nextafterf(f, INFINITY); return f; }
inline float bitsf(int i) { float f; memcpy(&f, &i, 4); return f; }
struct Item { Box box; double c[3]; double rad; int id; };
struct Sub { int code; Box box; double rmin; };

struct Builder {
    std::vector<Item> &items;
    PtBvhHost &out;
    int max_depth = 0;
    Builder(std::vector<Item> &it, PtBvhHost &o) : items(it), out(o) {}
    Sub leaf(int lo, int hi) {
        Sub s; s.box = empty(); s.rmin = DBL_MAX;
        const int first = (int)out.geom.size();
        for (int i = lo; i < hi; i++) { grow(s.box, items[i].box); s.rmin = std::min(s.rmin, items[i].rad); out.index.push_back(items[i].id); out.geom.push_back(f4{0, 0, 0, 0}); }
        s.code = pt_bvh_leaf_code(first, hi - lo);
        return s;
    }
    Sub build(int lo, int hi, int depth) {
        max_depth = std::max(max_depth, depth);
        if (hi - lo <= PT_BVH_LEAF_MAX) return leaf(lo, hi);
        Box cb = empty();
        for (int i = lo; i < hi; i++) for (int k = 0; k < 3; k++) { cb.lo[k] = std::min(cb.lo[k], items[i].c[k]); cb.hi[k] = std::max(cb.hi[k], items[i].c[k]); }
        int axis = 0;
        for (int k = 1; k < 3; k++) if (cb.hi[k] - cb.lo[k] > cb.hi[axis] - cb.lo[axis]) axis = k;
        int mid = (lo + hi) / 2;
        bool median = depth >= 24 || !(cb.hi[axis] > cb.lo[axis]);
        if (!median) {
            const int NB = 16;
            double best = DBL_MAX; int best_axis = -1, best_bin = -1;
            for (int ax = 0; ax < 3; ax++) {
                const double ext = cb.hi[ax] - cb.lo[ax];
                if (!(ext > 0)) continue;
                Box bb[NB]; int cnt[NB];
                for (int b = 0; b < NB; b++) { bb[b] = empty(); cnt[b] = 0; }
                for (int i = lo; i < hi; i++) {
                    int b = (int)((items[i].c[ax] - cb.lo[ax]) / ext * NB); if (b >= NB) b = NB - 1; if (b < 0) b = 0;
                    grow(bb[b], items[i].box); cnt[b]++;
                }
                double la[NB], ra[NB]; int lc[NB], rc[NB];
                Box acc = empty(); int c = 0;
                for (int b = 0; b < NB; b++) { if (cnt[b]) grow(acc, bb[b]); c += cnt[b]; la[b] = c ? area(acc) : 0; lc[b] = c; }
                acc = empty(); c = 0;
                for (int b = NB - 1; b >= 0; b--) { if (cnt[b]) grow(acc, bb[b]); c += cnt[b]; ra[b] = c ? area(acc) : 0; rc[b] = c; }
                for (int b = 0; b + 1 < NB; b++) {
                    if (!lc[b] || !rc[b + 1]) continue;
                    const double cost = la[b] * lc[b] + ra[b + 1] * rc[b + 1];
                    if (cost < best) { best = cost; best_axis = ax; best_bin = b; }
                }
            }
            if (best_axis < 0) median = true;
            else {
                const double ext = cb.hi[best_axis] - cb.lo[best_axis];
                auto it = std::partition(items.begin() + lo, items.begin() + hi, [&](const Item &a) {
                    int b = (int)((a.c[best_axis] - cb.lo[best_axis]) / ext * NB); if (b >= NB) b = NB - 1; if (b < 0) b = 0;
                    return b <= best_bin; });
                mid = (int)(it - items.begin());
                if (mid <= lo || mid >= hi) median = true;
            }
        }
        if (median) {
            mid = (lo + hi) / 2;
            std::nth_element(items.begin() + lo, items.begin() + mid, items.begin() + hi, [&](const Item &a, const Item &b) { return a.c[axis] < b.c[axis]; });
        }
        const int me = (int)out.nodes.size() / 4;
        out.nodes.resize(out.nodes.size() + 4);
        const Sub a = build(lo, mid, depth + 1), b = build(mid, hi, depth + 1);
        auto hinv = [](double rmin) { const double h = rmin > 0 ? 0.5 / rmin : 1e6; return (float)std::min(h * 1.000001, 1e6); };
        out.nodes[4 * me + 0] = f4{ down(a.box.lo[0]), up(a.box.hi[0]), down(a.box.lo[1]), up(a.box.hi[1]) };
        out.nodes[4 * me + 1] = f4{ down(b.box.lo[0]), up(b.box.hi[0]), down(b.box.lo[1]), up(b.box.hi[1]) };
        out.nodes[4 * me + 2] = f4{ down(a.box.lo[2]), up(a.box.hi[2]), down(b.box.lo[2]), up(b.box.hi[2]) };
        out.nodes[4 * me + 3] = f4{ bitsf(a.code), bitsf(b.code), hinv(a.rmin), hinv(b.rmin) };
        Sub s; s.code = me; s.box = a.box; grow(s.box, b.box); s.rmin = std::min(a.rmin, b.rmin);
        return s;
    }
};
}  // namespace bvh_detail

// geom[i] = (p.x, p.y, p.z, rad*rad) and rad[i] as in PtSoA (colr[i].w).
// skip (optional): entries with skip[i] != 0 are left out altogether (neither in the tree nor in the always-tested list).
inline void build_pt_bvh(const std::vector<f4> &geom, const std::vector<f4> &colr, PtBvhHost &out, const std::vector<char> *skip = nullptr) {
    using namespace bvh_detail;
    const int n = (int)geom.size();
    out = PtBvhHost();
    std::vector<double> radii;
    std::vector<char> finite((size_t)n);
    for (int i = 0; i < n; i++) {
        if (skip && (*skip)[i]) { finite[i] = 0; continue; }
        const double r = std::fabs((double)colr[i].w);
        finite[i] = std::isfinite(geom[i].x) && std::isfinite(geom[i].y) && std::isfinite(geom[i].z) && std::isfinite(geom[i].w) && std::isfinite(r) &&
                    std::fabs(geom[i].x) < 1e15f && std::fabs(geom[i].y) < 1e15f && std::fabs(geom[i].z) < 1e15f && r < 1e15;
        if (finite[i]) radii.push_back(r);
    }
    double big = INFINITY;
    if (!radii.empty()) {
        std::nth_element(radii.begin(), radii.begin() + radii.size() / 2, radii.end());
        big = 64.0 * radii[radii.size() / 2];
    }
    std::vector<Item> items;
    std::vector<int> bigs;
    for (int i = 0; i < n; i++) {
        if (skip && (*skip)[i]) continue;
        const double r = std::fabs((double)colr[i].w);
        // the test reads rad2 = geom.w = fl(rad*rad): its radius is sqrt(rad2); the box also covers the `rad` field
        const double reff = std::sqrt(std::max(0.0, (double)geom[i].w));
        const double rr = std::max(r, reff) * (1.0 + 1e-6);
        if (!finite[i] || rr > big) { bigs.push_back(i); continue; }
        Item it; it.id = i; it.rad = reff * (1.0 - 2e-6);
        for (int k = 0; k < 3; k++) { const double c = k == 0 ? geom[i].x : (k == 1 ? geom[i].y : geom[i].z); it.c[k] = c; it.box.lo[k] = c - rr; it.box.hi[k] = c + rr; }
        items.push_back(it);
    }
    // always-tested list, descending index like the reference's scan (any order gives the same result)
    std::sort(bigs.begin(), bigs.end(), [](int a, int b) { return a > b; });
    for (int i : bigs) { out.index.push_back(i); out.geom.push_back(geom[i]); }
    out.n_big = (int)bigs.size();
    if (!items.empty()) {
        double rmax = 0;
        for (const Item &it : items) rmax = std::max(rmax, std::max(it.box.hi[0] - it.c[0], it.rad));
        out.rmax2 = up(rmax * rmax * (1.0 + 1e-6));
        Builder B(items, out);
        const Sub root = B.build(0, (int)items.size(), 0);
        out.root = root.code; out.depth = B.max_depth;
        out.root_hinv = (float)std::min((root.rmin > 0 ? 0.5 / root.rmin : 1e6) * 1.000001, 1e6);
        for (int k = 0; k < 3; k++) { out.root_lo[k] = down(root.box.lo[k]); out.root_hi[k] = up(root.box.hi[k]); }
        for (size_t j = (size_t)out.n_big; j < out.index.size(); j++) out.geom[j] = geom[out.index[j]];
    }
    if (out.nodes.empty()) out.nodes.push_back(f4{0, 0, 0, 0});      // never read; keeps uploads non-empty
}

}  // namespace rtb
