// r306_lane.cuh -- per-lane state machine of the raytracer3.0.06 frame (BASELINE config 1; host/device).
//
// Follows R306/raytracer.cpp ("R306:"): Engine_Render (:301-530) and Engine_Raytrace (:30-271) over
// Primitive_Intersect (R306/scene.cpp:125-190).  The intersection arithmetic is the one of the 3.2.03 tracer
// (same expressions, same order), so the ray queries are the loops of whitted_lane.cuh; what differs is everything
// around them, and all of it is part of the result:
//   * no recursion and no queue: every sub-sample expands an implicit binary tree of 63 nodes in index order
//     (node i spawns its reflection into 2i+1 and its refraction into 2i+2, R306:398-465), each node traced by one
//     Engine_Raytrace call at depth 1, and the colours are folded bottom-up afterwards (:468-503): refraction child
//     first (Beer's law with the PARENT's hit distance), then the reflection child times the parent's colour and
//     m_Refl;
//   * a refraction child is traced whenever the parent's material has m_Refr > 0, even if the parent call found
//     total internal reflection and wrote no ray: it then re-traces whatever the caller's `refr_Ray` variable still
//     holds from an earlier call of the same sub-sample (or the camera ray), and its colour is added without
//     Beer's law (:435, :474).  Reproduced: L.vr is that variable;
//   * the refraction index handed down is the parent's *a_RIndex after the call (:219-220), a light adds (1,1,1)
//     (:55), "infinity" is 1e6 (:33), shading multiplies (c_prim * c_light) * diff and uses powf (:142-183), the
//     screen position is a running float sum (m_SX += m_DX, :523-525; host-built tables sx[], sy[]), rows 20 ..
//     h-71 only (:281, :309), and a pixel is 0x00RRGGBB (:520).
#pragma once
#include "whitted_lane.cuh"

namespace rtb {

#define R306_FAR 1000000.0f
#define R306_NODES 63
#define R306_PARENTS 31          /* nodes 31..62 are leaves: their children are never traced */

struct R306Frame {
    WFrame W;                     // geom / flags / runs / lights / mat_a = (colour, m_Refl) / mat_b = (m_Diff, m_Refr, m_RIndex, m_Spec) / rrad
    const float *sx, *sy;         // sx[x] = m_SX at column x, sy[y] = m_SY at row y (running sums, built on the host)
    int row0, row1;               // rows [row0, row1) are rendered: 20 and h - 70
};

// The per-sub-sample tree (lives in local memory: ~3 KB per lane).
struct R306Tree {
    float col[R306_NODES][3];
    float refl[R306_PARENTS], refr[R306_PARENTS], dist[R306_PARENTS], rindex[R306_PARENTS];
    int refl_idx[R306_PARENTS], refr_idx[R306_PARENTS];
    float refl_ray[R306_PARENTS][6], refr_ray[R306_PARENTS][6];
};

struct R306Lane {
    WLane q;                      // query machinery and the fields of the call in flight: (dx,dy,dz) ray direction, (px,py,pz) = pi,
                                  //  hit / hkind / dist, (cr,cg,cb) = a_Acc, li, phase, shadow batch
    int node;
    float ox, oy, oz;             // origin of the ray in flight
    float tr, tg, tb;             // total_acc of the pixel
    float in_rindex;              // *a_RIndex of the call in flight
    float vr[6];                  // Engine_Render's `refr_Ray` variable (origin, direction)
    unsigned long long traced;    // bit i: node i of the current sub-sample was traced (its record in the tree is valid)
};

RT_HD void r306_start_query(R306Lane &L) {
    WLane &q = L.q;
    q.qox = L.ox; q.qoy = L.oy; q.qoz = L.oz; q.qdx = q.dx; q.qdy = q.dy; q.qdz = q.dz;
    q.cumu = R306_FAR; q.qhit = -1; q.qkind = 0; q.phase = PH_NEAREST;
    q.cr = q.cg = q.cb = 0.f;                    // acc = tr_Color[i], which is still 0 when node i is traced
}

// R306:356-397: primary ray of sub-sample q.sub, node 0.
RT_HD void r306_start_subsample(R306Lane &L, const R306Frame &F) {
    WLane &q = L.q;
    const int tx = q.sub / 3 - 1, ty = q.sub % 3 - 1;
    const float SX = F.sx[q.x], SY = F.sy[q.y];
    float dx = f_sub(f_add(SX, f_mul(f_mul(F.W.DX, (float)tx), 0.5f)), 0.f);      // ( m_SX + m_DX * tx / 2.0f ) - camera.x
    float dy = f_sub(f_add(SY, f_mul(f_mul(F.W.DY, (float)ty), 0.5f)), 0.25f);
    float dz = f_sub(0.f, -7.0f);
    const float l = f_rcp(f_sqrt(f_add(f_add(f_mul(dx, dx), f_mul(dy, dy)), f_mul(dz, dz))));   // NORMALIZE, R306/common.h:19
    q.dx = f_mul(dx, l); q.dy = f_mul(dy, l); q.dz = f_mul(dz, l);
    L.ox = 0.f; L.oy = 0.25f; L.oz = -7.0f;
    L.vr[0] = L.ox; L.vr[1] = L.oy; L.vr[2] = L.oz; L.vr[3] = q.dx; L.vr[4] = q.dy; L.vr[5] = q.dz;    // refr_Ray = camera ray (:374-375)
    L.node = 0; L.in_rindex = 1.0f; L.traced = 1ull;
    r306_start_query(L);
}

RT_HD void r306_begin_pixel(R306Lane &L, const R306Frame &F, int x, int y) {
    L.q.x = x; L.q.y = y; L.q.sub = 0; L.tr = L.tg = L.tb = 0.f;
    r306_start_subsample(L, F);
}
// One sub-sample as a work unit of its own (r306_kernel<SPLIT>): Engine_Render folds each sub-sample's tree into one colour
// and adds the nine colours to the pixel in order (R306:504-506), so the nine trees are independent computations.  The
// lane delivers 0 + colour (L.tr after one sub-sample: exactly what the reference's total holds after its first addition)
// and r306_resolve_kernel adds the nine values in the reference's order.
RT_HD void r306_begin_subsample(R306Lane &L, const R306Frame &F, int x, int y, int sub) {
    L.q.x = x; L.q.y = y; L.q.sub = sub; L.tr = L.tg = L.tb = 0.f;
    r306_start_subsample(L, F);
}

// powf(v, 20) of R306:172 for v > 0 (glibc's algorithm, rt_math.cuh); a subnormal base underflows to 0 like in glibc.
RT_HD float r306_pow20(float v) { return v < 0x1p-126f ? 0.f : powf_glibc_unit(v, 20.0f); }

// Diffuse + specular of light l with unit vector (Lx,Ly,Lz) towards it, shade = 1 (R306:123-186).
RT_HD void r306_shade(R306Lane &L, const R306Frame &F, int l, float Lx, float Ly, float Lz) {
    WLane &q = L.q;
    const f4 ma = F.W.mat_a[q.hit], mb = F.W.mat_b[q.hit], lc = F.W.mat_a[l];
    float nx, ny, nz;
    w_normal(F.W, q.hit, q.px, q.py, q.pz, nx, ny, nz);
    if (mb.x > 0.f) {
        const float d = dot3(Lx, Ly, Lz, nx, ny, nz);
        if (d > 0.f) {
            const float diff = f_mul(f_mul(d, mb.x), 1.0f);
            q.cr = f_add(q.cr, f_mul(f_mul(ma.x, lc.x), diff));
            q.cg = f_add(q.cg, f_mul(f_mul(ma.y, lc.y), diff));
            q.cb = f_add(q.cb, f_mul(f_mul(ma.z, lc.z), diff));
        }
    }
    if (mb.w > 0.f) {
        const float k2 = f_mul(2.0f, dot3(Lx, Ly, Lz, nx, ny, nz));
        const float rx = f_sub(Lx, f_mul(k2, nx)), ry = f_sub(Ly, f_mul(k2, ny)), rz = f_sub(Lz, f_mul(k2, nz));
        const float vr = dot3(q.dx, q.dy, q.dz, rx, ry, rz);
        if (vr > 0.f) {
            const float spec = f_mul(f_mul(r306_pow20(vr), mb.w), 1.0f);
            q.cr = f_add(q.cr, f_mul(spec, lc.x));
            q.cg = f_add(q.cg, f_mul(spec, lc.y));
            q.cb = f_add(q.cb, f_mul(spec, lc.z));
        }
    }
}

// Unit vector to light l as the shading code forms it (R306:125-135): 0 when the light sits on the point.
RT_HD void r306_light_vector(const R306Frame &F, const WLane &q, int l, float &Lx, float &Ly, float &Lz, float &len) {
    const f4 lg = F.W.geom[l];
    const float ex = f_sub(lg.x, q.px), ey = f_sub(lg.y, q.py), ez = f_sub(lg.z, q.pz);
    len = f_sqrt(f_add(f_add(f_mul(ex, ex), f_mul(ey, ey)), f_mul(ez, ez)));
    const float inv = f_rcp(len);
    Lx = f_mul(ex, inv); Ly = f_mul(ey, inv); Lz = f_mul(ez, inv);
}

// Next batch of shadow rays (lights in index order, R306:69-118); a light that is not a sphere casts no shadow ray and
// is shaded on the spot, in order.  Sets PH_FINAL when no light is left.
RT_HD void r306_next_shadow_batch(R306Lane &L, const R306Frame &F) {
    WLane &q = L.q;
    for (;;) {
        if (q.li >= F.W.n_lights) { q.phase = PH_FINAL; return; }
        const int l = F.W.lights[q.li];
        if (F.W.flags[l] & W_FLAG_SPHERE) break;
        const f4 lg = F.W.lcenter[q.li];                           // m_Centre of a light that is not a sphere (R306:123-135)
        const float ex = f_sub(lg.x, q.px), ey = f_sub(lg.y, q.py), ez = f_sub(lg.z, q.pz);
        const float len = f_sqrt(f_add(f_add(f_mul(ex, ex), f_mul(ey, ey)), f_mul(ez, ez)));
        const float inv = f_rcp(len);
        float Lx = f_mul(ex, inv), Ly = f_mul(ey, inv), Lz = f_mul(ez, inv);
        if (!(len > 0.f)) Lx = Ly = Lz = 0.f;
        r306_shade(L, F, l, Lx, Ly, Lz);
        q.li++;
    }
    q.ns = 0; q.sblk = 0;
#pragma unroll
    for (int k = 0; k < W_SHADOW_BATCH; k++) {
        if (q.ns == k && q.li + k < F.W.n_lights && (F.W.flags[F.W.lights[q.li + k]] & W_FLAG_SPHERE)) {
            float Lx, Ly, Lz, len;
            r306_light_vector(F, q, F.W.lights[q.li + k], Lx, Ly, Lz, len);
            q.sox[k] = f_add(q.px, f_mul(Lx, W_EPS)); q.soy[k] = f_add(q.py, f_mul(Ly, W_EPS)); q.soz[k] = f_add(q.pz, f_mul(Lz, W_EPS));
            q.slx[k] = Lx; q.sly[k] = Ly; q.slz[k] = Lz; q.sreach[k] = len;
            q.ns = k + 1;
        }
    }
    q.phase = PH_SHADOW;
}

// Node `i` is complete: record what its children and the fold need (R306:386-397, :417-431, :447-461).
RT_HD void r306_store_node(R306Lane &L, R306Tree &T, float refl, int refl_idx, const float *refl_ray, float refr, int refr_idx) {
    const WLane &q = L.q;
    const int i = L.node;
    RT_CHECK(i >= 0 && i < R306_NODES, RT_CHK_TREE);
    T.col[i][0] = q.cr; T.col[i][1] = q.cg; T.col[i][2] = q.cb;
    if (i < R306_PARENTS) {
        T.refl[i] = refl; T.refl_idx[i] = refl_idx; T.refr[i] = refr; T.refr_idx[i] = refr_idx;
        T.rindex[i] = L.in_rindex; T.dist[i] = q.dist;
#pragma unroll
        for (int k = 0; k < 6; k++) { T.refl_ray[i][k] = refl_ray ? refl_ray[k] : 0.f; T.refr_ray[i][k] = L.vr[k]; }
    }
}

// After the nearest-hit round (R306:52-66).  Returns true when the node is already complete (miss or light).
RT_HD bool r306_after_nearest(R306Lane &L, const R306Frame &F, R306Tree &T) {
    WLane &q = L.q;
    q.dist = q.cumu; q.hit = q.qhit; q.hkind = q.qkind;
    if (q.hit < 0) { r306_store_node(L, T, 0.f, -1, nullptr, 0.f, -1); return true; }          // no hit: return -1, nothing written
    if (F.W.flags[q.hit] & W_FLAG_LIGHT) {                                                      // a light: a_Acc += 1
        q.cr = f_add(q.cr, 1.f); q.cg = f_add(q.cg, 1.f); q.cb = f_add(q.cb, 1.f);
        r306_store_node(L, T, 0.f, -1, nullptr, 0.f, -1);
        return true;
    }
    q.px = f_add(f_mul(q.qdx, q.dist), q.qox);
    q.py = f_add(f_mul(q.qdy, q.dist), q.qoy);
    q.pz = f_add(f_mul(q.qdz, q.dist), q.qoz);
    q.li = 0;
    r306_next_shadow_batch(L, F);
    return false;
}

RT_HD void r306_after_shadow(R306Lane &L, const R306Frame &F) {
    WLane &q = L.q;
#pragma unroll 1
    for (int k = 0; k < q.ns; k++) {
        float Lx = k == 0 ? q.slx[0] : (k == 1 ? q.slx[1] : q.slx[2]);
        float Ly = k == 0 ? q.sly[0] : (k == 1 ? q.sly[1] : q.sly[2]);
        float Lz = k == 0 ? q.slz[0] : (k == 1 ? q.slz[1] : q.slz[2]);
        const float len = k == 0 ? q.sreach[0] : (k == 1 ? q.sreach[1] : q.sreach[2]);
        if (!(len > 0.f)) Lx = Ly = Lz = 0.f;
        if (!((q.sblk >> k) & 1)) r306_shade(L, F, F.W.lights[q.li + k], Lx, Ly, Lz);
    }
    q.li += q.ns;
    r306_next_shadow_batch(L, F);
}

// The lights are done: refraction (R306:191-251), then reflection (:255-282), then the node record.
RT_HD void r306_finish_hit(R306Lane &L, const R306Frame &F, R306Tree &T) {
    WLane &q = L.q;
    const f4 ma = F.W.mat_a[q.hit], mb = F.W.mat_b[q.hit];
    float gx, gy, gz;
    w_normal(F.W, q.hit, q.px, q.py, q.pz, gx, gy, gz);
    const float refr = mb.y;
    int refr_idx = -1;
    if (refr > 0.f) {                                           // a_Depth is always 1 < TRACEDEPTH
        const float rindex = mb.z;
        const float n = f_div(L.in_rindex, rindex);
        L.in_rindex = rindex;
        const float sgn = (float)q.hkind;
        const float nx = f_mul(gx, sgn), ny = f_mul(gy, sgn), nz = f_mul(gz, sgn);
        const float cosI = -dot3(nx, ny, nz, q.dx, q.dy, q.dz);
        const float cosT2 = f_sub(1.0f, f_mul(f_mul(n, n), f_sub(1.0f, f_mul(cosI, cosI))));
        if (cosT2 > 0.0f) {
            const float kk = f_sub(f_mul(n, cosI), f_sqrt(cosT2));
            const float tx = f_add(f_mul(n, q.dx), f_mul(kk, nx)), ty = f_add(f_mul(n, q.dy), f_mul(kk, ny)), tz = f_add(f_mul(n, q.dz), f_mul(kk, nz));
            L.vr[0] = f_add(q.px, f_mul(tx, W_EPS)); L.vr[1] = f_add(q.py, f_mul(ty, W_EPS)); L.vr[2] = f_add(q.pz, f_mul(tz, W_EPS));
            L.vr[3] = tx; L.vr[4] = ty; L.vr[5] = tz;
            refr_idx = q.hit;
        }
    }
    const float refl = ma.w;
    int refl_idx = -1;
    float rr[6] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
    if (refl > 0.0f) {
        const float k2 = f_mul(2.0f, dot3(q.dx, q.dy, q.dz, gx, gy, gz));
        const float rx = f_sub(q.dx, f_mul(k2, gx)), ry = f_sub(q.dy, f_mul(k2, gy)), rz = f_sub(q.dz, f_mul(k2, gz));
        rr[0] = f_add(q.px, f_mul(rx, W_EPS)); rr[1] = f_add(q.py, f_mul(ry, W_EPS)); rr[2] = f_add(q.pz, f_mul(rz, W_EPS));
        rr[3] = rx; rr[4] = ry; rr[5] = rz;
        refl_idx = q.hit;
    }
    r306_store_node(L, T, refl, refl_idx, rr, refr, refr_idx);
}

// Bottom-up fold of the finished tree into node 0 (R306:468-503).  A node that was not traced holds colour 0 and no
// children in the reference's tables; here its record is simply never written or read (L.traced says which are valid): a
// pair below an untraced parent is skipped -- nothing ever reads that parent -- and an untraced child of a traced parent
// contributes the same +0 the reference adds.  Same operations on the same values; 95 % less local-memory traffic (most
// sub-samples trace node 0 only, and the old code cleared and re-read all 63 records: 5 GB of DRAM writes per frame).
RT_HD void r306_fold(const R306Frame &F, R306Tree &T, unsigned long long traced) {
    for (int i = R306_NODES - 1; i >= 2; i -= 2) {
        const int p = (i - 1) / 2;
        if (!((traced >> p) & 1ull)) continue;
        const bool ti = (traced >> i) & 1ull, tj = (traced >> (i - 1)) & 1ull;
        float ar = ti ? T.col[i][0] : 0.f, ag = ti ? T.col[i][1] : 0.f, ab = ti ? T.col[i][2] : 0.f;
        if (T.refr_idx[p] > -1 && T.refr[p] > 0.f) {
            const f4 c = F.W.mat_a[T.refr_idx[p]];
            const float nd = -T.dist[p];
            ar = f_mul(ar, expf_glibc(f_mul(f_mul(c.x, 0.15f), nd)));
            ag = f_mul(ag, expf_glibc(f_mul(f_mul(c.y, 0.15f), nd)));
            ab = f_mul(ab, expf_glibc(f_mul(f_mul(c.z, 0.15f), nd)));
        }
        T.col[p][0] = f_add(T.col[p][0], ar); T.col[p][1] = f_add(T.col[p][1], ag); T.col[p][2] = f_add(T.col[p][2], ab);
        float br = tj ? T.col[i - 1][0] : 0.f, bg = tj ? T.col[i - 1][1] : 0.f, bb = tj ? T.col[i - 1][2] : 0.f;
        if (T.refl_idx[p] > -1 && T.refl[p] > 0.f) {
            const f4 c = F.W.mat_a[T.refl_idx[p]];
            br = f_mul(f_mul(br, c.x), T.refl[p]); bg = f_mul(f_mul(bg, c.y), T.refl[p]); bb = f_mul(f_mul(bb, c.z), T.refl[p]);
        }
        T.col[p][0] = f_add(T.col[p][0], br); T.col[p][1] = f_add(T.col[p][1], bg); T.col[p][2] = f_add(T.col[p][2], bb);
    }
}

// Moves on to the next node that is traced (R306:398-465), the next sub-sample, or the end of the pixel (returns true).
RT_HD bool r306_next_node(R306Lane &L, const R306Frame &F, R306Tree &T, bool one_sub = false) {
    WLane &q = L.q;
    for (;;) {
        const int i = ++L.node;
        RT_CHECK(i >= 0 && i <= R306_NODES, RT_CHK_TREE);
        if (i >= R306_NODES) break;
        const int p = (i - 1) / 2;
        if (!((L.traced >> p) & 1ull)) continue;                      // the parent was not traced: neither is this node
        const bool traced = (i & 1) ? (T.refl[p] > 0.f) : (T.refr[p] > 0.f);
        if (traced) {
            const float *ray = (i & 1) ? T.refl_ray[p] : T.refr_ray[p];
            L.ox = ray[0]; L.oy = ray[1]; L.oz = ray[2]; q.dx = ray[3]; q.dy = ray[4]; q.dz = ray[5];
            L.in_rindex = T.rindex[p];
            L.traced |= 1ull << i;
            r306_start_query(L);
            return false;
        }
    }
    r306_fold(F, T, L.traced);
    L.tr = f_add(L.tr, T.col[0][0]); L.tg = f_add(L.tg, T.col[0][1]); L.tb = f_add(L.tb, T.col[0][2]);
    q.sub++;
    if (!one_sub && q.sub < 9) { r306_start_subsample(L, F); return false; }
    q.phase = PH_IDLE;
    return true;
}

// R306:512-520 with x86's float -> int conversion.
RT_HD uint32_t r306_pack_pixel(float r, float g, float b) {
    int ir = x86_float_to_int(f_mul(r, 28.0f)), ig = x86_float_to_int(f_mul(g, 28.0f)), ib = x86_float_to_int(f_mul(b, 28.0f));
    if (ir > 255) ir = 255;
    if (ig > 255) ig = 255;
    if (ib > 255) ib = 255;
    return ((uint32_t)ir << 16) + ((uint32_t)ig << 8) + (uint32_t)ib;
}

}  // namespace rtb
