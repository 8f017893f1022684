// whitted_bvh.cuh -- the exact hierarchy of pt_bvh.cuh under the Whitted tracer's sphere test (host/device).
//
// For scenes with hundreds or thousands of spheres (rt_whitted_from_spheres on the generated .scn scenes) the non-light
// spheres sit in the same conservative bounding-volume hierarchy as the path tracer's; planes, lights and spheres the tree
// does not take (huge, not finite) stay in the run tables and are tested by every query first, as before.  The sphere
// arithmetic is sphere_intersect (RNO:111-148) operation for operation -- v = o - c, b = -(v.d), det = (b*b - v.v) + r^2,
// the same roundings as the path tracer's test up to the sign of v, so the node margins of pt_bvh.cuh apply unchanged
// (accepted distance in {b - sqrt(det), b + sqrt(det)}, candidate only if det > 0).  Winner of a nearest query: the
// reference scans ascending with a strict '<' (RNO:185-193), i.e. the smallest distance and on a tie the LOWEST index;
// in any visiting order that is "t < cumu, or t == cumu and index < current".  A shadow ray only asks whether some
// non-light primitive has 0 < dist < distance to the light (RNO:232-240): order-free.
#pragma once
#include "whitted_lane.cuh"
#include "pt_bvh.cuh"

namespace rtb {

RT_HD void w_bvh_sphere_nearest(WLane &L, const f4 g, int idx) {
    const float vx = f_sub(L.qox, g.x), vy = f_sub(L.qoy, g.y), vz = f_sub(L.qoz, g.z);
    const float b = -dot3(vx, vy, vz, L.qdx, L.qdy, L.qdz);
    const float det = f_add(f_sub(f_mul(b, b), dot3(vx, vy, vz, vx, vy, vz)), g.w);
    if (!(det > 0.f)) return;
    const float sq = f_sqrt(det);
    const float i1 = f_sub(b, sq), i2 = f_add(b, sq);
    if (!(i2 > 0.f)) return;
    const bool inside = i1 < 0.f;
    const float t = inside ? i2 : i1;
    if (t < L.cumu || (t == L.cumu && idx < L.qhit)) { L.cumu = t; L.qhit = idx; L.qkind = inside ? -1 : 1; }
}

RT_HD bool w_bvh_sphere_blocks(float ox, float oy, float oz, float dx, float dy, float dz, float reach, const f4 g) {
    const float vx = f_sub(ox, g.x), vy = f_sub(oy, g.y), vz = f_sub(oz, g.z);
    const float b = -dot3(vx, vy, vz, dx, dy, dz);
    const float det = f_add(f_sub(f_mul(b, b), dot3(vx, vy, vz, vx, vy, vz)), g.w);
    if (!(det > 0.f)) return false;
    const float sq = f_sqrt(det);
    const float i1 = f_sub(b, sq), i2 = f_add(b, sq);
    return (i2 > 0.f) & ((i1 < 0.f ? i2 : i1) < reach);
}

// Nearest query of one lane over the tree, continuing from what the run tables found (L.cumu, L.qhit, L.qkind).
RT_HD void w_bvh_nearest(WLane &L, const PtBvh &B) {
    if (B.root == PT_BVH_NONE) return;
    int stack[PT_BVH_STACK];
    float stack_t[PT_BVH_STACK];
    PtTrav T;
    T.sp = 0;
    T.R = bvh_ray(L.qox, L.qoy, L.qoz, L.qdx, L.qdy, L.qdz, B);
    T.node = bvh_root(L.qox, L.qoy, L.qoz, L.cumu, B, T.R);
    while (T.node != PT_BVH_DONE) {
        if (pt_bvh_at_inner(T)) bvh_inner(L.qox, L.qoy, L.qoz, L.cumu, B, T, stack, stack_t);
        else {
            const int code = ~T.node, first = code >> 3, count = (code & 7) + 1;
            for (int j = 0; j < count; j++) w_bvh_sphere_nearest(L, B.geom[first + j], B.index[first + j]);
            bvh_pop(L.cumu, T, stack, stack_t);
        }
    }
}

// Is shadow ray (o, d, reach) blocked by a sphere of the tree?
RT_HD bool w_bvh_blocked(float ox, float oy, float oz, float dx, float dy, float dz, float reach, const PtBvh &B) {
    if (B.root == PT_BVH_NONE) return false;
    int stack[PT_BVH_STACK];
    float stack_t[PT_BVH_STACK];
    PtTrav T;
    T.sp = 0;
    T.R = bvh_ray(ox, oy, oz, dx, dy, dz, B);
    T.node = bvh_root(ox, oy, oz, reach, B, T.R);
    while (T.node != PT_BVH_DONE) {
        if (pt_bvh_at_inner(T)) bvh_inner(ox, oy, oz, reach, B, T, stack, stack_t);
        else {
            const int code = ~T.node, first = code >> 3, count = (code & 7) + 1;
            for (int j = 0; j < count; j++)
                if (w_bvh_sphere_blocks(ox, oy, oz, dx, dy, dz, reach, B.geom[first + j])) return true;
            bvh_pop(reach, T, stack, stack_t);
        }
    }
    return false;
}

// The tree part of a shadow round: the lane's (up to) three shadow rays that the run tables left unblocked.
RT_HD void w_bvh_shadow(WLane &L, const PtBvh &B) {
#pragma unroll 1
    for (int k = 0; k < W_SHADOW_BATCH; k++) {
        if (k >= L.ns || ((L.sblk >> k) & 1)) continue;
        const float ox = k == 0 ? L.sox[0] : (k == 1 ? L.sox[1] : L.sox[2]), oy = k == 0 ? L.soy[0] : (k == 1 ? L.soy[1] : L.soy[2]);
        const float oz = k == 0 ? L.soz[0] : (k == 1 ? L.soz[1] : L.soz[2]), dx = k == 0 ? L.slx[0] : (k == 1 ? L.slx[1] : L.slx[2]);
        const float dy = k == 0 ? L.sly[0] : (k == 1 ? L.sly[1] : L.sly[2]), dz = k == 0 ? L.slz[0] : (k == 1 ? L.slz[1] : L.slz[2]);
        const float reach = k == 0 ? L.sreach[0] : (k == 1 ? L.sreach[1] : L.sreach[2]);
        if (w_bvh_blocked(ox, oy, oz, dx, dy, dz, reach, B)) L.sblk |= 1 << k;
    }
}

}  // namespace rtb
