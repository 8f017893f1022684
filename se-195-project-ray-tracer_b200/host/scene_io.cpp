// scene_io.cpp -- host-side pieces the north star keeps from the reference's host programs:
// the .scn loader and camera set-up of smallptGPU, the scene tables of the Whitted tracer, the
// complex-scene generator and the two image writers.  Pure CPU code, part of librt_b200.so; none of
// it is on the rendering path.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/rt_b200.h"

namespace {

inline float dot(const rt_vec &a, const rt_vec &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline rt_vec unit(rt_vec v) {                      // vnorm, SPT/vec.h:41 (float overload of sqrt in the /TP build)
    const float l = 1.f / sqrtf(dot(v, v));
    rt_vec r = { l * v.x, l * v.y, l * v.z };
    return r;
}
inline rt_vec cross(const rt_vec &a, const rt_vec &b) {   // vxcross, SPT/vec.h:42
    rt_vec r = { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x };
    return r;
}

// --- Whitted scene builders, R323/scene.c:6-46.  Unlike the reference (which leaves the fields a
// primitive type does not use as stack garbage) unused fields are zero here.
rt_primitive blank() { rt_primitive p; memset(&p, 0, sizeof p); return p; }
void material(rt_primitive &p, float r, float g, float b, float refl, float refr, float refr_index, float diff, float spec) {
    p.m_color.x = r; p.m_color.y = g; p.m_color.z = b;
    p.m_refl = refl; p.m_diff = diff; p.m_refr = refr; p.m_refr_index = refr_index; p.m_spec = spec;
}
rt_primitive plane(float r, float g, float b, float refl, float refr, float ri, float diff, float spec, bool light,
                   float nx, float ny, float nz, float depth) {
    rt_primitive p = blank();
    material(p, r, g, b, refl, refr, ri, diff, spec);
    p.type = RT_PLANE; p.is_light = light ? 1 : 0;
    p.normal.x = nx; p.normal.y = ny; p.normal.z = nz; p.depth = depth;
    return p;
}
rt_primitive sphere(float r, float g, float b, float refl, float refr, float ri, float diff, float spec, bool light,
                    float cx, float cy, float cz, float radius) {
    rt_primitive p = blank();
    material(p, r, g, b, refl, refr, ri, diff, spec);
    p.type = RT_SPHERE; p.is_light = light ? 1 : 0;
    p.center.x = cx; p.center.y = cy; p.center.z = cz;
    p.radius = radius; p.sq_radius = radius * radius; p.r_radius = 1.0f / radius;    // R323/scene.c:41-43
    return p;
}

// --- complex-scene generator: the recursion of SPT/scene_build_complex.pl in doubles, printed the way
// perl prints numbers (%.15g), so that the text equals the script's output.
struct ComplexGen {
    FILE *f; double max_depth; unsigned count;
    void emit(double depth, double x, double y, double z, double rad) {
        const double k = depth / max_depth;
        const double col1 = 0.75 * k, col2 = 0.75 * (1.0 - k);
        if (f) fprintf(f, "sphere %.15g %.15g %.15g %.15g 0 0 0 %.15g 0 %.15g 0\n", rad, x, y, z, col2, col1);
        count++;
    }
    void rec(double depth, double x, double y, double z, double rad, int dir) {
        if (!(depth <= max_depth)) return;
        emit(depth, x, y, z, rad);
        const double nr = rad / 2.0;
        if (dir != 0) rec(depth + 1.0, x - rad - nr, y, z, nr, 1);
        if (dir != 1) rec(depth + 1.0, x + rad + nr, y, z, nr, 0);
        if (dir != 2) rec(depth + 1.0, x, y - rad - nr, z, nr, 3);
        if (dir != 3) rec(depth + 1.0, x, y + rad + nr, z, nr, 2);
        if (dir != 4) rec(depth + 1.0, x, y, z - rad - nr, nr, 5);
        if (dir != 5) rec(depth + 1.0, x, y, z + rad + nr, nr, 4);
    }
};

}  // namespace

extern "C" {

void rt_update_camera(rt_camera *cam, int w, int h) {       // SPT/displayfunc.cpp:182-195
    rt_vec d = { cam->target.x - cam->orig.x, cam->target.y - cam->orig.y, cam->target.z - cam->orig.z };
    cam->dir = unit(d);
    const rt_vec up = { 0.f, 1.f, 0.f };
    const float fov = (M_PI / 180.f) * 45.f;                 // formed in double, rounded to float once
    const rt_vec cx = unit(cross(cam->dir, up));
    const float kx = w * fov / h;
    cam->x.x = kx * cx.x; cam->x.y = kx * cx.y; cam->x.z = kx * cx.z;
    const rt_vec cy = unit(cross(cam->x, cam->dir));
    cam->y.x = fov * cy.x; cam->y.y = fov * cy.y; cam->y.z = fov * cy.z;
}

// One key press of the reference viewer, SPT/displayfunc.cpp:250-420.  MOVE_STEP is a float; ROTATE_STEP is
// (2.f * M_PI / 180.f), a double, so the rotation keys evaluate in double and round once per assignment.
int rt_viewer_key(int key, rt_camera *cam, int w, int h, rt_sphere *spheres, uint32_t n, uint32_t *current_sphere) {
    if (!cam) return RT_ERR_ARG;
    const float MOVE_STEP = 10.0f;
    const double ROTATE_STEP = 2.f * M_PI / 180.f;
    auto slide = [&](rt_vec dir, float k) {                 // vsmul + two vadd
        const rt_vec d = { k * dir.x, k * dir.y, k * dir.z };
        cam->orig.x = cam->orig.x + d.x; cam->orig.y = cam->orig.y + d.y; cam->orig.z = cam->orig.z + d.z;
        cam->target.x = cam->target.x + d.x; cam->target.y = cam->target.y + d.y; cam->target.z = cam->target.z + d.z;
    };
    auto moved = [&]() { rt_update_camera(cam, w, h); return (int)RT_KEY_CAMERA; };
    auto sphere_key = [&](float dx, float dy, float dz) {
        if (!spheres || !current_sphere || *current_sphere >= n) return (int)RT_ERR_ARG;
        rt_sphere &s = spheres[*current_sphere];
        if (dx != 0.f) s.p.x += dx;
        if (dy != 0.f) s.p.y += dy;
        if (dz != 0.f) s.p.z += dz;
        return (int)RT_KEY_SCENE;
    };
    if (key >= RT_KEY_SPECIAL) {
        rt_vec t = { cam->target.x - cam->orig.x, cam->target.y - cam->orig.y, cam->target.z - cam->orig.z };
        switch (key - RT_KEY_SPECIAL) {
            case 101: t.y = t.y * cos(-ROTATE_STEP) + t.z * sin(-ROTATE_STEP); t.z = -t.y * sin(-ROTATE_STEP) + t.z * cos(-ROTATE_STEP); break;   // UP
            case 103: t.y = t.y * cos(ROTATE_STEP) + t.z * sin(ROTATE_STEP); t.z = -t.y * sin(ROTATE_STEP) + t.z * cos(ROTATE_STEP); break;       // DOWN
            case 100: t.x = t.x * cos(-ROTATE_STEP) - t.z * sin(-ROTATE_STEP); t.z = t.x * sin(-ROTATE_STEP) + t.z * cos(-ROTATE_STEP); break;    // LEFT
            case 102: t.x = t.x * cos(ROTATE_STEP) - t.z * sin(ROTATE_STEP); t.z = t.x * sin(ROTATE_STEP) + t.z * cos(ROTATE_STEP); break;        // RIGHT
            case 104: cam->target.y += MOVE_STEP; return moved();                                                                                // PAGE_UP
            case 105: cam->target.y -= MOVE_STEP; return moved();                                                                                // PAGE_DOWN
            default: return RT_KEY_NONE;
        }
        cam->target.x = t.x + cam->orig.x; cam->target.y = t.y + cam->orig.y; cam->target.z = t.z + cam->orig.z;
        return moved();
    }
    switch (key) {
        case 'p': return RT_KEY_DUMP;
        case 27: return RT_KEY_QUIT;
        case ' ': rt_update_camera(cam, w, h); return RT_KEY_RESTART;
        case 'a': slide(unit(cam->x), -MOVE_STEP); return moved();
        case 'd': slide(unit(cam->x), MOVE_STEP); return moved();
        case 'w': slide(cam->dir, MOVE_STEP); return moved();
        case 's': slide(cam->dir, -MOVE_STEP); return moved();
        case 'r': cam->orig.y += MOVE_STEP; cam->target.y += MOVE_STEP; return moved();
        case 'f': cam->orig.y -= MOVE_STEP; cam->target.y -= MOVE_STEP; return moved();
        case '+': if (!current_sphere || !n) return RT_ERR_ARG; *current_sphere = (*current_sphere + 1) % n; return RT_KEY_SCENE;
        case '-': if (!current_sphere || !n) return RT_ERR_ARG; *current_sphere = (*current_sphere + (n - 1)) % n; return RT_KEY_SCENE;
        case '4': return sphere_key(-0.5f * MOVE_STEP, 0.f, 0.f);
        case '6': return sphere_key(0.5f * MOVE_STEP, 0.f, 0.f);
        case '8': return sphere_key(0.f, 0.f, -0.5f * MOVE_STEP);
        case '2': return sphere_key(0.f, 0.f, 0.5f * MOVE_STEP);
        case '9': return sphere_key(0.f, 0.5f * MOVE_STEP, 0.f);
        case '3': return sphere_key(0.f, -0.5f * MOVE_STEP, 0.f);
        default: return RT_KEY_NONE;
    }
}

int rt_whitted_from_spheres(const rt_sphere *s, uint32_t n, const rt_camera *cam, rt_primitive *out, int cap) {
    if (!s || !out || cap < 0 || (uint32_t)cap < n) return RT_ERR_ARG;
    // camera frame (doubles; this conversion has no reference counterpart to be bit-compatible with)
    double ex[3] = { 1, 0, 0 }, ey[3] = { 0, 1, 0 }, ez[3] = { 0, 0, 1 }, o[3] = { 0, 0, 0 }, scale = 1.0, shift[3] = { 0, 0, 0 };
    if (cam) {
        const double d[3] = { (double)cam->target.x - cam->orig.x, (double)cam->target.y - cam->orig.y, (double)cam->target.z - cam->orig.z };
        const double len = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        if (!(len > 0.0)) return RT_ERR_ARG;
        for (int k = 0; k < 3; k++) ez[k] = d[k] / len;
        double x[3] = { -ez[2], 0.0, ez[0] };                               // up x dir, up = (0,1,0): points to the viewer's right
        double xl = sqrt(x[0] * x[0] + x[2] * x[2]);
        if (!(xl > 1e-12)) { x[0] = 1.0; x[2] = 0.0; xl = 1.0; }
        for (int k = 0; k < 3; k++) ex[k] = x[k] / xl;
        ey[0] = ez[1] * ex[2] - ez[2] * ex[1]; ey[1] = ez[2] * ex[0] - ez[0] * ex[2]; ey[2] = ez[0] * ex[1] - ez[1] * ex[0];
        o[0] = cam->orig.x; o[1] = cam->orig.y; o[2] = cam->orig.z;
        scale = 14.0 / len;
        shift[1] = 0.25; shift[2] = -7.0;
    }
    for (uint32_t i = 0; i < n; i++) {
        const bool light = s[i].e.x != 0.f || s[i].e.y != 0.f || s[i].e.z != 0.f;
        float r = s[i].c.x, g = s[i].c.y, b = s[i].c.z;
        float refl = 0.f, refr = 0.f, ri = 1.f, diff = 0.f, spec = 0.f;
        if (light) {
            const float m = fmaxf(s[i].e.x, fmaxf(s[i].e.y, s[i].e.z));
            r = s[i].e.x / m; g = s[i].e.y / m; b = s[i].e.z / m;
        } else if (s[i].refl == RT_DIFF) { diff = 1.f; spec = 0.5f; }
        else if (s[i].refl == RT_SPEC) { refl = 1.f; }
        else { refr = 1.f; ri = 1.5f; refl = 0.1f; }
        const double v[3] = { s[i].p.x - o[0], s[i].p.y - o[1], s[i].p.z - o[2] };
        const double cx = scale * (v[0] * ex[0] + v[1] * ex[1] + v[2] * ex[2]) + shift[0];
        const double cy = scale * (v[0] * ey[0] + v[1] * ey[1] + v[2] * ey[2]) + shift[1];
        const double cz = scale * (v[0] * ez[0] + v[1] * ez[1] + v[2] * ez[2]) + shift[2];
        out[i] = sphere(r, g, b, refl, refr, ri, diff, spec, light, (float)cx, (float)cy, (float)cz, (float)(scale * s[i].rad));
    }
    return (int)n;
}

// Scene_InitScene + Primitive_Create, R306/scene.cpp:53-82, :217-272 (maxx = maxy = 0: no sphere grid).
int rt_r306_create_scene(rt_r306_primitive *out, int cap) {
    struct Row { int type; float a, b, c, rd, r, g, bl, refl, refr, ri, diff, spec; bool light; };
    const int S = RT_R306_SPHERE, P = RT_R306_PLANE;
    static const Row rows[] = {
        { P, 0.0f, 0.75f, 0.0f, 4.4f, 0.6f, 0.6f, 0.6f, 0.0f, 0.0f, 0.0f, 0.4f, 1.8f, false },        // floor plane
        { S, 0.0f, 6.5f, 22.0f, 0.35f, 0.85f, 0.85f, 0.85f, 0.0f, 0.0f, 0.0f, 1.0f, 1.0f, true },     // light source center
        { S, 3.4f, -3.40f, 23.0f, 2.5f, 0.08f, 0.08f, 0.08f, 1.9f, 1.0f, 2.3f, 0.0f, 0.0f, false },   // big sphere
        { S, -0.7f, -4.90f, 27.0f, 1.0f, 0.07f, 0.17f, 0.07f, 0.1f, 1.5f, 2.3f, 0.2f, 0.8f, false },  // small sphere 5
        { S, -3.4f, -3.40f, 29.0f, 2.5f, 1.0f, 1.0f, 1.0f, 0.8f, 0.0f, 0.0f, 0.0f, 0.0f, false },     // small sphere
        { S, 0.5f, -4.10f, 29.0f, 1.5f, 1.5f, 0.7f, 0.7f, 0.1f, 0.0f, 0.0f, 0.2f, 0.2f, false },      // small sphere 2
        { S, -6.0f, -4.10f, 32.0f, 1.5f, 0.7f, 0.7f, 1.7f, 0.2f, 0.0f, 0.0f, 0.2f, 0.2f, false },     // small sphere 3
        { S, -6.7f, -4.90f, 29.0f, 1.0f, 0.07f, 0.17f, 0.07f, 0.1f, 1.5f, 2.3f, 0.2f, 0.8f, false },  // small sphere 4
        { S, 6.4f, -4.90f, 18.0f, 1.0f, 0.18f, 0.18f, 0.18f, 1.7f, 1.0f, 2.6f, 1.8f, 0.0f, false },   // small sphere 6
        { P, 0.7f, 0.0f, 0.0f, 5.4f, 1.0f, 0.6f, 0.6f, 0.0f, 0.0f, 0.0f, 0.8f, 1.5f, false },         // left wall
        { P, -0.7f, 0.0f, 0.0f, 5.4f, 0.7f, 0.6f, 1.0f, 0.0f, 0.0f, 0.0f, 0.8f, 0.8f, false },        // right wall
        { P, 0.0f, -0.8f, 0.0f, 5.4f, 1.0f, 1.0f, 1.0f, 0.0f, 0.0f, 0.0f, 1.2f, 0.8f, false },        // top wall
        { P, 0.0f, 0.0f, -0.14f, 5.4f, 2.5f, 2.5f, 2.5f, 0.0f, 0.0f, 0.0f, 1.2f, 0.8f, false },       // back wall
        { P, 0.0f, 0.0f, 0.72f, 5.4f, 0.1f, 0.1f, 0.1f, 0.0f, 0.0f, 0.0f, 1.0f, 1.0f, false },        // front wall
        { S, -3.0f, 6.5f, 22.0f, 0.35f, 0.85f, 0.85f, 0.85f, 0.0f, 0.0f, 0.0f, 0.0f, 1.8f, true },    // light source right
        { S, 3.0f, 6.5f, 22.0f, 0.35f, 0.85f, 0.85f, 0.85f, 0.0f, 0.0f, 0.0f, 0.0f, 1.8f, true },     // light source left
        { S, -5.8f, -5.55f, 31.0f, 0.35f, 1.15f, 0.35f, 0.35f, 1.0f, 1.0f, 2.3f, 0.0f, 1.8f, true },  // light source ground back
    };
    const int n = (int)(sizeof rows / sizeof rows[0]);
    if (!out || cap < n) return RT_ERR_ARG;
    for (int i = 0; i < n; i++) {
        const Row &r = rows[i];
        rt_r306_primitive &p = out[i];
        memset(&p, 0, sizeof p);
        p.type = r.type; p.m_light = r.light ? 1 : 0;
        p.m_color.x = r.r; p.m_color.y = r.g; p.m_color.z = r.bl;
        p.m_refl = r.refl; p.m_refr = r.refr; p.m_rindex = r.ri; p.m_diff = r.diff; p.m_spec = r.spec;
        if (r.type == S) {
            p.centre.x = r.a; p.centre.y = r.b; p.centre.z = r.c;
            p.radius = r.rd; p.sq_radius = r.rd * r.rd; p.r_radius = r.rd > 0 ? 1.0f / r.rd : 0;
        } else {
            p.plane_n.x = r.a; p.plane_n.y = r.b; p.plane_n.z = r.c; p.plane_d = r.rd;
        }
    }
    return n;
}

int rt_read_scene(const char *path, rt_camera *cam_out, rt_sphere **spheres_out, uint32_t *n_out) {   // SPT/displayfunc.cpp:120-180
    if (!path || !cam_out || !spheres_out || !n_out) return RT_ERR_ARG;
    *spheres_out = nullptr; *n_out = 0;
    FILE *f = fopen(path, "r");
    if (!f) return RT_ERR_IO;
    memset(cam_out, 0, sizeof *cam_out);
    int c = fscanf(f, "camera %f %f %f  %f %f %f\n", &cam_out->orig.x, &cam_out->orig.y, &cam_out->orig.z,
                   &cam_out->target.x, &cam_out->target.y, &cam_out->target.z);
    unsigned n = 0;
    if (c != 6 || fscanf(f, "size %u\n", &n) != 1 || n == 0) { fclose(f); return RT_ERR_IO; }
    rt_sphere *s = (rt_sphere *)malloc(sizeof(rt_sphere) * n);
    if (!s) { fclose(f); return RT_ERR_IO; }
    for (unsigned i = 0; i < n; i++) {
        int mat = -1;
        c = fscanf(f, "sphere %f  %f %f %f  %f %f %f  %f %f %f  %d\n", &s[i].rad, &s[i].p.x, &s[i].p.y, &s[i].p.z,
                   &s[i].e.x, &s[i].e.y, &s[i].e.z, &s[i].c.x, &s[i].c.y, &s[i].c.z, &mat);
        if (c != 11 || mat < 0 || mat > 2) { free(s); fclose(f); return RT_ERR_IO; }   // the reference exit(-1)s here
        s[i].refl = mat;
    }
    fclose(f);
    *spheres_out = s; *n_out = n;
    return RT_OK;
}

int rt_write_complex_scene(const char *path, int max_depth) {
    if (!path || max_depth < 0 || max_depth > 8) return RT_ERR_ARG;
    ComplexGen count_only = { nullptr, (double)max_depth, 0 };
    count_only.rec(0.0, 0.0, 0.0, 0.0, 15.0, 2);
    FILE *f = fopen(path, "w");
    if (!f) return RT_ERR_IO;
    // header of SPT/scenes/complex.scn: camera, size, the light and the floor
    fprintf(f, "camera 20 80 150  0 15 0\nsize %u\n", count_only.count + 2);
    fprintf(f, "sphere 8     50 80 90   25 25 25  0 0 0           0\n");
    fprintf(f, "sphere 10000  0 -10050 0  0 0 0     0.75 0.75 0.75  0\n");
    ComplexGen gen = { f, (double)max_depth, 0 };
    gen.rec(0.0, 0.0, 0.0, 0.0, 15.0, 2);
    fclose(f);
    return RT_OK;
}

int rt_whitted_create_scene(int which, rt_primitive *out, int cap) {       // R323/scene.c:48-128
    std::vector<rt_primitive> v;
    if (which == 0) {
        const float light = 0.85f;
        v.push_back(plane(0.6f, 0.6f, 0.6f, 0.0f, 0.0f, 0.0f, 0.4f, 1.8f, false, 0.0f, 0.75f, 0.0f, 4.4f));          // floor
        v.push_back(sphere(0.08f, 0.08f, 0.08f, 0.2f, 1.0f, 1.4f, 0.0f, 0.0f, false, 3.4f, -3.4f, 23.0f, 2.5f));     // big glass sphere
        v.push_back(sphere(0.07f, 0.17f, 0.07f, 0.1f, 1.0f, 1.2f, 0.0f, 0.0f, false, -0.7f, -4.90f, 27.0f, 1.0f));
        v.push_back(sphere(1.0f, 1.0f, 1.0f, 0.8f, 0.0f, 0.0f, 0.0f, 0.0f, false, -3.4f, -3.4f, 29.0f, 2.5f));       // mirror
        v.push_back(sphere(1.5f, 0.7f, 0.7f, 0.1f, 0.0f, 0.0f, 0.2f, 0.2f, false, 0.5f, -4.1f, 29.0f, 1.5f));
        v.push_back(sphere(0.7f, 0.7f, 1.7f, 0.2f, 0.0f, 0.0f, 0.2f, 0.2f, false, -6.0f, -4.1f, 32.0f, 1.5f));
        v.push_back(sphere(0.07f, 0.17f, 0.07f, 0.3f, 1.0f, 1.2f, 0.2f, 0.8f, false, -6.7f, -4.90f, 29.0f, 1.0f));
        v.push_back(sphere(0.08f, 0.08f, 0.08f, 0.7f, 1.0f, 1.3f, 0.8f, 0.0f, false, 6.4f, -4.9f, 18.0f, 1.0f));
        v.push_back(plane(1.0f, 0.6f, 0.6f, 0.0f, 0.0f, 0.0f, 0.8f, 1.5f, false, 0.7f, 0.0f, 0.0f, 5.4f));           // left wall
        v.push_back(plane(0.7f, 0.6f, 1.0f, 0.0f, 0.0f, 0.0f, 0.8f, 0.8f, false, -0.7f, 0.0f, 0.0f, 5.4f));          // right wall
        v.push_back(plane(1.0f, 1.0f, 1.0f, 0.0f, 0.0f, 0.0f, 1.2f, 0.8f, false, 0.0f, -0.8f, 0.0f, 5.4f));          // ceiling
        v.push_back(plane(1.5f, 1.5f, 1.5f, 0.0f, 0.0f, 0.0f, 1.2f, 0.8f, false, 0.0f, 0.0f, -0.14f, 5.4f));         // back wall
        v.push_back(plane(0.1f, 0.1f, 0.1f, 0.0f, 0.0f, 0.0f, 1.0f, 1.0f, false, 0.0f, 0.0f, 0.72f, 5.4f));          // front wall
        v.push_back(sphere(light, light, light, 0.0f, 0.0f, 0.0f, 0.0f, 1.8f, true, 0.0f, 6.5f, 22.0f, 0.35f));      // three lights
        v.push_back(sphere(light, light, light, 0.0f, 0.0f, 0.0f, 0.0f, 1.8f, true, -3.0f, 6.5f, 22.0f, 0.35f));
        v.push_back(sphere(light, light, light, 0.0f, 0.0f, 0.0f, 0.0f, 1.8f, true, 3.0f, 6.5f, 22.0f, 0.35f));
        v.push_back(blank());      // n_primitives is 17 but only 16 slots are filled: an all-zero plane that nothing hits
    } else if (which == 1) {
        v.push_back(plane(0.4f, 0.3f, 0.3f, 0.0f, 0.0f, 1.0f, 1.0f, 0.8f, false, 0.0f, 1.0f, 0.0f, 4.4f));
        v.push_back(sphere(0.7f, 0.7f, 1.0f, 0.0f, 1.0f, 1.3f, 0.2f, 0.8f, false, 2.0f, 0.8f, 3.0f, 2.5f));
        v.push_back(sphere(0.7f, 0.7f, 1.0f, 0.5f, 0.0f, 1.0f, 0.1f, 0.8f, false, -5.5f, -0.5f, 7.0f, 2.0f));
        v.push_back(sphere(0.4f, 0.4f, 0.4f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, true, 0.0f, 5.0f, 5.0f, 0.1f));
        v.push_back(sphere(0.6f, 0.6f, 0.8f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, true, -3.0f, 5.0f, 1.0f, 0.1f));
        v.push_back(sphere(1.0f, 0.4f, 0.4f, 0.5f, 0.0f, 1.0f, 0.2f, 0.8f, false, -1.5f, -3.8f, 1.0f, 1.5f));
        v.push_back(plane(0.5f, 0.3f, 0.5f, 0.0f, 0.0f, 1.0f, 0.6f, 0.0f, false, 0.4f, 0.0f, -1.0f, 12.0f));
        v.push_back(plane(0.4f, 0.7f, 0.7f, 0.0f, 0.0f, 1.0f, 0.5f, 0.0f, false, 0.0f, -1.0f, 0.0f, 7.4f));
        for (int x = 0; x < 8; x++)
            for (int y = 0; y < 7; y++)
                v.push_back(sphere(0.3f, 1.0f, 0.4f, 0.0f, 0.0f, 1.0f, 0.6f, 0.6f, false, -4.5f + x * 1.5f, -4.3f + y * 1.5f, 10.0f, 0.3f));
    } else {
        return RT_ERR_ARG;
    }
    if (!out || cap < (int)v.size()) return RT_ERR_ARG;
    memcpy(out, v.data(), v.size() * sizeof(rt_primitive));
    return (int)v.size();
}

int rt_write_bmp(const char *path, const rt_uchar4 *px, int w, int h) {     // R323/bitmap.c:8-75
    if (!path || !px || w < 1 || h < 1) return RT_ERR_ARG;
    const int row = 3 * w, pad = (row % 4) ? 4 - row % 4 : 0;
    const uint32_t image = (uint32_t)(row + pad) * h, offset = 14 + 40, total = offset + image;
    FILE *f = fopen(path, "wb");
    if (!f) return RT_ERR_IO;
    unsigned char hdr[54];
    memset(hdr, 0, sizeof hdr);
    hdr[0] = 'B'; hdr[1] = 'M';
    auto le32 = [&](int at, uint32_t v) { for (int k = 0; k < 4; k++) hdr[at + k] = (unsigned char)(v >> (8 * k)); };
    auto le16 = [&](int at, uint32_t v) { hdr[at] = (unsigned char)v; hdr[at + 1] = (unsigned char)(v >> 8); };
    le32(2, total); le32(10, offset);
    le32(14, 40); le32(18, (uint32_t)w); le32(22, (uint32_t)h); le16(26, 1); le16(28, 24);
    le32(34, image); le32(38, 2835); le32(42, 2835);
    fwrite(hdr, 1, sizeof hdr, f);
    std::vector<unsigned char> line(row + pad, 0);
    for (int y = h - 1; y >= 0; --y) {          // bottom row first, BGR
        for (int x = 0; x < w; x++) {
            const rt_uchar4 &p = px[(size_t)y * w + x];
            line[3 * x] = p.z; line[3 * x + 1] = p.y; line[3 * x + 2] = p.x;
        }
        fwrite(line.data(), 1, line.size(), f);
    }
    fclose(f);
    return RT_OK;
}

int rt_write_ppm(const char *path, const uint32_t *pixels, int w, int h) {   // SPT/displayfunc.cpp:254-271
    if (!path || !pixels || w < 1 || h < 1) return RT_ERR_ARG;
    FILE *f = fopen(path, "w");
    if (!f) return RT_ERR_IO;
    fprintf(f, "P3\n%d %d\n%d\n", w, h, 255);
    for (int y = h - 1; y >= 0; --y)
        for (int x = 0; x < w; x++) {
            const uint32_t p = pixels[(size_t)y * w + x];
            fprintf(f, "%d %d %d ", (int)(p & 255u), (int)((p >> 8) & 255u), (int)((p >> 16) & 255u));
        }
    fclose(f);
    return RT_OK;
}

void rt_free(void *p) { free(p); }

}  // extern "C"
