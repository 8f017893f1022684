// oracle/_ref harness, part B -- TEST INFRASTRUCTURE ONLY, never linked into the product.
//
// Exposes the reference's own scene builder (scene.c: create_scene) and BMP writer (bitmap.c:
// write_bmp_file), both compiled unmodified from /root/reference by oracle/Makefile, and
// performs the Primitive -> Primitive_2 field-by-field conversion that the reference's main()
// does at raytracer.c:721-746.
#include "common.h"
#include "scene.h"
#include "bitmap.h"
#include <stdlib.h>
#include <string.h>

static_assert(sizeof(Primitive) == 96, "nested Primitive must be 96 bytes (unaligned cl_float4)");
static_assert(sizeof(Primitive_2) == 96, "flat Primitive_2 must be 96 bytes");

extern "C" {

// Returns the primitive count; fills at most `cap` entries of out (Primitive_2[cap]).
int ref_whitted_scene(void *out_v, int cap) {
    cl_uint n = 0;
    Primitive *src = create_scene(n);
    Primitive_2 *dst = (Primitive_2 *)out_v;
    for (cl_uint i = 0; i < n && (int)i < cap; i++) {
        Primitive_2 q; memset(&q, 0, sizeof q);
        const Primitive &p = src[i];
        q.m_color.x = p.material.color.s[0]; q.m_color.y = p.material.color.s[1];
        q.m_color.z = p.material.color.s[2]; q.m_color.w = p.material.color.s[3];
        q.m_refl = p.material.refl; q.m_diff = p.material.diff; q.m_refr = p.material.refr;
        q.m_refr_index = p.material.refr_index; q.m_spec = p.material.spec; q.dummy_3 = p.material.dummy_3;
        q.type = p.type; q.is_light = p.is_light;
        q.normal.x = p.normal.s[0]; q.normal.y = p.normal.s[1]; q.normal.z = p.normal.s[2]; q.normal.w = p.normal.s[3];
        q.center.x = p.center.s[0]; q.center.y = p.center.s[1]; q.center.z = p.center.s[2]; q.center.w = p.center.s[3];
        q.depth = p.depth; q.radius = p.radius; q.sq_radius = p.sq_radius; q.r_radius = p.r_radius;
        dst[i] = q;
    }
    free(src);
    return (int)n;
}

int ref_write_bmp(const unsigned char *rgba, int w, int h, const char *filename) {
    Pixel *px = (Pixel *)malloc(sizeof(Pixel) * (size_t)w * h);
    for (size_t i = 0; i < (size_t)w * h; i++)
        for (int c = 0; c < 4; c++) px[i].s[c] = rgba[4 * i + c];
    int rc = write_bmp_file(px, w, h, (char *)filename);
    free(px);
    return rc;
}

}  // extern "C"
