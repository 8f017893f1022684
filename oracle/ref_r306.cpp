// oracle/_ref harness for raytracer3.0.06.no_rec.samp -- TEST INFRASTRUCTURE ONLY, never linked into the product.
//
// BASELINE config 1: the reference's CPU Whitted render of its built-in sphere/plane scene at 800x600 (rows
// 20..529, 3x3 super-sampling, 63-node implicit ray tree), compiled UNMODIFIED from /root/reference
// (raytracer.cpp, scene.cpp, surface.cpp) by oracle/Makefile.  The harness repeats the set-up calls of the
// reference's main() (testapp.cpp:59-69, 122-124).  The engine keeps its state in globals, so it is single-threaded,
// like the reference.  Used as the config-1 CPU timing baseline only (not a GPU parity target).
#include "raytracer.h"
#include "scene.h"
#include "surface.h"
#include <string.h>

extern "C" {

// Renders one frame into dest (w*h Pixels, 0x00RRGGBB; rows outside 20..529 are left untouched).
void ref_r306_render(unsigned int *dest, int w, int h) {
    Engine_Constructor();
    TracedRays_init();
    Scene_InitScene();
    Engine_SetTarget((Pixel *)dest, w, h);
    Engine_InitRender();
    Engine_Render();
}

// The same with a caller-supplied scene table (n reference Primitive records of 96 bytes) instead of Scene_InitScene's.
void ref_r306_render_scene(unsigned int *dest, int w, int h, const void *prims, int n) {
    static Primitive table[256];
    Engine_Constructor();
    TracedRays_init();
    m_Scene = (Scene *)malloc(sizeof(Scene));
    memcpy(table, prims, sizeof(Primitive) * (size_t)(n > 256 ? 256 : n));
    m_Scene->m_Primitive = table;
    m_Scene->m_Primitives = n > 256 ? 256 : n;
    Engine_SetTarget((Pixel *)dest, w, h);
    Engine_InitRender();
    Engine_Render();
}

}  // extern "C"
