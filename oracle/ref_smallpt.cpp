// oracle/_ref harness for smallpt -- TEST INFRASTRUCTURE ONLY, never linked into the product.
//
// Pulls the reference's CPU path tracer (smallptgpu-v1.6/smallptCPU.cpp, which includes
// geomfunc.h / simplernd.h / scene.h) UNMODIFIED into this translation unit, compiled as C++
// like the reference's own vcxproj does (/TP), so that sqrt/cos/sin/pow/fabs bind to the
// single-precision overloads.  Including the .cpp (instead of linking it) gives this harness
// access to the file-static `colors` and `seeds` buffers, so it can own seed allocation: the
// reference's AllocateBuffers() hard-codes 640x480 (smallptCPU.cpp:67) and draws from libc rand().
// displayfunc.cpp (ReadScene, UpdateCamera, width/height/pixels) is compiled separately against
// the no-op GL/glut.h shim.  Nothing from the reference is copied into this repository.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include "smallptCPU.cpp"

// Globals and entry points that live in smallptGPU.cpp in the reference (not built here).
Camera camera;
int currentSample = 0;
Sphere *spheres = 0;
unsigned int sphereCount = 0;
int useOpenCL = 0;
unsigned int renderingFlags = 0;
void ReInitGPU(const int) {}
void ReInitSceneGPU() {}
void UpdateRenderingGPU() {}

static_assert(sizeof(Sphere) == 44, "Sphere must be 44 bytes");
static_assert(sizeof(Camera) == 60, "Camera must be 60 bytes");

static void drop_buffers() {
    free(colors); free(seeds); free(pixels);
    colors = 0; seeds = 0; pixels = 0;
}

void keyFunc(unsigned char key, int x, int y);      // displayfunc.cpp:252, :367 (C++ linkage)
void specialFunc(int key, int x, int y);

extern "C" {

// ReadScene + UpdateCamera of the reference (displayfunc.cpp:120-195). Returns the sphere count.
int ref_pt_load_scene(const char *scn_path, int w, int h) {
    width = w; height = h;
    ReadScene((char *)scn_path);
    UpdateCamera();
    return (int)sphereCount;
}

// Caller-supplied scene; camera orig/target are taken from cam and UpdateCamera() derives the rest.
void ref_pt_set_scene(const void *sph, int n, const void *cam, int w, int h) {
    width = w; height = h;
    spheres = (Sphere *)malloc(sizeof(Sphere) * n);
    memcpy(spheres, sph, sizeof(Sphere) * n);
    sphereCount = n;
    memcpy(&camera, cam, sizeof(Camera));
    UpdateCamera();
}

void ref_pt_get_scene(void *sph_out, void *cam_out) {
    if (sph_out) memcpy(sph_out, spheres, sizeof(Sphere) * sphereCount);
    if (cam_out) memcpy(cam_out, &camera, sizeof(Camera));
}

// n_passes calls of the reference's own UpdateRenderingCPU() (smallptCPU.cpp:77-132), starting
// from sample 0, on caller-owned seeds (2*w*h u32).  Outputs may be NULL.
void ref_pt_render(const unsigned *seeds_in, int n_passes, float *colors_out, unsigned *pixels_out,
                   unsigned *seeds_out) {
    const size_t np = (size_t)width * height;
    drop_buffers();
    colors = (Vec *)calloc(np, sizeof(Vec));
    seeds = (unsigned int *)malloc(sizeof(unsigned int) * np * 2);
    pixels = (unsigned int *)calloc(np, sizeof(unsigned int));
    memcpy(seeds, seeds_in, sizeof(unsigned int) * np * 2);
    currentSample = 0;
    for (int p = 0; p < n_passes; p++) UpdateRenderingCPU();
    if (colors_out) memcpy(colors_out, colors, sizeof(Vec) * np);
    if (pixels_out) memcpy(pixels_out, pixels, sizeof(unsigned int) * np);
    if (seeds_out) memcpy(seeds_out, seeds, sizeof(unsigned int) * np * 2);
}

// Row-range, multi-pass driver around the reference's per-pixel functions RadiancePathTracing
// (geomfunc.h:167-338) and RadianceDirectLighting (:340-483).  The pixel body below restates
// smallptCPU.cpp:84-124 (the reference has no CPU driver for the direct-lighting integrator and
// no way to render a row range); ref_pt_render() above is the unrestated loop it is checked against.
struct RowJob { int integrator, y0, y1, pass0, n_passes; float *colors; unsigned *seeds, *pix; };
static void *rows_thread(void *pv) {
    RowJob *j = (RowJob *)pv;
    const float invWidth = 1.f / width, invHeight = 1.f / height;
    for (int y = j->y0; y < j->y1; y++)
        for (int x = 0; x < width; x++) {
            const int i = (height - y - 1) * width + x, i2 = 2 * i;
            Vec c; c.x = j->colors[3 * i]; c.y = j->colors[3 * i + 1]; c.z = j->colors[3 * i + 2];
            for (int s = j->pass0; s < j->pass0 + j->n_passes; s++) {
                const float r1 = GetRandom(&j->seeds[i2], &j->seeds[i2 + 1]) - .5f;
                const float r2 = GetRandom(&j->seeds[i2], &j->seeds[i2 + 1]) - .5f;
                const float kcx = (x + r1) * invWidth - .5f;
                const float kcy = (y + r2) * invHeight - .5f;
                Vec rdir;
                vinit(rdir, camera.x.x * kcx + camera.y.x * kcy + camera.dir.x,
                      camera.x.y * kcx + camera.y.y * kcy + camera.dir.y,
                      camera.x.z * kcx + camera.y.z * kcy + camera.dir.z);
                Vec rorig;
                vsmul(rorig, 0.1f, rdir);
                vadd(rorig, rorig, camera.orig);
                vnorm(rdir);
                const Ray ray = {rorig, rdir};
                Vec r;
                if (j->integrator == 0)
                    RadiancePathTracing(spheres, sphereCount, &ray, &j->seeds[i2], &j->seeds[i2 + 1], &r);
                else
                    RadianceDirectLighting(spheres, sphereCount, &ray, &j->seeds[i2], &j->seeds[i2 + 1], &r);
                if (s == 0) c = r;
                else {
                    const float k1 = s, k2 = 1.f / (k1 + 1.f);
                    c.x = (c.x * k1 + r.x) * k2; c.y = (c.y * k1 + r.y) * k2; c.z = (c.z * k1 + r.z) * k2;
                }
            }
            j->colors[3 * i] = c.x; j->colors[3 * i + 1] = c.y; j->colors[3 * i + 2] = c.z;
            if (j->pix) j->pix[y * width + x] = toInt(c.x) | (toInt(c.y) << 8) | (toInt(c.z) << 16);
        }
    return 0;
}

// colors (3*w*h floats) and seeds (2*w*h u32) are updated in place; pixels may be NULL.
void ref_pt_render_mt(int integrator, int pass0, int n_passes, float *colors_io, unsigned *seeds_io,
                      unsigned *pixels_out, int threads) {
    if (threads < 1) threads = 1;
    pthread_t *t = (pthread_t *)malloc(sizeof(pthread_t) * threads);
    RowJob *j = (RowJob *)malloc(sizeof(RowJob) * threads);
    for (int k = 0; k < threads; k++) {
        j[k].integrator = integrator; j[k].pass0 = pass0; j[k].n_passes = n_passes;
        j[k].y0 = (int)((long)height * k / threads); j[k].y1 = (int)((long)height * (k + 1) / threads);
        j[k].colors = colors_io; j[k].seeds = seeds_io; j[k].pix = pixels_out;
        pthread_create(&t[k], 0, rows_thread, &j[k]);
    }
    for (int k = 0; k < threads; k++) pthread_join(t[k], 0);
    free(t); free(j);
}

// One key press through the reference's own GLUT callbacks (displayfunc.cpp:252-420).  ReInit(0) / ReInitScene()
// restart the image and, in the CPU build, render one pass at once -- so the render buffers must exist
// (ref_pt_render first; keep the image tiny).  Returns currentSample afterwards (1 after a camera key).
int ref_pt_key(int key, int special) {
    if (special) specialFunc(key, 0, 0); else keyFunc((unsigned char)key, 0, 0);
    return currentSample;
}

// Known-answer taps on the reference's own inline functions.
float ref_pt_get_random(unsigned *s0, unsigned *s1) { return GetRandom(s0, s1); }
float ref_pt_sphere_intersect(const void *sphere, const float *o3, const float *d3) {
    Ray r; r.o.x = o3[0]; r.o.y = o3[1]; r.o.z = o3[2]; r.d.x = d3[0]; r.d.y = d3[1]; r.d.z = d3[2];
    return SphereIntersect((const Sphere *)sphere, &r);
}

}  // extern "C"
