// rt_cli -- headless driver with the reference's own command line.
//
// The reference's smallptGPU takes   <0/1 cpu/gpu> <work-group size> <kernel file> <width> <height> <scene.scn>
// (argv parsing at SPT/smallptGPU.cpp:835-854, used by SPT/RUN_SCENE_*.bat, e.g.
//  "smallptGPU.exe 1 64 rendering_kernel.cl 640 480 scenes\cornell.scn") and dumps "image.ppm" when 'p' is pressed
// (SPT/displayfunc.cpp:254-271).  This driver accepts the same six arguments, runs the passes headless through
// the C ABI and writes the same PPM; the kernel-file argument selects the integrator exactly as the file did
// (rendering_kernel.cl -> path tracing, rendering_kernel_dl.cl -> direct lighting).  Naming raytracer_kernel.cl
// instead renders the Raytracer3.2.03 frame (R323/raytracer.c:705-797) and writes test.bmp; the scene argument
// is then CHOOSE_SCENE (0 or 1, R323/common.h:6).
//
//   rt_cli 1 64 rendering_kernel.cl 640 480 scenes/cornell.scn [passes=64] [out=image.ppm] [keys]
//   rt_cli 1 64 raytracer_kernel.cl 800 600 0 [ignored] [out=test.bmp]
//
// `keys` scripts an interactive session: each character is one key press of the reference viewer
// (SPT/displayfunc.cpp:252-420: a d w s r f move the camera, + - 4 6 8 2 9 3 select / move a sphere, space restarts,
// U D L R < > stand for the arrow and page keys), applied by rt_viewer_key after `passes` passes; the camera or scene
// is re-uploaded, the image restarts at sample 0 as in ReInit / ReInitScene, and `passes` more passes are rendered.
//
// Seeds follow the reference: 2*w*h draws of libc rand(), each raised to >= 2 (SPT/smallptGPU.cpp:105-110).
// The device argument must be 1: there is no CPU path.  The work-group argument is accepted and ignored (the
// persistent kernels choose their own CTA shape).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <vector>
#include "../../include/rt_b200.h"

static double now() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + t.tv_nsec * 1e-9; }

static int usage(const char *argv0) {
    fprintf(stderr, "Usage: %s <1 = gpu> <work-group size (ignored)> <rendering_kernel.cl | rendering_kernel_dl.cl | raytracer_kernel.cl> "
                    "<width> <height> <scene.scn | CHOOSE_SCENE> [passes] [output file]\n", argv0);
    return 2;
}

int main(int argc, char **argv) {
    if (argc < 7) return usage(argv[0]);
    if (atoi(argv[1]) != 1) { fprintf(stderr, "%s: device 0 (CPU) is not available: this renderer has no CPU path\n", argv[0]); return 2; }
    const char *kernel = argv[3];
    const int w = atoi(argv[4]), h = atoi(argv[5]);
    if (w < 1 || h < 1) return usage(argv[0]);
    const char *slash = strrchr(kernel, '/');
    if (!slash) slash = strrchr(kernel, '\\');
    const char *kname = slash ? slash + 1 : kernel;
    rt_ctx *ctx = nullptr;
    if (rt_init(&ctx, 0)) { fprintf(stderr, "%s\n", rt_last_error(nullptr)); return 1; }
    int rc = 0;
    if (!strcmp(kname, "raytracer_kernel.cl")) {
        std::vector<rt_primitive> prims(64);
        const int n = rt_whitted_create_scene(atoi(argv[6]), prims.data(), (int)prims.size());
        if (n < 0) { fprintf(stderr, "unknown CHOOSE_SCENE %s\n", argv[6]); rt_destroy(ctx); return 2; }
        std::vector<rt_uchar4> px((size_t)w * h);
        const double t0 = now();
        rc = rt_whitted_render(ctx, prims.data(), n, w, h, px.data(), nullptr);
        const int ms = (int)((now() - t0) * 1000.0);
        if (rc) fprintf(stderr, "%s\n", rt_last_error(ctx));
        else {
            printf("Runtime: %02d:%02d.%03d\n", (ms / 60000) % 100, (ms / 1000) % 60, ms % 1000);      // R323/raytracer.c:759-770
            rc = rt_write_bmp(argc > 8 ? argv[8] : "test.bmp", px.data(), w, h);
        }
    } else {
        const int integrator = !strcmp(kname, "rendering_kernel_dl.cl") ? 1 : 0;
        if (!integrator && strcmp(kname, "rendering_kernel.cl")) fprintf(stderr, "unknown kernel file '%s': using the path tracer\n", kname);
        const int passes = argc > 7 ? atoi(argv[7]) : 64;
        rt_camera cam; rt_sphere *spheres = nullptr; uint32_t n = 0;
        fprintf(stderr, "Reading scene: %s\n", argv[6]);
        if (rt_read_scene(argv[6], &cam, &spheres, &n)) { fprintf(stderr, "Failed to read scene: %s\n", argv[6]); rt_destroy(ctx); return 1; }
        fprintf(stderr, "Scene size: %u\n", n);
        rt_update_camera(&cam, w, h);
        std::vector<uint32_t> seeds((size_t)w * h * 2), pixels((size_t)w * h);
        for (auto &s : seeds) { s = (uint32_t)rand(); if (s < 2) s = 2; }
        rc = rt_pt_resize(ctx, w, h, seeds.data());
        if (!rc) rc = rt_pt_set_scene(ctx, spheres, n);
        if (!rc) rc = rt_pt_set_camera(ctx, &cam);
        const double t0 = now();
        if (!rc) rc = rt_pt_render(ctx, integrator, passes < 1 ? 1 : passes, pixels.data(), nullptr, nullptr);
        uint32_t current_sphere = 0;
        for (const char *k = argc > 9 ? argv[9] : ""; *k && !rc; ++k) {
            int key = (unsigned char)*k;
            switch (*k) { case 'U': key = RT_KEY_SPECIAL + 101; break; case 'D': key = RT_KEY_SPECIAL + 103; break; case 'L': key = RT_KEY_SPECIAL + 100; break;
                          case 'R': key = RT_KEY_SPECIAL + 102; break; case '<': key = RT_KEY_SPECIAL + 104; break; case '>': key = RT_KEY_SPECIAL + 105; break; default: break; }
            const int action = rt_viewer_key(key, &cam, w, h, spheres, n, &current_sphere);
            if (action == RT_KEY_CAMERA) rc = rt_pt_set_camera(ctx, &cam);
            else if (action == RT_KEY_SCENE) rc = rt_pt_set_scene(ctx, spheres, n);
            else if (action == RT_KEY_RESTART) {
                for (auto &s : seeds) { s = (uint32_t)rand(); if (s < 2) s = 2; }
                rc = rt_pt_resize(ctx, w, h, seeds.data());
            } else if (action == RT_KEY_QUIT) break;
            else if (action == RT_KEY_DUMP) { rc = rt_write_ppm("image.ppm", pixels.data(), w, h); continue; }
            else continue;
            if (!rc) rc = rt_pt_render(ctx, integrator, passes < 1 ? 1 : passes, pixels.data(), nullptr, nullptr);
        }
        const double dt = now() - t0;
        if (rc) fprintf(stderr, "%s\n", rt_last_error(ctx));
        else {
            // the reference's caption (SPT/smallptGPU.cpp:777-781)
            printf("Rendering time %.3f sec (pass %d)  Sample/sec  %.1fK\n", dt, rt_pt_current_sample(ctx), (double)w * h * passes / dt / 1000.0);
            rc = rt_write_ppm(argc > 8 ? argv[8] : "image.ppm", pixels.data(), w, h);
        }
        rt_free(spheres);
    }
    rt_destroy(ctx);
    return rc ? 1 : 0;
}
