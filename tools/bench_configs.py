"""Measures the BASELINE.json configs that are not bench.py's headline (kernel time by CUDA events, one line each).
C4: scene_build_complex.pl scenes ($maxDepth 4/5/6 -> 783 / 3 908 / 19 533 spheres) at 3840x2160, path tracing;
C5: cornell 7680x4320 (reduced spp on one GPU, scaled figure stated); plus Whitted at 4K.
Usage: python tools/bench_configs.py [quick]"""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
rt = g.load()
quick = len(sys.argv) > 1
r = rt.Renderer(0)
info = r.device_info()
peak = info["sm_count"] * 128 * 2 * 1.965e9


def pt_case(tag, spheres, cam, w, h, spp, integ=0):
    seeds = rt.reference_seeds(w, h, seed=1)
    r.pt_resize(w, h, seeds); r.pt_set_scene(spheres); r.pt_set_camera(cam)
    r.set_counting(True); r.pt_launch(integ, 1); c = r.counters(); r.set_counting(False)
    per = {k: v / max(c["samples"], 1) for k, v in c.items()}
    samples = w * h * spp
    # counters of pass 0 only: later passes have the same statistics (independent seeds)
    flop = 17.0 * per["sphere_tests"] * samples
    for mode, name in ((0, "loop over every sphere"), (1, "exact hierarchy")):
        if mode == 1 and spheres.size < 64:
            continue
        r.set_tuning(rt.TUNE_PT_BVH, mode)
        best = 1e30
        for _ in range(2):
            r.pt_resize(w, h, seeds); r.pt_set_camera(cam)
            r.timer_begin(); r.pt_launch(integ, spp); best = min(best, r.timer_end())
        line = (f"{tag} [{name}]: {spheres.size} spheres {w}x{h}x{spp}spp  {best:.1f} ms  {samples / best / 1e3:.1f} Msamples/s  "
                f"{samples * (per['nearest_queries'] + per['shadow_queries']) / best / 1e3:.0f} Mrays/s  {per['sphere_tests']:.0f} reference tests/sample")
        if mode == 0:
            line += f"  {flop / best / 1e9:.2f} TFLOP/s algorithmic = {100 * flop / (best * 1e-3) / peak:.1f}% of FP32 peak"
        else:
            line += f"  (reference-equivalent {flop / best / 1e9:.1f} TFLOP/s: the tests are culled, not executed)"
        print(line, flush=True)
    r.set_tuning(rt.TUNE_PT_BVH, -1)


with tempfile.TemporaryDirectory() as d:
    for depth, spp in [(4, 16), (5, 16), (6, 4)] if not quick else [(4, 2), (6, 1)]:
        p = os.path.join(d, f"complex{depth}.scn")
        rt.write_complex_scene(p, depth)
        w, h = (3840, 2160) if not quick else (960, 540)
        spheres, cam = rt.read_scene(p, w, h)
        pt_case(f"C4 complex maxDepth {depth}" + (" (chunked smem staging)" if spheres.size * 16 > 96 * 1024 else ""), spheres, cam, w, h, spp)
w, h = (7680, 4320) if not quick else (1920, 1080)
spheres, cam = rt.cornell_scene(w, h)
pt_case("C5 cornell 8K, 1 GPU, 16 of 1024 spp", spheres, cam, w, h, 16 if not quick else 2)
prims = rt.whitted_create_scene(0)
for (w, h) in [(1920, 1080), (3840, 2160)]:
    r.whitted_upload(prims, w, h)
    r.set_counting(True); r.whitted_launch(); c = r.counters(); r.set_counting(False)
    ts = []
    for _ in range(4):
        r.timer_begin(); r.whitted_launch(); ts.append(r.timer_end())
    t = min(ts[1:]); rays = c["nearest_queries"] + c["shadow_queries"]; flop = 16.0 * c["sphere_tests"] + 12.0 * c["plane_tests"]
    print(f"Whitted {w}x{h}: {t:.3f} ms  {rays / t / 1e3:.0f} Mrays/s  {flop / t / 1e9:.2f} TFLOP/s = {100 * flop / (t * 1e-3) / peak:.1f}% of FP32 peak", flush=True)
r.close()
