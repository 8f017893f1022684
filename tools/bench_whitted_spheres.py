"""Whitted tracer on the generated sphere scenes (rt_whitted_from_spheres): kernel time per frame."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
rt = g.load()
r = rt.Renderer(0)
w, h = (int(v) for v in os.environ.get("AB_SIZE", "1920x1080").split("x"))
with tempfile.TemporaryDirectory() as d:
    for depth in (int(v) for v in os.environ.get("AB_DEPTHS", "3,4,5,6").split(",")):
        p = os.path.join(d, f"c{depth}.scn")
        rt.write_complex_scene(p, depth)
        spheres, cam = rt.read_scene(p, w, h)
        prims = rt.whitted_from_spheres(spheres, cam)
        r.whitted_upload(prims, w, h)
        r.set_counting(True); r.whitted_launch(); c = r.counters(); r.set_counting(False)
        rays = c["nearest_queries"] + c["shadow_queries"]
        flop = 16.0 * c["sphere_tests"] + 12.0 * c["plane_tests"]
        for bvh, name in ((0, "run tables"), (1, "exact hierarchy")):
            r.set_tuning(rt.TUNE_WHITTED_BVH, bvh)
            r.whitted_launch(); r.sync()
            ts = []
            for _ in range(3):
                r.timer_begin(); r.whitted_launch(); ts.append(r.timer_end())
            print(f"Whitted {prims.size} primitives {w}x{h} [{name}]: {min(ts):.2f} ms  {rays / min(ts) / 1e3:.0f} Mrays/s  {rays / (w * h):.1f} rays/pixel  {flop / min(ts) / 1e9:.2f} TFLOP/s reference-equivalent", flush=True)
        r.set_tuning(rt.TUNE_WHITTED_BVH, -1)
r.close()
