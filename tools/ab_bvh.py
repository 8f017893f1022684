"""A/B timing of the large-scene path tracer: RT_B200_LIB=<so> python tools/ab_bvh.py -> one line of kernel times (ms),
hierarchy / loop, for the generated scenes of depth 2..6 at 1920x1080 x 4 spp."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
rt = g.load()
r = rt.Renderer(0)
w, h, spp = (int(v) for v in os.environ.get("AB_SIZE", "1920x1080x4").split("x"))
out = []
with tempfile.TemporaryDirectory() as d:
    for depth in (int(v) for v in os.environ.get("AB_DEPTHS", "2,3,4,5,6").split(",")):
        p = os.path.join(d, f"c{depth}.scn")
        rt.write_complex_scene(p, depth)
        spheres, cam = rt.read_scene(p, w, h)
        seeds = rt.reference_seeds(w, h, seed=1)
        res = []
        for mode in (1, 0):
            if mode == 0 and depth > 4 and len(sys.argv) < 2:
                res.append(float("nan")); continue
            r.set_tuning(rt.TUNE_PT_BVH, mode)
            r.pt_resize(w, h, seeds); r.pt_set_scene(spheres); r.pt_set_camera(cam)
            r.pt_launch(0, 1)                       # builds and uploads the hierarchy (host work, not timed here)
            best = 1e30
            for _ in range(3):
                r.pt_resize(w, h, seeds); r.pt_set_camera(cam)
                r.timer_begin(); r.pt_launch(0, spp); best = min(best, r.timer_end())
            res.append(best)
        out.append(f"{spheres.size}: {res[0]:.2f} / {res[1]:.2f}")
print("%-26s %s" % (os.path.basename(os.environ.get("RT_B200_LIB", "product")), " | ".join(out)))
r.close()
