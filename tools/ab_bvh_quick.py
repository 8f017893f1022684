"""Hierarchy kernels only: 783 / 3 908 / 19 533 spheres at 3840x2160 (path tracing, 16 / 16 / 4 spp) and the Whitted tracer on the same
tables at 1920x1080 -- kernel time by CUDA events, best of 3.  RT_B200_LIB=<so> python tools/ab_bvh_quick.py"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
rt = g.load()
r = rt.Renderer(0)
out = []
with tempfile.TemporaryDirectory() as d:
    for depth, spp in [(4, 16), (5, 16), (6, 4)]:
        p = os.path.join(d, f"c{depth}.scn")
        rt.write_complex_scene(p, depth)
        w, h = 3840, 2160
        sph, cam = rt.read_scene(p, w, h)
        seeds = rt.reference_seeds(w, h, seed=1)
        r.pt_resize(w, h, seeds); r.pt_set_scene(sph); r.pt_set_camera(cam)
        ts = []
        for _ in range(3):
            r.pt_resize(w, h, seeds); r.pt_set_camera(cam)
            r.timer_begin(); r.pt_launch(0, spp); ts.append(r.timer_end())
        out.append(f"pt {sph.size}: {min(ts):.2f} ms")
        sph2, cam2 = rt.read_scene(p, 1920, 1080)
        prims = rt.whitted_from_spheres(sph2, cam2)
        r.whitted_upload(prims, 1920, 1080)
        ts = []
        for _ in range(4):
            r.timer_begin(); r.whitted_launch(); ts.append(r.timer_end())
        out.append(f"whitted {sph.size}: {min(ts[1:]):.2f} ms")
print("%-24s %s" % (os.path.basename(os.environ.get("RT_B200_LIB", "product")), " | ".join(out)))
r.close()
