"""Where the hierarchy overtakes the run tables in the Whitted tracer: prefixes of the 158-sphere generated scene, 1920x1080, kernel ms."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
rt = g.load()
r = rt.Renderer(0)
w, h = 1920, 1080
with tempfile.TemporaryDirectory() as d:
    p = os.path.join(d, "c3.scn"); rt.write_complex_scene(p, 3)
    sph, cam = rt.read_scene(p, w, h)
for n in (12, 20, 33, 48, 64, 96, 158):
    prims = rt.whitted_from_spheres(sph[:n].copy(), cam)
    res = []
    for mode in (1, 0):
        r.set_tuning(rt.TUNE_WHITTED_BVH, mode)
        r.whitted_upload(prims, w, h); r.whitted_launch(); r.sync()
        ts = []
        for _ in range(4):
            r.timer_begin(); r.whitted_launch(); ts.append(r.timer_end())
        res.append(min(ts))
    print(f"{n:4d} spheres: hierarchy {res[0]:.2f} ms, run tables {res[1]:.2f} ms")
r.close()
r = rt.Renderer(0)
for scene in (0, 1):
    prims = rt.whitted_create_scene(scene)
    n_tree = int(((prims["type"] == 1) & (prims["is_light"] == 0)).sum())
    res = []
    for mode in (1, 0):
        r.set_tuning(rt.TUNE_WHITTED_BVH, mode)
        r.whitted_upload(prims, w, h); r.whitted_launch(); r.sync()
        ts = []
        for _ in range(4):
            r.timer_begin(); r.whitted_launch(); ts.append(r.timer_end())
        res.append(min(ts))
    print(f"scene {scene} ({prims.size} primitives, {n_tree} non-light spheres): hierarchy {res[0]:.2f} ms, run tables {res[1]:.2f} ms")
r.close()
