"""Scene 0 (config 2) at 1920x1080 with the seven non-light spheres behind the hierarchy's root box (RT_TUNE_WHITTED_BVH 1) against the run tables."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
rt = g.load()
r = rt.Renderer(0)
prims = rt.whitted_create_scene(0)
for bvh in (0, 1, 0, 1):
    r.set_tuning(rt.TUNE_WHITTED_BVH, bvh)
    r.whitted_upload(prims, 1920, 1080)
    for _ in range(3): r.whitted_launch()
    r.sync()
    ts = []
    for _ in range(6):
        r.timer_begin(); r.whitted_launch(); ts.append(r.timer_end())
    print(f"hierarchy {bvh}: {min(ts):.3f} ms")
r.close()
