"""Where the exact hierarchy overtakes the loop: prefixes of the 158-sphere generated scene, 1920x1080 x 4 spp, kernel ms (hierarchy / loop)."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
rt = g.load()
r = rt.Renderer(0)
w, h, spp = 1920, 1080, 4
with tempfile.TemporaryDirectory() as d:
    p = os.path.join(d, "c3.scn"); rt.write_complex_scene(p, 3)
    sph, cam = rt.read_scene(p, w, h)
seeds = rt.reference_seeds(w, h)
for n in (33, 48, 64, 80, 96, 128, 158):
    sc = sph[:n].copy()
    res = []
    for mode in (1, 0):
        r.set_tuning(rt.TUNE_PT_BVH, mode)
        r.pt_resize(w, h, seeds); r.pt_set_scene(sc); r.pt_set_camera(cam); r.pt_launch(0, 1)
        best = 1e9
        for _ in range(3):
            r.pt_resize(w, h, seeds); r.pt_set_camera(cam)
            r.timer_begin(); r.pt_launch(0, spp); best = min(best, r.timer_end())
        res.append(best)
    print(f"{n:4d} spheres: hierarchy {res[0]:.2f} ms, loop {res[1]:.2f} ms")
r.close()
