"""Cornell 1024x768 x 32 spp (both integrators) and the 783-sphere scene with and without the sin/cos table."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
rt = g.load()
r = rt.Renderer(0)
sph, cam = rt.cornell_scene(1024, 768)
seeds = rt.reference_seeds(1024, 768)
with tempfile.TemporaryDirectory() as d:
    p = os.path.join(d, "c4.scn"); rt.write_complex_scene(p, 4)
    sph4, cam4 = rt.read_scene(p, 1920, 1080)
seeds4 = rt.reference_seeds(1920, 1080)
for tab in (0, 1, 0, 1):
    r.set_tuning(rt.TUNE_PT_SINCOS_TABLE, tab)
    out = []
    for integ in (0, 1):
        best = 1e9
        for _ in range(3):
            r.pt_resize(1024, 768, seeds); r.pt_set_scene(sph); r.pt_set_camera(cam)
            r.timer_begin(); r.pt_launch(integ, 32); best = min(best, r.timer_end())
        out.append(best)
    r.pt_resize(1920, 1080, seeds4); r.pt_set_scene(sph4); r.pt_set_camera(cam4); r.pt_launch(0, 1)
    best = 1e9
    for _ in range(3):
        r.pt_resize(1920, 1080, seeds4); r.pt_set_camera(cam4)
        r.timer_begin(); r.pt_launch(0, 4); best = min(best, r.timer_end())
    print(f"sin/cos table {tab}: cornell pt 32 spp {out[0]:.3f} ms | dl 32 spp {out[1]:.3f} ms | 783 spheres 1080p x 4 spp {best:.3f} ms")
r.close()
