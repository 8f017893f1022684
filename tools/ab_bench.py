"""A/B timing of library variants: RT_B200_LIB=<so> python tools/ab_bench.py  -> one line of kernel times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
rt = g.load()
r = rt.Renderer(0)
prims = rt.whitted_create_scene(0)
r.whitted_upload(prims, 1920, 1080)
for _ in range(3): r.whitted_launch()
r.sync()
tw = []
for _ in range(5):
    r.timer_begin(); r.whitted_launch(); tw.append(r.timer_end())
r.set_tuning(rt.TUNE_WHITTED_COST_ORDER, 0)
tw0 = []
for _ in range(5):
    r.timer_begin(); r.whitted_launch(); tw0.append(r.timer_end())
r.set_tuning(rt.TUNE_WHITTED_COST_ORDER, 1)
sph, cam = rt.cornell_scene(1024, 768)
seeds = rt.reference_seeds(1024, 768)
tp = []
for integ in (0, 1):
    for _ in range(3):
        r.pt_resize(1024, 768, seeds); r.pt_set_scene(sph); r.pt_set_camera(cam)
        r.timer_begin(); r.pt_launch(integ, 32); tp.append((integ, r.timer_end()))
print("%-28s whitted1080p %.3f ms (screen order %.3f) | pt32spp %.3f ms | dl32spp %.3f ms" % (os.path.basename(os.environ.get("RT_B200_LIB", "product")), min(tw), min(tw0),
      min(t for i, t in tp if i == 0), min(t for i, t in tp if i == 1)))
r.close()
