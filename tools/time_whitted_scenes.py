import os, sys, tempfile
sys.path.insert(0, os.getcwd())
import __graft_entry__ as g
rt = g.load()
r = rt.Renderer(0)
scenes = [("scene1", rt.whitted_create_scene(1))]
with tempfile.TemporaryDirectory() as d:
    for depth in (4, 6):
        p = os.path.join(d, "c.scn"); rt.write_complex_scene(p, depth)
        sph, cam = rt.read_scene(p, 1920, 1080)
        scenes.append(("%d spheres" % sph.size, rt.whitted_from_spheres(sph, cam)))
for name, prims in scenes:
    out = []
    for (w, h) in [(960, 540), (1920, 1080), (3840, 2160)]:
        r.whitted_upload(prims, w, h)
        for _ in range(2): r.whitted_launch()
        r.sync()
        ts = []
        for _ in range(5):
            r.timer_begin(); r.whitted_launch(); ts.append(r.timer_end())
        out.append("%dx%d %.3f ms" % (w, h, min(ts)))
    print(name, prims.size, "prims:", " | ".join(out))
