import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np
import __graft_entry__ as g
rt = g.load()
r = rt.Renderer(0)
prims = rt.r306_create_scene()
for (w, h, split) in [(800, 600, 0), (800, 600, 1), (1920, 1080, 0), (1920, 1080, 1)]:
    r.set_tuning(rt.TUNE_R306_SPLIT, split)
    r.r306_upload(prims, w, h)
    for _ in range(3): r.r306_launch()
    r.sync()
    t = []
    for _ in range(5):
        r.timer_begin(); r.r306_launch(); t.append(r.timer_end())
    print("r306 %dx%d %s: kernels %.3f ms" % (w, h, "one sub-sample per work unit" if split else "one pixel per work unit", min(t)))
r.close()
