"""Whitted frame time with and without the shadow-candidate grid (RT_TUNE_WHITTED_GRID), CUDA events, best of 9 after warm-up;
the two frames must be the same bytes.  RT_B200_LIB=<so> python tools/ab_grid.py [w h]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
rt = g.load()
r = rt.Renderer(0)
sizes = [(int(sys.argv[1]), int(sys.argv[2]))] if len(sys.argv) > 2 else [(1920, 1080), (960, 540), (3840, 2160)]
for scene in (0,):
    prims = rt.whitted_create_scene(scene)
    for (w, h) in sizes:
        out, frames = [], []
        for rep in range(2):
            for grid, split in ((1, 1), (1, 2), (1, 0), (2, 0), (0, 0)):
                r.set_tuning(rt.TUNE_WHITTED_GRID, grid); r.set_tuning(rt.TUNE_WHITTED_SPLIT, split)
                r.whitted_upload(prims, w, h)
                for _ in range(3): r.whitted_launch()
                r.sync()
                ts = []
                for _ in range(9):
                    r.timer_begin(); r.whitted_launch(); ts.append(r.timer_end())
                out.append("grid=%d split=%d %.3f ms" % (grid, split, min(ts)))
                frames.append(r.whitted_download())
        same = all(np.array_equal(frames[0], f) for f in frames[1:])
        print("%-22s scene %d %dx%d: %s | same bytes: %s" % (os.path.basename(os.environ.get("RT_B200_LIB", "product")), scene, w, h, " | ".join(out), same))
r.close()
