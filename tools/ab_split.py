"""Whitted frame time against the share of the SMs' CTA slots given to whitted_split_kernel (RT_TUNE_WHITTED_SPLIT_BLOCKS) and the filler
share of the main kernel; CUDA events, best of 9.  python tools/ab_split.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
rt = g.load()
r = rt.Renderer(0)
prims = rt.whitted_create_scene(0)
for (w, h) in [(1920, 1080), (960, 540), (3840, 2160)]:
    r.whitted_upload(prims, w, h)
    out = []
    for blocks in (0, 10, 8, 6, 4, 3, 2):
        for filler in (10, 25):
            r.set_tuning(rt.TUNE_WHITTED_SPLIT_BLOCKS, blocks); r.set_tuning(rt.TUNE_WHITTED_FILLER_PCT, filler)
            for _ in range(3): r.whitted_launch()
            r.sync()
            ts = []
            for _ in range(9):
                r.timer_begin(); r.whitted_launch(); ts.append(r.timer_end())
            out.append("b%d/f%d %.3f" % (blocks, filler, min(ts)))
    print("%dx%d: %s" % (w, h, " | ".join(out)))
r.close()
