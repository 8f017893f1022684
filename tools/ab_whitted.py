"""Whitted 1080p kernel time under the scheduling knobs (CUDA events, best of 7 after warm-up): RT_B200_LIB=<so> python tools/ab_whitted.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
rt = g.load()
r = rt.Renderer(0)
prims = rt.whitted_create_scene(0)
size = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
r.whitted_upload(prims, *size)
out = []
for name, knobs in [("blocks+filler10", {rt.TUNE_WHITTED_COST_ORDER: 1, rt.TUNE_WHITTED_BLOCKS: 1, rt.TUNE_WHITTED_FILLER_PCT: 10}),
                    ("filler25", {rt.TUNE_WHITTED_FILLER_PCT: 25}), ("filler40", {rt.TUNE_WHITTED_FILLER_PCT: 40}), ("filler60", {rt.TUNE_WHITTED_FILLER_PCT: 60}),
                    ("filler0", {rt.TUNE_WHITTED_FILLER_PCT: 0}), ("lists", {rt.TUNE_WHITTED_COST_ORDER: 1, rt.TUNE_WHITTED_BLOCKS: 0}),
                    ("screen-order", {rt.TUNE_WHITTED_COST_ORDER: 0, rt.TUNE_WHITTED_BLOCKS: 0})]:
    for k, v in knobs.items():
        try:
            r.set_tuning(k, v)
        except rt.RtError:
            pass
    for _ in range(3): r.whitted_launch()
    r.sync()
    ts = []
    for _ in range(7):
        r.timer_begin(); r.whitted_launch(); ts.append(r.timer_end())
    out.append("%s %.3f ms" % (name, min(ts)))
print("%-24s whitted %dx%d: %s" % (os.path.basename(os.environ.get("RT_B200_LIB", "product")), size[0], size[1], " | ".join(out)))
r.close()
