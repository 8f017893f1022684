"""Debug: lane participation of the Whitted rounds (library built with -DW_ROUND_STATS): RT_B200_LIB=variants/librt_stats.so python tools/round_stats.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
rt = g.load()
r = rt.Renderer(0)
prims = rt.whitted_create_scene(0)
for order in (1, 0):
    r.set_tuning(rt.TUNE_WHITTED_COST_ORDER, order)
    r.whitted_upload(prims, 1920, 1080)
    r.set_counting(True); r.whitted_launch(); c = r.counters(); r.set_counting(False)
    print("cost order", order, "nearest rounds: lane slots %d, with a query %d (%.1f %%); shadow rounds: lane slots %d, with a batch %d (%.1f %%)" % (
        c["nearest_queries"], c["shadow_queries"], 100.0 * c["shadow_queries"] / c["nearest_queries"],
        c["sphere_tests"], c["plane_tests"], 100.0 * c["plane_tests"] / max(c["sphere_tests"], 1)))
r.close()
