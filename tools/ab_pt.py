"""Path tracer: step-aligned warps (RT_TUNE_PT_ALIGNED 1) versus the plain query loop (0): kernel times and bit-identity.
RT_B200_LIB=<so> python tools/ab_pt.py [spp]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
rt = g.load()
r = rt.Renderer(0)
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 32
scenes = {"cornell": rt.cornell_scene(1024, 768)}
rt.write_complex_scene("/tmp/ab_c4.scn", 4)
scenes["complex783"] = rt.read_scene("/tmp/ab_c4.scn", 1024, 768)
seeds = rt.reference_seeds(1024, 768)
for name, (sph, cam) in scenes.items():
    for integ in (0, 1):
        res = {}
        for aligned in (1, 0):
            r.set_tuning(rt.TUNE_PT_ALIGNED, aligned)
            t = []
            for _ in range(3):
                r.pt_resize(1024, 768, seeds); r.pt_set_scene(sph); r.pt_set_camera(cam)
                r.timer_begin(); r.pt_launch(integ, spp if name == "cornell" else max(1, spp // 8)); t.append(r.timer_end())
            out = r.pt_download()
            res[aligned] = (min(t), out)
        same = all(np.array_equal(res[0][1][k], res[1][1][k]) for k in ("pixels", "colors", "seeds"))
        print("%-12s %s  aligned %.3f ms  plain %.3f ms  identical: %s" % (name, "pt" if integ == 0 else "dl", res[1][0], res[0][0], same))
r.close()
