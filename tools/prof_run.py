"""Small fixed workload for ncu: one Whitted 1080p frame and one Cornell 1024x768 x 16 spp launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
rt = g.load()
r = rt.Renderer(0)
which = sys.argv[1] if len(sys.argv) > 1 else "both"
if which in ("both", "whitted"):
    prims = rt.whitted_create_scene(0)
    r.whitted_upload(prims, int(os.environ.get("PROF_W", 1920)), int(os.environ.get("PROF_H", 1080)))
    for _ in range(2):
        r.whitted_launch()
    r.sync()
if which in ("both", "pt"):
    sph, cam = rt.cornell_scene(1024, 768)
    r.pt_resize(1024, 768, rt.reference_seeds(1024, 768)); r.pt_set_scene(sph); r.pt_set_camera(cam)
    for _ in range(2):
        r.pt_launch(0, 16)
    r.sync()
if which == "bvh4":          # BASELINE config 4 itself: 783 spheres, 3840x2160 x 4 spp through the hierarchy, two launches
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "c4.scn")
        rt.write_complex_scene(p, 4)
        sph, cam = rt.read_scene(p, 3840, 2160)
    seeds = rt.reference_seeds(3840, 2160)
    r.pt_resize(3840, 2160, seeds); r.pt_set_scene(sph); r.pt_set_camera(cam)
    for _ in range(2):
        r.pt_launch(0, 4)
    r.sync()
if which == "bvh":           # the large-scene path tracer: 19 533 spheres, 1920x1080 x 4 spp, two launches
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "c6.scn")
        rt.write_complex_scene(p, 6)
        sph, cam = rt.read_scene(p, 1920, 1080)
    seeds = rt.reference_seeds(1920, 1080)
    r.pt_resize(1920, 1080, seeds); r.pt_set_scene(sph); r.pt_set_camera(cam)
    for _ in range(2):
        r.pt_launch(0, 4)
    r.sync()
r.close()
print("prof_run ok")
