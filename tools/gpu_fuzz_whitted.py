"""Differential fuzzing of the Whitted GPU path (default knobs: tables, split kernel, exact re-launch) against the oracle on the random
rooms of tools/cull_fuzz.py; pixels and hit IDs must be identical.  Usage: python tools/gpu_fuzz_whitted.py [n_scenes] [seed]"""
import ctypes, importlib.util, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
spec = importlib.util.spec_from_file_location("cull_fuzz", os.path.join(ROOT, "tools", "cull_fuzz.py"))
fz = importlib.util.module_from_spec(spec); spec.loader.exec_module(fz)
vp = fz.vp
n_scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 300
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 21
rt = g.load(); orc = g.oracle()
rs = np.random.RandomState(seed)
box = rt.whitted_create_scene(0)
bad = reported = split_scenes = 0
t0 = time.time()
with rt.Renderer(0) as r:
    for it in range(n_scenes):
        prims = fz.random_scene(rt, rs, box)
        w, h = [(96, 72), (160, 120), (61, 37), (256, 144)][it % 4]
        px, hits = r.whitted_render(prims, w, h, want_hit_ids=True)
        reported += r.whitted_redo_reports() > 0
        px_o, hits_o = np.zeros((h, w, 4), np.uint8), np.zeros((h, w, 9), np.int32)
        orc.oracle_whitted_render(vp(px_o), vp(hits_o), w, h, vp(prims), prims.size, 16, None)
        if not (np.array_equal(px, px_o) and np.array_equal(hits, hits_o)):
            bad += 1
            if bad <= 10:
                print(f"MISMATCH scene {it}: n={prims.size} {w}x{h}, {int(np.count_nonzero((px != px_o).any(axis=2)))} pixels differ", flush=True)
print(f"gpu fuzz: {n_scenes} random rooms, {reported} with pixels re-rendered by the exact launch: {bad} mismatches ({time.time() - t0:.0f} s)")
sys.exit(1 if bad else 0)
