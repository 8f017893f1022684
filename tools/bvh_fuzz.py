"""Differential fuzzing of the exact hierarchy (csrc/pt_bvh.cuh) against the reference-order loop over every sphere, on the
GPU through the C ABI: random sphere clouds (sizes from 1e-3 to 1e4, clusters, duplicates, nested and touching spheres,
mirrors and glass, several lights), random cameras (outside, inside the cloud, 1e3..1e5 units away), both integrators.
Colours, RNG state and pixels must be bit-identical.  Usage: python tools/bvh_fuzz.py [n_scenes] [seed]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
rt = g.load()
n_scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rs = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
r = rt.Renderer(0)
w, h, spp = 192, 144, 2
sph0, cam0 = rt.cornell_scene(w, h)
bad = 0
t0 = time.time()
for it in range(n_scenes):
    n = int(np.exp(rs.uniform(np.log(3), np.log(6000))))
    sc = np.zeros(n, sph0.dtype)
    extent = float(np.exp(rs.uniform(np.log(1.0), np.log(2000.0))))
    if rs.rand() < 0.5:                       # clustered
        k = rs.randint(1, 8)
        centres = rs.uniform(-extent, extent, (k, 3))
        sc["p"] = centres[rs.randint(0, k, n)] + rs.normal(0, extent * rs.uniform(0.01, 0.3), (n, 3))
    else:
        sc["p"] = rs.uniform(-extent, extent, (n, 3))
    lo, hi = np.log(extent * 1e-4), np.log(extent * rs.choice([0.02, 0.2, 1.0]))
    sc["rad"] = np.exp(rs.uniform(lo, hi, n))
    sc["c"] = rs.uniform(0.1, 0.95, (n, 3))
    sc["refl"] = rs.choice([0, 0, 0, 1, 2], n)
    nl = max(1, int(rs.randint(1, 4)))
    li = rs.choice(n, min(nl, n), replace=False)
    sc["e"][li] = rs.uniform(2, 30)
    sc["refl"][li] = 0
    if n > 10 and rs.rand() < 0.5:            # exact duplicates at other indices, a nested pair, a zero radius, a huge floor
        d = rs.choice(n, n // 10, replace=False); s = rs.choice(n, n // 10, replace=False)
        sc["p"][d] = sc["p"][s]; sc["rad"][d] = sc["rad"][s]
        sc["p"][1] = sc["p"][2]; sc["rad"][1] = sc["rad"][2] * 0.5
        sc["rad"][3] = 0.0
    if rs.rand() < 0.5:
        sc["rad"][0] = extent * 5000.0; sc["p"][0] = (0, -extent * 5001.0, 0); sc["e"][0] = 0
    cam = cam0.copy()
    mode = rs.randint(0, 4)
    dist = [extent * 3, extent * 0.2, extent * 1e3, extent * 1e5][mode]
    dirv = rs.normal(0, 1, 3); dirv /= np.linalg.norm(dirv)
    cam["orig"] = (dirv * dist).astype(np.float32)
    cam["target"] = rs.uniform(-extent, extent, 3).astype(np.float32) * 0.3
    rt.update_camera(cam, w, h)
    seeds = rt.reference_seeds(w, h, seed=1000 + it)
    for integ in (0, 1):
        outs = []
        for bvh in (1, 0):
            r.set_tuning(rt.TUNE_PT_BVH, bvh)
            r.pt_resize(w, h, seeds); r.pt_set_scene(sc); r.pt_set_camera(cam)
            outs.append(r.pt_render(integ, spp))
        same = all(np.array_equal(outs[0][k].reshape(-1).view(np.uint32), outs[1][k].reshape(-1).view(np.uint32)) for k in ("seeds", "colors", "pixels"))
        if not same:
            bad += 1
            nd = int(np.count_nonzero(outs[0]["pixels"].reshape(-1) != outs[1]["pixels"].reshape(-1)))
            print(f"MISMATCH scene {it} integrator {integ}: n={n} extent={extent:.3g} camera mode {mode}, {nd} pixels differ", flush=True)
print(f"{n_scenes} scenes x 2 integrators, {w}x{h} x {spp} spp: {bad} mismatches ({time.time() - t0:.0f} s)")
r.close()
sys.exit(1 if bad else 0)
