"""Differential fuzzing of the exact hierarchy (csrc/pt_bvh.cuh) against the reference-order loop over every sphere, on the
GPU through the C ABI: random sphere clouds (sizes from 1e-3 to 1e4, clusters, duplicates, nested and touching spheres,
mirrors and glass, several lights), random cameras (outside, inside the cloud, 1e3..1e5 units away), both integrators.
Colours, RNG state and pixels must be bit-identical.  Then the same for the Whitted tracer (pixels and hit IDs).  Usage: python tools/bvh_fuzz.py [n_scenes] [seed]"""
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
print(f"path tracer: {n_scenes} scenes x 2 integrators, {w}x{h} x {spp} spp: {bad} mismatches ({time.time() - t0:.0f} s)")

# ---- Whitted tracer: random primitive tables (spheres of mixed sizes around the fixed camera at (0, 0.25, -7), some planes,
# 1-4 sphere lights, mirrors and glass), hierarchy (RT_TUNE_WHITTED_BVH 1) against the run tables (0): pixels and hit IDs.
box = rt.whitted_create_scene(0)
wbad = 0
t0 = time.time()
ww, wh = 160, 120
for it in range(n_scenes):
    n = int(np.exp(rs.uniform(np.log(4), np.log(1500))))
    t = np.zeros(n, box.dtype)
    t[:] = box[2]                                  # a sphere record as template
    extent = float(np.exp(rs.uniform(np.log(2.0), np.log(60.0))))
    c = rs.uniform(-extent, extent, (n, 3)); c[:, 2] = rs.uniform(2.0, 2.0 + 2 * extent, n)
    rad = np.exp(rs.uniform(np.log(extent * 2e-3), np.log(extent * rs.choice([0.03, 0.15, 0.5])), n)).astype(np.float32)
    t["center"][:, :3] = c.astype(np.float32)
    t["radius"] = rad; t["sq_radius"] = rad * rad; t["r_radius"] = np.float32(1.0) / rad
    t["m_color"][:, :3] = rs.uniform(0.1, 1.0, (n, 3))
    t["m_diff"] = rs.uniform(0, 1, n) * (rs.rand(n) < 0.8); t["m_spec"] = rs.uniform(0, 1.5, n) * (rs.rand(n) < 0.6)
    t["m_refl"] = rs.uniform(0, 0.9, n) * (rs.rand(n) < 0.3)
    refr = rs.rand(n) < 0.15
    t["m_refr"] = np.where(refr, rs.uniform(0.3, 1.0, n), 0); t["m_refr_index"] = np.where(refr, rs.uniform(1.1, 1.6, n), 0)
    t["is_light"] = 0
    parts = [t]
    if rs.rand() < 0.6:
        parts.insert(rs.randint(0, 2), box[[0, 8, 9, 10, 11, 12]][: rs.randint(1, 7)])        # some of the walls
    lights = box[13:16][: rs.randint(1, 4)].copy()
    lights["center"][:, :3] += rs.uniform(-2, 2, (lights.size, 3)).astype(np.float32)
    parts.insert(rs.randint(0, len(parts) + 1), lights)
    prims = np.concatenate(parts)
    if prims.size > 12 and rs.rand() < 0.5:        # exact duplicates at other indices (ties: the lower index must win)
        sp = np.flatnonzero((prims["type"] == 1) & (prims["is_light"] == 0))
        d, s_ = rs.choice(sp, sp.size // 8), rs.choice(sp, sp.size // 8)
        for f in ("center", "radius", "sq_radius", "r_radius"):
            prims[f][d] = prims[f][s_]
    outs = []
    for bvh in (0, 1):
        r.set_tuning(rt.TUNE_WHITTED_BVH, bvh)
        outs.append(r.whitted_render(prims, ww, wh, want_hit_ids=True))
    if not (np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])):
        wbad += 1
        print(f"WHITTED MISMATCH scene {it}: n={prims.size} extent={extent:.3g}, {int(np.count_nonzero(outs[0][1] != outs[1][1]))} hit IDs differ", flush=True)
r.set_tuning(rt.TUNE_WHITTED_BVH, -1)
print(f"Whitted: {n_scenes} scenes, {ww}x{wh}: {wbad} mismatches ({time.time() - t0:.0f} s)")
r.close()
sys.exit(1 if bad or wbad else 0)
