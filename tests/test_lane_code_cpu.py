"""The device lane code (csrc/*.cuh), compiled for the host by tests/devsim, against the oracle.
CPU only.  This checks the state machines the CUDA kernels are built from -- not the product path;
the GPU parity tests (test_gpu_parity.py) are the ones that call librt_b200.so."""
import ctypes

import numpy as np
import pytest

from conftest import vp, load_smallpt_golden


def test_sincos_bit_exact_on_the_whole_input_domain(devsim, orc):
    """Every argument the path can pass to sin/cos is 2*pi*GetRandom(): 2^23 values.  All of them."""
    i = np.arange(1 << 23, dtype=np.uint32)
    u = ((i | 0x40000000).view(np.float32) - np.float32(2)) / np.float32(2)
    a = (np.float32(2.0) * np.float32(3.14159265358979323846) * u).astype(np.float32)
    s, c, s2, c2 = (np.zeros_like(a) for _ in range(4))
    devsim.devsim_sincos(vp(a), vp(s), vp(c), ctypes.c_long(a.size))
    orc.oracle_libm_sincosf(vp(a), vp(s2), vp(c2), ctypes.c_long(a.size))
    assert np.array_equal(s.view(np.uint32), s2.view(np.uint32))
    assert np.array_equal(c.view(np.uint32), c2.view(np.uint32))


def test_expf_and_gamma_match_host_libm(devsim, orc):
    rs = np.random.RandomState(0)
    x = np.concatenate([-(rs.rand(2_000_000) * 70), -np.logspace(-6, 2, 20000), [0.0, -0.0, -100, -104, -200]]).astype(np.float32)
    a, b = np.zeros_like(x), np.zeros_like(x)
    devsim.devsim_expf(vp(x), vp(a), ctypes.c_long(x.size))
    orc.oracle_libm_expf(vp(x), vp(b), ctypes.c_long(x.size))
    assert np.count_nonzero(a.view(np.uint32) != b.view(np.uint32)) <= 2     # documented: < 1e-8 of inputs may differ by 1 ulp
    v = np.concatenate([rs.rand(2_000_000) * 1.2 - 0.1, np.logspace(-30, 0, 20000), [0, 1, 1e-40, 2, -1, 0.5]]).astype(np.float32)
    g1, g2 = np.zeros(v.size, np.int32), np.zeros(v.size, np.int32)
    devsim.devsim_to_int_gamma(vp(v), vp(g1), ctypes.c_long(v.size))
    orc.oracle_libm_to_int_gamma(vp(v), vp(g2), ctypes.c_long(v.size))
    assert np.array_equal(g1, g2)


def test_pow20_close_to_libm_pow(devsim, orc):
    for v in np.random.RandomState(1).rand(2000).astype(np.float32):
        a, b = devsim.devsim_pow20(float(v)), orc.oracle_libm_pow20(float(v))
        assert a == b or abs(a - b) <= 12 * np.spacing(b)


def test_whitted_lanes_equal_oracle(devsim, orc, rt):
    prims = rt.whitted_create_scene(0)
    for (w, h) in [(120, 90), (37, 29)]:
        px, hits, ctr = np.zeros((h, w, 4), np.uint8), np.zeros((h, w, 9), np.int32), np.zeros(5, np.uint64)
        devsim.devsim_whitted(vp(px), vp(hits), w, h, vp(prims), prims.size, 0, 1, 8, vp(ctr), None, 1)
        px_o, hits_o, ctr_o = np.zeros((h, w, 4), np.uint8), np.zeros((h, w, 9), np.int32), np.zeros(5, np.uint64)
        orc.oracle_whitted_render(vp(px_o), vp(hits_o), w, h, vp(prims), prims.size, 4, vp(ctr_o))
        assert np.array_equal(px, px_o) and np.array_equal(hits, hits_o)
        assert list(ctr[:4]) == list(ctr_o[:4]) and ctr[4] == w * h * 9


def test_whitted_hot_runs_skip_only_never_hit_primitives(devsim, orc, rt):
    """Timed launches walk the runs without never-hit primitives (the zeroed slot of scene 0, unknown types): same
    pixels and hit IDs as the oracle, which tests every primitive."""
    for scene in (0, 1):
        prims = rt.whitted_create_scene(scene)
        if scene == 1:
            prims = prims.copy(); prims["type"][5] = 7          # an unknown type: intersect() returns MISS (RNO:150-160)
        w, h = (96, 72) if scene == 0 else (40, 30)
        px, hits = np.zeros((h, w, 4), np.uint8), np.zeros((h, w, 9), np.int32)
        devsim.devsim_whitted(vp(px), vp(hits), w, h, vp(prims), prims.size, 0, 1, 8, None, None, 2)
        px_o, hits_o = np.zeros((h, w, 4), np.uint8), np.zeros((h, w, 9), np.int32)
        orc.oracle_whitted_render(vp(px_o), vp(hits_o), w, h, vp(prims), prims.size, 4, None)
        assert np.array_equal(hits, hits_o) and np.array_equal(px, px_o)


def test_whitted_lanes_sharded_equal_unsharded(devsim, rt):
    prims = rt.whitted_create_scene(0)
    w, h = 50, 41
    full = np.zeros((h, w, 4), np.uint8)
    devsim.devsim_whitted(vp(full), None, w, h, vp(prims), prims.size, 0, 1, 8, None, None, 1)
    parts = np.zeros((h, w, 4), np.uint8)
    for rank in range(3):
        devsim.devsim_whitted(vp(parts), None, w, h, vp(prims), prims.size, rank, 3, 4, None, None, 1)
    assert np.array_equal(full, parts)


@pytest.mark.parametrize("scene", ["cornell", "caustic3", "simple", "complex"])
def test_smallpt_lanes_equal_reference_fixture(devsim, rt, scene):
    g = load_smallpt_golden(rt, scene)
    w, h = g["w"], g["h"]
    for integ, tag in [(0, "pt"), (1, "dl")]:
        col, sd, pix = np.zeros(3 * w * h, np.float32), g["seeds_in"].copy(), np.zeros(w * h, np.uint32)
        devsim.devsim_pt(integ, vp(g["spheres"]), g["spheres"].size, vp(g["camera"]), w, h, 0, g["passes"], 0,
                         vp(col), vp(sd), vp(pix), 0, 1, 8, None, 1 << 30)
        assert np.array_equal(col.view(np.uint32), g[tag + "_colors"]), (scene, tag)
        assert np.array_equal(sd, g[tag + "_seeds"]) and np.array_equal(pix, g[tag + "_pixels"]), (scene, tag)


def test_smallpt_lane_counters_equal_oracle(devsim, orc, rt):
    g = load_smallpt_golden(rt, "cornell")
    w, h = g["w"], g["h"]
    col, sd, ctr = np.zeros(3 * w * h, np.float32), g["seeds_in"].copy(), np.zeros(5, np.uint64)
    devsim.devsim_pt(0, vp(g["spheres"]), g["spheres"].size, vp(g["camera"]), w, h, 0, 3, 0, vp(col), vp(sd), None, 0, 1, 8, vp(ctr), 1 << 30)
    col_o, sd_o, ctr_o = np.zeros(3 * w * h, np.float32), g["seeds_in"].copy(), np.zeros(4, np.uint64)
    orc.oracle_pt_render(0, vp(g["spheres"]), g["spheres"].size, vp(g["camera"]), w, h, 0, 3, vp(col_o), vp(sd_o), None, 2, vp(ctr_o))
    assert (ctr[4], ctr[0], ctr[1], ctr[2]) == tuple(ctr_o)


def test_r306_lanes_equal_the_reference_frame(devsim, rt):
    """raytracer3.0.06 (config 1) through the lane code on the host == the frame the reference's own Engine_Render
    produced (fixture from oracle/_ref), and == that code run now where it is built."""
    import hashlib, json, os, zlib
    from conftest import GOLDEN, graft
    g = json.load(open(os.path.join(GOLDEN, "r306_golden.json")))
    prims = rt.r306_create_scene()
    assert prims.size == 17
    for (w, h) in [(160, 120), (203, 131)]:
        img = np.zeros((h, w), np.uint32)
        devsim.devsim_r306(vp(img), w, h, vp(prims), prims.size)
        assert hashlib.sha256(img.tobytes()).hexdigest() == g["frames"][f"{w}x{h}"], (w, h)
    want = np.frombuffer(zlib.decompress(open(os.path.join(GOLDEN, "r306_160x120.u32.zlib"), "rb").read()), np.uint32).reshape(120, 160)
    img = np.zeros((120, 160), np.uint32)
    devsim.devsim_r306(vp(img), 160, 120, vp(prims), prims.size)
    assert np.array_equal(img, want)
    ref = os.path.join(graft.ORACLE_DIR, "_ref", "libref_r306.so")
    if os.path.exists(ref):
        L = ctypes.CDLL(ref)
        a, b = np.zeros((150, 97), np.uint32), np.zeros((150, 97), np.uint32)
        devsim.devsim_r306(vp(a), 97, 150, vp(prims), prims.size)
        L.ref_r306_render(vp(b), 97, 150)
        assert np.array_equal(a, b)
        # other shapes of the scene table through the reference's own engine: one light, no light, a light that is not a
        # sphere (shaded without a shadow ray towards its zero m_Centre), lights first, a primitive of unknown type
        variants = []
        v = prims.copy(); v["m_light"][[14, 15, 16]] = 0; variants.append(v)
        v = prims.copy(); v["m_light"][:] = 0; variants.append(v)
        v = prims.copy(); v["m_light"][11] = 1; variants.append(v)
        variants.append(prims[[1, 14, 0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15, 16]].copy())
        v = prims.copy(); v["type"][5] = 3; variants.append(v)
        for k, v in enumerate(variants):
            a, b = np.zeros((131, 80), np.uint32), np.zeros((131, 80), np.uint32)
            devsim.devsim_r306(vp(a), 80, 131, vp(v), v.size)
            L.ref_r306_render_scene(vp(b), 80, 131, vp(v), v.size)
            assert np.array_equal(a, b), k


def _devsim_pt(devsim, integ, sph, cam, w, h, passes, seeds, mode):
    col, sd, pix = np.zeros(3 * w * h, np.float32), seeds.copy(), np.zeros(w * h, np.uint32)
    devsim.devsim_pt(integ, vp(sph), sph.size, vp(cam), w, h, 0, passes, 0, vp(col), vp(sd), vp(pix), 0, 1, 8, None, mode)
    return col.view(np.uint32), sd, pix


def _complex_scene(rt, tmp_path, depth, w, h):
    path = str(tmp_path / f"complex{depth}.scn")
    rt.write_complex_scene(path, depth)
    return rt.read_scene(path, w, h)


@pytest.mark.parametrize("depth,w,h,passes", [(2, 64, 48, 3), (3, 48, 36, 2), (4, 40, 30, 2)])
def test_bvh_traversal_equals_the_plain_sphere_loop_on_generated_scenes(devsim, rt, tmp_path, depth, w, h, passes):
    """pt_query_bvh (exact culling, csrc/pt_bvh.cuh) against the reference-order loop over every sphere: colours, RNG state
    and pixels bit-identical, both integrators.  The plain loop itself equals the reference (fixtures above)."""
    sph, cam = _complex_scene(rt, tmp_path, depth, w, h)
    seeds = rt.reference_seeds(w, h)
    st = np.zeros(5, np.int32)
    devsim.devsim_bvh_stats(vp(sph), sph.size, vp(st))
    assert st[1] == sph.size and st[2] >= 1 and st[0] >= 1          # every sphere once; the floor is in the always-tested list
    for integ in (0, 1):
        a = _devsim_pt(devsim, integ, sph, cam, w, h, passes, seeds, -2)
        b = _devsim_pt(devsim, integ, sph, cam, w, h, passes, seeds, 0)
        for x, y in zip(a, b):
            assert np.array_equal(x, y), (depth, integ)


def test_bvh_traversal_keeps_the_reference_tie_rule_and_odd_spheres(devsim, rt, tmp_path):
    """Exact ties (duplicated spheres: the HIGHER index must win, SPT/geomfunc.h:80-88), mirrors and glass among the
    small spheres, a zero-radius sphere, touching and nested spheres, a camera inside the cloud, and a camera so far away that the
    reference's own discriminant is mostly rounding noise (the hierarchy must reproduce that noise, not the geometry)."""
    sph, cam = _complex_scene(rt, tmp_path, 3, 40, 30)
    rs = np.random.RandomState(7)
    extra = sph[2:].copy()
    rs.shuffle(extra)
    dup = extra[:60].copy()                                   # exact copies at other indices, other colours
    dup["c"] = (0.9, 0.1, 0.1)
    scene = np.concatenate([sph, dup, dup[:20]])
    scene["refl"][10:40:3] = 1
    scene["refl"][11:40:3] = 2
    scene["rad"][50] = 0.0
    scene["p"][60] = scene["p"][61]
    perm = np.concatenate([[0, 1], 2 + rs.permutation(scene.size - 2)])
    scene = scene[perm].copy()
    for cam_pos in (None, (3.0, 21.0, 4.0), (2100.0, 16000.0, 7700.0)):      # default, inside the cloud, 18 000 units away
        c = cam.copy()
        if cam_pos:
            c["orig"] = cam_pos
            rt.update_camera(c, 40, 30)
        seeds = rt.reference_seeds(40, 30, seed=3)
        for integ in (0, 1):
            a = _devsim_pt(devsim, integ, scene, c, 40, 30, 3, seeds, -2)
            b = _devsim_pt(devsim, integ, scene, c, 40, 30, 3, seeds, 0)
            for x, y in zip(a, b):
                assert np.array_equal(x, y), (cam_pos, integ)


def test_bvh_traversal_on_random_clouds(devsim, rt):
    """Random sphere clouds of mixed sizes (no floor, several lights, a few huge spheres), rays that leave the scene."""
    rs = np.random.RandomState(11)
    for n in (5, 37, 300):
        sph, cam = rt.cornell_scene(32, 24)
        scene = np.zeros(n, sph.dtype)
        scene["p"] = np.stack([rs.uniform(0, 100, n), rs.uniform(0, 80, n), rs.uniform(0, 150, n)], 1)
        scene["rad"] = np.exp(rs.uniform(np.log(0.05), np.log(9.0), n))
        scene["c"] = rs.uniform(0.2, 0.9, (n, 3))
        scene["refl"] = rs.randint(0, 3, n)
        lights = rs.choice(n, max(1, n // 40), replace=False)
        scene["e"][lights] = 12
        if n > 30:
            scene["rad"][:2] = (600.0, 5000.0); scene["p"][1, 1] = -5000.0
        seeds = rt.reference_seeds(32, 24, seed=n)
        for integ in (0, 1):
            a = _devsim_pt(devsim, integ, scene, cam, 32, 24, 4, seeds, -2)
            b = _devsim_pt(devsim, integ, scene, cam, 32, 24, 4, seeds, 0)
            for x, y in zip(a, b):
                assert np.array_equal(x, y), (n, integ)


def test_whitted_hierarchy_equals_oracle_on_generated_sphere_scenes(devsim, orc, rt, tmp_path):
    """The Whitted tracer on .scn sphere scenes with the non-light spheres in the exact hierarchy (whitted_bvh.cuh): pixels
    and hit IDs equal the oracle's, which tests every primitive -- also with planes, several lights, glass and mirrors."""
    for depth, w, h in [(3, 48, 36), (4, 40, 30)]:
        sph, cam = _complex_scene(rt, tmp_path, depth, w, h)
        prims = rt.whitted_from_spheres(sph, cam)
        variants = [prims]
        v = prims.copy()
        box = rt.whitted_create_scene(0)
        v = np.concatenate([box[[0, 8, 9, 10, 11, 12]], v, box[13:16]])          # the six walls and the three lights of scene 0 around it
        v["m_refl"][10:60:4] = 0.6; v["m_refr"][11:60:4] = 0.8; v["m_refr_index"][11:60:4] = 1.3
        variants.append(v)
        for k, pv in enumerate(variants):
            px, hits = np.zeros((h, w, 4), np.uint8), np.zeros((h, w, 9), np.int32)
            devsim.devsim_whitted(vp(px), vp(hits), w, h, vp(pv), pv.size, 0, 1, 8, None, None, 3)
            px_o, hits_o = np.zeros((h, w, 4), np.uint8), np.zeros((h, w, 9), np.int32)
            orc.oracle_whitted_render(vp(px_o), vp(hits_o), w, h, vp(pv), pv.size, 4, None)
            assert np.array_equal(hits, hits_o), (depth, k)
            assert np.array_equal(px, px_o), (depth, k)


def test_whitted_shadow_culls_change_nothing(devsim, orc, rt, tmp_path):
    """The timed path of the Whitted tracer (hot runs, no counting, the shadow-round culls of whitted_lane.cuh with the tables
    of build_w_cull) against the oracle, which tests every primitive: the reference's two scenes, scene 0 with a light moved to
    within a hair of a wall (that wall loses its cull entry), with a light below the floor (both sides lit: no entry either), a
    sphere scene walked through the hierarchy with walls around it, and a handful of random rooms (tools/cull_fuzz.py runs
    thousands, plus a self-check that tables without the margins are caught)."""
    import importlib.util, os
    box = rt.whitted_create_scene(0)
    near = box.copy(); near["center"][13, 1] = np.float32(6.7495)            # 0.0004 below the ceiling plane y = 6.75
    below = box.copy(); below["center"][14, 1] = np.float32(-9.0)
    st = np.zeros(4, np.int32)
    devsim.devsim_whitted_cull_stats(vp(box), box.size, vp(st))
    assert list(st[:3]) == [1, 6, 1]                                          # all six walls and the sphere run are cullable
    devsim.devsim_whitted_cull_stats(vp(near), near.size, vp(st))
    assert st[1] == 5
    devsim.devsim_whitted_cull_stats(vp(below), below.size, vp(st))
    assert st[1] == 5 and st[2] == 0                                          # the floor separates the lights; no box face has them all beyond it
    cases = [(box, 96, 72, 4), (rt.whitted_create_scene(1), 48, 36, 4), (near, 64, 48, 4), (below, 64, 48, 4)]
    sph, cam = _complex_scene(rt, tmp_path, 3, 40, 30)
    v = np.concatenate([box[[0, 8, 9, 10, 11, 12]], rt.whitted_from_spheres(sph, cam), box[13:16]])
    cases.append((v, 40, 30, 5))
    spec = importlib.util.spec_from_file_location("cull_fuzz", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "cull_fuzz.py"))
    fz = importlib.util.module_from_spec(spec); spec.loader.exec_module(fz)
    rs = np.random.RandomState(5)
    cases += [(fz.random_scene(rt, rs, box), 40, 30, 4) for _ in range(40)]
    for k, (prims, w, h, mode) in enumerate(cases):
        px, hits = np.zeros((h, w, 4), np.uint8), np.zeros((h, w, 9), np.int32)
        devsim.devsim_whitted(vp(px), vp(hits), w, h, vp(prims), prims.size, 0, 1, 8, None, None, mode)
        px_o, hits_o = np.zeros((h, w, 4), np.uint8), np.zeros((h, w, 9), np.int32)
        orc.oracle_whitted_render(vp(px_o), vp(hits_o), w, h, vp(prims), prims.size, 4, None)
        assert np.array_equal(hits, hits_o), k
        assert np.array_equal(px, px_o), k


def test_whitted_blocked_light_times_an_overflowed_specular_term_is_nan_like_the_reference(devsim, orc, rt, ref_whitted):
    """The reference multiplies both shading terms by shade = 0 for a blocked light (RNO:250, 270); when a ray's direction has blown up
    (refraction directions are not re-normalised), pow(V.R, 20) is inf and inf * 0 = NaN: the accumulator turns NaN and the pixel black.
    Round 1 skipped blocked lights outright and drew such pixels lit.  The scene is the random room of tools/cull_fuzz.py that exposed it
    (tests/golden/whitted_blocked_light_nan.npy); the reference's own code, the oracle and the lane code must agree, counting and timed."""
    import os
    from conftest import GOLDEN
    prims = np.load(os.path.join(GOLDEN, "whitted_blocked_light_nan.npy"))
    w, h = 64, 48
    px_o, hits_o = np.zeros((h, w, 4), np.uint8), np.zeros((h, w, 9), np.int32)
    orc.oracle_whitted_render(vp(px_o), vp(hits_o), w, h, vp(prims), prims.size, 4, None)
    px_r = np.zeros((h, w, 4), np.uint8)
    # the room is open and the reference reads prims[-1] after a miss (RNO:370-432): give it an all-zero record there (spawns nothing,
    # what the oracle defines) instead of whatever the allocator left in front of the table
    padded = np.zeros(prims.size + 1, prims.dtype); padded[1:] = prims
    ref_whitted.ref_whitted_render(vp(px_r), w, h, ctypes.c_void_p(padded.ctypes.data + prims.dtype.itemsize), prims.size)
    assert np.array_equal(px_o, px_r)
    assert not px_o[29, 29, :3].any() and not px_o[30, 29, :3].any()      # the two pixels whose accumulator is NaN
    devsim.devsim_whitted_redo_pixels.restype = ctypes.c_long
    for mode in (1, 2, 4):
        px, hits = np.zeros((h, w, 4), np.uint8), np.zeros((h, w, 9), np.int32)
        devsim.devsim_whitted(vp(px), vp(hits), w, h, vp(prims), prims.size, 0, 1, 8, None, None, mode)
        assert np.array_equal(px, px_o) and np.array_equal(hits, hits_o), mode
    # the timed mode reported those pixels (and few others) for the exact pass; the reference's own scene reports none
    assert 2 <= devsim.devsim_whitted_redo_pixels() < w * h // 8
    box = rt.whitted_create_scene(0)
    px, hits = np.zeros((h, w, 4), np.uint8), np.zeros((h, w, 9), np.int32)
    devsim.devsim_whitted(vp(px), vp(hits), w, h, vp(box), box.size, 0, 1, 8, None, None, 4)
    assert devsim.devsim_whitted_redo_pixels() == 0


def test_whitted_one_lane_per_subsample_with_ordered_logs_equals_one_lane_per_pixel(devsim, orc, rt):
    """whitted_split_kernel's arithmetic on the lane simulator (mode 6): every sub-sample of a pixel traced on its own with the addends of
    its rays logged (w_finalize<.., SPLIT>, zero triples skipped), the nine logs added in order -- the same bytes and hit IDs as the
    reference's one running accumulator, on the reference's scenes, through the grid and the primary-ray tiles, and on the room whose
    accumulators turn NaN."""
    import os
    from conftest import GOLDEN
    for prims, (w, h) in [(rt.whitted_create_scene(0), (160, 90)), (rt.whitted_create_scene(0), (61, 37)),
                          (np.load(os.path.join(GOLDEN, "whitted_blocked_light_nan.npy")), (64, 48))]:
        px_o, hits_o = np.zeros((h, w, 4), np.uint8), np.zeros((h, w, 9), np.int32)
        orc.oracle_whitted_render(vp(px_o), vp(hits_o), w, h, vp(prims), prims.size, 4, None)
        px, hits = np.zeros((h, w, 4), np.uint8), np.zeros((h, w, 9), np.int32)
        devsim.devsim_whitted(vp(px), vp(hits), w, h, vp(prims), prims.size, 0, 1, 8, None, None, 6)
        assert np.array_equal(px, px_o) and np.array_equal(hits, hits_o), (w, h)


def test_whitted_tile_and_grid_words_contain_every_accepted_hit_and_every_blocker(devsim, rt):
    """The per-scene tables of whitted_lane.cuh checked directly (not through a rendered frame): for every primary ray of a frame the hit the
    reference's loop accepts must be in the ray's tile word, and every primitive that blocks a shadow ray of the hit point must be in the
    word of the point's grid cell -- on the reference's scene and on random rooms of tools/cull_fuzz.py.  The words must also stay SHARP
    on the reference's scene (a table of all-ones would pass the first two checks and cost the kernels their speed)."""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("cull_fuzz", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "cull_fuzz.py"))
    fz = importlib.util.module_from_spec(spec); spec.loader.exec_module(fz)
    box = rt.whitted_create_scene(0)
    out = np.zeros(7, np.int64)
    devsim.devsim_whitted_table_check(vp(box), box.size, 240, 135, vp(out))
    assert out[6] == 1 and out[0] == 0 and out[3] == 0, out
    assert out[1] / out[2] < 2.0            # primitives a primary ray still tests (of 16): 1.26 on this frame
    assert out[4] / out[5] < 1.7            # primitives a shadow batch still tests (of 13): 1.13
    rs = np.random.RandomState(5)
    with_grid = 0
    for _ in range(40):
        prims = fz.random_scene(rt, rs, box)
        devsim.devsim_whitted_table_check(vp(prims), prims.size, 64, 48, vp(out))
        assert out[0] == 0 and out[3] == 0, out
        with_grid += int(out[6])
    assert with_grid >= 30


def test_hierarchy_builder_on_degenerate_inputs(devsim, rt):
    """build_pt_bvh: every sphere ends up exactly once in the tree or in the always-tested list, and the depth stays below
    the traversal stack (64) -- one sphere, a thousand identical ones, NaN / inf / huge entries, a line of 5 000, 200 000
    random ones."""
    sph, _ = rt.cornell_scene(32, 24)
    rs = np.random.RandomState(3)

    def stats(s):
        st = np.zeros(5, np.int32)
        devsim.devsim_bvh_stats(vp(s), s.size, vp(st))
        return st

    one = np.zeros(1, sph.dtype); one["rad"] = 1
    st = stats(one); assert st[1] == 1 and st[3] < 64
    same = np.zeros(1000, sph.dtype); same["rad"] = 1; same["p"] = (1, 2, 3)
    st = stats(same); assert st[1] == 1000 and st[3] < 64 and st[2] == 0
    odd = np.zeros(100, sph.dtype); odd["rad"] = rs.rand(100) + 0.1; odd["p"] = rs.rand(100, 3) * 10
    odd["p"][3, 0] = np.nan; odd["rad"][5] = np.inf; odd["p"][7, 1] = 1e30
    st = stats(odd); assert st[1] == 100 and st[2] >= 3          # the three odd ones are tested by every query
    line = np.zeros(5000, sph.dtype); line["rad"] = 0.1; line["p"][:, 0] = np.arange(5000)
    st = stats(line); assert st[1] == 5000 and st[3] < 64
    big = np.zeros(200000, sph.dtype); big["rad"] = rs.rand(200000) * 0.5 + 0.01; big["p"] = rs.rand(200000, 3) * 1000
    st = stats(big); assert st[1] == 200000 and st[3] < 64 and st[4] >= 200000 // 8
