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
