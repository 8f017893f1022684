"""The smallpt oracle (oracle/oracle_smallpt.c) against outputs of the reference's own CPU path.  CPU only."""
import ctypes
import json
import os

import numpy as np
import pytest

from conftest import vp, f32, GOLDEN, load_smallpt_golden


def run_oracle(orc, integ, g, passes=None, pass0=0, threads=4, colors=None, seeds=None):
    w, h = g["w"], g["h"]
    col = np.zeros(3 * w * h, np.float32) if colors is None else colors
    sd = g["seeds_in"].copy() if seeds is None else seeds
    pix = np.zeros(w * h, np.uint32)
    ctr = np.zeros(4, np.uint64)
    orc.oracle_pt_render(integ, vp(g["spheres"]), g["spheres"].size, vp(g["camera"]), w, h, pass0,
                         g["passes"] if passes is None else passes, vp(col), vp(sd), vp(pix), threads, vp(ctr))
    return col, sd, pix, ctr


@pytest.mark.parametrize("scene", ["cornell", "caustic3", "simple", "complex"])
def test_oracle_matches_reference_fixture(orc, rt, scene):
    """colors / pixels / RNG state after N passes == what the reference's smallptCPU.cpp loop
    (UpdateRenderingCPU, unmodified) and RadianceDirectLighting produced -- bit for bit."""
    g = load_smallpt_golden(rt, scene)
    for integ, tag in [(0, "pt"), (1, "dl")]:
        col, sd, pix, _ = run_oracle(orc, integ, g)
        assert np.array_equal(col.view(np.uint32), g[tag + "_colors"]), (scene, tag)
        assert np.array_equal(sd, g[tag + "_seeds"]), (scene, tag)
        assert np.array_equal(pix, g[tag + "_pixels"]), (scene, tag)


def test_known_answers_from_compiled_reference(orc):
    kat = json.load(open(os.path.join(GOLDEN, "smallpt_kat.json")))
    for case in kat["get_random"]:
        s0, s1 = ctypes.c_uint32(case["start"][0]), ctypes.c_uint32(case["start"][1])
        for a, b, fbits in case["calls"]:
            f = orc.oracle_pt_get_random(ctypes.byref(s0), ctypes.byref(s1))
            assert ("%08x" % s0.value, "%08x" % s1.value) == (a, b)
            assert np.float32(f).view(np.uint32) == int(fbits, 16)
    for case in kat["sphere_intersect"]:
        sph = np.zeros(11, np.float32)
        sph[:4] = [f32(v) for v in case["sphere"]]
        o = np.array([f32(v) for v in case["o"]], np.float32)
        d = np.array([f32(v) for v in case["d"]], np.float32)
        t = orc.oracle_pt_sphere_intersect(vp(sph), vp(o), vp(d))
        assert np.float32(t).view(np.uint32) == int(case["t"], 16)


def test_oracle_equals_compiled_reference_live(orc, rt, ref_smallpt):
    """Same comparison against oracle/_ref run now (build container only), at another size and pass count."""
    g = load_smallpt_golden(rt, "cornell")
    w, h, passes = 33, 21, 3
    ref_smallpt.ref_pt_set_scene(vp(g["spheres"]), g["spheres"].size, vp(g["camera"]), w, h)
    cam = np.zeros(1, rt.CAMERA_DTYPE)
    ref_smallpt.ref_pt_get_scene(None, vp(cam))
    seeds = rt.reference_seeds(w, h, seed=7)
    col_r = np.zeros(3 * w * h, np.float32); pix_r = np.zeros(w * h, np.uint32); sd_r = np.zeros(2 * w * h, np.uint32)
    ref_smallpt.ref_pt_render(vp(seeds), passes, vp(col_r), vp(pix_r), vp(sd_r))
    g2 = dict(g, w=w, h=h, passes=passes, camera=cam, seeds_in=seeds)
    col, sd, pix, _ = run_oracle(orc, 0, g2)
    assert np.array_equal(col.view(np.uint32), col_r.view(np.uint32))
    assert np.array_equal(sd, sd_r) and np.array_equal(pix, pix_r)
    # UpdateCamera restatement == the reference's
    mine = g["camera"].copy(); mine["dir"] = 0; mine["x"] = 0; mine["y"] = 0
    rt.update_camera(mine, w, h)
    assert mine.tobytes() == cam.tobytes()
    orc_cam = g["camera"].copy()
    orc.oracle_pt_update_camera(vp(orc_cam), w, h)
    assert orc_cam.tobytes() == cam.tobytes()


def test_passes_can_be_split_and_threads_do_not_matter(orc, rt):
    """5 passes at once == 2 + 3 passes (progressive state is colors + seeds + sample index)."""
    g = load_smallpt_golden(rt, "cornell")
    a = run_oracle(orc, 0, g, passes=5, threads=1)
    col, sd, _, _ = run_oracle(orc, 0, g, passes=2, threads=3)
    b = run_oracle(orc, 0, g, passes=3, pass0=2, threads=7, colors=col, seeds=sd)
    assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


def test_work_per_sample_matches_survey(orc, rt):
    """cornell = 6.80 Intersect + 2.67 IntersectP per sample, 84.2 sphere tests."""
    g = load_smallpt_golden(rt, "cornell")
    _, _, _, ctr = run_oracle(orc, 0, g, passes=8)
    s = float(ctr[0])
    assert abs(ctr[1] / s - 6.80) < 0.15 and abs(ctr[2] / s - 2.67) < 0.15 and abs(ctr[3] / s - 84.2) < 2.0
