"""The Whitted oracle (oracle/oracle_whitted.c) against the reference's own artefacts.  CPU only."""
import hashlib

import numpy as np

from conftest import vp, f32


def render_oracle(orc, prims, w, h, threads=8):
    px = np.zeros((h, w, 4), np.uint8)
    hits = np.zeros((h, w, 9), np.int32)
    ctr = np.zeros(5, np.uint64)
    orc.oracle_whitted_render(vp(px), vp(hits), w, h, vp(prims), prims.size, threads, vp(ctr))
    return px, hits, ctr


def test_oracle_reproduces_reference_golden_image(orc, rt, whitted_golden):
    """800x600 render == the reference's shipped test.bmp, every byte (R323/test.bmp, md5 in the fixture)."""
    prims = rt.whitted_create_scene(0)
    px, hits, ctr = render_oracle(orc, prims, 800, 600)
    assert np.array_equal(px[:, :, :3], whitted_golden["rgb"])
    assert not px[:, :, 3].any()
    # known-answer hit IDs recorded from the compiled reference raytrace()
    centre = hits[:, :, 4]
    assert [int(np.count_nonzero(centre == k)) for k in range(-1, 17)] == whitted_golden["centre_hit_histogram_ids_-1_to_16"]
    assert hashlib.sha256(hits.tobytes()).hexdigest() == whitted_golden["all9_hit_id_sha256"]
    for k in whitted_golden["centre_primary_rays"]:
        assert centre[k["y"], k["x"]] == k["hit"]
        assert list(px[k["y"], k["x"], :3]) == k["bmp_rgb"]
    # work per pixel quoted in SURVEY.md 6 / BASELINE.md
    n = 800 * 600
    assert abs(ctr[0] / n - 15.30) < 0.01 and abs(ctr[1] / n - 45.82) < 0.01
    assert abs(ctr[2] / n - 423.3) < 0.2 and abs(ctr[3] / n - 363.7) < 0.2


def test_bmp_writer_reproduces_reference_file(orc, rt, whitted_golden, tmp_path):
    """rt_write_bmp(oracle frame) has the md5 of the reference's test.bmp."""
    prims = rt.whitted_create_scene(0)
    px, _, _ = render_oracle(orc, prims, 800, 600)
    path = tmp_path / "t.bmp"
    rt.write_bmp(str(path), px)
    assert hashlib.md5(path.read_bytes()).hexdigest() == whitted_golden["bmp_md5"]


def test_oracle_equals_compiled_reference(orc, rt, ref_whitted):
    """oracle == oracle/_ref (raytracer_non_OpenCL.c compiled unmodified) on other sizes, incl. odd ones."""
    prims = rt.whitted_create_scene(0)
    for (w, h) in [(160, 120), (203, 77), (64, 64), (1, 1), (7, 3)]:
        px, hits, _ = render_oracle(orc, prims, w, h, threads=3)
        ref_px = np.zeros((h, w, 4), np.uint8)
        ref_whitted.ref_whitted_render(vp(ref_px), w, h, vp(prims), prims.size)
        ref_hits = np.zeros((h, w, 9), np.int32)
        ref_whitted.ref_whitted_primary_hits(vp(ref_hits), None, None, w, h, vp(prims), prims.size)
        assert np.array_equal(px, ref_px), (w, h)
        assert np.array_equal(hits, ref_hits), (w, h)


def test_scene_builder_matches_reference_create_scene(rt, ref_whitted):
    """rt_whitted_create_scene(0) == the reference's create_scene + Primitive_2 copy, on every field a
    primitive type uses (the reference leaves the others as stack garbage)."""
    mine = rt.whitted_create_scene(0)
    ref = np.zeros(64, rt.PRIMITIVE_DTYPE)
    n = ref_whitted.ref_whitted_scene(vp(ref), 64)
    assert n == mine.size == 17
    ref = ref[:n]
    for f in ["m_refl", "m_diff", "m_refr", "m_refr_index", "m_spec", "type", "is_light"]:
        assert np.array_equal(mine[f][:16], ref[f][:16]), f
    assert np.array_equal(mine["m_color"][:16, :3], ref["m_color"][:16, :3])
    sph = mine["type"] == 1
    for f in ["radius", "sq_radius", "r_radius"]:
        assert np.array_equal(mine[f][sph], ref[f][sph]), f
    assert np.array_equal(mine["center"][sph][:, :3], ref["center"][sph][:, :3])
    pl = (mine["type"] == 0) & (np.arange(17) < 16)
    assert np.array_equal(mine["normal"][pl][:, :3], ref["normal"][pl][:, :3])
    assert np.array_equal(mine["depth"][pl], ref["depth"][pl])
    assert not mine[16].tobytes().strip(b"\0")          # the 17th slot is all zero


def test_threads_do_not_change_results(orc, rt):
    prims = rt.whitted_create_scene(0)
    a = render_oracle(orc, prims, 97, 61, threads=1)
    b = render_oracle(orc, prims, 97, 61, threads=5)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2][:4], b[2][:4])


def test_miss_is_defined(orc, rt):
    """An open scene (floor + one light): rays that leave it contribute nothing and spawn nothing."""
    prims = rt.whitted_create_scene(0)[[0, 13]].copy()
    px, hits, _ = render_oracle(orc, prims, 40, 30, threads=1)
    assert (hits == -1).any() and (hits >= 0).any()
    assert not px[hits[:, :, :].max(axis=2) == -1][:, :3].any()
