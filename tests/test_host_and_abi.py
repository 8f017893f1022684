"""Host-side logic and the C-ABI surface.  CPU only: no compute entry point is called here."""
import ctypes
import hashlib
import json
import os
import re

import numpy as np

from conftest import vp, GOLDEN, ROOT, load_smallpt_golden


def test_library_exports_every_declared_symbol(rt):
    """Every function include/rt_b200.h declares resolves in librt_b200.so, and the binding lists all of them."""
    hdr = open(os.path.join(ROOT, "include", "rt_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(rt.SYMBOLS), declared ^ set(rt.SYMBOLS)
    handle = ctypes.CDLL(rt.LIB_PATH)
    for name in declared:
        assert getattr(handle, name) is not None


def test_no_device_is_a_loud_error_not_a_fallback(rt):
    """Without a GPU rt_init must fail (RT_ERR_NO_DEVICE); with one it must succeed.  Never a CPU path."""
    import torch
    if torch.cuda.is_available():
        rt.Renderer(0).close()
        return
    try:
        rt.Renderer(0)
    except rt.RtError as e:
        assert e.code == rt.RT_ERR_NO_DEVICE and "no CPU fallback" in str(e)
    else:
        raise AssertionError("Renderer() succeeded without a CUDA device")


def test_product_library_does_not_link_the_oracle(rt):
    import subprocess
    out = subprocess.run(["ldd", rt.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "devsim" not in out and "libref" not in out
    syms = subprocess.run(["nm", "-D", rt.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle_" not in syms and "devsim_" not in syms


def test_complex_scene_generator_matches_perl_script(rt, tmp_path):
    """rt_write_complex_scene == header of SPT/scenes/complex.scn + `perl scene_build_complex.pl` (md5 fixture)."""
    md5 = json.load(open(os.path.join(GOLDEN, "complex_scene_md5.json")))
    for depth, want in md5.items():
        p = tmp_path / f"c{depth}.scn"
        rt.write_complex_scene(str(p), int(depth))
        lines = p.read_text().splitlines(keepends=True)
        assert lines[0] == "camera 20 80 150  0 15 0\n" and lines[1] == f"size {want['spheres'] + 2}\n"
        assert hashlib.md5("".join(lines[4:]).encode()).hexdigest() == want["md5"], depth
        spheres, cam = rt.read_scene(str(p), 64, 48)
        assert spheres.size == want["spheres"] + 2
    # depth 4 is the shipped complex.scn: same spheres as the fixture taken from the reference's file
    g = load_smallpt_golden(rt, "complex")
    p = tmp_path / "c4.scn"
    spheres, cam = rt.read_scene(str(p), g["w"], g["h"])
    assert spheres.tobytes() == g["spheres"].tobytes()
    assert cam.tobytes() == g["camera"].tobytes()


def test_read_scene_rejects_bad_files(rt, tmp_path):
    bad = tmp_path / "bad.scn"
    bad.write_text("camera 1 2 3  4 5 6\nsize 1\nsphere 1  0 0 0  0 0 0  1 1 1  7\n")
    for path in [str(bad), str(tmp_path / "missing.scn")]:
        try:
            rt.read_scene(path, 8, 8)
        except rt.RtError as e:
            assert e.code == rt.RT_ERR_IO
        else:
            raise AssertionError("bad scene accepted")


def test_ppm_writer_format(rt, tmp_path):
    """P3, bottom row first, 'r g b ' triplets (SPT/displayfunc.cpp:254-271)."""
    px = np.array([[0x030201, 0x060504], [0x090807, 0x0c0b0a]], np.uint32)
    p = tmp_path / "i.ppm"
    rt.write_ppm(str(p), px)
    assert p.read_text() == "P3\n2 2\n255\n7 8 9 10 11 12 1 2 3 4 5 6 "


def test_work_items_cover_each_owned_pixel_exactly_once(devsim):
    """Row-tile sharding: over all ranks every pixel is visited once; each rank only touches its tiles."""
    for (w, h, world, tile) in [(64, 48, 1, 8), (61, 37, 2, 8), (33, 50, 4, 4), (8, 5, 8, 8), (100, 99, 3, 5)]:
        total = np.zeros((h, w), np.int32)
        for rank in range(world):
            v = np.zeros((h, w), np.int32)
            devsim.devsim_cover(w, h, rank, world, tile, vp(v))
            rows = np.nonzero(v.any(axis=1))[0]
            assert all((y // tile) % world == rank for y in rows)
            total += v
        assert (total == 1).all(), (w, h, world, tile)


def test_builtin_cornell_equals_scn_fixture(rt):
    """rt.cornell_scene() (CornellSpheres[] of SPT/scene.h) == ReadScene(SPT/scenes/cornell.scn) from the fixture."""
    g = load_smallpt_golden(rt, "cornell")
    spheres, cam = rt.cornell_scene(g["w"], g["h"])
    assert spheres.tobytes() == g["spheres"].tobytes() and cam.tobytes() == g["camera"].tobytes()


def test_viewer_keys_equal_the_reference_callbacks(rt, cornell):
    """rt_viewer_key against a session scripted through the reference's own keyFunc / specialFunc (fixture made by
    tests/golden/make_golden.py viewer_keys from oracle/_ref): camera bits and sphere table after every key press."""
    g = json.load(open(os.path.join(GOLDEN, "viewer_keys.json")))
    spheres, cam = cornell
    cam = cam.copy()
    rt.update_camera(cam, g["w"], g["h"])
    v = rt.ViewerState(spheres, cam, g["w"], g["h"])
    for i, step in enumerate(g["steps"]):
        k = step["key"]
        action = v.key(rt.KEY_SPECIAL + int(k[1:]) if k.startswith("S") else k)
        want = {"+": rt.KEY_SCENE, "-": rt.KEY_SCENE, "h": rt.KEY_NONE}.get(k, rt.KEY_SCENE if k in "468293" else rt.KEY_CAMERA)
        assert action == want, (i, k, action)
        assert ["%08x" % x for x in v.cam.view(np.uint32).reshape(-1)] == step["camera"], (i, k)
        assert hashlib.sha256(v.spheres.view(np.float32).tobytes()).hexdigest() == step["spheres_sha256"], (i, k)
    assert v.key(" ") == rt.KEY_RESTART and v.key("p") == rt.KEY_DUMP and v.key(chr(27)) == rt.KEY_QUIT


def test_bench_reference_arm_prints_the_contract_line_without_the_product_library():
    """`bench.py --impl reference` (the driver's CPU arm): one JSON line with the contract's keys, produced without loading librt_b200.so
    (the arm may only execute oracle/), on the same metric / unit / workload string as the GPU arm prints."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, LD_DEBUG="files")                          # the dynamic loader names every library it maps on stderr
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "Mrays/s" and line["unit"] == "Mrays/s" and line["value"] > 0
    assert line["config"]["workload"].startswith("Raytracer3.2.03 Whitted scene") and line["config"]["workload"].endswith("1920x1080")
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "librt_b200.so" not in p.stderr and "liboracle.so" in p.stderr
