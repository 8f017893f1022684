import ctypes
import json
import os
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def vp(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


@pytest.fixture(scope="session")
def built():
    """Native artefacts; built on demand so a fresh checkout can run the CPU suite directly."""
    need = [os.path.join(graft.PKG_DIR, "librt_b200.so"), os.path.join(graft.ORACLE_DIR, "liboracle.so"),
            os.path.join(graft.DEVSIM_DIR, "libdevsim.so")]
    if not all(os.path.exists(p) for p in need):
        graft.build()
    return True


@pytest.fixture(scope="session")
def rt(built):
    return graft.load()


@pytest.fixture(scope="session")
def orc(built):
    o = graft.oracle()
    o.oracle_pt_get_random.restype = ctypes.c_float
    o.oracle_pt_sphere_intersect.restype = ctypes.c_float
    o.oracle_libm_pow20.restype = ctypes.c_double
    o.oracle_libm_pow20.argtypes = [ctypes.c_float]
    return o


@pytest.fixture(scope="session")
def devsim(built):
    d = ctypes.CDLL(os.path.join(graft.DEVSIM_DIR, "libdevsim.so"))
    d.devsim_pow20.restype = ctypes.c_double
    d.devsim_pow20.argtypes = [ctypes.c_float]
    d.devsim_get_random.restype = ctypes.c_float
    d.devsim_cover.restype = ctypes.c_uint32
    return d


def _ref(name):
    p = os.path.join(graft.ORACLE_DIR, "_ref", name)
    if not os.path.exists(p):
        pytest.skip(f"{p} not built (needs /root/reference, build container only)")
    return ctypes.CDLL(p)


@pytest.fixture(scope="session")
def ref_whitted(built):
    return _ref("libref_whitted.so")


@pytest.fixture(scope="session")
def ref_smallpt(built):
    return _ref("libref_smallpt.so")


@pytest.fixture(scope="session")
def whitted_golden():
    g = json.load(open(os.path.join(GOLDEN, "whitted_golden.json")))
    raw = zlib.decompress(open(os.path.join(GOLDEN, "whitted_test_bmp.rgb.zlib"), "rb").read())
    g["rgb"] = np.frombuffer(raw, np.uint8).reshape(g["height"], g["width"], 3)
    return g


def load_smallpt_golden(rt, scene):
    z = np.load(os.path.join(GOLDEN, f"smallpt_{scene}.npz"))
    g = {k: z[k] for k in z.files}
    g["spheres"] = g["spheres"].view(np.uint8).view(rt.SPHERE_DTYPE).copy()
    g["camera"] = g["camera"].view(np.uint8).view(rt.CAMERA_DTYPE).copy()
    g["w"], g["h"], g["passes"] = int(g["w"]), int(g["h"]), int(g["passes"])
    return g


@pytest.fixture(scope="session")
def cornell(rt):
    """Cornell box: scene + camera of SPT/scenes/cornell.scn, taken from the golden fixture."""
    g = load_smallpt_golden(rt, "cornell")
    return g["spheres"], g["camera"]


@pytest.fixture(scope="session")
def gpu(rt):
    r = rt.Renderer(0)     # raises without a CUDA device: gpu tests must never pass on a fallback
    yield r
    r.close()


def f32(hexstr):
    return np.array([int(hexstr, 16)], np.uint32).view(np.float32)[0]
