"""ctypes binding of librt_b200.so -- the reference-side stub a Python caller would use.

The product is the shared library (CUDA kernels for sm_100a behind the C ABI of
``include/rt_b200.h``); this module only declares its entry points for ``ctypes`` and wraps them in
a thin class whose method names and argument meaning mirror the reference's host entry points
(``UpdateRenderingGPU`` -> :meth:`Renderer.pt_render`, ``raytracer_non_kernel`` ->
:meth:`Renderer.whitted_render`, ``ReadScene``/``UpdateCamera``/``create_scene`` -> module
functions).  There is no Python or CPU implementation of the rendering path here: if the library
or a CUDA device is missing, loading / ``Renderer()`` raises.

The directory name contains hyphens (it is the reference's repository name), so import it with
``importlib`` -- see ``load()`` in ``__graft_entry__.py``.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# RT_B200_LIB lets tools/variants.sh point the binding at an A/B build; the default is the in-tree product.
LIB_PATH = os.environ.get("RT_B200_LIB") or os.path.join(_HERE, "librt_b200.so")

# ----------------------------------------------------------------------------- POD layouts
SPHERE_DTYPE = np.dtype([("rad", "<f4"), ("p", "<f4", 3), ("e", "<f4", 3), ("c", "<f4", 3), ("refl", "<i4")])
CAMERA_DTYPE = np.dtype([("orig", "<f4", 3), ("target", "<f4", 3), ("dir", "<f4", 3), ("x", "<f4", 3), ("y", "<f4", 3)])
PRIMITIVE_DTYPE = np.dtype([
    ("m_color", "<f4", 4), ("m_refl", "<f4"), ("m_diff", "<f4"), ("m_refr", "<f4"), ("m_refr_index", "<f4"),
    ("m_spec", "<f4"), ("dummy_3", "<f4"), ("type", "<i4"), ("is_light", "u1"), ("pad_", "u1", 3),
    ("normal", "<f4", 4), ("center", "<f4", 4), ("depth", "<f4"), ("radius", "<f4"), ("sq_radius", "<f4"),
    ("r_radius", "<f4")])
R306_PRIMITIVE_DTYPE = np.dtype([                    # R306/raytracer.h:19-34 Primitive (+ Material)
    ("type", "<i4"), ("m_light", "<i4"), ("centre", "<f4", 3), ("sq_radius", "<f4"), ("radius", "<f4"), ("r_radius", "<f4"),
    ("plane_n", "<f4", 3), ("plane_d", "<f4"), ("plane_cell", "<f4", 4),
    ("m_color", "<f4", 3), ("m_refl", "<f4"), ("m_refr", "<f4"), ("m_diff", "<f4"), ("m_spec", "<f4"), ("m_rindex", "<f4")])
assert SPHERE_DTYPE.itemsize == 44 and CAMERA_DTYPE.itemsize == 60 and PRIMITIVE_DTYPE.itemsize == 96
assert R306_PRIMITIVE_DTYPE.itemsize == 96


class Counters(C.Structure):
    _fields_ = [("nearest_queries", C.c_uint64), ("shadow_queries", C.c_uint64), ("sphere_tests", C.c_uint64),
                ("plane_tests", C.c_uint64), ("samples", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


RT_DIFF_, RT_SPEC_, RT_REFR_ = 0, 1, 2
RT_OK, RT_ERR_NO_DEVICE, RT_ERR_CUDA, RT_ERR_ARG, RT_ERR_STATE, RT_ERR_CAPACITY, RT_ERR_IO = 0, -1, -2, -3, -4, -5, -6
TUNE_PT_MAX_RESIDENT_BYTES, TUNE_PT_CHUNK_SPHERES, TUNE_MAX_BLOCKS_PER_SM, TUNE_WHITTED_COST_ORDER, TUNE_PT_ALIGNED, TUNE_PT_BVH, TUNE_WHITTED_BVH, TUNE_R306_SPLIT, TUNE_PT_SINCOS_TABLE, TUNE_WHITTED_BLOCKS, TUNE_WHITTED_FILLER_PCT, TUNE_WHITTED_STAGE_CAP, TUNE_WHITTED_REDO_CAP, TUNE_WHITTED_GRID, TUNE_WHITTED_SPLIT, TUNE_WHITTED_SPLIT_BLOCKS = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15
BUF_WHITTED_PIXELS, BUF_WHITTED_HITS, BUF_PT_PIXELS, BUF_PT_COLORS, BUF_PT_SEEDS = 0, 1, 2, 3, 4
IPC_HANDLE_BYTES = 80

# Every symbol include/rt_b200.h declares: name -> (restype, argtypes).
_VP, _I, _U32, _U64 = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64
SYMBOLS = {
    "rt_init": (_I, [C.POINTER(_VP), _I]),
    "rt_destroy": (None, [_VP]),
    "rt_last_error": (C.c_char_p, [_VP]),
    "rt_device_info": (_I, [_VP, C.POINTER(_I), C.POINTER(_I), C.c_char_p, _I]),
    "rt_set_shard": (_I, [_VP, _I, _I, _I]),
    "rt_set_counting": (_I, [_VP, _I]),
    "rt_get_counters": (_I, [_VP, C.POINTER(Counters)]),
    "rt_get_counters_ex": (_I, [_VP, _VP]),
    "rt_set_tuning": (_I, [_VP, _I, _I]),
    "rt_whitted_render": (_I, [_VP, _VP, _I, _I, _I, _VP, _VP]),
    "rt_whitted_upload": (_I, [_VP, _VP, _I, _I, _I, _I]),
    "rt_whitted_launch": (_I, [_VP]),
    "rt_whitted_download": (_I, [_VP, _VP, _VP]),
    "rt_r306_render": (_I, [_VP, _VP, _I, _I, _I, _VP]),
    "rt_r306_upload": (_I, [_VP, _VP, _I, _I, _I]),
    "rt_r306_launch": (_I, [_VP]),
    "rt_r306_download": (_I, [_VP, _VP]),
    "rt_r306_create_scene": (_I, [_VP, _I]),
    "rt_pt_resize": (_I, [_VP, _I, _I, _VP]),
    "rt_pt_set_scene": (_I, [_VP, _VP, _U32]),
    "rt_pt_set_camera": (_I, [_VP, _VP]),
    "rt_pt_render": (_I, [_VP, _I, _I, _VP, _VP, _VP]),
    "rt_pt_launch": (_I, [_VP, _I, _I]),
    "rt_pt_download": (_I, [_VP, _VP, _VP, _VP]),
    "rt_pt_current_sample": (_I, [_VP]),
    "rt_pt_restore": (_I, [_VP, _VP, _VP, _I]),
    "rt_pt_set_accumulate_sums": (_I, [_VP, _I]),
    "rt_pt_resolve_sums": (_I, [_VP, _I]),
    "rt_sync": (_I, [_VP]),
    "rt_timer_begin": (_I, [_VP]),
    "rt_timer_end": (_I, [_VP, C.POINTER(C.c_float)]),
    "rt_launch_count": (_U64, [_VP]),
    "rt_whitted_redo_reports": (_I, [_VP, C.POINTER(_U32)]),
    "rt_device_buffer": (_VP, [_VP, _I, C.POINTER(_U64)]),
    "rt_selftest_math": (_I, [_VP, _I, _VP, _VP, _U64]),
    "rt_ipc_export": (_I, [_VP, _I, _VP]),
    "rt_ipc_import": (_I, [_VP, _I, _VP]),
    "rt_ipc_close": (_I, [_VP]),
    "rt_whitted_download_rows": (_I, [_VP, _VP]),
    "rt_pt_download_rows": (_I, [_VP, _VP]),
    "rt_host_register": (_I, [_VP, _VP, _U64]),
    "rt_host_unregister": (_I, [_VP, _VP]),
    "rt_debug_check_flags": (C.c_longlong, [_VP]),
    "rt_stream": (_VP, [_VP]),
    "rt_set_stream": (_I, [_VP, _VP]),
    "rt_update_camera": (None, [_VP, _I, _I]),
    "rt_viewer_key": (_I, [_I, _VP, _I, _I, _VP, _U32, C.POINTER(_U32)]),
    "rt_read_scene": (_I, [C.c_char_p, _VP, C.POINTER(_VP), C.POINTER(_U32)]),
    "rt_write_complex_scene": (_I, [C.c_char_p, _I]),
    "rt_whitted_create_scene": (_I, [_I, _VP, _I]),
    "rt_whitted_from_spheres": (_I, [_VP, _U32, _VP, _VP, _I]),
    "rt_write_bmp": (_I, [C.c_char_p, _VP, _I, _I]),
    "rt_write_ppm": (_I, [C.c_char_p, _VP, _I, _I]),
    "rt_free": (None, [_VP]),
}

_lib = None


def lib():
    """Loads librt_b200.so (once).  Raises if it has not been built -- there is no other path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(or `make -C se-195-project-ray-tracer_b200`). The renderer has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


class RtError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"rt_b200 error {code}: {message}")
        self.code = code


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ----------------------------------------------------------------------------- host-side scene helpers
def update_camera(cam, w, h):
    """UpdateCamera (SPT/displayfunc.cpp:182-195) on a CAMERA_DTYPE record array of length 1."""
    assert cam.dtype == CAMERA_DTYPE and cam.size == 1
    lib().rt_update_camera(_ptr(cam), w, h)
    return cam


KEY_NONE, KEY_CAMERA, KEY_SCENE, KEY_RESTART, KEY_DUMP, KEY_QUIT = 0, 1, 2, 3, 4, 5
KEY_SPECIAL = 0x100
KEY_UP, KEY_DOWN, KEY_LEFT, KEY_RIGHT, KEY_PAGE_UP, KEY_PAGE_DOWN = (KEY_SPECIAL + c for c in (101, 103, 100, 102, 104, 105))


class ViewerState:
    """What the reference viewer keeps between key presses (SPT/displayfunc.cpp): camera, sphere table, selection."""

    def __init__(self, spheres, cam, w, h):
        self.spheres, self.cam, self.w, self.h = spheres.copy(), cam.copy(), w, h
        self.current = C.c_uint32(0)

    def key(self, key):
        """keyFunc / specialFunc for one key (a character or one of the KEY_* specials); returns the KEY_* action."""
        code = ord(key) if isinstance(key, str) else int(key)
        rc = lib().rt_viewer_key(code, _ptr(self.cam), self.w, self.h, _ptr(self.spheres), self.spheres.size, C.byref(self.current))
        if rc < 0:
            raise RtError(rc, f"rt_viewer_key({key!r})")
        return rc


def read_scene(path, w, h):
    """ReadScene + UpdateCamera (SPT/displayfunc.cpp:120-195): returns (spheres, camera)."""
    cam = np.zeros(1, CAMERA_DTYPE)
    sp, n = C.c_void_p(), C.c_uint32()
    rc = lib().rt_read_scene(os.fsencode(path), _ptr(cam), C.byref(sp), C.byref(n))
    if rc != RT_OK:
        raise RtError(rc, f"cannot read scene {path}")
    try:
        spheres = np.frombuffer(C.string_at(sp.value, n.value * 44), dtype=SPHERE_DTYPE).copy()
    finally:
        lib().rt_free(sp)
    update_camera(cam, w, h)
    return spheres, cam


def write_complex_scene(path, max_depth):
    rc = lib().rt_write_complex_scene(os.fsencode(path), max_depth)
    if rc != RT_OK:
        raise RtError(rc, f"cannot write {path}")


def whitted_create_scene(which=0):
    """create_scene of R323/scene.c:48-128 as flat Primitive_2 records."""
    out = np.zeros(64, PRIMITIVE_DTYPE)
    n = lib().rt_whitted_create_scene(which, _ptr(out), out.size)
    if n < 0:
        raise RtError(n, "rt_whitted_create_scene")
    return out[:n].copy()


def whitted_from_spheres(spheres, cam=None):
    """A smallpt sphere table as Primitive_2 records for the Whitted tracer (rt_whitted_from_spheres); with a camera
    the scene is moved into the Whitted tracer's fixed eye frame."""
    spheres = np.ascontiguousarray(spheres)
    assert spheres.dtype == SPHERE_DTYPE
    out = np.zeros(spheres.size, PRIMITIVE_DTYPE)
    n = lib().rt_whitted_from_spheres(_ptr(spheres), spheres.size, _ptr(cam), _ptr(out), out.size)
    if n < 0:
        raise RtError(n, "rt_whitted_from_spheres")
    return out[:n]


def r306_create_scene():
    """Scene_InitScene of R306/scene.cpp:217-272 as flat Primitive records."""
    out = np.zeros(32, R306_PRIMITIVE_DTYPE)
    n = lib().rt_r306_create_scene(_ptr(out), out.size)
    if n < 0:
        raise RtError(n, "rt_r306_create_scene")
    return out[:n].copy()


def write_bmp(path, pixels):
    h, w = pixels.shape[:2]
    rc = lib().rt_write_bmp(os.fsencode(path), _ptr(np.ascontiguousarray(pixels)), w, h)
    if rc != RT_OK:
        raise RtError(rc, f"cannot write {path}")


def write_ppm(path, pixels_u32):
    h, w = pixels_u32.shape
    rc = lib().rt_write_ppm(os.fsencode(path), _ptr(np.ascontiguousarray(pixels_u32)), w, h)
    if rc != RT_OK:
        raise RtError(rc, f"cannot write {path}")


def cornell_scene(w, h):
    """The reference's built-in scene: CornellSpheres[] (SPT/scene.h:32-42) and the default camera of
    mainGPU (SPT/smallptGPU.cpp:856-857), in float arithmetic like the C initialisers.  Bit-identical
    to what ReadScene makes of SPT/scenes/cornell.scn (checked by tests/test_host_and_abi.py)."""
    f = np.float32
    W = f(1e4)
    rows = [(W, (W + f(1), 40.8, 81.6), (0, 0, 0), (.75, .25, .25), RT_DIFF_), (W, (-W + f(99), 40.8, 81.6), (0, 0, 0), (.25, .25, .75), RT_DIFF_),
            (W, (50, 40.8, W), (0, 0, 0), (.75, .75, .75), RT_DIFF_), (W, (50, 40.8, -W + f(270)), (0, 0, 0), (0, 0, 0), RT_DIFF_),
            (W, (50, W, 81.6), (0, 0, 0), (.75, .75, .75), RT_DIFF_), (W, (50, -W + f(81.6), 81.6), (0, 0, 0), (.75, .75, .75), RT_DIFF_),
            (16.5, (27, 16.5, 47), (0, 0, 0), (.9, .9, .9), RT_SPEC_), (16.5, (73, 16.5, 78), (0, 0, 0), (.9, .9, .9), RT_REFR_),
            (7, (50, f(81.6) - f(15), 81.6), (12, 12, 12), (0, 0, 0), RT_DIFF_)]
    spheres = np.zeros(len(rows), SPHERE_DTYPE)
    for i, row in enumerate(rows):
        spheres[i] = row
    cam = np.zeros(1, CAMERA_DTYPE)
    cam["orig"] = (50, 45, 205.6)
    cam["target"] = (50, f(45) - f(0.042612), 204.6)
    update_camera(cam, w, h)
    return spheres, cam


def reference_seeds(w, h, seed=1):
    """The reference's seed rule (SPT/smallptGPU.cpp:105-110): 2*w*h draws, each clamped to >= 2.
    The reference draws from libc rand(); the C ABI takes seeds as an INPUT, so any generator will do --
    numpy's is used here so that fixtures do not depend on a libc."""
    s = np.random.RandomState(seed).randint(0, 2 ** 31 - 1, size=2 * w * h, dtype=np.int64).astype(np.uint32)
    return np.maximum(s, 2).astype(np.uint32)


# ----------------------------------------------------------------------------- multi-GPU plumbing
# One process per GPU (torchrun); torch.distributed carries the only two exchanges the path has
# (SURVEY.md 8e): the gather of the interleaved row tiles to rank 0, and the sum of the per-rank
# accumulation buffers in the sample-sharded mode.  Works on CUDA tensors over NCCL/NVLink and on CPU
# tensors over gloo (the world_size-2 tests).
class DeviceArray:
    """Wraps a device pointer handed out by rt_device_buffer for torch.as_tensor(..., device="cuda")."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def pick_tile_rows(h, world):
    """Largest tile height of 8, 4, 2, 1 rows that deals the frame's tiles evenly over the ranks (else 8)."""
    for t in (8, 4, 2, 1):
        if h % t == 0 and (h // t) % world == 0:
            return t
    return 8


def owned_rows(h, rank, world, tile_rows):
    """Row indices rendered by `rank` under rt_set_shard(rank, world, tile_rows)."""
    return [y for y in range(h) if (y // tile_rows) % world == rank]


def _tile_views(frame, world, tile_rows):
    """Per-rank (strided view of the full tiles, view of the ragged last tile or None)."""
    h = frame.shape[0]
    n_full = h // tile_rows
    body = frame[:n_full * tile_rows].reshape(n_full, -1)           # [tile, tile_rows * row_elems], a view
    tail_owner = n_full % world if h % tile_rows else -1
    return [(body[q::world], frame[n_full * tile_rows:].reshape(-1) if q == tail_owner else None) for q in range(world)]


def gather_row_tiles(frame, rank, world, tile_rows, staging=None):
    """frame: contiguous [h, ...] tensor whose rows owned by this rank are valid.  After the call rank 0's
    frame is complete.  Each rank sends only its own rows, as one message ((world-1)/world of the frame
    crosses the links); rank 0 de-interleaves with strided copies."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return frame
    if frame.dtype == torch.uint32:                                  # torch has few uint32 kernels
        frame = frame.view(torch.int32)
    views = _tile_views(frame, world, tile_rows)
    sizes = [v.numel() + (t.numel() if t is not None else 0) for v, t in views]
    if rank == 0:
        if staging is None:
            staging = [torch.empty(sizes[q], dtype=frame.dtype, device=frame.device) for q in range(world)]
        for r_ in dist.batch_isend_irecv([dist.P2POp(dist.irecv, staging[q], q) for q in range(1, world)]):
            r_.wait()
        for q in range(1, world):
            v, t = views[q]
            v.copy_(staging[q][:v.numel()].view(v.shape))
            if t is not None:
                t.copy_(staging[q][v.numel():])
    else:
        v, t = views[rank]
        mine = v.reshape(-1) if t is None else torch.cat([v.reshape(-1), t])
        for r_ in dist.batch_isend_irecv([dist.P2POp(dist.isend, mine.contiguous(), 0)]):
            r_.wait()
    return frame


def gather_staging(frame, world, tile_rows):
    """Pre-allocated receive buffers for gather_row_tiles on rank 0 (so timed steps do not allocate)."""
    import torch
    if frame.dtype == torch.uint32:
        frame = frame.view(torch.int32)
    views = _tile_views(frame, world, tile_rows)
    return [torch.empty(v.numel() + (t.numel() if t is not None else 0), dtype=frame.dtype, device=frame.device) for v, t in views]


def share_rank0_framebuffer(renderer, which, rank, world):
    """Fused frame assembly: rank 0 exports its pixel buffer (CUDA IPC), the others map it and will render their
    rows straight into it.  Collective over the default process group; returns True if EVERY rank succeeded (else
    nothing is redirected and the caller falls back to gather_row_tiles)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return False
    dev = torch.device("cuda", torch.cuda.current_device())
    handle = torch.zeros(IPC_HANDLE_BYTES, dtype=torch.uint8, device=dev)
    ok = torch.ones(1, dtype=torch.int32, device=dev)
    if rank == 0:
        try:
            handle.copy_(torch.from_numpy(renderer.ipc_export(which)))
        except RtError:
            ok.zero_()
    dist.broadcast(handle, 0)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) and rank != 0:
        try:
            renderer.ipc_import(which, handle.cpu().numpy())
        except RtError:
            ok.zero_()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if not int(ok.item()):
        renderer.ipc_close()
        return False
    return True


def allreduce_sums(colors):
    """Sample-sharded mode: adds the per-rank float accumulation buffers in place (ncclAllReduce, sum)."""
    import torch.distributed as dist
    dist.all_reduce(colors, op=dist.ReduceOp.SUM)
    return colors


# ----------------------------------------------------------------------------- the renderer
class Renderer:
    """One context = one GPU.  Method names follow the reference's host entry points."""

    def __init__(self, device=0):
        self._lib = lib()
        self._ctx = C.c_void_p()
        rc = self._lib.rt_init(C.byref(self._ctx), device)
        if rc != RT_OK:
            msg = self._lib.rt_last_error(None).decode()
            self._ctx = None
            raise RtError(rc, msg)
        self.pt_size = None
        self.whitted_size = None

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.rt_destroy(self._ctx)
            self._ctx = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != RT_OK:
            raise RtError(rc, self._lib.rt_last_error(self._ctx).decode())

    # -- context
    def device_info(self):
        sm, khz = C.c_int(), C.c_int()
        name = C.create_string_buffer(256)
        self._ck(self._lib.rt_device_info(self._ctx, C.byref(sm), C.byref(khz), name, 256))
        return {"sm_count": sm.value, "sm_clock_khz": khz.value, "name": name.value.decode()}

    def set_shard(self, rank, world, tile_rows=8):
        self._ck(self._lib.rt_set_shard(self._ctx, rank, world, tile_rows))

    def set_counting(self, enabled):
        self._ck(self._lib.rt_set_counting(self._ctx, int(bool(enabled))))

    def counters(self):
        c = Counters()
        self._ck(self._lib.rt_get_counters(self._ctx, C.byref(c)))
        return c.as_dict()

    def counters_ex(self):
        """Raw counter block (8 x u64) of the last counting launch; [5] = shadow rays a timed Whitted launch traces."""
        out = np.zeros(8, np.uint64)
        self._ck(self._lib.rt_get_counters_ex(self._ctx, _ptr(out)))
        return out

    def set_tuning(self, key, value):
        self._ck(self._lib.rt_set_tuning(self._ctx, key, value))

    def sync(self):
        self._ck(self._lib.rt_sync(self._ctx))

    def timer_begin(self):
        self._ck(self._lib.rt_timer_begin(self._ctx))

    def timer_end(self):
        ms = C.c_float()
        self._ck(self._lib.rt_timer_end(self._ctx, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(self._lib.rt_launch_count(self._ctx))

    def whitted_redo_reports(self):
        """Shadow batches the last timed Whitted launch handed to the exact launch (see rt_whitted_redo_reports)."""
        n = C.c_uint32(0)
        self._ck(self._lib.rt_whitted_redo_reports(self._ctx, C.byref(n)))
        return int(n.value)

    def set_stream(self, cuda_stream_handle):
        """Issue all work on the given cudaStream_t (an int, e.g. torch.cuda.current_stream().cuda_stream)."""
        self._ck(self._lib.rt_set_stream(self._ctx, C.c_void_p(cuda_stream_handle)))

    def selftest_math(self, op, x):
        """Device-side evaluation of the kernels' elementary functions (see rt_selftest_math)."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = {0: np.zeros((x.size, 2), np.float32), 1: np.zeros(x.size, np.float32), 2: np.zeros(x.size, np.int32),
               3: np.zeros((x.size, 2), np.float32), 4: np.zeros(x.size, np.float64)}[op]
        self._ck(self._lib.rt_selftest_math(self._ctx, op, _ptr(x), _ptr(out), x.size))
        return out

    def ipc_export(self, which):
        h = np.zeros(IPC_HANDLE_BYTES, np.uint8)
        self._ck(self._lib.rt_ipc_export(self._ctx, which, _ptr(h)))
        return h

    def ipc_import(self, which, handle):
        h = np.ascontiguousarray(handle, dtype=np.uint8)
        assert h.size == IPC_HANDLE_BYTES
        self._ck(self._lib.rt_ipc_import(self._ctx, which, _ptr(h)))

    def ipc_close(self):
        self._ck(self._lib.rt_ipc_close(self._ctx))

    def debug_check_flags(self):
        """Mask of failed device-side bounds checks since the last call (-DRT_DEVICE_CHECKS builds); -1 if not compiled in."""
        return int(self._lib.rt_debug_check_flags(self._ctx))

    def device_buffer(self, which):
        n = C.c_uint64()
        p = self._lib.rt_device_buffer(self._ctx, which, C.byref(n))
        return p, n.value

    # -- Whitted (raytracer_non_kernel(pixels, width, height, primitives, n_primitives))
    def whitted_render(self, prims, w, h, want_hit_ids=False, pixels_out=None, hits_out=None):
        prims = np.ascontiguousarray(prims)
        assert prims.dtype == PRIMITIVE_DTYPE
        pixels = pixels_out if pixels_out is not None else np.zeros((h, w, 4), np.uint8)
        hits = None
        if want_hit_ids:
            hits = hits_out if hits_out is not None else np.zeros((h, w, 9), np.int32)
        self._ck(self._lib.rt_whitted_render(self._ctx, _ptr(prims), prims.size, w, h, _ptr(pixels), _ptr(hits)))
        self.whitted_size = (w, h)
        return (pixels, hits) if want_hit_ids else pixels

    def r306_render(self, prims, w, h, dest=None):
        """Engine_Render of raytracer3.0.06: w*h Pixels 0x00RRGGBB; rows outside 20 .. h-71 keep what `dest` held."""
        prims = np.ascontiguousarray(prims)
        assert prims.dtype == R306_PRIMITIVE_DTYPE
        dest = dest if dest is not None else np.zeros((h, w), np.uint32)
        self._ck(self._lib.rt_r306_render(self._ctx, _ptr(prims), prims.size, w, h, _ptr(dest)))
        return dest

    def r306_upload(self, prims, w, h):
        prims = np.ascontiguousarray(prims)
        assert prims.dtype == R306_PRIMITIVE_DTYPE
        self._ck(self._lib.rt_r306_upload(self._ctx, _ptr(prims), prims.size, w, h))

    def r306_launch(self):
        self._ck(self._lib.rt_r306_launch(self._ctx))

    def whitted_upload(self, prims, w, h, want_hit_ids=False):
        prims = np.ascontiguousarray(prims)
        assert prims.dtype == PRIMITIVE_DTYPE
        self._ck(self._lib.rt_whitted_upload(self._ctx, _ptr(prims), prims.size, w, h, int(want_hit_ids)))
        self.whitted_size = (w, h)

    def whitted_launch(self):
        self._ck(self._lib.rt_whitted_launch(self._ctx))

    def whitted_download(self, want_hit_ids=False, pixels_out=None):
        w, h = self.whitted_size
        pixels = pixels_out if pixels_out is not None else np.zeros((h, w, 4), np.uint8)
        hits = np.zeros((h, w, 9), np.int32) if want_hit_ids else None
        self._ck(self._lib.rt_whitted_download(self._ctx, _ptr(pixels), _ptr(hits)))
        return (pixels, hits) if want_hit_ids else pixels

    # -- multi-GPU read-back into one host frame shared by all ranks
    def whitted_download_rows(self, frame):
        """Copies the rows this rank owns into `frame` (a contiguous (h, w, 4) uint8 array, e.g. over shared memory)."""
        assert frame.flags["C_CONTIGUOUS"] and frame.nbytes == self.whitted_size[0] * self.whitted_size[1] * 4
        self._ck(self._lib.rt_whitted_download_rows(self._ctx, _ptr(frame)))

    def pt_download_rows(self, frame):
        assert frame.flags["C_CONTIGUOUS"] and frame.nbytes == self.pt_size[0] * self.pt_size[1] * 4
        self._ck(self._lib.rt_pt_download_rows(self._ctx, _ptr(frame)))

    def host_register(self, array):
        self._ck(self._lib.rt_host_register(self._ctx, _ptr(array), array.nbytes))

    def host_unregister(self, array):
        self._ck(self._lib.rt_host_unregister(self._ctx, _ptr(array)))

    # -- smallpt (AllocateBuffers / ReInitSceneGPU / ReInitGPU / UpdateRenderingGPU)
    def pt_resize(self, w, h, seeds):
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        assert seeds.size == 2 * w * h
        self._ck(self._lib.rt_pt_resize(self._ctx, w, h, _ptr(seeds)))
        self.pt_size = (w, h)

    def pt_set_scene(self, spheres):
        spheres = np.ascontiguousarray(spheres)
        assert spheres.dtype == SPHERE_DTYPE
        self._ck(self._lib.rt_pt_set_scene(self._ctx, _ptr(spheres), spheres.size))

    def pt_set_camera(self, cam):
        assert cam.dtype == CAMERA_DTYPE and cam.size == 1
        self._ck(self._lib.rt_pt_set_camera(self._ctx, _ptr(cam)))

    def pt_render(self, integrator, n_passes, want=("pixels", "colors", "seeds"), pixels_out=None):
        w, h = self.pt_size
        pixels = (pixels_out if pixels_out is not None else np.zeros((h, w), np.uint32)) if "pixels" in want else None
        colors = np.zeros((h, w, 3), np.float32) if "colors" in want else None
        seeds = np.zeros(2 * w * h, np.uint32) if "seeds" in want else None
        self._ck(self._lib.rt_pt_render(self._ctx, integrator, n_passes, _ptr(pixels), _ptr(colors), _ptr(seeds)))
        return {"pixels": pixels, "colors": colors, "seeds": seeds}

    def pt_launch(self, integrator, n_passes):
        self._ck(self._lib.rt_pt_launch(self._ctx, integrator, n_passes))

    def pt_download(self, want=("pixels", "colors", "seeds")):
        w, h = self.pt_size
        pixels = np.zeros((h, w), np.uint32) if "pixels" in want else None
        colors = np.zeros((h, w, 3), np.float32) if "colors" in want else None
        seeds = np.zeros(2 * w * h, np.uint32) if "seeds" in want else None
        self._ck(self._lib.rt_pt_download(self._ctx, _ptr(pixels), _ptr(colors), _ptr(seeds)))
        return {"pixels": pixels, "colors": colors, "seeds": seeds}

    def pt_restore(self, colors, seeds, current_sample):
        colors = np.ascontiguousarray(colors, dtype=np.float32)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        self._ck(self._lib.rt_pt_restore(self._ctx, _ptr(colors), _ptr(seeds), int(current_sample)))

    def pt_current_sample(self):
        return int(self._lib.rt_pt_current_sample(self._ctx))

    def pt_set_accumulate_sums(self, enabled):
        self._ck(self._lib.rt_pt_set_accumulate_sums(self._ctx, int(bool(enabled))))

    def pt_resolve_sums(self, total_samples):
        self._ck(self._lib.rt_pt_resolve_sums(self._ctx, total_samples))
