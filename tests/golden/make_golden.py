#!/usr/bin/env python3
"""Regenerates tests/golden/* from the reference (run in the build container, where /root/reference
and the compiled oracle/_ref exist).  The fixtures are OUTPUTS of the reference's own code and data:

  whitted_test_bmp.rgb.zlib   the reference's golden image R323/test.bmp (800x600), decoded to top-down
                              RGB bytes and zlib-compressed; whitted_golden.json holds the .bmp md5.
  whitted_golden.json         known-answer vectors from the compiled reference raytrace():
                              eight centre-sample primary rays and the 800x600 hit-ID histogram.
  smallpt_<scene>.npz         oracle/_ref (the reference's smallptCPU.cpp / geomfunc.h compiled as C++)
                              run on small frames: scene, camera, seeds in; colors / pixels / seeds out,
                              for the path-tracing and the direct-lighting integrator.
  smallpt_kat.json            GetRandom and SphereIntersect known answers from the compiled reference.
  complex_scene_md5.json      md5 of `perl scene_build_complex.pl` output for $maxDepth 1..5.
  r306_golden.json            raytracer3.0.06 (BASELINE config 1) frames rendered by the reference's own Engine_Render
                              (oracle/_ref/libref_r306.so): sha256 at 800x600 and two odd sizes, plus the 160x120
                              frame itself (r306_160x120.u32.zlib) for diffing.
  viewer_keys.json            a scripted session through the reference viewer's own keyFunc / specialFunc
                              (displayfunc.cpp compiled in oracle/_ref): camera and sphere table after every key.
"""
import ctypes, hashlib, json, os, re, struct, subprocess, sys, zlib
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
R323 = os.path.join(REF, "Raytracer3.2.03/raytracer/OpenCL Raytracer")
SPT = os.path.join(REF, "smallptgpu-v1.6")
vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)


def f32hex(x):
    return "%08x" % struct.unpack("<I", struct.pack("<f", x))[0]


def whitted():
    bmp = open(os.path.join(R323, "test.bmp"), "rb").read()
    off, = struct.unpack_from("<I", bmp, 10)
    w, h = struct.unpack_from("<ii", bmp, 18)
    assert (w, h) == (800, 600) and struct.unpack_from("<H", bmp, 28)[0] == 24
    rows = np.frombuffer(bmp, np.uint8, count=w * h * 3, offset=off).reshape(h, w, 3)   # 2400 B rows: no padding
    rgb = rows[::-1, :, ::-1].copy()                                                  # bottom-up BGR -> top-down RGB
    open(os.path.join(HERE, "whitted_test_bmp.rgb.zlib"), "wb").write(zlib.compress(rgb.tobytes(), 9))
    L = ctypes.CDLL(os.path.join(ROOT, "oracle/_ref/libref_whitted.so"))
    prims = np.zeros(96 * 64, np.uint8)
    n = L.ref_whitted_scene(vp(prims), 64)
    kat = []
    for (x, y) in [(0, 0), (400, 30), (400, 300), (250, 420), (560, 430), (430, 470), (120, 480), (700, 500)]:
        dist = ctypes.c_float(); col = (ctypes.c_float * 3)(); res = ctypes.c_int()
        hid = L.ref_whitted_probe(x, y, w, h, vp(prims), n, ctypes.byref(dist), col, ctypes.byref(res))
        kat.append({"x": x, "y": y, "hit": hid, "result": res.value, "dist": f32hex(dist.value),
                    "col": [f32hex(c) for c in col], "bmp_rgb": [int(v) for v in rgb[y, x]]})
    hits = np.zeros((h, w, 9), np.int32)
    L.ref_whitted_primary_hits(vp(hits), None, None, w, h, vp(prims), n)
    centre = hits[:, :, 4]
    hist = [int(np.count_nonzero(centre == k)) for k in range(-1, 17)]
    json.dump({"bmp_md5": hashlib.md5(bmp).hexdigest(), "width": w, "height": h, "n_primitives": n,
               "centre_primary_rays": kat, "centre_hit_histogram_ids_-1_to_16": hist,
               "all9_hit_id_sha256": hashlib.sha256(hits.tobytes()).hexdigest()},
              open(os.path.join(HERE, "whitted_golden.json"), "w"), indent=1)


def smallpt():
    L = ctypes.CDLL(os.path.join(ROOT, "oracle/_ref/libref_smallpt.so"))
    L.ref_pt_get_random.restype = ctypes.c_float
    L.ref_pt_sphere_intersect.restype = ctypes.c_float
    for scene, (w, h, passes) in {"cornell": (48, 36, 5), "caustic3": (40, 30, 4), "simple": (40, 30, 4),
                                  "complex": (20, 15, 2)}.items():
        n = L.ref_pt_load_scene(os.path.join(SPT, "scenes", scene + ".scn").encode(), w, h)
        sph = np.zeros(n * 11, np.float32); cam = np.zeros(15, np.float32)
        L.ref_pt_get_scene(vp(sph), vp(cam))
        seeds = np.maximum(np.random.RandomState(195).randint(0, 2 ** 31 - 1, size=2 * w * h).astype(np.uint32), 2)
        out = {"spheres": sph.view(np.uint32), "camera": cam.view(np.uint32), "seeds_in": seeds,
               "w": w, "h": h, "passes": passes}
        col = np.zeros(3 * w * h, np.float32); pix = np.zeros(w * h, np.uint32); sd = np.zeros(2 * w * h, np.uint32)
        L.ref_pt_render(vp(seeds), passes, vp(col), vp(pix), vp(sd))           # the reference's own loop
        out.update(pt_colors=col.view(np.uint32).copy(), pt_pixels=pix.copy(), pt_seeds=sd.copy())
        col = np.zeros(3 * w * h, np.float32); pix = np.zeros(w * h, np.uint32); sd = seeds.copy()
        L.ref_pt_render_mt(1, 0, passes, vp(col), vp(sd), vp(pix), 1)          # reference RadianceDirectLighting
        out.update(dl_colors=col.view(np.uint32).copy(), dl_pixels=pix.copy(), dl_seeds=sd.copy())
        np.savez_compressed(os.path.join(HERE, f"smallpt_{scene}.npz"), **out)
    kat = {"get_random": [], "sphere_intersect": []}
    for s0, s1 in [(2, 2), (1804289383, 846930886), (0xffffffff, 0x12345678)]:
        a, b = ctypes.c_uint(s0), ctypes.c_uint(s1)
        seq = []
        for _ in range(4):
            f = L.ref_pt_get_random(ctypes.byref(a), ctypes.byref(b))
            seq.append(["%08x" % a.value, "%08x" % b.value, f32hex(f)])
        kat["get_random"].append({"start": [s0, s1], "calls": seq})
    sphere = np.array([16.5, 27, 16.5, 47, 0, 0, 0, .9, .9, .9, 0], np.float32); sphere[10:].view(np.int32)[0] = 1
    d = np.array([-0.14, -0.17, -0.975], np.float32); d = (d * np.float32(1.0 / np.sqrt(np.float32(d @ d)))).astype(np.float32)
    for o, dd in [((50, 45, 205.6), d), ((50, 45, 205.6), (0, 0, -1)), ((27, 16.5, 47), (0, 0, -1))]:
        o = np.array(o, np.float32); dd = np.array(dd, np.float32)
        t = L.ref_pt_sphere_intersect(vp(sphere), vp(o), vp(dd))
        kat["sphere_intersect"].append({"sphere": [f32hex(v) for v in sphere[:4]], "o": [f32hex(v) for v in o],
                                        "d": [f32hex(v) for v in dd], "t": f32hex(t)})
    json.dump(kat, open(os.path.join(HERE, "smallpt_kat.json"), "w"), indent=1)


def complex_scene():
    src = open(os.path.join(SPT, "scene_build_complex.pl")).read()
    md5 = {}
    for depth in range(1, 6):
        txt = subprocess.run(["perl", "-e", re.sub(r"\$maxDepth = [0-9.]+;", f"$maxDepth = {depth}.0;", src)],
                             check=True, capture_output=True, text=True).stdout
        md5[str(depth)] = {"md5": hashlib.md5(txt.encode()).hexdigest(), "spheres": txt.count("\n")}
    json.dump(md5, open(os.path.join(HERE, "complex_scene_md5.json"), "w"), indent=1)


VIEWER_KEYS = ["a", "a", "w", "S103", "d", "r", "S100", "S100", "S101", "s", "f", "S102", "S104", "S105", "+", "+", "4", "9", "-", "8", "2", "6", "3",
               "S101", "S101", "w", "+", "+", "+", "+", "+", "+", "+", "+", "6", "h"]


def viewer_keys():
    L = ctypes.CDLL(os.path.join(ROOT, "oracle/_ref/libref_smallpt.so"))
    w, h = 8, 6
    n = L.ref_pt_load_scene(os.path.join(SPT, "scenes", "cornell.scn").encode(), w, h)
    seeds = np.full(2 * w * h, 7, np.uint32)
    L.ref_pt_render(vp(seeds), 1, None, None, None)                 # the key handlers re-render: buffers must exist
    steps = []
    for k in VIEWER_KEYS:
        L.ref_pt_key(int(k[1:]) if k.startswith("S") else ord(k), 1 if k.startswith("S") else 0)
        sph = np.zeros(n * 11, np.float32); cam = np.zeros(15, np.float32)
        L.ref_pt_get_scene(vp(sph), vp(cam))
        steps.append({"key": k, "camera": ["%08x" % v for v in cam.view(np.uint32)],
                      "spheres_sha256": hashlib.sha256(sph.tobytes()).hexdigest()})
    json.dump({"scene": "cornell.scn", "w": w, "h": h, "steps": steps}, open(os.path.join(HERE, "viewer_keys.json"), "w"), indent=1)


def r306():
    L = ctypes.CDLL(os.path.join(ROOT, "oracle/_ref/libref_r306.so"))
    out = {"pixel_format": "0x00RRGGBB, rows 20 .. h-71 rendered, the rest zero", "frames": {}}
    for (w, h) in [(160, 120), (203, 131), (800, 600), (1003, 377)]:
        img = np.zeros((h, w), np.uint32)
        L.ref_r306_render(vp(img), w, h)
        out["frames"][f"{w}x{h}"] = hashlib.sha256(img.tobytes()).hexdigest()
        if (w, h) == (160, 120):
            open(os.path.join(HERE, "r306_160x120.u32.zlib"), "wb").write(zlib.compress(img.tobytes(), 9))
    json.dump(out, open(os.path.join(HERE, "r306_golden.json"), "w"), indent=1)


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference (build container)")
    only = sys.argv[1:]
    for name, fn in (("whitted", whitted), ("smallpt", smallpt), ("complex_scene", complex_scene), ("viewer_keys", viewer_keys), ("r306", r306)):
        if not only or name in only:
            fn()
    print("golden fixtures written to", HERE)
