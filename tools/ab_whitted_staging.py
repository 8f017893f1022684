"""Does staging the Whitted scene tables in shared memory matter?  The same frames with the stage mode capped (RT_TUNE_WHITTED_STAGE_CAP):
scene 0 (17 primitives) and scene 0 plus 40 / 400 / 2 000 extra walls outside the room (never hit, but tested by every ray: a table far
beyond the per-CTA share of shared memory).  Kernel time by CUDA events, best of 5; all frames must be identical.
Usage: python tools/ab_whitted_staging.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
rt = g.load()
r = rt.Renderer(0)
box = rt.whitted_create_scene(0)[:16]
rs = np.random.RandomState(3)
for extra in (0, 40, 400, 2000):
    walls = np.zeros(extra, box.dtype)
    if extra:
        walls[:] = box[8]
        nrm = rs.normal(0, 1, (extra, 3)); nrm /= np.linalg.norm(nrm, axis=1)[:, None]
        walls["normal"][:, :3] = nrm.astype(np.float32)
        walls["depth"] = rs.uniform(60, 90, extra).astype(np.float32)          # 60-90 units from the origin: outside the room, behind its walls
    prims = np.concatenate([box, walls])
    w, h = (1920, 1080) if extra <= 40 else (480, 270)
    ref, out = None, []
    for cap in (-1, 2, 1, 0):
        r.set_tuning(rt.TUNE_WHITTED_STAGE_CAP, cap)
        try:
            img = r.whitted_render(prims, w, h)
        except rt.RtError as e:
            out.append(f"cap {cap}: {e}"); continue
        if ref is None: ref = img
        same = np.array_equal(img, ref)
        r.whitted_upload(prims, w, h)
        ts = []
        for _ in range(5):
            r.timer_begin(); r.whitted_launch(); ts.append(r.timer_end())
        out.append(f"cap {cap}: {min(ts):.3f} ms{'' if same else ' DIFFERENT IMAGE'}")
    print(f"{prims.size} primitives, {w}x{h}: " + " | ".join(out), flush=True)
r.set_tuning(rt.TUNE_WHITTED_STAGE_CAP, -1)
r.close()
