"""Diagnostic: where does the GPU Whitted frame differ from the oracle / the host build of the lane code?"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
rt = g.load(); orc = g.oracle()
dev = ctypes.CDLL(os.path.join(g.DEVSIM_DIR, "libdevsim.so"))
vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
prims = rt.whitted_create_scene(0)
r = rt.Renderer(0)
px = r.whitted_render(prims, w, h)
po = np.zeros((h, w, 4), np.uint8); orc.oracle_whitted_render(vp(po), None, w, h, vp(prims), prims.size, 16, None)
pd = np.zeros((h, w, 4), np.uint8); dev.devsim_whitted(vp(pd), None, w, h, vp(prims), prims.size, 0, 1, 8, None, None, 1)
ref_path = os.path.join(g.ORACLE_DIR, "_ref", "libref_whitted.so")
bad = np.argwhere((px != po).any(axis=2))
print("gpu!=oracle pixels:", len(bad), "of", w * h, " devsim!=oracle:", int((pd != po).any(axis=2).sum()), " gpu!=devsim:", int((px != pd).any(axis=2).sum()))
for y, x in bad[:20]:
    print((x, y), "gpu", px[y, x, :3], "oracle", po[y, x, :3], "devsim", pd[y, x, :3])
if len(bad):
    d = np.abs(px.astype(int) - po.astype(int)); print("max abs diff", d.max())
