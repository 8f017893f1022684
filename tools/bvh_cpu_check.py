"""One-off CPU check of the exact hierarchy (csrc/pt_bvh.cuh) on the larger generated scenes through the lane simulator
(tests/devsim): identical results and the number of sphere tests it executes per query."""
import sys, time, ctypes, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
rt = g.load()
d = ctypes.CDLL(os.path.join(ROOT, 'tests/devsim/libdevsim.so'))
vp = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
for depth, w, h, passes in [(4, 160, 120, 2), (5, 120, 90, 2), (6, 64, 48, 1)]:
    path = os.path.join(ROOT, f'gpurun_out/complex{depth}.scn')
    rt.write_complex_scene(path, depth)
    sph, cam = rt.read_scene(path, w, h)
    seeds = rt.reference_seeds(w, h)
    st = np.zeros(5, np.int32); d.devsim_bvh_stats(vp(sph), sph.size, vp(st))
    res = []
    vis = np.zeros(2, np.int64)
    for mode in (-2, 0):
        if mode == -2: d.devsim_bvh_visits(vp(vis))
        col, sd, pix, ctr = np.zeros(3*w*h, np.float32), seeds.copy(), np.zeros(w*h, np.uint32), np.zeros(5, np.uint64)
        t = time.time()
        d.devsim_pt(0, vp(sph), sph.size, vp(cam), w, h, 0, passes, 0, vp(col), vp(sd), vp(pix), 0, 1, 8, vp(ctr), mode)
        if mode == -2: d.devsim_bvh_visits(vp(vis))
        res.append((col.view(np.uint32).copy(), sd, pix, ctr, time.time()-t))
    same = all(np.array_equal(a, b) for a, b in zip(res[0][:3], res[1][:3]))
    q = float(res[0][3][0] + res[0][3][1])
    print(f'depth {depth}: {sph.size} spheres, nodes {st[0]} big {st[2]} depth {st[3]} leaves {st[4]}; identical={same}; inner visits/query {vis[0]/q:.1f} leaf visits/query {vis[1]/q:.1f} tests/query bvh {res[0][3][2]/q:.1f} vs plain {res[1][3][2]/q:.1f}; cpu time {res[0][4]:.2f}s vs {res[1][4]:.2f}s')
