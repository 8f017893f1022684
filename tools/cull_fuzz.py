"""Differential fuzzing of the Whitted shadow-round culls (whitted_lane.cuh "Exact culls of the shadow round", tables from
build_w_cull in scene_soa.h) on the CPU: the lane code as a timed launch runs it (tests/devsim, mode 4: hot runs, culls on)
against the oracle, on random rooms -- planes of random orientation and length of normal, spheres, 1-3 sphere lights placed
anywhere including a hair's breadth from a wall, between two walls' sides, inside the sphere cluster.  Pixels and hit IDs must
be identical.  `--self-check` runs the same scenes against a devsim built with -DW_CULL_TEST_NO_MARGIN (T = 0, boxes not
grown, every face usable): that build MUST be caught, otherwise the fuzz proves nothing.
Since round 2 mode 4 includes the shadow-candidate grid, the primary-ray tiles and the exact re-render of reported pixels; `--split` runs
mode 6 (every pixel one lane per sub-sample with ordered logs, as whitted_split_kernel renders them).
Usage: python tools/cull_fuzz.py [n_scenes] [seed] [--self-check] [--split]"""
import ctypes, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g


def vp(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def random_scene(rt, rs, box):
    """A Primitive_2 table around the Whitted tracer's fixed eye (0, 0.25, -7)."""
    eye = np.array([0.0, 0.25, -7.0])
    parts = []
    kind = rs.randint(0, 4)
    if kind == 0:                                    # the reference's room, walls possibly pushed around
        walls = box[[0, 8, 9, 10, 11, 12]].copy()
        walls["depth"] += rs.uniform(-0.5, 2.0, walls.size).astype(np.float32)
        parts.append(walls[rs.permutation(walls.size)[: rs.randint(2, 7)]])
    else:                                            # random planes that keep the eye on their positive side
        k = rs.randint(1, 9)
        pl = np.zeros(k, box.dtype); pl[:] = box[0]
        for i in range(k):
            nrm = rs.normal(0, 1, 3)
            if rs.rand() < 0.4:
                nrm = np.eye(3)[rs.randint(0, 3)] * rs.choice([-1, 1])
            nrm = nrm / np.linalg.norm(nrm) * np.exp(rs.uniform(np.log(0.05), np.log(4.0)))
            dist = np.exp(rs.uniform(np.log(0.5), np.log(40.0)))           # distance of the plane from the eye, eye in front
            depth = dist * np.linalg.norm(nrm) - nrm @ eye
            pl["normal"][i, :3] = nrm.astype(np.float32); pl["depth"][i] = np.float32(depth)
            pl["m_color"][i, :3] = rs.uniform(0.2, 1.0, 3)
            pl["m_diff"][i] = rs.uniform(0.2, 1.2); pl["m_spec"][i] = rs.uniform(0, 1.5) * (rs.rand() < 0.7)
            pl["m_refl"][i] = rs.uniform(0, 0.6) * (rs.rand() < 0.3)
        parts.append(pl)
    ns = rs.randint(0, 12)
    if ns:
        sp = np.zeros(ns, box.dtype); sp[:] = box[2]
        c = rs.uniform(-8, 8, (ns, 3)); c[:, 2] = rs.uniform(0.0, 35.0, ns)
        if rs.rand() < 0.5:
            c[:, 1] = rs.uniform(-6, -1, ns)         # a cluster low in the room, like the reference's
        rad = np.exp(rs.uniform(np.log(0.02), np.log(3.0), ns)).astype(np.float32)
        sp["center"][:, :3] = c.astype(np.float32)
        sp["radius"] = rad; sp["sq_radius"] = rad * rad; sp["r_radius"] = np.float32(1.0) / rad
        sp["m_color"][:, :3] = rs.uniform(0.05, 1.5, (ns, 3))
        sp["m_diff"] = rs.uniform(0, 1, ns) * (rs.rand(ns) < 0.8); sp["m_spec"] = rs.uniform(0, 1.5, ns) * (rs.rand(ns) < 0.6)
        sp["m_refl"] = rs.uniform(0, 0.9, ns) * (rs.rand(ns) < 0.4)
        refr = rs.rand(ns) < 0.25
        sp["m_refr"] = np.where(refr, rs.uniform(0.3, 1.0, ns), 0); sp["m_refr_index"] = np.where(refr, rs.uniform(1.1, 1.6, ns), 0)
        sp["is_light"] = 0
        parts.append(sp)
    nl = rs.randint(1, 4)
    lights = box[13:16][:nl].copy()
    for i in range(nl):
        mode = rs.randint(0, 5)
        if mode == 0:
            p = box[13 + i]["center"][:3] + rs.uniform(-1, 1, 3)
        elif mode == 1:                               # a hair's breadth from a plane of the scene (either side)
            planes = np.concatenate([q for q in parts if q["type"][0] == 0])
            q = planes[rs.randint(0, planes.size)]
            nrm = q["normal"][:3].astype(np.float64); nn = np.linalg.norm(nrm)
            foot = rs.uniform(-6, 6, 3); foot[2] = rs.uniform(0, 30)
            foot = foot - nrm * ((nrm @ foot + q["depth"]) / (nn * nn))
            p = foot + nrm / nn * rs.choice([-1, 1]) * np.exp(rs.uniform(np.log(1e-4), np.log(0.5)))
        elif mode == 2 and ns:                        # inside / right next to the sphere cluster
            p = parts[-1]["center"][rs.randint(0, ns), :3] + rs.uniform(-3, 3, 3)
        else:
            p = rs.uniform(-7, 7, 3); p[2] = rs.uniform(-6, 34)
        lights["center"][i, :3] = np.asarray(p, np.float32)
        rad = np.float32(np.exp(rs.uniform(np.log(0.01), np.log(0.6))))
        lights["radius"][i] = rad; lights["sq_radius"][i] = rad * rad; lights["r_radius"][i] = np.float32(1.0) / rad
    parts.insert(rs.randint(0, len(parts) + 1), lights)
    order = rs.permutation(len(parts))
    return np.concatenate([parts[i] for i in order])


def build_no_margin():
    out = os.path.join(g.DEVSIM_DIR, "libdevsim_nomargin.so")
    subprocess.run(["g++", "-O2", "-fPIC", "-fno-fast-math", "-ffp-contract=off", "-std=c++17", "-shared", "-w", "-DW_CULL_TEST_NO_MARGIN",
                    "-x", "c++", "devsim.cpp", "-o", out, "-lm"], cwd=g.DEVSIM_DIR, check=True)
    return ctypes.CDLL(out)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n_scenes = int(args[0]) if args else 200
    seed = int(args[1]) if len(args) > 1 else 1
    self_check = "--self-check" in sys.argv
    mode = 6 if "--split" in sys.argv else 4
    rt = g.load(); orc = g.oracle()
    dev = build_no_margin() if self_check else ctypes.CDLL(os.path.join(g.DEVSIM_DIR, "libdevsim.so"))
    rs = np.random.RandomState(seed)
    box = rt.whitted_create_scene(0)
    w, h = 64, 48
    bad = culled_scenes = 0
    t0 = time.time()
    for it in range(n_scenes):
        prims = random_scene(rt, rs, box)
        st = np.zeros(4, np.int32)
        dev.devsim_whitted_cull_stats(vp(prims), prims.size, vp(st))
        culled_scenes += int(st[0])
        px, hits = np.zeros((h, w, 4), np.uint8), np.zeros((h, w, 9), np.int32)
        dev.devsim_whitted(vp(px), vp(hits), w, h, vp(prims), prims.size, 0, 1, 8, None, None, mode)
        px_o, hits_o = np.zeros((h, w, 4), np.uint8), np.zeros((h, w, 9), np.int32)
        orc.oracle_whitted_render(vp(px_o), vp(hits_o), w, h, vp(prims), prims.size, 8, None)
        if not (np.array_equal(px, px_o) and np.array_equal(hits, hits_o)):
            bad += 1
            if bad <= 10:
                print(f"MISMATCH scene {it}: n={prims.size}, {int(np.count_nonzero((px != px_o).any(axis=2)))} pixels differ", flush=True)
    print(f"cull fuzz{' (NO-MARGIN self-check build)' if self_check else ''}{' (split mode)' if mode == 6 else ''}: {n_scenes} scenes {w}x{h}, culls active in {culled_scenes}: "
          f"{bad} mismatches ({time.time() - t0:.0f} s)")
    if self_check:
        sys.exit(0 if bad > 0 else 1)               # the broken build must be caught
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
