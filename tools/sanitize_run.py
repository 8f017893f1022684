"""Small workload for a checked run -- under compute-sanitizer where that is available, else against a library built with
-DRT_DEVICE_CHECKS (tools/checked_build.sh; the GPU pool of this project refuses compute-sanitizer): one launch of every kernel shape the library has, at
sizes a sanitizer run finishes in a minute -- Whitted 160x120 (lists, screen blocks + filler, hit IDs, the shadow-round culls), the Whitted
hierarchy on a 158-sphere table, the 3.0.06 frame (split and unsplit), Cornell 64x48 x 4 spp (step-aligned and plain), a 158-sphere scene
forced through the chunked staging, the 783-sphere scene through the path-tracer hierarchy, sum mode + resolve, the math self-test.
`python tools/sanitize_run.py ipc` is the two-rank part: rank 1 stores its rows into rank 0's frame through CUDA IPC.
Usage: compute-sanitizer --tool memcheck python tools/sanitize_run.py [ipc]"""
import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
rt = g.load()


def single():
    r = rt.Renderer(0)
    prims = rt.whitted_create_scene(0)
    for blocks, filler, order in ((1, 25, 1), (1, 0, 1), (0, 0, 1), (0, 0, 0)):
        r.set_tuning(rt.TUNE_WHITTED_BLOCKS, blocks); r.set_tuning(rt.TUNE_WHITTED_FILLER_PCT, filler); r.set_tuning(rt.TUNE_WHITTED_COST_ORDER, order)
        px, hits = r.whitted_render(prims, 160, 120, want_hit_ids=True)
    r.set_tuning(rt.TUNE_WHITTED_BLOCKS, 1); r.set_tuning(rt.TUNE_WHITTED_FILLER_PCT, 25); r.set_tuning(rt.TUNE_WHITTED_COST_ORDER, 1)
    # the tables, the split kernel and the exact re-launch (with a report list that overflows, too), a frame size with padding
    import numpy as np
    nan_scene = np.load(os.path.join(g.ROOT, "tests", "golden", "whitted_blocked_light_nan.npy"))
    for grid, split, cap in ((1, 1, 65536), (1, 0, 65536), (2, 0, 65536), (0, 0, 65536), (1, 1, 1)):
        r.set_tuning(rt.TUNE_WHITTED_GRID, grid); r.set_tuning(rt.TUNE_WHITTED_SPLIT, split); r.set_tuning(rt.TUNE_WHITTED_REDO_CAP, cap)
        r.whitted_render(prims, 333, 250, want_hit_ids=True)
        r.whitted_render(nan_scene, 97, 61, want_hit_ids=True)
    r.set_tuning(rt.TUNE_WHITTED_GRID, 1); r.set_tuning(rt.TUNE_WHITTED_SPLIT, 1); r.set_tuning(rt.TUNE_WHITTED_REDO_CAP, 65536)
    r.set_counting(True); r.whitted_upload(prims, 64, 48); r.whitted_launch(); r.counters(); r.set_counting(False)
    r.whitted_render(rt.whitted_create_scene(1), 96, 72, want_hit_ids=True)                 # 64 primitives: the hierarchy path
    with tempfile.TemporaryDirectory() as d:
        p3, p4 = os.path.join(d, "c3.scn"), os.path.join(d, "c4.scn")
        rt.write_complex_scene(p3, 3); rt.write_complex_scene(p4, 4)
        sph3, cam3 = rt.read_scene(p3, 64, 48)
        sph4, cam4 = rt.read_scene(p4, 64, 48)
    r.whitted_render(rt.whitted_from_spheres(sph3, cam3), 64, 48)
    r306 = rt.r306_create_scene()
    for split in (1, 0):
        r.set_tuning(rt.TUNE_R306_SPLIT, split)
        r.r306_render(r306, 160, 140)
    sph, cam = rt.cornell_scene(64, 48)
    seeds = rt.reference_seeds(64, 48)
    for aligned in (1, 0):
        r.set_tuning(rt.TUNE_PT_ALIGNED, aligned)
        for integ in (0, 1):
            r.pt_resize(64, 48, seeds); r.pt_set_scene(sph); r.pt_set_camera(cam)
            r.pt_render(integ, 4)
    r.set_tuning(rt.TUNE_PT_ALIGNED, -1)
    r.set_tuning(rt.TUNE_PT_BVH, 0); r.set_tuning(rt.TUNE_PT_MAX_RESIDENT_BYTES, 1024); r.set_tuning(rt.TUNE_PT_CHUNK_SPHERES, 50)   # 158 spheres in ragged chunks of 50
    r.pt_resize(64, 48, seeds); r.pt_set_scene(sph3); r.pt_set_camera(cam3); r.pt_render(0, 2)
    r.set_tuning(rt.TUNE_PT_BVH, -1); r.set_tuning(rt.TUNE_PT_MAX_RESIDENT_BYTES, 96 * 1024); r.set_tuning(rt.TUNE_PT_CHUNK_SPHERES, 3072)
    r.pt_resize(64, 48, seeds); r.pt_set_scene(sph4); r.pt_set_camera(cam4); r.pt_render(0, 2)            # 783 spheres: pt_bvh_kernel
    r.pt_set_accumulate_sums(True)
    r.pt_resize(64, 48, seeds); r.pt_set_scene(sph); r.pt_set_camera(cam); r.pt_launch(0, 3); r.pt_resolve_sums(3); r.pt_download()
    r.pt_set_accumulate_sums(False)
    r.selftest_math(0, np.linspace(0, 6.28, 1000, dtype=np.float32)); r.selftest_math(3, np.linspace(0, 100, 1000, dtype=np.float32))
    flags = r.debug_check_flags()
    r.close()
    print("sanitize_run single ok; device-side bounds checks:", "not compiled in" if flags == -1 else ("clean (0)" if flags == 0 else f"FAILED, mask {flags:#x}"))


def _ipc(rank, world, port):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r = rt.Renderer(0)
    prims = rt.whitted_create_scene(0)
    r.set_shard(rank, world, 8)
    r.whitted_upload(prims, 160, 120)
    box = [r.ipc_export(rt.BUF_WHITTED_PIXELS).tobytes() if rank == 0 else None]
    dist.broadcast_object_list(box, 0)
    if rank != 0:
        r.ipc_import(rt.BUF_WHITTED_PIXELS, np.frombuffer(box[0], np.uint8))
    dist.barrier()
    r.whitted_launch(); r.sync()
    dist.barrier()
    flags = r.debug_check_flags()
    print(f"rank {rank}: device-side bounds checks:", "not compiled in" if flags == -1 else ("clean (0)" if flags == 0 else f"FAILED, mask {flags:#x}"), flush=True)
    r.ipc_close()
    r.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "ipc":
        import socket
        import torch.multiprocessing as mp
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
        mp.spawn(_ipc, args=(2, port), nprocs=2, join=True)
        print("sanitize_run ipc ok")
    else:
        single()
