#!/usr/bin/env python3
"""bench.py -- the measurement contract.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): the Raytracer3.2.03 Whitted scene (CHOOSE_SCENE 0, 17 primitive slots), 3x3
super-sampling, TRACEDEPTH 5, at 1920x1080 on one B200.  A "step" is one frame.  With N > 1 GPUs the frame grows with N (weak
scaling: 1920x1080 pixels per GPU; the Whitted view window is fixed, so this is the same picture at a higher resolution), is
sharded by interleaved row tiles (rt_set_shard) and is assembled on rank 0 over NVLink inside the step.

Metric: Mrays/s.  rays = raytrace() calls + shadow rays of the REFERENCE's algorithm on this frame ("reference-equivalent" rays,
counted by a separate counting launch; identical in the oracle, tests/test_gpu_parity.py).  The timed kernel produces the same
pixels but does not trace every one of them (no shadow rays for hits on a material with neither a diffuse nor a specular term)
nor execute every primitive test (exact culls, hierarchy): `config.rays_traced_per_frame` is what it really traces, and
`roofline.algorithmic_flop_per_launch` is the reference's test count x 16 / 12 FLOP, stated as such.

  value     frames resident in HBM: K x (kernel [+ NVLink frame assembly]) timed with CUDA events on the stream the kernels run on
            (torch's current stream, injected with rt_set_stream), L2 flushed between steps outside the events, max over ranks.
  e2e       the reference-facing call with HOST buffers, host clock.  N = 1: rt_whitted_render(ctx, prims, n, w, h, pixels, NULL)
            = scene upload + kernels + read-back of the frame, into pinned memory (`e2e`) and into pageable memory as the
            reference's caller allocates it (`e2e_pageable`).  N > 1: every rank uploads, renders its rows and copies them over
            its OWN PCIe link into one host frame shared by the ranks (rt_whitted_download_rows), then a host barrier.
  roofline  FP32-FMA bound: algorithmic FLOP = 16 x sphere tests + 12 x plane tests per frame, divided by the
            kernel's mean duration, against 148 SM x 128 lanes x 2 x f_max.
  cpu_baseline / --impl reference
            the reference's OWN CPU code (oracle/_ref: raytracer_non_OpenCL.c compiled unmodified) on the box's host cores, one
            frame per thread on a bounded sample; `cpu_baseline_1core` is the same on one thread (the reference program is
            single-threaded).  The reference arm never loads the product library.

Extra blocks (not the headline): "pt" = BASELINE configs[2] (smallpt cornell.scn 1024x768 x 256 spp, Msamples/s), "c4" =
configs[3] (783-sphere generated scene, 3840x2160 x 16 spp), "config1" = configs[0]; at N > 1: "strong" (fixed 1920x1080 and
3840x2160 frames over the N GPUs) and "c5" = configs[4] (cornell 7680x4320 x 1024 spp, image-sharded and sample-sharded +
all-reduce, each with its check).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
ORACLE_DIR = os.path.join(ROOT, "oracle")

BASE_W, BASE_H = 1920, 1080
SAMPLE_W, SAMPLE_H = 480, 270          # CPU sample frame: 1/16 of the pixels, same rays per pixel
PT_W, PT_H, PT_SPP = 1024, 768, 256
WEAK_SIZES = {1: (1920, 1080), 2: (1920, 2160), 4: (3840, 2160), 8: (3840, 4320)}
C5_W, C5_H, C5_SPP = 7680, 4320, 1024
WORKLOAD = "Raytracer3.2.03 Whitted scene (CHOOSE_SCENE 0, 17 primitive slots, 3x3 AA, TRACEDEPTH 5) {w}x{h}"


def vp(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "sm_max_mhz": d.get("sm_max_mhz", 1965.0), "source": "MEASURED_PEAKS.json"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback (B200_PROFILING.md)"}


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the newest committed `ncu --set full` summary
    (profiles/r*_ncu_full_summary.json, made by tools/ncu_summary.py from the 1920x1080 frame); None when there is none."""
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def to_bytes(text):
        v, u = text.split()
        return float(v) * unit[u]

    prof = os.path.join(ROOT, "profiles")
    for name in sorted((f for f in os.listdir(prof) if f.endswith("_ncu_full_summary.json")), reverse=True) if os.path.isdir(prof) else []:
        try:
            for m in json.load(open(os.path.join(prof, name))):
                if kernel in m.get("Kernel Name", "") and "dram__bytes_read.sum" in m:
                    return int(to_bytes(m["dram__bytes_read.sum"]) + to_bytes(m["dram__bytes_write.sum"])), name
        except (OSError, ValueError, KeyError, TypeError, AttributeError):
            continue
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.mark = [], None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def begin(self):
        self.mark = time.perf_counter()

    def end(self):
        t1 = time.perf_counter()
        if self.proc:
            time.sleep(0.12)
            self.proc.terminate()
        rows = [r for t, r in self.rows if self.mark is not None and self.mark <= t <= t1 + 0.12] or [r for _, r in self.rows[-3:]]
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(rows)}


# --------------------------------------------------------------------------------------------- CPU legs (oracle/ only, never the product)
def cpu_libs():
    """(liboracle.so, oracle/_ref/libref_whitted.so or None, libref_smallpt.so or None): the checker libraries, loaded directly."""
    orc = ctypes.CDLL(os.path.join(ORACLE_DIR, "liboracle.so"))

    def opt(name):
        p = os.path.join(ORACLE_DIR, "_ref", name)
        return ctypes.CDLL(p) if os.path.exists(p) else None

    return orc, opt("libref_whitted.so"), opt("libref_smallpt.so")


def cpu_whitted_scene(orc, ref):
    """The config-2 scene table (17 x 96 bytes) from the reference's own create_scene (oracle/_ref) or the oracle's restatement."""
    buf = np.zeros(64 * 96, np.uint8)
    n = ref.ref_whitted_scene(vp(buf), 64) if ref is not None else orc.oracle_whitted_scene0(vp(buf), 64)
    return buf[: n * 96].copy(), n


def time_reference(steps, warmup, threads):
    """K steps of: `threads` host threads each render one SAMPLE_W x SAMPLE_H frame with the reference's own raytracer_non_kernel
    (oracle/_ref).  Returns (Mrays/s, ms per step, kind, description).  Loads nothing but oracle/."""
    orc, ref, _ = cpu_libs()
    prims, n = cpu_whitted_scene(orc, ref)
    ctr = np.zeros(5, np.uint64)
    orc.oracle_whitted_render(vp(np.zeros((SAMPLE_H, SAMPLE_W, 4), np.uint8)), None, SAMPLE_W, SAMPLE_H, vp(prims), n, host_threads(), vp(ctr))
    rays = int(ctr[0] + ctr[1])
    px = np.zeros((threads, SAMPLE_H, SAMPLE_W, 4), np.uint8)
    if ref is not None:
        run = lambda: ref.ref_whitted_render_mt(vp(px), SAMPLE_W, SAMPLE_H, vp(prims), n, threads)
        kind = "reference"
    else:       # the oracle port (only if oracle/_ref was not shipped)
        run = lambda: [orc.oracle_whitted_render(vp(px[0]), None, SAMPLE_W, SAMPLE_H, vp(prims), n, threads, None) for _ in range(threads)]
        kind = "port"
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = time.perf_counter() - t0
    mrays = rays * threads * steps / dt / 1e6
    sample = (f"{threads} host thread(s) x {steps} steps, each thread one {SAMPLE_W}x{SAMPLE_H} frame of the config-2 scene "
              f"({rays / (SAMPLE_W * SAMPLE_H):.1f} rays/pixel, same as 1080p) with the reference's raytracer_non_kernel, g++ -O2")
    return mrays, dt / steps * 1e3, kind, sample


def run_reference(args, rank):
    if rank != 0:
        return
    threads = host_threads()
    steps = max(1, min(args.steps, 8))          # bounded: a step is ~1 s of wall clock on every core
    warmup = max(1, min(args.warmup, 1))
    mrays, ms, kind, sample = time_reference(steps, warmup, threads)
    w, h = WEAK_SIZES.get(args.gpus, WEAK_SIZES[1])
    print(json.dumps({
        "impl": "reference", "metric": "Mrays/s", "value": round(mrays, 3), "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(w=w, h=h), "parallelism": f"{threads} host threads",
                   "sampling": f"the CPU arm times {SAMPLE_W}x{SAMPLE_H} frames of that scene (rays per pixel do not depend on the resolution) and quotes Mrays/s"},
        "cpu_baseline": {"value": round(mrays, 3), "unit": "Mrays/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": round(mrays, 3), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "native_libraries": "oracle/ only (liboracle.so for the ray count, oracle/_ref/libref_whitted.so for the scene table and the timed code)",
    }))


def cpu_pt_baseline(spheres, cam_fn, sw, sh, sp, threads, label):
    """Msamples/s of the reference's RadiancePathTracing (oracle/_ref) on `threads` host threads over a bounded sample."""
    orc, _, ref = cpu_libs()
    cam_s = cam_fn(sw, sh)
    sd = np.maximum(np.random.RandomState(1).randint(0, 2 ** 31 - 1, size=2 * sw * sh, dtype=np.int64).astype(np.uint32), 2).astype(np.uint32)
    col = np.zeros(3 * sw * sh, np.float32)
    if ref is not None:
        ref.ref_pt_set_scene(vp(spheres), spheres.size, vp(cam_s), sw, sh)
        t0 = time.perf_counter()
        ref.ref_pt_render_mt(0, 0, sp, vp(col), vp(sd), None, threads)
        kind = "reference"
    else:
        t0 = time.perf_counter()
        orc.oracle_pt_render(0, vp(spheres), spheres.size, vp(cam_s), sw, sh, 0, sp, vp(col), vp(sd), None, threads, None)
        kind = "port"
    dt = time.perf_counter() - t0
    return {"value": round(sw * sh * sp / dt / 1e6, 3), "unit": "Msamples/s", "cores": threads, "kind": kind,
            "sample": f"{label} {sw}x{sh} x {sp} spp, rows split over {threads} host thread(s), reference RadiancePathTracing"}


# --------------------------------------------------------------------------------------------- our arm
class Frame:
    """One Whitted frame size on this rank: upload, the fused (CUDA-IPC) or NCCL frame assembly, a torch view of the framebuffer."""

    def __init__(self, rt, r, torch, dist, prims, w, h, rank, world):
        self.rt, self.r, self.torch, self.dist, self.rank, self.world = rt, r, torch, dist, rank, world
        self.w, self.h = w, h
        self.tile = rt.pick_tile_rows(h, world)
        r.set_shard(rank, world, self.tile)
        r.whitted_upload(prims, w, h)
        ptr, _ = r.device_buffer(rt.BUF_WHITTED_PIXELS)
        self.fb = torch.as_tensor(rt.DeviceArray(ptr, (h, w), "<i4"), device="cuda")
        # Preferred: fused -- the render kernels of ranks 1.. store their rows straight into rank 0's framebuffer through a CUDA-IPC
        # peer mapping (NVLink), and one tiny all-reduce orders "all kernels done".  Fallback: NCCL send/recv + strided de-interleave.
        self.fused = rt.share_rank0_framebuffer(r, rt.BUF_WHITTED_PIXELS, rank, world)
        self.staging = rt.gather_staging(self.fb, world, self.tile) if (rank == 0 and world > 1 and not self.fused) else None
        self.flag = torch.zeros(1, dtype=torch.int32, device="cuda")

    def assemble(self):
        if self.world == 1:
            return
        if self.fused:
            self.dist.all_reduce(self.flag)
        else:
            self.rt.gather_row_tiles(self.fb, self.rank, self.world, self.tile, self.staging)

    def check_against_one_gpu(self):
        """The assembled frame equals a 1-GPU render of the same frame (rank 0 renders it all; untimed)."""
        torch, dist, r = self.torch, self.dist, self.r
        self.fb.zero_()
        torch.cuda.synchronize(); dist.barrier()
        r.whitted_launch(); self.assemble()
        torch.cuda.synchronize(); dist.barrier()
        verdict = None
        if self.rank == 0:
            assembled = self.fb.clone()
            r.set_shard(0, 1, self.tile)
            r.whitted_launch(); r.sync()
            verdict = "bit-identical to the 1-GPU frame" if bool((self.fb == assembled).all()) else "MISMATCH"
            r.set_shard(self.rank, self.world, self.tile)
        dist.barrier()
        return verdict

    def time_steps(self, steps, warmup, flush, stream, sampler=None):
        """(ms per step, kernel ms per step) over `steps` frames, CUDA events, max over ranks."""
        torch, dist, r = self.torch, self.dist, self.r
        for _ in range(warmup):
            r.whitted_launch(); self.assemble()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.begin()
        launches0 = r.launch_count()
        for a, k, b in ev:
            flush.zero_()                      # L2 flush, outside the events
            a.record(stream)
            r.whitted_launch()
            k.record(stream)
            self.assemble()
            b.record(stream)
        torch.cuda.synchronize()
        self.timed_launches = r.launch_count() - launches0           # kernels of librt_b200.so launched inside the timed region
        if self.world > 1:
            dist.barrier()
        t = torch.tensor([sum(a.elapsed_time(b) for a, _, b in ev) / steps, sum(a.elapsed_time(k) for a, k, _ in ev) / steps], dtype=torch.float64, device="cuda")
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0].item()), float(t[1].item())

    def close(self):
        if self.world > 1:
            self.torch.cuda.synchronize(); self.dist.barrier()
            self.r.ipc_close()
            self.dist.barrier()


def count_frame(r, dist, torch, world):
    """Counting launch of the uploaded frame: (reference-equivalent rays, rays the timed launch traces, sphere tests, plane tests) over
    all ranks, and this rank's algorithmic FLOP."""
    r.set_counting(True)
    r.whitted_launch()
    cnt = r.counters()
    ex = r.counters_ex()
    r.set_counting(False)
    local = torch.tensor([cnt["nearest_queries"] + cnt["shadow_queries"], cnt["nearest_queries"] + int(ex[5]), cnt["sphere_tests"], cnt["plane_tests"]],
                         dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(local)
    rays, traced, sph, pl = (float(v) for v in local.tolist())
    return rays, traced, sph, pl, 16.0 * cnt["sphere_tests"] + 12.0 * cnt["plane_tests"]


def shared_host_frame(dist, rank, world, nbytes):
    """One host buffer of `nbytes` mapped by every rank's process (POSIX shared memory; rank 0 creates and names it)."""
    from multiprocessing import shared_memory
    box = [None]
    shm = None
    if rank == 0:
        shm = shared_memory.SharedMemory(create=True, size=nbytes)
        box[0] = shm.name
    if world > 1:
        dist.broadcast_object_list(box, 0)
    if rank != 0:
        shm = shared_memory.SharedMemory(name=box[0])
        try:        # attaching registers the segment with this process's resource tracker too (CPython < 3.13); rank 0 owns and unlinks it
            from multiprocessing import resource_tracker
            resource_tracker.unregister(shm._name, "shared_memory")
        except Exception:
            pass
    return shm


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as graft
    rt = graft.load()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    r = rt.Renderer(local_rank)
    stream = torch.cuda.current_stream()
    r.set_stream(stream.cuda_stream)
    info = r.device_info()
    peaks = measured_peaks()
    fp32_peak = info["sm_count"] * 128 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12       # TFLOP/s
    w, h = WEAK_SIZES.get(world, (BASE_W, BASE_H * world))
    prims = rt.whitted_create_scene(0)
    flush = torch.empty(384 * 1024 * 1024, dtype=torch.uint8, device="cuda")       # > 126 MB L2
    steps, warmup = args.steps, max(args.warmup, 3)

    # ---- headline: weak-scaling frame, device-resident steps
    fr = Frame(rt, r, torch, dist, prims, w, h, rank, world)
    rays_frame, traced_frame, sph_tests, pl_tests, flop_local = count_frame(r, dist, torch, world)
    frame_check = fr.check_against_one_gpu() if world > 1 else None
    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.6 if sampler else 0.0)
    ms_per_step, kern_ms = fr.time_steps(steps, warmup, flush, stream, sampler)
    clocks = sampler.end() if sampler else None
    launches = fr.timed_launches
    value = rays_frame / (ms_per_step * 1e-3) / 1e6
    tile, fused = fr.tile, fr.fused
    fr.close()                                                         # from here on every rank renders into its own buffer again

    # ---- e2e through the reference-facing calls with host buffers
    scene_bytes = prims.size * (3 * 16 + 4 + 4) + 4 * int(prims["is_light"].sum())
    e2e = {}
    r.set_shard(rank, world, tile)
    if world == 1:
        pinned = torch.empty((h, w, 4), dtype=torch.uint8).pin_memory()
        for tag, dest in (("e2e", pinned.numpy()), ("e2e_pageable", np.zeros((h, w, 4), np.uint8))):
            for _ in range(3):
                r.whitted_render(prims, w, h, pixels_out=dest)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(steps):
                r.whitted_render(prims, w, h, pixels_out=dest)
            dt = time.perf_counter() - t0
            e2e[tag] = {"value": round(rays_frame * steps / dt / 1e6, 1), "unit": "Mrays/s", "h2d_bytes_per_step": int(scene_bytes), "d2h_bytes_per_step": int(w * h * 4),
                        "host_buffer": "pinned (cudaHostAlloc)" if tag == "e2e" else "pageable (numpy / malloc, as the reference's caller allocates it: R323/raytracer.c:18-42)"}
    else:
        shm = shared_host_frame(dist, rank, world, w * h * 4)
        host = np.ndarray((h, w, 4), np.uint8, buffer=shm.buf)
        r.host_register(host)

        def e2e_step():
            r.whitted_upload(prims, w, h)
            r.whitted_launch()
            r.whitted_download_rows(host)                 # this rank's rows -> their place in the shared host frame, over its own PCIe link

        for _ in range(3):
            e2e_step()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            e2e_step()
        dist.barrier()                                     # all rows are in the host frame
        e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        host_check = None
        if rank == 0:                                      # the host frame equals the frame assembled on the device (untimed)
            r.set_shard(0, 1, tile); r.whitted_launch(); r.sync()
            host_check = "host frame bit-identical to the 1-GPU frame" if np.array_equal(host, r.whitted_download()) else "MISMATCH"
            r.set_shard(rank, world, tile)
        dist.barrier()
        e2e["e2e"] = {"value": round(rays_frame * steps / float(e2e_s.item()) / 1e6, 1), "unit": "Mrays/s", "h2d_bytes_per_step": int(scene_bytes) * world,
                      "d2h_bytes_per_step": int(w * h * 4), "host_buffer": "one POSIX shared-memory frame, page-locked in every rank's process (rt_host_register); "
                      "each rank copies its own row tiles (rt_whitted_download_rows), host barrier at the end", "frame_check": host_check}
        r.host_unregister(host)
        del host
        shm.close()
        if rank == 0:
            shm.unlink()

    out = None
    if rank == 0:
        achieved = flop_local / (kern_ms * 1e-3) / 1e12
        # one rt_whitted_launch = the tile pre-pass, whitted_split_kernel (pixels whose rays have children, one lane per sub-sample) and
        # whitted_wall_kernel (pure wall blocks) side by side on two streams, and the exact re-launch (whitted_kernel<.., EXACT>) over the reported pixels
        parts = [ncu_traffic(k) for k in ("whitted_split_kernel", "whitted_wall_kernel", "whitted_kernel<")]
        traffic_src = next((src for _, src in parts if src), None)
        traffic = sum(t for t, src in parts if t is not None and src == traffic_src) if traffic_src else None
        out = {
            "metric": "Mrays/s", "value": round(value, 1), "unit": "Mrays/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD.format(w=w, h=h),
                       "rays_per_frame": int(rays_frame), "rays_are": "reference-equivalent: raytrace() calls + shadow rays of the reference's algorithm on this frame (counting launch == oracle)",
                       "rays_traced_per_frame": int(traced_frame), "rays_traced_are": "what the timed kernel traces: it starts no shadow rays for hits on a material with neither a diffuse nor a specular term (they add exactly 0)",
                       "mrays_traced_per_s": round(traced_frame / (ms_per_step * 1e-3) / 1e6, 1),
                       "parallelism": f"row tiles of {tile} rows interleaved over {world} GPU(s)" + ((", rows stored into rank 0's frame by the render kernels through CUDA-IPC peer memory (NVLink) + 4-byte all-reduce as barrier" if fused else ", NCCL send/recv gather to rank 0") if world > 1 else ""),
                       "frame_check": frame_check,
                       "l2": "flushed between timed steps (384 MB memset outside the per-step CUDA events)"},
            "e2e": e2e["e2e"],
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "fp32_fma", "achieved": round(achieved, 3), "peak": round(fp32_peak, 2), "unit": "TFLOP/s",
                         "frac": round(achieved / fp32_peak, 4), "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": "whitted_split_kernel + whitted_wall_kernel (concurrent on two streams; kernel_ms = CUDA events around one rt_whitted_launch: tile pre-pass, both render kernels, exact re-launch)", "kernel_ms": round(kern_ms, 4),
                         "algorithmic_flop_per_launch": int(flop_local),
                         "algorithmic_flop_is": "16 FLOP x sphere tests + 12 FLOP x plane tests of the REFERENCE's algorithm on this frame (SURVEY 8d); the kernel culls part of them exactly",
                         "peak_source": f"{info['sm_count']} SMs x 128 FP32 lanes x 2 FLOP x {peaks['sm_max_mhz']:.0f} MHz (sm_max_mhz of {peaks['source']}; that file has no FP32 entry)",
                         "hbm_side": {"algorithmic_bytes": int(w * h * 4 // world), "achieved_gbs": round(w * h * 4 / world / (kern_ms * 1e-3) / 1e9, 2), "peak_gbs": peaks["hbm_gbs"]}},
        }
        if "e2e_pageable" in e2e:
            out["e2e_pageable"] = e2e["e2e_pageable"]

    if world > 1:
        strong = bench_strong(rt, r, torch, dist, prims, rank, world, flush, stream, min(steps, 20))
        c5 = bench_c5(rt, r, torch, dist, rank, world, stream)
        if rank == 0:
            out["strong"] = strong
            out["c5"] = c5
    else:
        # extra blocks, N = 1 only
        out["pt"] = bench_pt(rt, r, info, peaks, torch, stream)
        threads = host_threads()
        mrays, ms, kind, sample = time_reference(2, 1, threads)
        out["cpu_baseline"] = {"value": round(mrays, 3), "unit": "Mrays/s", "cores": threads, "kind": kind, "sample": sample}
        mrays1, _, kind1, sample1 = time_reference(1, 0, 1)
        out["cpu_baseline_1core"] = {"value": round(mrays1, 3), "unit": "Mrays/s", "cores": 1, "kind": kind1, "sample": sample1,
                                     "note": "the reference's CPU path is single-threaded (R323/raytracer_non_OpenCL.c:285-450)"}
        out["config1"] = time_config1(rt, r)
        out["c4"] = bench_c4(rt, r, info, peaks, torch, stream)
    if rank == 0:
        print(json.dumps(out))
    r.close()
    if world > 1:
        dist.destroy_process_group()


def bench_strong(rt, r, torch, dist, prims, rank, world, flush, stream, steps):
    """Strong scaling beside the weak-scaling headline: FIXED frames (1920x1080 and 3840x2160) over the N GPUs, device-timed like `value`."""
    res = {}
    for (w, h) in ((1920, 1080), (3840, 2160)):
        fr = Frame(rt, r, torch, dist, prims, w, h, rank, world)
        rays, traced, _, _, _ = count_frame(r, dist, torch, world)
        ms, kern = fr.time_steps(steps, 3, flush, stream)
        check = fr.check_against_one_gpu()
        fr.close()
        if rank == 0:
            res[f"{w}x{h}"] = {"mrays_per_s": round(rays / (ms * 1e-3) / 1e6, 1), "ms_per_step": round(ms, 4), "kernel_ms_max_over_ranks": round(kern, 4),
                               "rays_per_frame": int(rays), "steps": steps, "frame_check": check, "scaling": "strong"}
    return res


def bench_c5(rt, r, torch, dist, rank, world, stream):
    """BASELINE configs[4]: cornell.scn 7680x4320 x 1024 spp on N GPUs, one launch per mode.
    image-sharded   interleaved row tiles, every pixel runs the 1-GPU instruction stream; the 8-bit frame is assembled on rank 0 by the render
                    kernels' peer stores.  Check: rows spread over the frame (every 67th tile) re-rendered by ONE GPU are byte-identical.
    sample-sharded  every rank renders the whole frame for 1024 / N passes with its own seeds as SUMS, ncclAllReduce, resolve.  Judged by RMSE
                    (SURVEY 8e): against the image-sharded running-mean image (an independent 1024-spp estimate); and, exactly, the all-reduced
                    sums of the check rows against one GPU adding the N ranks' passes one after the other (float rounding only)."""
    w, h, spp = C5_W, C5_H, C5_SPP
    spheres, cam = rt.cornell_scene(w, h)
    tile = rt.pick_tile_rows(h, world)
    n_px = w * h
    res = {"workload": f"smallpt cornell.scn {w}x{h} x {spp} spp, path tracing, {world} GPUs", "samples": n_px * spp}
    seeds = rt.reference_seeds(w, h, seed=1)
    check_world, check_rank = 67, 5                  # rows of tiles 5, 72, 139, ...: spread over the frame, owned by different ranks
    check_rows = np.array(rt.owned_rows(h, check_rank, check_world, tile))

    def dev(which, shape, typ):
        ptr, _ = r.device_buffer(which)
        return torch.as_tensor(rt.DeviceArray(ptr, shape, typ), device="cuda")

    def timed(fn):
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream)
        torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- image-sharded
    r.set_shard(rank, world, tile)
    r.pt_resize(w, h, seeds); r.pt_set_scene(spheres); r.pt_set_camera(cam)
    fused = rt.share_rank0_framebuffer(r, rt.BUF_PT_PIXELS, rank, world)
    pixels = dev(rt.BUF_PT_PIXELS, (h, w), "<i4")
    colors = dev(rt.BUF_PT_COLORS, (h, w, 3), "<f4")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    staging = rt.gather_staging(pixels, world, tile) if (rank == 0 and not fused) else None

    def image_sharded():
        r.pt_launch(0, spp)
        if fused:
            dist.all_reduce(flag)
        else:
            rt.gather_row_tiles(pixels, rank, world, tile, staging)

    ms = timed(image_sharded)
    rows_t = torch.as_tensor(check_rows, device="cuda")
    assembled_rows = pixels[rows_t].clone() if rank == 0 else None
    # the running-mean image of the whole frame on every rank (rows a rank does not own are still 0 in its buffer): the RMSE reference
    mean_image = colors.clone()
    dist.all_reduce(mean_image)
    if fused:
        torch.cuda.synchronize(); dist.barrier(); r.ipc_close(); dist.barrier()
    verdict, noise_floor = None, None
    if rank == 0:       # one GPU renders the check rows alone, all 1024 passes
        r.set_shard(check_rank, check_world, tile)
        r.pt_resize(w, h, seeds); r.pt_set_camera(cam)
        r.pt_launch(0, spp); r.sync()
        same = bool((pixels[rows_t] == assembled_rows).all())
        same_col = bool((colors[(h - 1) - rows_t].view(torch.int32) == mean_image[(h - 1) - rows_t].view(torch.int32)).all())
        verdict = f"{len(check_rows)} rows ({len(check_rows) * w} pixels, every {check_world}th tile) re-rendered by one GPU: 8-bit frame " \
                  f"{'byte-identical' if same else 'MISMATCH'}, float radiance {'bit-identical' if same_col else 'MISMATCH'}"
        # noise floor for the RMSE judgement below: the same rows once more in the reference's own mode, other seeds
        r.pt_resize(w, h, rt.reference_seeds(w, h, seed=999)); r.pt_set_camera(cam)
        r.pt_launch(0, spp); r.sync()
        d0 = colors[(h - 1) - rows_t].double() - mean_image[(h - 1) - rows_t].double()
        noise_floor = float(torch.sqrt((d0 * d0).mean()).item())
    dist.barrier()
    res["image_sharded"] = {"ms": round(ms, 2), "msamples_per_s": round(n_px * spp / ms / 1e3, 1),
                            "assembly": "render kernels store into rank 0's frame through CUDA-IPC peer memory + 4-byte all-reduce" if fused else "NCCL send/recv gather",
                            "check": verdict}

    # ---- sample-sharded: full frame per rank, spp / N passes, rank-distinct seeds, sums -> all-reduce -> resolve
    per_rank = spp // world
    r.set_shard(0, 1, tile)
    r.pt_set_accumulate_sums(True)
    my_seeds = rt.reference_seeds(w, h, seed=100 + rank)
    r.pt_resize(w, h, my_seeds); r.pt_set_camera(cam)
    flat = dev(rt.BUF_PT_COLORS, (n_px * 3,), "<f4")

    def sample_sharded():
        r.pt_launch(0, per_rank)
        rt.allreduce_sums(flat)
        r.pt_resolve_sums(per_rank * world)

    ms2 = timed(sample_sharded)
    sums = dev(rt.BUF_PT_COLORS, (h, w, 3), "<f4")
    diff = sums.double() / float(per_rank * world) - mean_image.double()
    rmse = float(torch.sqrt((diff * diff).mean()).item())
    mean_level = float(mean_image.double().mean().item())
    pix_ss = dev(rt.BUF_PT_PIXELS, (h, w), "<i4")
    verdict2, rmse_rows = None, None
    if rank == 0:
        sums_rows = sums[(h - 1) - rows_t].clone()
        d1 = sums_rows.double() / float(per_rank * world) - mean_image[(h - 1) - rows_t].double()
        rmse_rows = float(torch.sqrt((d1 * d1).mean()).item())
        # exact side of the check: ONE GPU adds the N ranks' passes of the check rows one after the other
        r.set_shard(check_rank, check_world, tile)
        for q in range(world):
            sq = rt.reference_seeds(w, h, seed=100 + q)
            if q == 0:
                r.pt_resize(w, h, sq); r.pt_set_camera(cam)
            else:                                         # keep the sums, swap the seeds, restart the pass counter
                ptr, _ = r.device_buffer(rt.BUF_PT_SEEDS)
                torch.as_tensor(rt.DeviceArray(ptr, (2 * n_px,), "<i4"), device="cuda").copy_(torch.from_numpy(sq.view(np.int32)))
                r.pt_set_camera(cam)
            r.pt_launch(0, per_rank)
        r.sync()
        seq = sums[(h - 1) - rows_t]
        rel = float(((seq - sums_rows).abs() / seq.abs().clamp_min(1e-6)).max().item())
        verdict2 = f"all-reduced sums of the check rows vs one GPU adding the {world} ranks' passes sequentially: max relative difference {rel:.2e} " \
                   f"({'within float rounding (bound 1e-5)' if rel < 1e-5 else 'TOO LARGE'})"
    r.pt_set_accumulate_sums(False)
    r.set_shard(rank, world, tile)
    dist.barrier()
    res["sample_sharded"] = {"ms": round(ms2, 2), "msamples_per_s": round(n_px * per_rank * world / ms2 / 1e3, 1), "passes_per_rank": per_rank,
                             "collective": f"ncclAllReduce(sum, f32, {3 * n_px} floats = {3 * n_px * 4 / 1e6:.0f} MB) + resolve",
                             "rmse_vs_running_mean_image": round(rmse, 6), "mean_radiance": round(mean_level, 4), "check": verdict2}
    if rank == 0:
        ok = rmse_rows <= 1.25 * noise_floor
        res["sample_sharded"]["rmse_judgement"] = {
            "rows": f"the {len(check_rows)} check rows", "sample_sharded_vs_running_mean": round(rmse_rows, 6),
            "running_mean_other_seeds_vs_running_mean": round(noise_floor, 6),
            "bound": "the sample-sharded image may differ from the reference-mode image by no more than 1.25 x what a second reference-mode render with other seeds "
                     f"does (both are independent {spp}-spp estimates; Monte-Carlo noise, heavy-tailed through the glass sphere): {'within' if ok else 'EXCEEDED'}"}
    return res


def time_config1(rt=None, r=None):
    """BASELINE configs[0]: the reference's own CPU render (raytracer3.0.06: Engine_InitRender + Engine_Render, unmodified,
    oracle/_ref/libref_r306.so) of its built-in scene at 800x600, one frame, one core (its engine keeps state in globals),
    next to the same frame from rt_r306_render on the GPU (bit-identical, tests/test_gpu_parity.py)."""
    out = {"workload": "raytracer3.0.06 Whitted render of its built-in scene, 800x600 (rows 20..529), 3x3 AA, 63-node ray tree, 1 frame"}
    frame = None
    path = os.path.join(ORACLE_DIR, "_ref", "libref_r306.so")
    if os.path.exists(path):
        lib = ctypes.CDLL(path)
        frame = np.zeros((600, 800), np.uint32)
        t0 = time.perf_counter()
        lib.ref_r306_render(vp(frame), 800, 600)
        out["cpu"] = {"seconds_per_frame": round(time.perf_counter() - t0, 3), "cores": 1, "kind": "reference",
                      "rows_rendered": int(frame.any(axis=1).sum())}
    else:
        out["cpu"] = {"unavailable": "oracle/_ref/libref_r306.so not shipped"}
    if r is not None:
        prims = rt.r306_create_scene()
        r.r306_upload(prims, 800, 600)
        for _ in range(3):
            r.r306_launch()
        r.sync()
        ts = []
        for _ in range(5):
            r.timer_begin(); r.r306_launch(); ts.append(r.timer_end())
        t0 = time.perf_counter()
        img = r.r306_render(prims, 800, 600)
        e2e = time.perf_counter() - t0
        out["gpu"] = {"kernel_ms": round(min(ts), 3), "e2e_ms": round(e2e * 1e3, 3),
                      "equals_cpu_frame": bool((img == frame).all()) if frame is not None else None}
        if frame is not None:
            out["gpu"]["speedup_vs_cpu_1core_e2e"] = round(out["cpu"]["seconds_per_frame"] / e2e, 1)
    return out


def bench_pt(rt, r, info, peaks, torch, stream):
    spheres, cam = rt.cornell_scene(PT_W, PT_H)
    seeds = rt.reference_seeds(PT_W, PT_H, seed=1)
    res = {}
    for integ, tag in [(0, "path_tracing"), (1, "direct_lighting")]:
        r.pt_resize(PT_W, PT_H, seeds); r.pt_set_scene(spheres); r.pt_set_camera(cam)
        r.set_counting(True)
        r.pt_launch(integ, 8)
        c = r.counters()
        r.set_counting(False)
        per = {k: v / c["samples"] for k, v in c.items()}
        times = []
        for it in range(4):
            r.pt_resize(PT_W, PT_H, seeds); r.pt_set_camera(cam)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            r.pt_launch(integ, PT_SPP)
            b.record(stream)
            torch.cuda.synchronize()
            times.append(a.elapsed_time(b))
        ms = min(times[1:])
        samples = PT_W * PT_H * PT_SPP
        flop = 17.0 * per["sphere_tests"] * samples
        fp32_peak = info["sm_count"] * 128 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12
        seeds_pin = torch.from_numpy(seeds.view(np.int32)).pin_memory().numpy().view(np.uint32)
        pix_pin = torch.empty((PT_H, PT_W), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
        t0 = time.perf_counter()
        r.pt_resize(PT_W, PT_H, seeds_pin); r.pt_set_scene(spheres); r.pt_set_camera(cam)
        r.pt_render(integ, PT_SPP, want=("pixels",), pixels_out=pix_pin)
        e2e_s = time.perf_counter() - t0
        res[tag] = {"workload": f"smallpt cornell.scn {PT_W}x{PT_H} x {PT_SPP} spp, one launch", "msamples_per_s": round(samples / ms / 1e3, 1),
                    "mrays_per_s": round(samples * (per["nearest_queries"] + per["shadow_queries"]) / ms / 1e3, 1), "kernel_ms": round(ms, 3),
                    "rays_per_sample": round(per["nearest_queries"] + per["shadow_queries"], 3), "sphere_tests_per_sample": round(per["sphere_tests"], 2),
                    "fp32_tflops_algorithmic": round(flop / ms / 1e9, 3), "frac_of_fp32_peak": round(flop / ms / 1e9 / fp32_peak, 4),
                    "e2e_msamples_per_s": round(samples / e2e_s / 1e6, 1), "e2e_includes": "rt_pt_resize (seed upload 6.3 MB) + rt_pt_set_scene + rt_pt_set_camera + rt_pt_render (kernel + pixel read-back 3.1 MB), pinned host buffers"}
    # CPU side: the reference's own RadiancePathTracing (oracle/_ref) on a bounded sample, all cores and one core
    cam_fn = lambda sw, sh: rt.cornell_scene(sw, sh)[1]
    res["cpu_baseline"] = cpu_pt_baseline(spheres, cam_fn, 256, 192, 8, host_threads(), "cornell")
    res["cpu_baseline_1core"] = cpu_pt_baseline(spheres, cam_fn, 128, 96, 8, 1, "cornell")
    res["cpu_baseline_1core"]["note"] = "smallptCPU is single-threaded (SPT/smallptCPU.cpp:77-132)"
    return res


def bench_c4(rt, r, info, peaks, torch, stream):
    """BASELINE configs[3]: the scene_build_complex.pl scene ($maxDepth 4 = 783 spheres) at 3840x2160 x 16 spp, path tracing:
    the reference-order loop over every sphere (shared-memory staging) and the exact hierarchy (csrc/pt_bvh.cuh), which gives
    the same bits; the reference's own CPU code on a bounded sample next to them."""
    import tempfile
    w, h, spp = 3840, 2160, 16
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "complex4.scn")
        rt.write_complex_scene(path, 4)
        spheres, cam = rt.read_scene(path, w, h)
        cam_fn = lambda sw, sh, path=path: rt.read_scene(path, sw, sh)[1]
        cams = {(384, 216): cam_fn(384, 216), (128, 72): cam_fn(128, 72)}
    seeds = rt.reference_seeds(w, h, seed=1)
    r.pt_resize(w, h, seeds); r.pt_set_scene(spheres); r.pt_set_camera(cam)
    r.set_counting(True); r.pt_launch(0, 1); c = r.counters(); r.set_counting(False)
    per = {k: v / c["samples"] for k, v in c.items()}
    samples = w * h * spp
    flop = 17.0 * per["sphere_tests"] * samples
    fp32_peak = info["sm_count"] * 128 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12
    res = {"workload": f"scene_build_complex.pl $maxDepth 4: {spheres.size} spheres, {w}x{h} x {spp} spp, path tracing, one launch",
           "reference_sphere_tests_per_sample": round(per["sphere_tests"], 1), "rays_per_sample": round(per["nearest_queries"] + per["shadow_queries"], 3)}
    pixels = {}
    for mode, tag in ((0, "loop_over_every_sphere"), (1, "exact_hierarchy")):
        r.set_tuning(rt.TUNE_PT_BVH, mode)
        times = []
        for it in range(3):
            r.pt_resize(w, h, seeds); r.pt_set_camera(cam)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); r.pt_launch(0, spp); b.record(stream)
            torch.cuda.synchronize()
            times.append(a.elapsed_time(b))
        ms = min(times[1:])
        pixels[tag] = r.pt_download(want=("colors",))["colors"].copy()
        e2e = {}
        for tag_h, sd_h, px_h in (("pageable", seeds, None), ("pinned", torch.from_numpy(seeds.view(np.int32)).pin_memory().numpy().view(np.uint32),
                                                             torch.empty((h, w), dtype=torch.int32).pin_memory().numpy().view(np.uint32))):
            t0 = time.perf_counter()                  # end to end: seeds up, scene (an unchanged table keeps its hierarchy), kernel, pixels down
            r.pt_resize(w, h, sd_h); r.pt_set_scene(spheres); r.pt_set_camera(cam)
            r.pt_render(0, spp, want=("pixels",), pixels_out=px_h)
            e2e[tag_h] = time.perf_counter() - t0
        changed = spheres.copy(); changed["c"][0] *= np.float32(0.5)      # a CHANGED table: tables re-uploaded, hierarchy rebuilt on the host
        t0 = time.perf_counter()
        r.pt_set_scene(changed); r.pt_launch(0, 1); r.sync()
        t_new_scene = time.perf_counter() - t0
        r.pt_set_scene(spheres)
        e2e_s = e2e["pinned"]
        res[tag] = {"kernel_ms": round(ms, 2), "msamples_per_s": round(samples / ms / 1e3, 1),
                    "mrays_per_s": round(samples * (per["nearest_queries"] + per["shadow_queries"]) / ms / 1e3, 1),
                    "e2e_ms": round(e2e_s * 1e3, 1), "e2e_pageable_ms": round(e2e["pageable"] * 1e3, 1),
                    "e2e_includes": "rt_pt_resize (seed upload 66 MB) + rt_pt_set_scene (same table: tables and hierarchy kept) + rt_pt_render (kernel, pixel read-back 33 MB); pinned / pageable host buffers",
                    "new_scene_plus_1spp_ms": round(t_new_scene * 1e3, 1)}
        if mode == 0:
            res[tag]["fp32_tflops_algorithmic"] = round(flop / ms / 1e9, 2)
            res[tag]["frac_of_fp32_peak"] = round(flop / ms / 1e9 / fp32_peak, 4)
    r.set_tuning(rt.TUNE_PT_BVH, -1)
    res["hierarchy_equals_loop"] = bool(np.array_equal(pixels["loop_over_every_sphere"].view(np.uint32), pixels["exact_hierarchy"].view(np.uint32)))
    res["cpu_baseline"] = cpu_pt_baseline(spheres, lambda sw, sh: cams[(sw, sh)], 384, 216, 4, host_threads(), "the same scene")
    res["cpu_baseline_1core"] = cpu_pt_baseline(spheres, lambda sw, sh: cams[(sw, sh)], 128, 72, 4, 1, "the same scene")
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus:
        if args.gpus > 1:
            sys.exit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
