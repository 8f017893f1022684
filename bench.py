#!/usr/bin/env python3
"""bench.py -- the measurement contract.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): the Raytracer3.2.03 Whitted scene (CHOOSE_SCENE 0, 17 primitive
slots), 3x3 super-sampling, TRACEDEPTH 5, at 1920x1080 on one B200.  A "step" is one frame.  With N > 1
GPUs the frame grows with N (weak scaling: 1920x1080 pixels per GPU; the Whitted view window is fixed, so
this is the same picture at a higher resolution), is sharded by interleaved row tiles
(rt_set_shard) and is gathered to rank 0 over NVLink (NCCL send/recv) inside the step.

Metric: Mrays/s, rays = raytrace() calls + shadow rays, counted by a separate counting launch (the counts
are a property of scene and resolution, identical in the oracle -- tests/test_gpu_parity.py).

  value     frames resident in HBM: K x (kernel [+ NVLink gather]) timed with CUDA events on the stream the
            kernels run on (torch's current stream, injected with rt_set_stream), L2 flushed between steps
            outside the events, max over ranks.
  e2e       the reference-facing call rt_whitted_render(ctx, prims, n, w, h, pixels, NULL) with HOST buffers
            (scene upload + kernel + read-back of the frame into pinned host memory) timed by the host clock.
  roofline  FP32-FMA bound: algorithmic FLOP = 16 x sphere tests + 12 x plane tests per frame,
            divided by the kernel's mean duration, against 148 SM x 128 lanes x 2 x f_max.
  cpu_baseline / --impl reference
            the reference's OWN CPU code (oracle/_ref: raytracer_non_OpenCL.c compiled unmodified) on the
            box's host cores, one frame per thread on a bounded sample.

Extra (not the headline): "pt" = BASELINE configs[2], smallpt cornell.scn 1024x768 x 256 spp, Msamples/s;
"c4" = BASELINE configs[3], the 783-sphere generated scene at 3840x2160 x 16 spp (loop over every sphere and exact hierarchy).
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
import __graft_entry__ as graft  # noqa: E402

BASE_W, BASE_H = 1920, 1080
SAMPLE_W, SAMPLE_H = 480, 270          # CPU sample frame: 1/16 of the pixels, same rays per pixel
PT_W, PT_H, PT_SPP = 1024, 768, 256
WEAK_SIZES = {1: (1920, 1080), 2: (1920, 2160), 4: (3840, 2160), 8: (3840, 4320)}


def vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


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
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full` capture
    (profiles/r01_ncu_full_summary.json, made by tools/ncu_summary.py from the 1920x1080 frame); None when it is missing."""
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def to_bytes(text):
        v, u = text.split()
        return float(v) * unit[u]

    try:
        for m in json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_full_summary.json"))):
            if kernel in m.get("Kernel Name", "") and "dram__bytes_read.sum" in m:
                return int(to_bytes(m["dram__bytes_read.sum"]) + to_bytes(m["dram__bytes_write.sum"]))
    except (OSError, ValueError, KeyError, TypeError, AttributeError):
        pass
    return None


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


# --------------------------------------------------------------------------------------------- reference arm
def rays_per_frame_cpu(orc, prims, w, h):
    ctr = np.zeros(5, np.uint64)
    px = np.zeros((h, w, 4), np.uint8)
    orc.oracle_whitted_render(vp(px), None, w, h, vp(prims), prims.size, host_threads(), vp(ctr))
    return int(ctr[0] + ctr[1])


def time_reference(rt, steps, warmup, threads):
    """K steps of: `threads` host threads each render one SAMPLE_W x SAMPLE_H frame with the reference's own
    raytracer_non_kernel (oracle/_ref).  Returns (Mrays/s, ms per step, description)."""
    ref_path = os.path.join(graft.ORACLE_DIR, "_ref", "libref_whitted.so")
    prims = rt.whitted_create_scene(0)
    orc = graft.oracle()
    rays = rays_per_frame_cpu(orc, prims, SAMPLE_W, SAMPLE_H)
    px = np.zeros((threads, SAMPLE_H, SAMPLE_W, 4), np.uint8)
    if os.path.exists(ref_path):
        ref = ctypes.CDLL(ref_path)
        run = lambda: ref.ref_whitted_render_mt(vp(px), SAMPLE_W, SAMPLE_H, vp(prims), prims.size, threads)
        kind = "reference"
    else:       # the oracle port (only if oracle/_ref was not shipped)
        run = lambda: [orc.oracle_whitted_render(vp(px[0]), None, SAMPLE_W, SAMPLE_H, vp(prims), prims.size, threads, None) for _ in range(threads)]
        kind = "port"
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = time.perf_counter() - t0
    mrays = rays * threads * steps / dt / 1e6
    sample = (f"{threads} host threads x {steps} steps, each thread one {SAMPLE_W}x{SAMPLE_H} frame of the config-2 scene "
              f"({rays / (SAMPLE_W * SAMPLE_H):.1f} rays/pixel, same as 1080p) with the reference's raytracer_non_kernel, g++ -O2")
    return mrays, dt / steps * 1e3, kind, sample


def run_reference(args, rank):
    if rank != 0:
        return
    rt = graft.load()
    threads = host_threads()
    steps = max(1, min(args.steps, 8))          # bounded: a step is ~1 s of wall clock on every core
    warmup = max(1, min(args.warmup, 1))
    mrays, ms, kind, sample = time_reference(rt, steps, warmup, threads)
    w, h = WEAK_SIZES.get(args.gpus, WEAK_SIZES[1])
    print(json.dumps({
        "impl": "reference", "metric": "Mrays/s", "value": round(mrays, 3), "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"Raytracer3.2.03 Whitted scene (17 primitive slots, 3x3 AA, depth 5) {w}x{h}; CPU arm timed on a bounded sample",
                   "parallelism": f"{threads} host threads"},
        "cpu_baseline": {"value": round(mrays, 3), "unit": "Mrays/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": round(mrays, 3), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    rt = graft.load()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    r = rt.Renderer(local_rank)
    stream = torch.cuda.current_stream()
    r.set_stream(stream.cuda_stream)
    info = r.device_info()
    peaks = measured_peaks()
    w, h = WEAK_SIZES.get(world, (BASE_W, BASE_H * world))
    tile = rt.pick_tile_rows(h, world)
    prims = rt.whitted_create_scene(0)
    r.set_shard(rank, world, tile)

    # ---- work per frame (counting launch, not timed)
    r.set_counting(True)
    r.whitted_upload(prims, w, h)
    r.whitted_launch()
    cnt = r.counters()
    r.set_counting(False)
    local = torch.tensor([cnt["nearest_queries"] + cnt["shadow_queries"], cnt["sphere_tests"], cnt["plane_tests"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(local)
    rays_frame, sph_tests, pl_tests = (float(v) for v in local.tolist())
    flop_frame = 16.0 * sph_tests + 12.0 * pl_tests
    flop_local = 16.0 * cnt["sphere_tests"] + 12.0 * cnt["plane_tests"]

    # ---- device-resident steps
    ptr, nbytes = r.device_buffer(rt.BUF_WHITTED_PIXELS)
    fb = torch.as_tensor(rt.DeviceArray(ptr, (h, w), "<i4"), device="cuda")      # the context's framebuffer, as a torch view
    flush = torch.empty(384 * 1024 * 1024, dtype=torch.uint8, device="cuda")       # > 126 MB L2
    # Frame assembly on rank 0.  Preferred: fused -- the render kernels of ranks 1.. store their rows straight into
    # rank 0's framebuffer through a CUDA-IPC peer mapping (NVLink), and one tiny all-reduce orders "all kernels done".
    # Fallback if the mapping cannot be made: NCCL send/recv of each rank's rows + strided de-interleave.
    fused = rt.share_rank0_framebuffer(r, rt.BUF_WHITTED_PIXELS, rank, world)
    staging = rt.gather_staging(fb, world, tile) if (rank == 0 and world > 1 and not fused) else None
    done_flag = torch.zeros(1, dtype=torch.int32, device="cuda")

    def gather():
        if world == 1:
            return
        if fused:
            dist.all_reduce(done_flag)
        else:
            rt.gather_row_tiles(fb, rank, world, tile, staging)

    frame_check = None
    if world > 1:           # the assembled frame equals a 1-GPU render of the same frame (checked once, untimed)
        fb.zero_()
        torch.cuda.synchronize(); dist.barrier()
        r.whitted_launch(); gather()
        torch.cuda.synchronize(); dist.barrier()
        if rank == 0:
            assembled = fb.cpu().numpy().copy()
            r.set_shard(0, 1, tile)
            r.whitted_launch(); r.sync()
            frame_check = "bit-identical to the 1-GPU frame" if (fb.cpu().numpy() == assembled).all() else "MISMATCH"
            r.set_shard(rank, world, tile)
        dist.barrier()

    def step():
        r.whitted_launch()
        gather()

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.6 if sampler else 0.0)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = r.launch_count()
    if sampler:
        sampler.begin()
    for a, k, b in ev:
        flush.zero_()                      # L2 flush, outside the events
        a.record(stream)
        r.whitted_launch()
        k.record(stream)
        gather()
        b.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.end() if sampler else None
    launches = r.launch_count() - launches0
    step_ms = sum(a.elapsed_time(b) for a, _, b in ev)
    kern_ms = sum(a.elapsed_time(k) for a, k, _ in ev) / args.steps
    t = torch.tensor([step_ms, kern_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, kern_ms_max = (float(v) for v in t.tolist())
    ms_per_step = step_ms / args.steps
    value = rays_frame / (ms_per_step * 1e-3) / 1e6

    # ---- e2e through the reference-facing call with host buffers
    pinned = torch.empty((h, w, 4), dtype=torch.uint8).pin_memory()
    pinned_np = pinned.numpy()
    scene_bytes = prims.size * (3 * 16 + 4 + 4) + 4 * int(prims["is_light"].sum())

    def e2e_step():
        if world == 1:
            r.whitted_render(prims, w, h, pixels_out=pinned_np)
        else:
            r.whitted_upload(prims, w, h)
            r.whitted_launch()
            gather()
            if rank == 0:
                r.whitted_download(pixels_out=pinned_np)
            else:
                r.sync()

    for _ in range(3):
        e2e_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = rays_frame * args.steps / float(e2e_s.item()) / 1e6

    out = None
    if rank == 0:
        fp32_peak = info["sm_count"] * 128 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12       # TFLOP/s
        achieved = flop_local / (kern_ms_max * 1e-3) / 1e12
        out = {
            "metric": "Mrays/s", "value": round(value, 1), "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"Raytracer3.2.03 Whitted scene (CHOOSE_SCENE 0, 17 primitive slots, 3x3 AA, TRACEDEPTH 5) {w}x{h}",
                       "rays_per_frame": int(rays_frame), "parallelism": f"row tiles of {tile} rows interleaved over {world} GPU(s)" + ((", rows stored into rank 0's frame by the render kernels through CUDA-IPC peer memory (NVLink) + 4-byte all-reduce as barrier" if fused else ", NCCL send/recv gather to rank 0") if world > 1 else ""),
                       "frame_check": frame_check,
                       "l2": "flushed between timed steps (384 MB memset outside the per-step CUDA events)"},
            "e2e": {"value": round(e2e_value, 1), "unit": "Mrays/s", "h2d_bytes_per_step": int(scene_bytes), "d2h_bytes_per_step": int(w * h * 4)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "fp32_fma", "achieved": round(achieved, 3), "peak": round(fp32_peak, 2), "unit": "TFLOP/s",
                         "frac": round(achieved / fp32_peak, 4), "traffic": ncu_traffic("whitted_kernel"),
                         "kernel": "whitted_kernel<false>", "kernel_ms": round(kern_ms_max, 4),
                         "algorithmic_flop_per_launch": int(flop_local),
                         "peak_source": f"{info['sm_count']} SMs x 128 FP32 lanes x 2 FLOP x {peaks['sm_max_mhz']:.0f} MHz (sm_max_mhz of {peaks['source']}; that file has no FP32 entry)",
                         "hbm_side": {"algorithmic_bytes": int(w * h * 4 // world), "achieved_gbs": round(w * h * 4 / world / (kern_ms_max * 1e-3) / 1e9, 2), "peak_gbs": peaks["hbm_gbs"]}},
        }

    # ---- extra: smallpt cornell 1024x768 x 256 spp (BASELINE configs[2]); N = 1 only
    if world == 1:
        out["pt"] = bench_pt(rt, r, info, peaks, torch, stream)
        threads = host_threads()
        mrays, ms, kind, sample = time_reference(rt, 2, 1, threads)
        out["cpu_baseline"] = {"value": round(mrays, 3), "unit": "Mrays/s", "cores": threads, "kind": kind, "sample": sample}
        out["config1"] = time_config1(rt, r)
        out["c4"] = bench_c4(rt, r, info, peaks, torch, stream)
    if rank == 0:
        print(json.dumps(out))
    r.close()
    if world > 1:
        dist.destroy_process_group()


def time_config1(rt=None, r=None):
    """BASELINE configs[0]: the reference's own CPU render (raytracer3.0.06: Engine_InitRender + Engine_Render, unmodified,
    oracle/_ref/libref_r306.so) of its built-in scene at 800x600, one frame, one core (its engine keeps state in globals),
    next to the same frame from rt_r306_render on the GPU (bit-identical, tests/test_gpu_parity.py)."""
    out = {"workload": "raytracer3.0.06 Whitted render of its built-in scene, 800x600 (rows 20..529), 3x3 AA, 63-node ray tree, 1 frame"}
    frame = None
    path = os.path.join(graft.ORACLE_DIR, "_ref", "libref_r306.so")
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


def cornell_scene(rt, w, h):
    return rt.cornell_scene(w, h)


def bench_pt(rt, r, info, peaks, torch, stream):
    spheres, cam = cornell_scene(rt, PT_W, PT_H)
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
        t0 = time.perf_counter()
        r.pt_resize(PT_W, PT_H, seeds); r.pt_set_camera(cam)
        out = r.pt_render(integ, PT_SPP, want=("pixels",))
        e2e_s = time.perf_counter() - t0
        res[tag] = {"workload": f"smallpt cornell.scn {PT_W}x{PT_H} x {PT_SPP} spp, one launch", "msamples_per_s": round(samples / ms / 1e3, 1),
                    "mrays_per_s": round(samples * (per["nearest_queries"] + per["shadow_queries"]) / ms / 1e3, 1), "kernel_ms": round(ms, 3),
                    "rays_per_sample": round(per["nearest_queries"] + per["shadow_queries"], 3), "sphere_tests_per_sample": round(per["sphere_tests"], 2),
                    "fp32_tflops_algorithmic": round(flop / ms / 1e9, 3), "frac_of_fp32_peak": round(flop / ms / 1e9 / fp32_peak, 4),
                    "e2e_msamples_per_s": round(samples / e2e_s / 1e6, 1), "e2e_includes": "seed upload 6.3 MB + kernel + pixel read-back 3.1 MB"}
    # CPU side: the reference's own RadiancePathTracing (oracle/_ref) on a bounded sample
    ref_path = os.path.join(graft.ORACLE_DIR, "_ref", "libref_smallpt.so")
    threads = host_threads()
    sw, sh, sp = 256, 192, 8
    sph, cam_s = cornell_scene(rt, sw, sh)
    sd = rt.reference_seeds(sw, sh, seed=1)
    col = np.zeros(3 * sw * sh, np.float32)
    if os.path.exists(ref_path):
        ref = ctypes.CDLL(ref_path)
        ref.ref_pt_set_scene(vp(sph), sph.size, vp(cam_s), sw, sh)
        t0 = time.perf_counter()
        ref.ref_pt_render_mt(0, 0, sp, vp(col), vp(sd), None, threads)
        kind = "reference"
    else:
        orc = graft.oracle()
        t0 = time.perf_counter()
        orc.oracle_pt_render(0, vp(sph), sph.size, vp(cam_s), sw, sh, 0, sp, vp(col), vp(sd), None, threads, None)
        kind = "port"
    dt = time.perf_counter() - t0
    res["cpu_baseline"] = {"value": round(sw * sh * sp / dt / 1e6, 3), "unit": "Msamples/s", "cores": threads, "kind": kind,
                           "sample": f"cornell {sw}x{sh} x {sp} spp, rows split over {threads} host threads, reference RadiancePathTracing"}
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
        sw, sh, sp = 384, 216, 4
        sph_s, cam_s = rt.read_scene(path, sw, sh)
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
        t0 = time.perf_counter()                      # end to end: scene tables (+ hierarchy build), seeds up, kernel, pixels down
        r.pt_resize(w, h, seeds); r.pt_set_scene(spheres); r.pt_set_camera(cam)
        r.pt_render(0, spp, want=("pixels",))
        e2e_s = time.perf_counter() - t0
        res[tag] = {"kernel_ms": round(ms, 2), "msamples_per_s": round(samples / ms / 1e3, 1),
                    "mrays_per_s": round(samples * (per["nearest_queries"] + per["shadow_queries"]) / ms / 1e3, 1),
                    "e2e_ms": round(e2e_s * 1e3, 1), "e2e_includes": "scene upload (and the host build of the hierarchy), seed upload 66 MB, kernel, pixel read-back 33 MB"}
        if mode == 0:
            res[tag]["fp32_tflops_algorithmic"] = round(flop / ms / 1e9, 2)
            res[tag]["frac_of_fp32_peak"] = round(flop / ms / 1e9 / fp32_peak, 4)
    r.set_tuning(rt.TUNE_PT_BVH, -1)
    res["hierarchy_equals_loop"] = bool(np.array_equal(pixels["loop_over_every_sphere"].view(np.uint32), pixels["exact_hierarchy"].view(np.uint32)))
    ref_path = os.path.join(graft.ORACLE_DIR, "_ref", "libref_smallpt.so")
    threads = host_threads()
    sd = rt.reference_seeds(sw, sh, seed=1)
    col = np.zeros(3 * sw * sh, np.float32)
    if os.path.exists(ref_path):
        ref = ctypes.CDLL(ref_path)
        ref.ref_pt_set_scene(vp(sph_s), sph_s.size, vp(cam_s), sw, sh)
        t0 = time.perf_counter()
        ref.ref_pt_render_mt(0, 0, sp, vp(col), vp(sd), None, threads)
        kind = "reference"
    else:
        orc = graft.oracle()
        t0 = time.perf_counter()
        orc.oracle_pt_render(0, vp(sph_s), sph_s.size, vp(cam_s), sw, sh, 0, sp, vp(col), vp(sd), None, threads, None)
        kind = "port"
    dt = time.perf_counter() - t0
    res["cpu_baseline"] = {"value": round(sw * sh * sp / dt / 1e6, 3), "unit": "Msamples/s", "cores": threads, "kind": kind,
                           "sample": f"the same scene {sw}x{sh} x {sp} spp, rows split over {threads} host threads, reference RadiancePathTracing"}
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
