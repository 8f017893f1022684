"""BASELINE config 5 under torchrun: cornell.scn 7680x4320, path tracing, N GPUs, both sharding modes.
  image-sharded : interleaved row tiles, per-pixel seeds global (bit-identical to 1 GPU), pixels stored into rank 0's
                  frame through CUDA-IPC peer memory (fallback: NCCL gather);
  sample-sharded: every rank renders the whole frame for spp/N passes with its own seeds into SUMS, ncclAllReduce of
                  the float accumulation buffers, resolve (RMSE-judged mode).
Usage: torchrun --nproc-per-node N tools/bench_c5_multigpu.py [spp=64] [w=7680] [h=4320]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import __graft_entry__ as g

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
w = int(sys.argv[2]) if len(sys.argv) > 2 else 7680
h = int(sys.argv[3]) if len(sys.argv) > 3 else 4320
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rt = g.load()
r = rt.Renderer(local)
stream = torch.cuda.current_stream()
r.set_stream(stream.cuda_stream)
spheres, cam = rt.cornell_scene(w, h)
tile = rt.pick_tile_rows(h, world)
flag = torch.zeros(1, dtype=torch.int32, device="cuda")


def timed(fn):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream); fn(); b.record(stream)
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---- image-sharded
seeds = rt.reference_seeds(w, h, seed=1)
r.set_shard(rank, world, tile)
r.pt_resize(w, h, seeds); r.pt_set_scene(spheres); r.pt_set_camera(cam)
fused = rt.share_rank0_framebuffer(r, rt.BUF_PT_PIXELS, rank, world)
ptr, _ = r.device_buffer(rt.BUF_PT_PIXELS)
fb = torch.as_tensor(rt.DeviceArray(ptr, (h, w), "<i4"), device="cuda")


def image_sharded():
    r.pt_launch(0, spp)
    if world > 1:
        if fused:
            dist.all_reduce(flag)
        else:
            rt.gather_row_tiles(fb, rank, world, tile)


r.pt_launch(0, 1); r.pt_set_camera(cam)        # warm-up, then reset the sample counter (seeds advance; irrelevant for timing)
ms = timed(image_sharded)
if rank == 0:
    covered = float((fb != 0).float().mean().item())
    print(f"C5 image-sharded  {world} GPU(s) {w}x{h} x {spp} spp: {ms:.1f} ms  {w * h * spp / ms / 1e3:.0f} Msamples/s"
          f"  ({'fused IPC peer stores' if fused else 'NCCL gather'}; non-black pixels on rank 0: {100 * covered:.1f} %)", flush=True)
r.ipc_close()

# ---- sample-sharded
r.set_shard(0, 1, 8)
r.pt_set_accumulate_sums(True)
r.pt_resize(w, h, rt.reference_seeds(w, h, seed=100 + rank)); r.pt_set_camera(cam)
cptr, _ = r.device_buffer(rt.BUF_PT_COLORS)
colors = torch.as_tensor(rt.DeviceArray(cptr, (h * w * 3,), "<f4"), device="cuda")
per_rank = max(1, spp // world)


def sample_sharded():
    r.pt_launch(0, per_rank)
    if world > 1:
        rt.allreduce_sums(colors)
    r.pt_resolve_sums(per_rank * world)


ms = timed(sample_sharded)
if rank == 0:
    print(f"C5 sample-sharded {world} GPU(s) {w}x{h} x {per_rank * world} spp ({per_rank}/rank): {ms:.1f} ms  "
          f"{w * h * per_rank * world / ms / 1e3:.0f} Msamples/s  (ncclAllReduce of {colors.numel() * 4 / 1e6:.0f} MB of float sums + resolve)", flush=True)
r.pt_set_accumulate_sums(False)
r.close()
if world > 1:
    dist.destroy_process_group()
