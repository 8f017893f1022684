"""The N > 1 path on CPU: two gloo ranks run the row-tile sharding, the gather to rank 0 and the
sample-sharded all-reduce, with the host build of the lane code standing in for the kernels."""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402


def vp(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rt = graft.load()
    dev = ctypes.CDLL(os.path.join(graft.DEVSIM_DIR, "libdevsim.so"))
    # --- image-sharded Whitted frame, gathered to rank 0
    prims = rt.whitted_create_scene(0)
    w, h = 64, 44                                   # 44 rows: 5 full tiles of 8 + a ragged one; uneven over 2 ranks
    tile = 8
    frame = np.zeros((h, w, 4), np.uint8)
    dev.devsim_whitted(vp(frame), None, w, h, vp(prims), prims.size, rank, world, tile, None, None, 1)
    mine = rt.owned_rows(h, rank, world, tile)
    others = [y for y in range(h) if y not in mine]
    assert frame[mine].any() and not frame[others].any()          # a rank touches its own rows only
    t = torch.from_numpy(frame)
    rt.gather_row_tiles(t, rank, world, tile)
    # --- image-sharded path tracing: colors gathered the same way (float rows), flipped index handled by the caller
    sph, cam = rt.cornell_scene(32, 24)
    seeds = rt.reference_seeds(32, 24, seed=4)
    col, sd, pix = np.zeros(3 * 32 * 24, np.float32), seeds.copy(), np.zeros((24, 32), np.uint32)
    dev.devsim_pt(0, vp(sph), sph.size, vp(cam), 32, 24, 0, 3, 0, vp(col), vp(sd), vp(pix), rank, world, 4, None, 1 << 30)
    tp = torch.from_numpy(pix)
    rt.gather_row_tiles(tp, rank, world, 4)
    # --- sample-sharded: every rank renders the full frame with its own seeds into SUMS, then all-reduce
    seeds_r = rt.reference_seeds(32, 24, seed=100 + rank)
    sums = np.zeros(3 * 32 * 24, np.float32)
    dev.devsim_pt(0, vp(sph), sph.size, vp(cam), 32, 24, 0, 4, 1, vp(sums), vp(seeds_r.copy()), None, 0, 1, 8, None, 1 << 30)
    local = sums.copy()
    ts = torch.from_numpy(sums)
    rt.allreduce_sums(ts)
    np.save(os.path.join(out_dir, f"local_{rank}.npy"), local)
    if rank == 0:
        np.save(os.path.join(out_dir, "whitted.npy"), frame)
        np.save(os.path.join(out_dir, "pt_pixels.npy"), pix)
        np.save(os.path.join(out_dir, "sums.npy"), sums)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_gather_and_allreduce(tmp_path):
    if not os.path.exists(os.path.join(graft.DEVSIM_DIR, "libdevsim.so")):
        graft.build()
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rt = graft.load()
    orc = graft.oracle()
    # gathered Whitted frame == unsharded oracle frame
    prims = rt.whitted_create_scene(0)
    ref = np.zeros((44, 64, 4), np.uint8)
    orc.oracle_whitted_render(vp(ref), None, 64, 44, vp(prims), prims.size, 2, None)
    assert np.array_equal(np.load(tmp_path / "whitted.npy"), ref)
    # gathered path-traced pixels == unsharded oracle pixels (same seeds, global indices)
    sph, cam = rt.cornell_scene(32, 24)
    seeds = rt.reference_seeds(32, 24, seed=4)
    col, sd, pix = np.zeros(3 * 32 * 24, np.float32), seeds.copy(), np.zeros(32 * 24, np.uint32)
    orc.oracle_pt_render(0, vp(sph), sph.size, vp(cam), 32, 24, 0, 3, vp(col), vp(sd), vp(pix), 2, None)
    assert np.array_equal(np.load(tmp_path / "pt_pixels.npy").reshape(-1), pix)
    # all-reduced sums == sum of the two ranks' buffers; the mean is an 8-spp estimate of the same image
    l0, l1 = np.load(tmp_path / "local_0.npy"), np.load(tmp_path / "local_1.npy")
    assert np.array_equal(np.load(tmp_path / "sums.npy"), l0 + l1)
    assert not np.array_equal(l0, l1)


def test_tile_choice_and_ownership():
    rt = graft.load()
    assert rt.pick_tile_rows(1080, 1) == 8 and rt.pick_tile_rows(2160, 2) == 8
    assert rt.pick_tile_rows(2160, 4) == 4 and rt.pick_tile_rows(4320, 8) == 4
    for h, world, tile in [(44, 2, 8), (1080, 8, 8), (7, 3, 2)]:
        seen = sorted(y for q in range(world) for y in rt.owned_rows(h, q, world, tile))
        assert seen == list(range(h))
