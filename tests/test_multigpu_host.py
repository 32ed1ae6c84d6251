"""Host-side logic of the multi-GPU paths, exercised on CPU with gloo and world_size 2:
frames of a flythrough dealt to ranks and gathered to rank 0 (bench.py's N>1 path), and the
scanline-band sharding of one frame (XnartFrame.shardIndex/shardCount).  The renderer behind the
shard API here is the oracle — this file tests the partition/gather plumbing, not the kernels."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from xna_ray_trace_b200 import scenes
    from xna_ray_trace_b200.parallel import band_rows, frame_owner, gather_bands, gather_frames

    s = scenes.crate_scene()
    o = O.Oracle()
    scenes.upload(o, s)
    W, H = 96, 64
    # (1) one frame split into scanline bands
    f = s.frame(W, H, shard_index=rank, shard_count=world, shard_band_rows=8)
    part = torch.from_numpy(o.render_frame(f).astype(np.int64))
    full = gather_bands(part, H, world, 8, rank, dst=0)
    # uneven shares (9 bands of 8 rows over 2 ranks: 5 + 4) and a height that is no multiple of the band (row lists)
    gathered_extra = {}
    for h2, band in ((72, 8), (70, 8)):
        f2 = s.frame(W, h2, shard_index=rank, shard_count=world, shard_band_rows=band)
        p2 = torch.from_numpy(o.render_frame(f2).astype(np.int64))
        gathered_extra[h2] = gather_bands(p2, h2, world, band, rank, dst=0)
    # (2) a flythrough: frame i belongs to rank i % world
    n_frames = 5
    mine = {}
    for i in range(n_frames):
        if frame_owner(i, world) == rank:
            cam = scenes.orbit_camera(i, n_frames=n_frames, radius=72.0, height_y=35.0, target=(0.0, 11.8, 0.0))
            mine[i] = torch.from_numpy(o.render_frame(s.frame(W, H, camera=cam)).astype(np.int64))
    frames = gather_frames(mine, n_frames, (H, W), torch.int64, world, rank, dst=0)
    if rank == 0:
        np.save(os.path.join(out_dir, "full.npy"), full.numpy())
        for h2, g in gathered_extra.items():
            np.save(os.path.join(out_dir, "full_%d.npy" % h2), g.numpy())
        np.save(os.path.join(out_dir, "frames.npy"), torch.stack(frames).numpy())
    assert sorted(band_rows(H, rank, world, 8)) == [y for y in range(H) if (y // 8) % world == rank]
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path, oracle_mod):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from xna_ray_trace_b200 import scenes

    s = scenes.crate_scene()
    o = oracle_mod.Oracle()
    scenes.upload(o, s)
    W, H = 96, 64
    want = o.render_frame(s.frame(W, H)).astype(np.int64)
    assert np.array_equal(np.load(tmp_path / "full.npy"), want)
    for h2 in (72, 70):
        assert np.array_equal(np.load(tmp_path / ("full_%d.npy" % h2)), o.render_frame(s.frame(W, h2)).astype(np.int64)), h2
    frames = np.load(tmp_path / "frames.npy")
    for i in range(5):
        cam = scenes.orbit_camera(i, n_frames=5, radius=72.0, height_y=35.0, target=(0.0, 11.8, 0.0))
        assert np.array_equal(frames[i], o.render_frame(s.frame(W, H, camera=cam)).astype(np.int64)), i
