#!/usr/bin/env python
"""tools/tilesplit.py — BASELINE.json configs[2] / [4]: ONE frame split into scanline bands across the
GPUs of the box (XnartFrame.shardIndex/shardCount/shardBandRows: bands dealt round-robin, scene replicated),
finished bands gathered to rank 0 over NCCL.  Strong scaling: total work fixed as N grows.

  python tools/tilesplit.py grid 3840 2160
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/tilesplit.py grid 3840 2160
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scene", choices=["grid", "soup", "crate"])
    ap.add_argument("width", type=int)
    ap.add_argument("height", type=int)
    ap.add_argument("--tris", dest="n", type=int, default=None, help="triangle / grid-size parameter of the scene generator")
    ap.add_argument("--depth", type=int, default=None)
    ap.add_argument("--frames", type=int, default=5)
    ap.add_argument("--band-rows", type=int, default=8)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    from xna_ray_trace_b200 import capi, scenes
    from xna_ray_trace_b200.parallel import gather_bands

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    s = {"grid": lambda: scenes.grid_scene(args.n or 707), "soup": lambda: scenes.soup_scene(args.n or 1_000_000),
         "crate": scenes.crate_scene}[args.scene]()
    if args.depth is not None:
        s.max_reflections = args.depth
    x = capi.Xnart(devices=(local,))
    scenes.upload(x, s)
    W, H = args.width, args.height
    f = s.frame(W, H, shard_index=rank, shard_count=world, shard_band_rows=args.band_rows) if world > 1 else s.frame(W, H)
    buf = torch.zeros((H, W), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()

    def one():
        x.render_frame_device(f, buf.data_ptr(), stream.cuda_stream)
        return gather_bands(buf, H, world, args.band_rows, rank, dst=0) if world > 1 else buf

    full = one()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.frames):
        full = one()
    e1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.frames
    x.render_frame_device(f, buf.data_ptr(), None)
    c = x.counters()
    t = torch.tensor([ms, float(c.rays), c.msLevels + c.msResolve], dtype=torch.float64, device=dev)
    if world > 1:
        tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = t.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        tn = t.clone(); dist.all_reduce(tn, op=dist.ReduceOp.MIN)
        ms, rays, slow, fast = float(tm[0]), float(ts[1]), float(tm[2]), float(tn[2])
    else:
        rays, slow, fast = float(t[1]), float(t[2]), float(t[2])
    if rank == 0:
        checksum = int(full.to(torch.int64).bitwise_and(0xFFFFFFFF).sum().item())
        print(json.dumps({"workload": "%s %dx%d depth %d, one frame split into %d-row bands" % (s.name, W, H, s.max_reflections, args.band_rows),
                          "n_gpus": world, "ms_per_frame": ms, "frames_per_s": 1e3 / ms, "Mrays_per_s": rays / (ms * 1e-3) / 1e6,
                          "rays_per_frame": rays, "render_ms_slowest_rank": slow, "render_ms_fastest_rank": fast, "scaling": "strong",
                          "checksum": checksum}))
    x.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
