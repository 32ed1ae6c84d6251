#!/usr/bin/env python
"""All-to-one ingest rate of the frame shipping of bench.py (N > 1): every rank copies `--mb` MB into its slot of an
IPC-shared buffer of rank 0 with xnart_copy_async (copy engines over NVLink), `--streams` pieces on as many streams.
  torchrun --nproc-per-node 8 tools/ingestbench.py [--mb 1012] [--reps 10]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from xna_ray_trace_b200 import capi

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=1012)
ap.add_argument("--reps", type=int, default=10)
args = ap.parse_args()
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
nbytes = args.mb << 20
src = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
box = [None]
if rank == 0:
    recv = capi.alloc_device(lr, world * nbytes)
    box[0] = capi.ipc_export(lr, recv)
dist.broadcast_object_list(box, src=0)
dst = recv if rank == 0 else capi.ipc_open(lr, box[0])
flag = torch.zeros(1, dtype=torch.int32, device="cuda")
for nstreams, senders in ((1, world), (2, world), (4, world), (8, world), (1, 2), (1, 4)):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    piece = nbytes // nstreams
    active = rank < senders and rank != 0
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        if active:
            for k, st in enumerate(streams):
                capi.copy_async(lr, dst + rank * nbytes + k * piece, src.data_ptr() + k * piece, piece, st.cuda_stream)
    torch.cuda.synchronize()
    dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        n = min(senders, world) - 1
        print("senders %d, %d stream(s) each: %.1f GB/s into rank 0" % (n, nstreams, n * nbytes * args.reps / dt / 1e9), flush=True)
dist.barrier()
dist.destroy_process_group()
