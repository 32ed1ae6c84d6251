#!/usr/bin/env python
"""tools/ipc_check.py — two processes on ONE GPU: rank 1 writes a pattern into a buffer of rank 0 through the
CUDA IPC utilities of libxnart (xnart_ipc_export / xnart_ipc_open / xnart_copy_async).  Mechanics check for the
peer-to-peer frame gather of bench.py (gloo only: NCCL refuses two ranks on one GPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def worker(rank, port):
    from xna_ray_trace_b200 import capi

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=2)
    torch.cuda.set_device(0)
    n = 1 << 20
    box = [None]
    if rank == 0:
        buf = capi.alloc_device(0, n * 4)
        box[0] = capi.ipc_export(0, buf)
    dist.broadcast_object_list(box, src=0)
    if rank == 1:
        remote = capi.ipc_open(0, box[0])
        src = torch.arange(n, dtype=torch.int32, device="cuda:0") * 3 + 1
        capi.copy_async(0, remote, src.data_ptr(), n * 4, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        out = torch.zeros(n, dtype=torch.int32, device="cuda:0")
        capi.copy_async(0, out.data_ptr(), buf, n * 4, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        ok = bool((out.cpu().numpy() == np.arange(n, dtype=np.int64) * 3 + 1).all())
        print("ipc_check:", "OK" if ok else "MISMATCH")
    dist.barrier()
    if rank == 1:
        capi.ipc_close(0, remote)
    dist.barrier()
    if rank == 0:
        capi.free_device(0, buf)
    dist.destroy_process_group()


if __name__ == "__main__":
    mp.spawn(worker, args=(29791,), nprocs=2, join=True)
