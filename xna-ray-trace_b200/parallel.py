"""Multi-GPU partitioning of the trace path (one process per GPU, torch.distributed).

The path shards along its two natural axes only (BASELINE.json north_star): scanline bands of one
frame, and whole frames of a flythrough.  The scene is replicated on every GPU; there is no
collective inside traversal — the only exchange is the gather of finished bands/frames to rank 0
(NCCL over NVLink on the GPU box, gloo in the CPU tests).  The reference already treats scanlines as
independent work items (RayTracer.cs:49-52, 399-426).
"""
import torch
import torch.distributed as dist


def band_rows(height, shard_index, shard_count, band_rows_):
    """Scanlines owned by a shard: bands of `band_rows_` rows dealt round-robin (XnartFrame.shard*)."""
    return [y for y in range(height) if (y // band_rows_) % max(1, shard_count) == shard_index]


def frame_owner(frame_index, world):
    """Flythrough frame f is rendered by rank f % world."""
    return frame_index % world


def gather_bands(part, height, world, band_rows_, rank, dst=0):
    """Assemble one frame from per-rank renders that filled only their own bands.

    `part` is a full-size (H, W) tensor whose foreign rows are don't-care.  Each rank contributes its
    rows as one contiguous slab (equal sizes are not required); rank `dst` returns the full frame.
    When H is a multiple of the band height the bands of a rank are a strided view of the frame
    (`frame.view(H/b, b, W)[rank::world]`): one copy out, one collective, one copy in per rank."""
    if world == 1:
        return part
    if height % band_rows_ == 0:
        nb = height // band_rows_
        v = part.view(nb, band_rows_, *part.shape[1:])
        counts = [len(range(r, nb, world)) for r in range(world)]
        cmax = max(counts)
        mine = v[rank::world]
        if counts[rank] < cmax:
            # ranks with one band fewer pad their share, so that ONE equal-size collective serves uneven splits too
            # (2160 rows in 8-row bands over 8 ranks: 34 or 33 bands each) instead of a send / receive per rank
            pad = torch.zeros((cmax - counts[rank],) + tuple(v.shape[1:]), dtype=part.dtype, device=part.device)
            mine = torch.cat([mine, pad])
        else:
            mine = mine.contiguous()
        if rank == dst:
            recv = torch.empty((world, cmax) + tuple(v.shape[1:]), dtype=part.dtype, device=part.device)
            dist.gather(mine, list(recv.unbind(0)), dst=dst)
            full = torch.empty_like(part)
            fv = full.view(nb, band_rows_, *part.shape[1:])
            whole = (nb // world) * world          # bands every rank owns the same number of: one strided copy
            if whole:
                fv[:whole].view(nb // world, world, *v.shape[1:]).copy_(recv[:, : nb // world].transpose(0, 1))
            for r in range(nb - whole):            # the last, incomplete round
                fv[whole + r] = recv[r, nb // world]
            return full
        dist.gather(mine, None, dst=dst)
        return None
    rows = torch.tensor(band_rows(height, rank, world, band_rows_), device=part.device, dtype=torch.long)
    mine = part.index_select(0, rows).contiguous()
    counts = [len(band_rows(height, r, world, band_rows_)) for r in range(world)]
    if rank == dst:
        bufs = [torch.empty((counts[r],) + tuple(part.shape[1:]), dtype=part.dtype, device=part.device) for r in range(world)]
        dist.gather(mine, bufs, dst=dst) if len(set(counts)) == 1 else _gather_uneven(mine, bufs, rank, world, dst)
        full = torch.empty_like(part)
        for r in range(world):
            rr = torch.tensor(band_rows(height, r, world, band_rows_), device=part.device, dtype=torch.long)
            full.index_copy_(0, rr, bufs[r])
        return full
    if len(set(counts)) == 1:
        dist.gather(mine, None, dst=dst)
    else:
        _gather_uneven(mine, None, rank, world, dst)
    return None


def _gather_uneven(mine, bufs, rank, world, dst):
    if rank == dst:
        bufs[dst].copy_(mine)
        reqs = [dist.irecv(bufs[r], src=r) for r in range(world) if r != dst]
        for q in reqs:
            q.wait()
    else:
        dist.send(mine, dst=dst)


def gather_frames(mine, n_frames, shape, dtype, world, rank, dst=0, device=None):
    """Collect the frames of a flythrough on rank `dst`.  `mine` maps frame index -> (H, W) tensor for the
    frames this rank rendered (frame_owner).  Rounds of `world` frames are gathered with one collective each."""
    device = device or (next(iter(mine.values())).device if mine else torch.device("cpu"))
    out = [None] * n_frames if rank == dst else None
    for base in range(0, n_frames, world):
        idx = base + rank
        send = mine[idx] if idx < n_frames else torch.zeros(shape, dtype=dtype, device=device)
        if world == 1:
            out[idx] = send
            continue
        if rank == dst:
            bufs = [torch.empty(shape, dtype=dtype, device=device) for _ in range(world)]
            dist.gather(send.contiguous(), bufs, dst=dst)
            for r in range(world):
                if base + r < n_frames:
                    out[base + r] = bufs[r]
        else:
            dist.gather(send.contiguous(), None, dst=dst)
    return out
