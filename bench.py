#!/usr/bin/env python
"""bench.py — headline measurement of the trace path (BASELINE.json: Mrays/s, frames/s at 1080p).

Own arm (default):   python bench.py --gpus N --steps K --warmup W
    step   = every rank renders --frames-per-step (128) 1920x1080 frames of the Free_crate scene (configs[1]:
             depth 8, Point+Wrap, 1 spot light), each from its own orbit camera — frames of a flythrough are
             the path's natural multi-GPU axis (weak scaling) — and ships them to rank 0 over NVLink.
             Scene resident on the GPU, output stays on the device.  20 steps = 2560 frames per GPU, about half a second of device time.
    value  = rays traced by all ranks / device time (CUDA events on the launching stream, max over ranks), Mrays/s
    e2e    = same metric through the reference-facing C-ABI call xnart_render_frames with HOST buffers: the
             XnartFrame records go in, the finished frames come back to page-locked host memory inside the
             timed region.
    roofline / cpu_baseline: see DESIGN.md ("Measurement").  The ceilings the kernels actually run against
             (un-fused FP32 issue rate, L1, L2) are MEASURED here with libxnart's microbenchmarks.
    extra keys (partitioned configs of BASELINE.json, every N): tile_split_4k (configs[2]), flythrough_1M
             (configs[3], AVI written by rank 0), soup_10M (configs[4]); N=1 also: deep_scene.
Reference arm:       python bench.py --impl reference ...
    times the CPU restatement of the reference (oracle, all host threads) on whole frames of the same
    workload.  The reference's own C# cannot be built here (no mono/dotnet).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H = 1920, 1080
WORKLOAD = "free_crate_1080p_depth8"
ORBIT = 64   # camera positions on the orbit


def orbit_camera(scenes, i, n=ORBIT):
    rot = 2.0 * math.pi * (i % n) / n
    return scenes.CameraData((math.sin(rot) * 72.0, 35.0, math.cos(rot) * 72.0), (0.0, 11.8, 0.0))


def bench_config(scene, frames_per_step):
    """The workload description shared by both arms (so the driver sees the same `config`)."""
    return {"workload": WORKLOAD, "width": W, "height": H, "max_reflections": scene.max_reflections,
            "triangles": scene.n_triangles, "lights": len(scene.lights), "filtering": "point", "address_mode": "wrap",
            "cameras": "orbit of %d positions around the crate" % ORBIT}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


def profile_facts(key):
    """Counters of the committed ncu --set full capture of a kernel (profiles/traffic.json, written by
    tools/ncu_traffic.py from the .ncu-rep): DRAM bytes per launch and the utilisation of each unit."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(tp)).get(key)
    except Exception:
        return None


class ClockSampler:
    """SM clock and throttle reasons sampled WHILE the timed region runs (B200_PROFILING.md), from NVML directly
    (pynvml, ~200 Hz); nvidia-smi is the fallback."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None
        self._nvml, self._h = None, None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        sm = float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM))
        try:
            mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self._h))
        except Exception:
            mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
        return [str(sm), str(self._max)] + [("Active" if mask & bit else "Not Active") for _, bit in self.BITS]

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self.rows.append(self._sample_nvml())
                    self._stop.wait(0.005)
                    continue
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvml" if self._nvml is not None else "nvidia-smi"}


def alg_bytes(c):
    """Algorithmic bytes of SURVEY.md §8(d): 40 B per ray + 32 B per slab test + 4 B per triangle
    reference + 24 B per culled and 48 B per full triangle test + 84 B per shaded hit + 4 B per pixel."""
    rays = c.raysPrimary + c.raysShadow + c.raysReflection + c.raysRefraction
    return 40 * rays + 32 * c.nBox + 4 * c.nTriRef + 24 * c.nTriCulled + 48 * c.nTriFull + 84 * c.hits + 4 * c.pixels


def alg_flops(c):
    """Algorithmic FP32 operations of SURVEY.md §8(d): 18 per slab test, 8 per culled and 51 per full triangle
    test, 120 per shaded hit."""
    return 18 * c.nBox + 8 * c.nTriCulled + 51 * c.nTriFull + 120 * c.hits


def bind_to_gpu_numa_node(local_rank):
    """Page-locked frame buffers should live on the NUMA node the GPU hangs off (8 ranks each writing tens of
    GB/s into host memory otherwise all land wherever the ranks happened to start): restrict this process to the
    CPUs NVML names as the GPU's ideal affinity before any pinned allocation is made.  Best effort; returns the
    number of CPUs bound to, or None."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w in range(words) for b in range(64) if (int(mask[w]) >> b) & 1}
        allowed = cpus & os.sched_getaffinity(0)
        if allowed and len(allowed) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, allowed)
            return len(allowed)
    except Exception:
        pass
    return None


def cpu_baseline(scene, scenes, n_threads, rows, cam_index=0, frames=1, budget_s=None, oracle=None):
    """Oracle (port of the reference's CPU path) on `frames` frames (scanlines `rows` of each), stopping early
    once `budget_s` seconds are spent; returns Mrays/s, seconds, rays, frames rendered."""
    from oracle import oracle as O

    o = oracle
    if o is None:
        o = O.Oracle()
        scenes.upload(o, scene)
    tot_t, tot_rays, done = 0.0, 0, 0
    for k in range(frames):
        f = scene.frame(W, H, camera=orbit_camera(scenes, cam_index + k))
        t0 = time.perf_counter()
        o.render_frame(f, n_threads=n_threads, rows=rows)
        tot_t += time.perf_counter() - t0
        tot_rays += o.counters().rays
        done += 1
        if budget_s is not None and tot_t >= budget_s:
            break
    if oracle is None:
        o.close()
    return tot_rays / tot_t / 1e6, tot_t, tot_rays, done


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port, all host threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from xna_ray_trace_b200 import scenes

    from oracle import oracle as O

    scene = scenes.crate_scene()
    cores = os.cpu_count() or 1
    # one step = a bounded sample of the own arm's step: FRAMES_PER_STEP whole 1080p frames of the same orbit
    # (about a quarter of a second of work for all host threads), so the default --steps/--warmup finish in seconds
    FRAMES_PER_STEP = 4
    rows = (0, H)
    o = O.Oracle()
    scenes.upload(o, scene)
    for i in range(args.warmup):
        cpu_baseline(scene, scenes, cores, rows, i, frames=1, oracle=o)
    tot_rays, tot_t = 0, 0.0
    for i in range(args.steps):
        _, dt, rays, _ = cpu_baseline(scene, scenes, cores, rows, i * FRAMES_PER_STEP, frames=FRAMES_PER_STEP, oracle=o)
        tot_rays += rays
        tot_t += dt
    o.close()
    v = tot_rays / tot_t / 1e6
    line = {
        "impl": "reference", "metric": "Mrays/s (primary+secondary)", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(scene, args.frames_per_step),
        "note": "the reference's C# cannot be built here (no mono/dotnet): CPU restatement (oracle) on the host cores; "
                "a step is a bounded sample (%d whole frames) of the own arm's %d-frame step" % (FRAMES_PER_STEP, args.frames_per_step),
        "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port",
                         "sample": "%d whole 1080p frames per step (orbit cameras), %d threads, %.1f s in total" % (FRAMES_PER_STEP, cores, tot_t)},
        "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "frames_per_s": float(FRAMES_PER_STEP) * args.steps / tot_t,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="xnart", choices=["xnart", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=128, help="1080p frames every GPU renders per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-deep-scene", action="store_true", help="skip the 1M-triangle measurement at N=1")
    ap.add_argument("--extras", default="tile_split_4k,flythrough_1M,soup_10M",
                    help="comma-separated partitioned configs to add as extra keys ('' = none)")
    ap.add_argument("--fly-frames", type=int, default=600)
    ap.add_argument("--soup-triangles", type=int, default=10_000_000)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and os.environ.get("OMP_NUM_THREADS", "1") == "1":
        # torchrun pins every rank to ONE OpenMP thread; the native octree builder (csrc/scene_build.cpp) is OpenMP
        # code and every rank builds its own copy of the scene: give each rank its share of the host cores
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // world))

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libxnart has no CPU fallback")
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None

    from xna_ray_trace_b200 import capi, parallel, records, scenes

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    FPS = max(1, args.frames_per_step)
    INFLIGHT = max(1, int(os.environ.get("XNART_BENCH_INFLIGHT", "2")))
    scene = scenes.crate_scene()
    x = capi.Xnart(devices=(local_rank,), frames_in_flight=INFLIGHT)
    scenes.upload(x, scene)
    stream = torch.cuda.current_stream()

    # Every step renders FPS frames into one half of a ring; finished halves are shipped to rank 0 while the next
    # step is being rendered (XNART_BENCH_GATHER):
    #   "p2p"   every rank copies its half straight into a receive buffer of rank 0 (CUDA IPC handle from
    #           xnart_ipc_export, device-to-device xnart_copy_async on a copy stream: copy engines over NVLink, no
    #           SMs) and then, on the same stream, the step's number into its slot of a stamp array of rank 0 — the
    #           completion signal, also moved by a copy engine.  A rank may reuse a ring half once its OWN copy of it
    #           has finished (a local event); rank 0 checks the stamps when the timed region ends.  (A one-element
    #           all-reduce per step was the signal before: its kernel had to find a free SM between persistent
    #           k_path launches that fill the machine — 18.9 vs 17.6 ms per step at 8 GPUs.)
    #   "nccl"  one async dist.gather per step; its receive kernels on rank 0 share the SMs with the rendering.
    ring = torch.zeros((2, FPS, H, W), dtype=torch.int32, device=dev)
    half_bytes = FPS * H * W * 4
    pending = [None, None]
    gather_mode = os.environ.get("XNART_BENCH_GATHER", "p2p") if world > 1 else "none"
    gathered = [None, None]
    recv_ptr, remote_ptr = [None, None], [None, None]
    copy_stream = torch.cuda.Stream(device=dev) if world > 1 else None
    done_ev = [None, None]          # p2p: this rank's copy of ring[half] (and its stamp) has finished
    stamp_ptr = None                # p2p: world x int64 on rank 0, slot r = number of the last step rank r delivered
    if gather_mode == "p2p":
        box = [None]
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        try:
            if rank == 0:
                recv_ptr = [capi.alloc_device(local_rank, world * half_bytes) for _ in range(2)]
                stamp_ptr = capi.alloc_device(local_rank, 8 * world)
                capi.copy_async(local_rank, stamp_ptr, torch.zeros(world, dtype=torch.int64, device=dev).data_ptr(), 8 * world, stream.cuda_stream)
                torch.cuda.synchronize()
                box[0] = [capi.ipc_export(local_rank, p) for p in recv_ptr] + [capi.ipc_export(local_rank, stamp_ptr)]
        except Exception as e:
            print("bench: CUDA IPC export failed (%s)" % e, file=sys.stderr)
            box[0] = None
        dist.broadcast_object_list(box, src=0)
        try:
            if box[0] is None:
                raise RuntimeError("no handles")
            if rank != 0:
                remote_ptr = [capi.ipc_open(local_rank, h) for h in box[0][:2]]
                stamp_ptr = capi.ipc_open(local_rank, box[0][2])
        except Exception as e:
            print("bench: CUDA IPC open failed on rank %d (%s)" % (rank, e), file=sys.stderr)
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:      # every rank falls back together
            gather_mode = "nccl"
    if gather_mode == "nccl":
        gathered = [[torch.zeros((FPS, H, W), dtype=torch.int32, device=dev) for _ in range(world)] for _ in range(2)] if rank == 0 else [None, None]

    warm = max(args.warmup, 3)
    # camera matrices are host-side inputs (Camera.View/Projection in the reference): built up front
    step_frames = [[scene.frame(W, H, camera=orbit_camera(scenes, (s * FPS + k) * world + rank)) for k in range(FPS)]
                   for s in range(warm + args.steps)]

    stamps = torch.arange(1, warm + args.steps + 2, dtype=torch.int64, device=dev)   # stamps[s] = s + 1: the source of the stamp copies

    def ship(half, s):
        if world == 1:
            return
        if gather_mode == "nccl":
            pending[half] = dist.gather(ring[half], gathered[half], dst=0, async_op=True)
            return
        copy_stream.wait_stream(stream)
        dst = (recv_ptr if rank == 0 else remote_ptr)[half] + rank * half_bytes
        capi.copy_async(local_rank, dst, ring[half].data_ptr(), half_bytes, copy_stream.cuda_stream)
        capi.copy_async(local_rank, stamp_ptr + 8 * rank, stamps[s].data_ptr(), 8, copy_stream.cuda_stream)   # after the frames, in stream order
        done_ev[half] = torch.cuda.Event()
        done_ev[half].record(copy_stream)

    def step(s):
        half = s & 1
        if pending[half] is not None:
            pending[half].wait()   # the render stream waits until this half's previous transfer has finished
            pending[half] = None
        if done_ev[half] is not None:
            stream.wait_event(done_ev[half])
            done_ev[half] = None
        x.render_frames_device(step_frames[s], ring[half].data_ptr(), stream.cuda_stream)
        ship(half, s)

    def drain():
        for b in range(2):
            if pending[b] is not None:
                pending[b].wait()
                pending[b] = None
            if done_ev[b] is not None:
                stream.wait_event(done_ev[b])
                done_ev[b] = None

    def barrier():
        drain()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for s in range(warm):
        step(s)
    barrier()
    launches0 = x.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        ev0.record(stream)
        for s in range(args.steps):
            step(warm + s)
        drain()              # the last transfers belong to the timed region
        ev1.record(stream)
        barrier()
    my_ms = ev0.elapsed_time(ev1)
    launches = x.kernel_launches() - launches0
    gather_ok = None
    if world > 1 and rank == 0:
        # every rank's frames must have landed on rank 0: different cameras -> different, non-empty frames
        last = (warm + args.steps - 1) & 1
        if gather_mode == "p2p":
            g = torch.zeros((world, FPS, H, W), dtype=torch.int32, device=dev)
            capi.copy_async(local_rank, g.data_ptr(), recv_ptr[last], world * half_bytes, stream.cuda_stream)
            got_stamps = torch.zeros(world, dtype=torch.int64, device=dev)
            capi.copy_async(local_rank, got_stamps.data_ptr(), stamp_ptr, 8 * world, stream.cuda_stream)
            torch.cuda.synchronize()
            stamps_ok = bool((got_stamps == warm + args.steps).all().item())   # every rank delivered its last step
        else:
            g = torch.stack(gathered[last])
        sums = [int(g[r, 0].to(torch.int64).bitwise_and(0xFFFFFF).sum().item()) for r in range(world)]
        gather_ok = all(v > 0 for v in sums) and len(set(sums)) == world and (gather_mode != "p2p" or stamps_ok)
        del g
    # the frames the timed region left behind (its last step) must be the frames an untimed render of that step gives
    last_half = (warm + args.steps - 1) & 1
    timed_sum = int(ring[last_half].to(torch.int64).bitwise_and(0xFFFFFF).sum().item())
    # rays of the timed steps: exact device counters of an untimed re-render of the same frames
    my_rays = 0.0
    for s in range(args.steps):
        x.render_frames_device(step_frames[warm + s], ring[last_half ^ 1].data_ptr(), None)
        my_rays += float(x.counters().rays)
    checksum = int(ring[last_half ^ 1].to(torch.int64).bitwise_and(0xFFFFFF).sum().item())
    frames_verified = checksum == timed_sum and len(torch.unique(ring[last_half][0])) > 1000
    t = torch.tensor([my_ms, my_rays], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, total_rays = float(tmax[0]), float(tsum[1])
    else:
        ms, total_rays = my_ms, my_rays
    value = total_rays / (ms * 1e-3) / 1e6

    # ---- e2e: the C ABI with HOST buffers.  xnart_render_frames, one call per step: FPS frame records in (184 B
    # each), FPS finished frames out to page-locked host memory, everything inside the timed region (wall clock);
    # the context keeps INFLIGHT frames in flight and overlaps the device->host copies with the rendering.
    pin_n = capi.PinnedFrame(H, W, frames=FPS)
    x.render_frames(step_frames[0], out=pin_n.array)   # warm-up (allocates the context's second frame buffer)
    barrier()
    with ClockSampler(local_rank) as clocks2:
        t0 = time.perf_counter()
        for s in range(args.steps):
            x.render_frames(step_frames[warm + s], out=pin_n.array)
        e2e_s = time.perf_counter() - t0
    e2e_checksum = int(pin_n.array.astype(np.uint64).sum())
    # one frame per call, synchronous (what RenderInternal + SetData would be replaced by)
    one = step_frames[warm][:min(FPS, 32)]
    pin = capi.PinnedFrame(H, W)
    for f in one[:3]:
        x.render_frame(f, out=pin)
    one_rays = 0
    t0 = time.perf_counter()
    for f in one:
        x.render_frame(f, out=pin)
    e2e1_s = time.perf_counter() - t0
    for f in one:
        x.render_frame_device(f, ring[0, 0].data_ptr(), None)
        one_rays += x.counters().rays
    te = torch.tensor([e2e_s, e2e1_s, float(one_rays)], dtype=torch.float64, device=dev)
    if world > 1:
        temax = te.clone()
        dist.all_reduce(temax, op=dist.ReduceOp.MAX)
        tesum = te.clone()
        dist.all_reduce(tesum, op=dist.ReduceOp.SUM)
        e2e_s, e2e1_s, one_rays = float(temax[0]), float(temax[1]), float(tesum[2])
    e2e_value = total_rays / e2e_s / 1e6
    e2e1_value = one_rays / e2e1_s / 1e6
    clocks.rows += clocks2.rows
    pin_n.free()
    pin.free()

    # ---- roofline of the dominant kernel -----------------------------------------------------------
    # Free_crate frames are ONE launch of the persistent path kernel each (csrc/path.cu) — every other launch of a
    # step is a memset.  What binds it is the issue rate of un-fused FP32 / compare / select instructions (ncu:
    # issue active 75 %, DRAM 0.06 %; profiles/), so `frac` is algorithmic FP32 operations per second against the
    # MEASURED FMUL+FADD rate of this GPU; the HBM / L2 / L1 views of the algorithmic bytes are listed beside it.
    roof, cpu, deep, peaks_out = None, None, None, None
    if rank == 0:
        peaks, how = measured_peaks()
        fp32_peak = capi.measure_peak(local_rank, capi.PEAK_FP32_NOFMA) / 1e12
        l1_peak = capi.measure_peak(local_rank, capi.PEAK_L1_READ) / 1e9
        l2_peak = capi.measure_peak(local_rank, capi.PEAK_L2_READ) / 1e9
        hbm_read = capi.measure_peak(local_rank, capi.PEAK_HBM_READ, reps=3) / 1e9
        sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
        peaks_out = {"fp32_noFMA_Tops": fp32_peak, "fp32_noFMA_Tops_nominal": sms * 128 * 1.965e9 / 1e12, "l1_read_GBs": l1_peak,
                     "l2_read_GBs": l2_peak, "hbm_read_GBs": hbm_read, "hbm_copy_GBs": peaks["hbm_gbs"], "hbm_copy_source": how,
                     "how": "xnart_measure_peak (csrc/microbench.cu), best of 5 launches, on this GPU in this process"}

        def roofline_of(cc, seconds, launches_n, kernel, key, share):
            """cc: instrumented counters (== the oracle's) of the work that took `seconds` in `launches_n` launches."""
            ab, af = alg_bytes(cc), alg_flops(cc)
            facts = profile_facts(key) or {}
            views = {"issue_fp32": {"achieved": af / seconds / 1e12, "peak": fp32_peak, "unit": "TFLOP/s"},
                     "l1": {"achieved": ab / seconds / 1e9, "peak": l1_peak, "unit": "GB/s"},
                     "l2": {"achieved": ab / seconds / 1e9, "peak": l2_peak, "unit": "GB/s"},
                     "hbm": {"achieved": ab / seconds / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s"}}
            for v in views.values():
                v["frac"] = v["achieved"] / v["peak"] if v["peak"] else None
            b = views["issue_fp32"]
            return {"bound": "issue", "achieved": b["achieved"], "peak": b["peak"], "unit": b["unit"], "frac": b["frac"],
                    "traffic": facts.get("dram_bytes_per_launch"), "peak_source": "measured (xnart_measure_peak: FMUL+FADD, no FMA)",
                    "kernel": kernel, "launches_per_step": launches_n, "avg_launch_ms": 1e3 * seconds / launches_n,
                    "alg_flops_per_launch": af / launches_n, "alg_bytes_per_launch": ab / launches_n,
                    "kernel_share_of_step": share, "views": views, "ncu": facts,
                    "work_per_ray": {"slab_tests": cc.nBox / max(1, cc.rays), "triangle_refs": cc.nTriRef / max(1, cc.rays),
                                     "full_triangle_tests": cc.nTriFull / max(1, cc.rays)},
                    "note": "achieved = algorithmic FP32 operations (SURVEY.md 8d: 18 per slab test, 8 / 51 per culled / full triangle "
                            "test, 120 per shaded hit; counted by the instrumented kernels == the oracle) of one step / the step's "
                            "device time, frames in flight as timed; views.hbm is the algorithmic-byte rate against the HBM copy "
                            "peak: the scene is read from the constant bank / caches, so it is NOT the binding unit (traffic = "
                            "DRAM bytes per launch from the committed ncu capture)"}

        x.set_option(records.OPT_COUNTERS, 1)
        x.render_frames_device(step_frames[warm], ring[0].data_ptr(), None)
        cc = x.counters()
        x.set_option(records.OPT_COUNTERS, 0)
        roof = roofline_of(cc, my_ms * 1e-3 / max(1, args.steps), FPS,
                           "k_path<false> (persistent path kernel: one launch per frame, scene in the constant bank)", "k_path", 1.0)

        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v_all, dt_all, _, n_all = cpu_baseline(scene, scenes, cores, (0, H), 0, frames=400, budget_s=10.0)
            v_one, dt_one, _, n_one = cpu_baseline(scene, scenes, 1, (0, H), 0, frames=20, budget_s=4.0)
            cpu = {"value": v_all, "unit": "Mrays/s", "cores": cores, "kind": "port",
                   "sample": "%d whole 1080p frames of the same workload (orbit cameras), %d threads, %.1f s" % (n_all, cores, dt_all),
                   "single_thread_as_shipped": {"value": v_one, "cores": 1, "seconds": dt_one, "frames": n_one,
                                                "note": "RayTracer.cs:99 hard-codes one worker"}}
    x.close()
    del ring
    if gather_mode == "p2p":
        torch.cuda.synchronize()
        if rank != 0:
            for ptr in remote_ptr:
                capi.ipc_close(local_rank, ptr)
        dist.barrier()
        if rank == 0:
            for ptr in recv_ptr:
                capi.free_device(local_rank, ptr)
    gathered = None
    torch.cuda.empty_cache()

    # ---- the partitioned configs of BASELINE.json (every N) and the 1M-triangle roofline (N = 1) ------------------
    extras = [e for e in args.extras.split(",") if e]
    extra_out = {}
    need_grid = ("tile_split_4k" in extras) or ("flythrough_1M" in extras) or (world == 1 and not args.no_deep_scene)
    if need_grid:
        gs = scenes.grid_scene(707)
        xg = capi.Xnart(devices=(local_rank,), frames_in_flight=INFLIGHT)
        scenes.upload(xg, gs)
        fbuf = torch.zeros((H, W), dtype=torch.int32, device=dev)

        if world == 1 and not args.no_deep_scene and rank == 0:
            # second target of north_star: the 1M-triangle octree scene (configs[2] geometry) at 1080p, depth 4, 2 lights
            gframes = [gs.frame(W, H, camera=scenes.CameraData((np.sin(0.13 * i) * 80.0, 40.0, np.cos(0.13 * i) * 80.0), (0.0, 0.0, 0.0))) for i in range(12)]
            gbuf = torch.zeros((len(gframes), H, W), dtype=torch.int32, device=dev)
            xg.render_frames_device(gframes[:3], gbuf.data_ptr(), stream.cuda_stream)
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            g0.record(stream)
            xg.render_frames_device(gframes, gbuf.data_ptr(), stream.cuda_stream)   # INFLIGHT frames side by side
            g1.record(stream)
            torch.cuda.synchronize()
            gsec = g0.elapsed_time(g1) * 1e-3
            xg.render_frames_device(gframes, gbuf.data_ptr(), None)
            grays = xg.counters().rays
            l0 = xg.kernel_launches()
            xg.set_option(records.OPT_COUNTERS, 1)
            xg.render_frames_device(gframes, gbuf.data_ptr(), None)
            gcc = xg.counters()
            xg.set_option(records.OPT_COUNTERS, 0)
            gl = xg.kernel_launches() - l0
            ntrace = len(gframes) * 2 * (gs.max_reflections + 1)
            groof = roofline_of(gcc, gsec, ntrace, "k_trace<false> (persistent trace kernel, 2 query batches per recursion level)", "k_trace",
                                (profile_facts("k_trace") or {}).get("share_of_frame"))
            groof["launches_of_all_kernels"] = gl
            deep = {"workload": "grid_1M_triangles_1080p_depth%d" % gs.max_reflections, "triangles": gs.n_triangles,
                    "lights": len(gs.lights), "value": grays / gsec / 1e6, "unit": "Mrays/s",
                    "ms_per_frame": 1e3 * gsec / len(gframes), "frames_per_s": len(gframes) / gsec, "rays_per_frame": grays / len(gframes),
                    "frames_in_flight": INFLIGHT, "roofline": groof}
            del gbuf

        if "tile_split_4k" in extras:
            extra_out["tile_split_4k"] = tile_split(xg, gs, 3840, 2160, world, rank, dev, parallel, torch, dist, frames=5,
                                                    label="configs[2]: 1M-triangle grid, 3840x2160, depth %d, ONE frame split across the GPUs" % gs.max_reflections)
        if "flythrough_1M" in extras:
            extra_out["flythrough_1M"] = flythrough_avi(xg, gs, args.fly_frames, world, rank, dev, torch, dist)
        xg.close()
        del fbuf
        torch.cuda.empty_cache()
    if "soup_10M" in extras:
        t0 = time.perf_counter()
        ss = scenes.soup_scene(args.soup_triangles)
        xs = capi.Xnart(devices=(local_rank,))
        scenes.upload(xs, ss)
        build_s = time.perf_counter() - t0
        r = tile_split(xs, ss, W, H, world, rank, dev, parallel, torch, dist, frames=3,
                       label="configs[4]: %d-triangle random soup, depth-%d mirror reflections, 1920x1080, ONE frame split across the GPUs"
                             % (ss.n_triangles, ss.max_reflections))
        if r is not None:
            r["scene_build_and_upload_s"] = build_s
        extra_out["soup_10M"] = r
        xs.close()

    if rank == 0:
        cfg = bench_config(scene, FPS)
        line = {
            "metric": "Mrays/s (primary+secondary)", "value": value, "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "run": {"frames_per_step_per_gpu": FPS, "frames_in_flight": INFLIGHT,
                    "parallelism": ("frames sharded across GPUs (%d per GPU per step); every step's frames are shipped to rank 0 (%s) while the next step renders"
                                    % (FPS, "device-to-device copies into an IPC-shared buffer of rank 0, each followed by a step stamp in rank 0's stamp array; no collective in the data path"
                                       if gather_mode == "p2p" else "async NCCL gather")) if world > 1 else "single GPU",
                    "l2": "each step writes %d MB of frames (ring of 2 steps) — larger than the 126 MB L2 — from %d different cameras; the "
                          "scene itself is 1.5 KB in the constant bank and cannot be made larger than a cache" % (half_bytes >> 20, ORBIT),
                    "cpus_bound_near_gpu": numa, "cpu_cores": os.cpu_count()},
            "frames_per_s": world * FPS * args.steps / (ms * 1e-3),
            "rays_per_frame": total_rays / (world * FPS * max(1, args.steps)),
            "checksum": checksum, "timed_frames_match_rerender": bool(frames_verified),
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 184 * FPS, "d2h_bytes_per_step": W * H * 4 * FPS,
                    "frames_per_s": world * FPS * args.steps / e2e_s, "checksum": e2e_checksum,
                    "api": "xnart_render_frames (one call per step: %d frames, %d in flight per GPU, page-locked host frames)" % (FPS, INFLIGHT),
                    "one_frame_per_call": {"value": e2e1_value, "frames_per_s": world * len(one) / e2e1_s,
                                           "api": "xnart_render_frame, synchronous, %d calls" % len(one)}},
            "gpu_launches": int(launches),
            "gather_check": gather_ok,
            "clocks": clocks.summary(),
            "peaks": peaks_out,
            "roofline": roof,
            "cpu_baseline": cpu,
            "deep_scene": deep,
        }
        line.update(extra_out)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def tile_split(x, scene, width, height, world, rank, dev, parallel, torch, dist, frames, label, band_rows=8):
    """ONE frame split into scanline bands dealt round-robin to the ranks (XnartFrame.shardIndex/shardCount/shardBandRows;
    RayTracer.cs:49-52 treats scanlines as independent work items), finished bands gathered to rank 0 over NCCL.
    Strong scaling: the frame is fixed as N grows.  Timed on the device (CUDA events, max over ranks), gather included."""
    f = scene.frame(width, height, shard_index=rank, shard_count=world, shard_band_rows=band_rows) if world > 1 else scene.frame(width, height)
    buf = torch.zeros((height, width), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()

    def one():
        x.render_frame_device(f, buf.data_ptr(), stream.cuda_stream)
        return parallel.gather_bands(buf, height, world, band_rows, rank, dst=0) if world > 1 else buf

    full = one()
    full = one()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(frames):
        full = one()
    e1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / frames
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
    if rank != 0:
        return None
    checksum = int(full.to(torch.int64).bitwise_and(0xFFFFFFFF).sum().item())
    return {"workload": label, "n_gpus": world, "scaling": "strong", "ms_per_frame": ms, "frames_per_s": 1e3 / ms,
            "value": rays / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "rays_per_frame": rays, "band_rows": band_rows,
            "render_ms_slowest_rank": slow, "render_ms_fastest_rank": fast, "frames_timed": frames, "checksum": checksum}


def flythrough_avi(x, scene, n_frames, world, rank, dev, torch, dist):
    """configs[3]: an n-frame 1080p orbit of the 1M-triangle scene, frame f rendered by rank f % N, converted to 24bpp DIB
    rows on the device, gathered to rank 0 and appended to an uncompressed OpenDML AVI by the AviManager-style sink
    (a writer thread, so file I/O overlaps the rendering).  frames/s = frames / wall time until the file is closed."""
    from xna_ray_trace_b200 import avi, flythrough

    cams = flythrough.orbit_cameras(n_frames)
    path, mgr, sink, err = None, None, None, None
    if rank == 0:
        try:
            tmpdir = os.environ.get("XNART_BENCH_TMP")
            if not tmpdir:
                # 3.7 GB of frames: a RAM-backed directory when there is one with room, else the system temp directory
                try:
                    st = os.statvfs("/dev/shm")
                    if st.f_bavail * st.f_frsize > (6 << 30):
                        tmpdir = "/dev/shm"
                except OSError:
                    tmpdir = None
            fd, path = tempfile.mkstemp(suffix=".avi", prefix="xnart_fly_", dir=tmpdir or None)
            os.close(fd)
            mgr = avi.AviManager(path, False, io_threads=8)
            sink = mgr.AddVideoStream(False, 30, 0, W, H, avi.PIXELFORMAT_24BPP_RGB)   # 30 fps as Game1.cs:198
        except Exception as e:
            err, sink = "cannot create the AVI file: %s" % e, None
    # warm-up round outside the timed region (arena allocation, NCCL channels)
    # (rank 0 gets a sink that discards: the page-locked staging buffers of the real run are allocated here)
    flythrough.render_flythrough(x, scene, cams[:8 * world], W, H, sink=flythrough.NullSink(W, H) if rank == 0 else None,
                                 world=world, rank=rank, batch=8, dib24=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    try:
        n_mine, rays, _ = flythrough.render_flythrough(x, scene, cams, W, H, sink=sink, world=world, rank=rank, batch=8, dib24=True)
    except Exception as e:   # the sink failed (disk full ...): the frames were still rendered and gathered
        err = "AVI sink failed: %s" % e
        n_mine, rays = len([f for f in range(n_frames) if f % world == rank]), 0
    render_s = time.perf_counter() - t0
    size = None
    if rank == 0 and mgr is not None:
        try:
            mgr.Close()
            size = os.path.getsize(path)
        except Exception as e:
            err = "closing the AVI failed: %s" % e
    total_s = time.perf_counter() - t0
    if path:
        try:
            os.remove(path)
        except OSError:
            pass
    t = torch.tensor([total_s, render_s, float(rays)], dtype=torch.float64, device=dev)
    if world > 1:
        tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = t.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        total_s, render_s, rays = float(tm[0]), float(tm[1]), float(ts[2])
    if rank != 0:
        return None
    return {"workload": "configs[3]: %d-frame 1920x1080 orbit of the 1M-triangle grid, depth %d, frames sharded across the GPUs, "
                        "gathered to rank 0 into an uncompressed 24bpp OpenDML AVI" % (n_frames, scene.max_reflections),
            "n_gpus": world, "scaling": "strong", "frames": n_frames, "seconds": total_s, "frames_per_s": n_frames / total_s,
            "value": rays / total_s / 1e6, "unit": "Mrays/s", "avi_bytes": size, "avi_error": err,
            "avi_dir": os.path.dirname(path) if path else None, "avi_write_busy_s": getattr(flythrough.render_flythrough, "last_sink_busy_s", None),
            "seconds_until_last_frame_delivered": render_s}


if __name__ == "__main__":
    sys.exit(main())
