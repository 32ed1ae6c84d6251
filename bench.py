#!/usr/bin/env python
"""bench.py — headline measurement of the trace path (BASELINE.json: Mrays/s, frames/s at 1080p).

Own arm (default):   python bench.py --gpus N --steps K --warmup W
    step   = every rank renders ONE 1920x1080 frame of the Free_crate scene (configs[1]: depth 8,
             Point+Wrap, 1 spot light) from its own orbit camera — frames of a flythrough are the
             path's natural multi-GPU axis (weak scaling) — and the finished frames are gathered to
             rank 0 over NCCL.  Scene resident in HBM, output stays on the device.
    value  = rays traced by all ranks / device time (CUDA events, max over ranks), Mrays/s
    e2e    = same metric through the reference-facing C-ABI call xnart_render_frame with HOST
             buffers: the XnartFrame goes in, the frame comes back to pinned host memory inside the
             timed region.
    roofline / cpu_baseline: see DESIGN.md ("Measurement").  "deep_scene" adds the same measurement for the
             1M-triangle octree scene of BASELINE.json (N=1 only; not part of `value`).
Reference arm:       python bench.py --impl reference ...
    times the CPU restatement of the reference (oracle, all host threads) on a bounded band of
    scanlines of the same frames.  The reference's own C# cannot be built here (no mono/dotnet).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H = 1920, 1080
WORKLOAD = "free_crate_1080p_depth8"


def orbit_camera(scenes, i, n=64):
    import math

    rot = 2.0 * math.pi * (i % n) / n
    return scenes.CameraData((math.sin(rot) * 72.0, 35.0, math.cos(rot) * 72.0), (0.0, 11.8, 0.0))


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled WHILE the timed region runs (B200_PROFILING.md).  The timed
    regions here last a few milliseconds, so the samples come from NVML directly (pynvml, ~1 kHz) — one
    `nvidia-smi` process start takes longer than the whole region; nvidia-smi is the fallback."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    # nvmlClocksEventReasons bits
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
                    self._stop.wait(0.001)
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
    return 18 * c.nBox + 8 * c.nTriCulled + 51 * c.nTriFull + 120 * c.hits


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
    # one step = FRAMES_PER_STEP whole 1080p frames (about half a second of work for all host threads),
    # so that the default --steps/--warmup finish well within a few minutes
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
    frames_equiv = float(FRAMES_PER_STEP)
    line = {
        "impl": "reference", "metric": "Mrays/s (primary+secondary)", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "width": W, "height": H, "max_reflections": 8,
                   "note": "reference C# cannot be built here (no mono/dotnet): CPU restatement (oracle) on host cores"},
        "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port",
                         "sample": "%d whole 1080p frames per step (orbit cameras), %d threads, %.1f s in total" % (FRAMES_PER_STEP, cores, tot_t)},
        "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "frames_per_s": frames_equiv * args.steps / tot_t,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="xnart", choices=["xnart", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-deep-scene", action="store_true", help="skip the secondary 1M-triangle measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from xna_ray_trace_b200 import capi, records, scenes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libxnart has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    scene = scenes.crate_scene()
    x = capi.Xnart(devices=(local_rank,))
    scenes.upload(x, scene)
    if os.environ.get("XNART_BENCH_FUSE") is not None:  # experiment switch: 1 = k_frame instead of one launch per level
        x.set_option(records.OPT_FUSE_FRAME, int(os.environ["XNART_BENCH_FUSE"]))
    # Frames in flight: a context renders one frame at a time on one stream (one path arena); INFLIGHT contexts
    # on INFLIGHT streams render consecutive frames of the flythrough concurrently, so the kernels of one frame
    # fill the launch gaps and the nearly empty deep levels of the other.
    INFLIGHT = max(1, int(os.environ.get("XNART_BENCH_INFLIGHT", "3")))
    ctxs = [x]
    for _ in range(1, INFLIGHT):
        xi = capi.Xnart(devices=(local_rank,))
        scenes.upload(xi, scene)
        if os.environ.get("XNART_BENCH_FUSE") is not None:
            xi.set_option(records.OPT_FUSE_FRAME, int(os.environ["XNART_BENCH_FUSE"]))
        ctxs.append(xi)
    rstreams = [torch.cuda.Stream(device=dev) for _ in ctxs] if INFLIGHT > 1 else []

    # Frames are rendered into a ring of 2 x GROUP device buffers; every GROUP steps ONE async NCCL gather
    # ships the last GROUP frames of every rank to rank 0 (on NCCL's stream) while the next group is being
    # rendered.  Grouping keeps the host-side cost of the collective (one Python/NCCL call) off the
    # critical path: a 1080p crate frame takes only ~0.6 ms of GPU time.
    GROUP = 4 if world > 1 else 1
    ring = torch.zeros((2, GROUP, H, W), dtype=torch.int32, device=dev)
    frame_dev = ring[0, 0]
    pending = [None, None]
    stream = torch.cuda.current_stream()

    # How the finished frames reach rank 0, once per group of GROUP steps (XNART_BENCH_GATHER):
    #   "nccl"  one async dist.gather (NCCL over NVLink); its receive kernels on rank 0 share the SMs with rank 0's
    #           own rendering.
    #   "p2p"   every rank copies its group straight into a receive buffer of rank 0 (CUDA IPC handle from
    #           xnart_ipc_export, device-to-device xnart_copy_async on a copy stream: copy engines, no SMs); a
    #           one-element all-reduce ordered after the copy is the completion signal.
    gather_mode = os.environ.get("XNART_BENCH_GATHER", "p2p") if world > 1 else "none"   # 8 GPUs: p2p 95.1, nccl 92.1 Grays/s
    gathered = [None, None]
    recv_ptr, remote_ptr = [None, None], [None, None]
    half_bytes = GROUP * H * W * 4
    copy_stream = torch.cuda.Stream(device=dev) if world > 1 else None
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    if gather_mode == "p2p":
        box = [None]
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        try:
            if rank == 0:
                recv_ptr = [capi.alloc_device(local_rank, world * half_bytes) for _ in range(2)]
                box[0] = [capi.ipc_export(local_rank, p) for p in recv_ptr]
        except Exception as e:
            print("bench: CUDA IPC export failed (%s)" % e, file=sys.stderr)
            box[0] = None
        dist.broadcast_object_list(box, src=0)
        try:
            if box[0] is None:
                raise RuntimeError("no handles")
            if rank != 0:
                remote_ptr = [capi.ipc_open(local_rank, h) for h in box[0]]
        except Exception as e:
            print("bench: CUDA IPC open failed on rank %d (%s)" % (rank, e), file=sys.stderr)
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:      # every rank falls back together
            gather_mode = "nccl"
    if gather_mode == "nccl":
        gathered = [[torch.zeros((GROUP, H, W), dtype=torch.int32, device=dev) for _ in range(world)] for _ in range(2)] if rank == 0 else [None, None]

    warm = max(args.warmup, 3)
    # camera matrices are host-side inputs (Camera.View/Projection in the reference): built up front
    all_frames = [scene.frame(W, H, camera=orbit_camera(scenes, i * world + rank)) for i in range(warm + args.steps)]
    state = {"n": 0}

    def join_render_streams():
        for rs in rstreams:
            stream.wait_stream(rs)

    def flush_group(half):
        if world == 1:
            return
        join_render_streams()   # the transfer is ordered after every frame of the group
        if gather_mode == "nccl":
            pending[half] = dist.gather(ring[half], gathered[half], dst=0, async_op=True)
            return
        copy_stream.wait_stream(stream)
        dst = (recv_ptr if rank == 0 else remote_ptr)[half] + rank * half_bytes
        capi.copy_async(local_rank, dst, ring[half].data_ptr(), half_bytes, copy_stream.cuda_stream)
        with torch.cuda.stream(copy_stream):
            pending[half] = dist.all_reduce(flag, async_op=True)   # done => every rank's group has landed

    def step(i):
        k = state["n"]
        half, slot = (k // GROUP) & 1, k % GROUP
        if slot == 0 and pending[half] is not None:
            pending[half].wait()   # the render stream waits until this half's previous gather has finished
            pending[half] = None
        if INFLIGHT > 1:
            c = k % INFLIGHT
            if slot < INFLIGHT:
                rstreams[c].wait_stream(stream)   # buffer reuse / timing start are ordered through the main stream
            ctxs[c].render_frame_device(all_frames[i], ring[half, slot].data_ptr(), rstreams[c].cuda_stream)
        else:
            x.render_frame_device(all_frames[i], ring[half, slot].data_ptr(), stream.cuda_stream)
        state["n"] = k + 1
        if slot == GROUP - 1:
            flush_group(half)

    def drain():
        k = state["n"]
        if world > 1 and k % GROUP != 0:      # a partial group: ship it as it is (stale tail slots are ignored)
            flush_group((k // GROUP) & 1)
            state["n"] = k + (GROUP - k % GROUP)
        for b in range(2):
            if pending[b] is not None:
                pending[b].wait()
                pending[b] = None
        join_render_streams()

    def barrier():
        drain()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warm):
        step(i)
    barrier()
    rays_per_step = []
    launches0 = sum(xi.kernel_launches() for xi in ctxs)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        ev0.record(stream)
        for i in range(args.steps):
            step(warm + i)
        drain()              # the last gathers belong to the timed region
        ev1.record(stream)
        barrier()
    ms = ev0.elapsed_time(ev1)
    launches = sum(xi.kernel_launches() for xi in ctxs) - launches0
    gather_ok = None
    if world > 1 and rank == 0:
        # every rank's frames must have landed on rank 0: different cameras -> different, non-empty frames
        last = ((state["n"] // GROUP) - 1) & 1
        if gather_mode == "p2p":
            g = torch.zeros((world, GROUP, H, W), dtype=torch.int32, device=dev)
            capi.copy_async(local_rank, g.data_ptr(), recv_ptr[last], world * half_bytes, stream.cuda_stream)
            torch.cuda.synchronize()
        else:
            g = torch.stack(gathered[last])
        sums = [int(g[r, 0].to(torch.int64).bitwise_and(0xFFFFFF).sum().item()) for r in range(world)]
        gather_ok = all(v > 0 for v in sums) and len(set(sums)) == world
    # rays per frame (device counters are per render; re-render each timed camera outside the timed region)
    lvl_ms, res_ms = [], []
    for i in range(args.steps):
        x.render_frame_device(all_frames[warm + i], frame_dev.data_ptr(), None)
        c = x.counters()
        rays_per_step.append(c.rays)
        lvl_ms.append(c.msLevels)
        res_ms.append(c.msResolve)
    my_rays = float(sum(rays_per_step))
    t = torch.tensor([ms, my_rays], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, total_rays = float(tmax[0]), float(tsum[1])
    else:
        total_rays = my_rays
    value = total_rays / (ms * 1e-3) / 1e6

    # ---- e2e: the C-ABI with HOST buffers.  Headline: the flythrough call xnart_render_frames — K frame
    # descriptions in (184 B each), K finished frames out to page-locked host memory, everything inside the
    # timed region; the context keeps INFLIGHT frames in flight and overlaps the device->host copies with the
    # rendering of the next frames.  Beside it: K synchronous xnart_render_frame calls (one frame per call).
    frames = all_frames[warm:]
    E2E_BATCH = 32
    xe = capi.Xnart(devices=(local_rank,), frames_in_flight=INFLIGHT)
    scenes.upload(xe, scene)
    pin_n = capi.PinnedFrame(H, W, frames=max(4, min(E2E_BATCH, len(frames))))
    xe.render_frames(all_frames[:min(4, len(all_frames))], out=pin_n.array[:min(4, len(all_frames))])   # warm-up
    barrier()
    with ClockSampler(local_rank) as clocks2:
        t0 = time.perf_counter()
        for b0 in range(0, len(frames), E2E_BATCH):
            chunk = frames[b0:b0 + E2E_BATCH]
            xe.render_frames(chunk, out=pin_n.array[:len(chunk)])
        e2e_s = time.perf_counter() - t0
    checksum = int(pin_n.array[:len(frames[-E2E_BATCH:])].astype(np.uint64).sum()) if frames else 0
    # one frame per call
    pin = capi.PinnedFrame(H, W)
    for i in range(3):
        x.render_frame(all_frames[i], out=pin)
    t0 = time.perf_counter()
    for f in frames:
        x.render_frame(f, out=pin)
    e2e1_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s, e2e1_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = total_rays / float(te[0]) / 1e6
    e2e1_value = total_rays / float(te[1]) / 1e6
    clocks.rows += clocks2.rows  # both timed regions are short: pool their samples
    xe.close()
    pin_n.free()

    # ---- roofline of the dominant kernel -----------------------------------------------------------
    # Shallow scenes: k_level, one launch per recursion level (D+1 per frame) + the resolve launches; the levels
    # are timed on the device with CUDA events on the launching stream (XnartCounters.msLevels).  Deep scenes:
    # the persistent trace kernel k_trace, 2 launches per level.
    roof = None
    cpu = None
    deep = None
    if rank == 0:
        peaks, how = measured_peaks()
        fp32_peak = 148 * 128 * 1.965e9 / 1e12

        def roofline_of(xc, frame_list, kernel, launches_per_frame, traffic_key):
            lv, rs = [], []
            for f in frame_list:
                xc.render_frame_device(f, frame_dev.data_ptr(), None)
                c = xc.counters()
                lv.append(c.msLevels); rs.append(c.msResolve)
            xc.set_option(records.OPT_COUNTERS, 1)
            xc.render_frame_device(frame_list[0], frame_dev.data_ptr(), None)
            cc = xc.counters()
            xc.set_option(records.OPT_COUNTERS, 0)
            xc.render_frame_device(frame_list[0], frame_dev.data_ptr(), None)
            lvl0 = xc.counters().msLevels * 1e-3  # same frame as the counters
            ab, af = alg_bytes(cc), alg_flops(cc)
            achieved = ab / lvl0 / 1e9
            traffic = None
            tp = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(tp):
                try:
                    traffic = json.load(open(tp)).get(traffic_key)
                except Exception:
                    traffic = None
            return {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                    "traffic": traffic, "peak_source": how, "kernel": kernel, "launches_per_step": launches_per_frame,
                    "alg_bytes_per_launch": ab / launches_per_frame, "avg_launch_ms": 1e3 * lvl0 / launches_per_frame,
                    "alg_fp32_Tops": af / lvl0 / 1e12, "fp32_noFMA_peak_Tops_nominal": fp32_peak,
                    "fp32_frac_of_noFMA_nominal": af / lvl0 / 1e12 / fp32_peak,
                    "kernel_share_of_step": float(np.mean(lv)) / (float(np.mean(lv)) + float(np.mean(rs)) + 1e-12),
                    "work_per_ray": {"slab_tests": cc.nBox / max(1, cc.rays), "triangle_refs": cc.nTriRef / max(1, cc.rays)},
                    "note": "achieved = algorithmic bytes (SURVEY.md 8d, counted by the instrumented kernels == oracle) / device time of the "
                            "kernel's trace phase; the scene is cache resident, so the FP32 view is reported beside it"}, float(np.mean(lv)) + float(np.mean(rs)), cc

        if os.environ.get("XNART_BENCH_FUSE") == "1":
            roof, _, _ = roofline_of(x, all_frames[warm:warm + min(4, args.steps)], "k_frame<false> (levels phase of the fused frame kernel)", 1, "k_frame")
        else:
            roof, _, _ = roofline_of(x, all_frames[warm:warm + min(4, args.steps)], "k_level<false, 6, leaf-only, 128 threads> (one launch per recursion level)",
                                     scene.max_reflections + 1, "k_level")

        if world == 1 and not args.no_deep_scene:
            # second target of BASELINE.json's north_star: the 1M-triangle octree scene (configs[2] geometry) at 1080p
            gs = scenes.grid_scene(707)
            xg = capi.Xnart(devices=(local_rank,), frames_in_flight=INFLIGHT)
            scenes.upload(xg, gs)
            gframes = [gs.frame(W, H, camera=scenes.CameraData((np.sin(0.13 * i) * 80.0, 40.0, np.cos(0.13 * i) * 80.0), (0.0, 0.0, 0.0))) for i in range(12)]
            gbuf = torch.zeros((len(gframes), H, W), dtype=torch.int32, device=dev)
            xg.render_frames_device(gframes[:3], gbuf.data_ptr(), stream.cuda_stream)
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            g0.record(stream)
            xg.render_frames_device(gframes, gbuf.data_ptr(), stream.cuda_stream)   # INFLIGHT frames side by side
            g1.record(stream)
            torch.cuda.synchronize()
            gms = g0.elapsed_time(g1) / len(gframes)
            grays = []
            for f in gframes:
                xg.render_frame_device(f, frame_dev.data_ptr(), None)
                grays.append(xg.counters().rays)
            groof, _, gcc = roofline_of(xg, gframes[:3], "k_trace<false> (persistent trace kernel, 2 launches per level)",
                                        2 * (gs.max_reflections + 1), "k_trace")
            deep = {"workload": "grid_1M_triangles_1080p_depth%d" % gs.max_reflections, "triangles": gs.n_triangles,
                    "lights": len(gs.lights), "value": float(np.mean(grays)) / (gms * 1e-3) / 1e6, "unit": "Mrays/s",
                    "ms_per_frame": gms, "frames_per_s": 1e3 / gms, "rays_per_frame": float(np.mean(grays)), "frames_in_flight": INFLIGHT,
                    "roofline": groof}
            xg.close()

        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v_all, dt_all, _, n_all = cpu_baseline(scene, scenes, cores, (0, H), warm, frames=400, budget_s=10.0)
            v_one, dt_one, _, n_one = cpu_baseline(scene, scenes, 1, (0, H), warm, frames=20, budget_s=4.0)
            cpu = {"value": v_all, "unit": "Mrays/s", "cores": cores, "kind": "port",
                   "sample": "%d whole 1080p frames of the same workload (orbit cameras), %d threads, %.1f s" % (n_all, cores, dt_all),
                   "single_thread_as_shipped": {"value": v_one, "cores": 1, "seconds": dt_one, "frames": n_one,
                                                "note": "RayTracer.cs:99 hard-codes one worker"}}

    if rank == 0:
        line = {
            "metric": "Mrays/s (primary+secondary)", "value": value, "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "width": W, "height": H, "max_reflections": scene.max_reflections,
                       "triangles": scene.n_triangles, "lights": len(scene.lights), "frames_per_step_per_gpu": 1,
                       "parallelism": ("frames sharded across GPUs (one per GPU per step); every %d steps the group is shipped to rank 0 (%s), overlapping the next frames" % (GROUP, "device-to-device copies into an IPC-shared buffer of rank 0 + 1-element all-reduce" if gather_mode == "p2p" else "async NCCL gather")) if world > 1 else "single GPU",
                       "frames_in_flight": INFLIGHT,
                       "l2": "per-frame path arena (~0.9 GB) exceeds L2; a different camera every step"},
            "frames_per_s": world * args.steps / (ms * 1e-3),
            "rays_per_frame": total_rays / (world * max(1, args.steps)),
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 184, "d2h_bytes_per_step": W * H * 4,
                    "frames_per_s": world * args.steps / float(te[0]), "checksum": checksum,
                    "api": "xnart_render_frames (flythrough call, %d frames in flight per GPU, pinned host frames)" % INFLIGHT,
                    "one_frame_per_call": {"value": e2e1_value, "frames_per_s": world * args.steps / float(te[1]),
                                           "api": "xnart_render_frame, synchronous"}},
            "gpu_launches": int(launches),
            "gather_check": gather_ok,
            "clocks": clocks.summary(),
            "roofline": roof,
            "cpu_baseline": cpu,
            "deep_scene": deep,
        }
        print(json.dumps(line))
    pin.free()
    if gather_mode == "p2p":
        torch.cuda.synchronize()
        if rank != 0:
            for ptr in remote_ptr:
                capi.ipc_close(local_rank, ptr)
        dist.barrier()
        if rank == 0:
            for ptr in recv_ptr:
                capi.free_device(local_rank, ptr)
    for xi in ctxs:
        xi.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
