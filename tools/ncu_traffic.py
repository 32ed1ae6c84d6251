#!/usr/bin/env python
"""tools/ncu_traffic.py — profiles/traffic.json from `ncu --set full` reports of the SHIPPED kernels.

  python tools/ncu_traffic.py k_path=gpurun_out/r2x_kpath.ncu-rep k_trace=gpurun_out/r2x_ktrace.ncu-rep [share:k_trace=0.97]

For every key the report's captured launches are averaged: DRAM bytes per launch (`roofline.traffic` of bench.py)
and the utilisation of the units a trace kernel can run out of (issue slots, L1 data pipe, L2, DRAM) — the evidence
for `roofline.bound`.  Existing keys of other kernels are kept.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "traffic.json")

METRICS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "lanes_per_instruction",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_data_pipe_pct",
    "lts__t_sectors.avg.pct_of_peak_sustained_elapsed": "l2_sectors_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "launch__registers_per_thread": "registers",
    "sm__icc_request_hit_rate.pct": "icache_hit_pct",
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def read(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = {}
    names = set()
    for name, key in METRICS.items():
        if name not in hdr:
            continue
        i = hdr.index(name)
        vals = []
        for r in data:
            try:
                vals.append(float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0))
            except ValueError:
                pass
        if vals:
            out[key] = sum(vals) / len(vals)
    if "Kernel Name" in hdr:
        k = hdr.index("Kernel Name")
        names = sorted({r[k] for r in data})
    out["launches_captured"] = len(data)
    out["kernel_names"] = list(names)
    out["dram_bytes_per_launch"] = out.get("dram_read", 0.0) + out.get("dram_write", 0.0)
    out["source"] = os.path.relpath(rep, ROOT)
    return out


def main():
    cur = {}
    if os.path.exists(OUT):
        try:
            cur = json.load(open(OUT))
        except Exception:
            cur = {}
    for a in sys.argv[1:]:
        k, _, v = a.partition("=")
        if k.startswith("share:"):
            cur.setdefault(k[6:], {})["share_of_frame"] = float(v)
        else:
            keep = cur.get(k, {}).get("share_of_frame")
            cur[k] = read(v)
            if keep is not None:
                cur[k]["share_of_frame"] = keep
    json.dump(cur, open(OUT, "w"), indent=1, sort_keys=True)
    print(json.dumps(cur, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
