#!/usr/bin/env python
"""tools/ncu_summary.py — turn ncu output brought back from the GPU box into the small, tracked
summaries under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_launches.md
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/r01_k_level_full.md
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio",
    "local_load_bytes", "local_store_bytes", "smsp__inst_executed_op_ldl.sum", "smsp__inst_executed_op_stl.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, rows = rows[start], rows[start + 1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
    tot, cnt = collections.defaultdict(float), collections.Counter()
    seq = []
    for r in rows:
        name = r[ki].split("(")[0].strip()
        v = float(r[vi].replace(",", "")) / 1e3
        tot[name] += v
        cnt[name] += 1
        seq.append((name, v, r[gi], r[bi]))
    T = sum(tot.values())
    out = ["# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)", "",
           "source: `%s`, %d launches, %.1f us total" % (src, len(rows), T), "",
           "| kernel | launches | total us | share | avg us |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        out.append("| `%s` | %d | %.1f | %.3f | %.1f |" % (k, cnt[k], v, v / T, v / cnt[k]))
    out += ["", "First launches in order (one frame = the level kernels then the resolve kernels):", "", "```"]
    for name, v, g, b in seq[:40]:
        out.append("%-40s %8.1f us  grid %s block %s" % (name[:40], v, g, b))
    out.append("```")
    open(dst, "w").write("\n".join(out) + "\n")


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = ["# ncu --set full summary", "", "source: `%s` (%d launches captured)" % (src, len(data)), ""]
    out.append("| metric | unit | " + " | ".join("launch %d" % i for i in range(len(data))) + " |")
    out.append("|---|---|" + "---:|" * len(data))
    ki = hdr.index("Kernel Name")
    out.append("| Kernel Name | | " + " | ".join("`%s`" % d[ki].split("(")[0] for d in data) + " |")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            out.append("| %s | %s | %s |" % (k, units[i], " | ".join(d[i] for d in data)))
    # hottest CUDA source lines by warp-stall samples (needs -lineinfo at compile time)
    srcp = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    out += ["", "## hottest source lines (warp stall samples and executed warp instructions per CUDA line, first captured launch)", "", "```"]
    try:
        rd = list(csv.reader(io.StringIO(srcp)))
        agg = collections.defaultdict(lambda: [0, 0])
        fname, seen_kernels = "?", 0
        h = None
        for r in rd:
            if len(r) >= 2 and r[0] == "File Path":
                fname = r[1].split("/")[-1]
                continue
            if len(r) >= 2 and r[0] == "Function Name":
                seen_kernels += 1 if (h is not None and False) else 0
                continue
            if len(r) > 5 and r[0] == "Line No":
                h = r
                ci, ii = h.index("# Samples"), h.index("Instructions Executed")
                continue
            if h is None or len(r) <= max(ci, ii):
                continue
            if r[2] != "-":
                continue  # SASS row; the CUDA row carries the per-line totals
            try:
                agg[(fname, int(r[0]), r[1].strip())][0] += int(float(r[ci] or 0))
                agg[(fname, int(r[0]), r[1].strip())][1] += int(float(r[ii] or 0))
            except ValueError:
                continue
        total = sum(v[0] for v in agg.values()) or 1
        tinst = sum(v[1] for v in agg.values()) or 1
        out.append("samples%  inst%   file:line  source")
        for (fn, ln, text), (n, ins) in sorted(agg.items(), key=lambda x: -x[1][0])[:30]:
            out.append("%6.2f%% %6.2f%%  %s:%d  %s" % (100.0 * n / total, 100.0 * ins / tinst, fn, ln, text[:120]))
    except Exception as e:  # the source page layout differs between ncu versions
        out.append("(source page not parsed: %r)" % (e,))
    out.append("```")
    open(dst, "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
