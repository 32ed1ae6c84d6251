#!/usr/bin/env python
"""tools/ncu_regions.py — per-region view of one kernel launch from an ncu report's source page:
executed warp instructions, average active threads and stall samples per block of SASS instructions,
plus the top memory instructions.  Development aid for the trace kernels.

  python tools/ncu_regions.py gpurun_out/prof.ncu-rep [launch-index] [block]
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    step = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    print(rows[0][1] if rows and len(rows[0]) > 1 else "?")
    hdr = rows[1]
    H = {h: i for i, h in enumerate(hdr)}
    seen, data = set(), []
    for r in rows[2:]:
        if len(r) != len(hdr) or r[0] in seen:
            continue
        seen.add(r[0])
        data.append(r)

    def f(r, k):
        try:
            return int(r[H[k]])
        except Exception:
            return 0

    tot = sum(f(r, "Instructions Executed") for r in data)
    tt = sum(f(r, "Thread Instructions Executed") for r in data)
    print("%d SASS instructions; warp-inst %d thread-inst %d avg active threads %.2f" % (len(data), tot, tt, tt / max(tot, 1)))
    for i in range(0, len(data), step):
        ch = data[i:i + step]
        wi = sum(f(r, "Instructions Executed") for r in ch)
        ti = sum(f(r, "Thread Instructions Executed") for r in ch)
        smp = sum(f(r, "# Samples") for r in ch)
        marks = set()
        for r in ch:
            t = r[H["Source"]].strip().split()
            if not t:
                continue
            op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
            if any(x in op for x in ("256", "STL.128", "LDL", "ATOM", "VOTE", "MUFU", "WARPSYNC", "LDG", "STG")):
                marks.add(op)
        print("%4d-%4d  warp-inst %6.2f%%  avg-thr %5.1f  samples %7d  %s" % (i, i + step, 100.0 * wi / max(tot, 1), ti / max(wi, 1), smp, " ".join(sorted(marks))))
    mem = [r for r in data if any(x in r[H["Source"]] for x in ("LDL", "STL", "LDG", "STG", "ATOM"))]
    mem.sort(key=lambda r: -f(r, "Instructions Executed"))
    print("\ntop memory instructions")
    for r in mem[:14]:
        print("  %-62s inst %10d avg-thr %5s" % (r[H["Source"]].strip()[:62], f(r, "Instructions Executed"), r[H["Avg. Threads Executed"]][:5]))


if __name__ == "__main__":
    main()
