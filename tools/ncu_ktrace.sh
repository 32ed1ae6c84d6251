#!/bin/bash
# ncu --set full of the whole-ray build of the persistent trace kernel on a 1080p frame of the 1M-triangle grid:
# the first three launches of the second frame (level-0 rays, level-0 shadow rays, level-1 rays).
cd "$(dirname "$0")/.."
python tools/prof_frames.py grid 1920 1080 2 > gpurun_out/r3r_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:k_trace<\(bool\)0, \(bool\)0>' -s 10 -c 3 -o gpurun_out/r3r_ktrace \
    python tools/prof_frames.py grid 1920 1080 2 > gpurun_out/r3r_ncu.log 2>&1
tail -2 gpurun_out/r3r_ncu.log
