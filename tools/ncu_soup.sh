#!/bin/bash
# ncu --set full of the persistent trace kernel on the 10M-triangle soup (configs[4]), 1080p: level-0 rays, level-0 shadow rays, level-1 rays
cd "$(dirname "$0")/.."
python tools/prof_frames.py soup 1920 1080 2 > gpurun_out/r2s_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:k_trace<\(bool\)0, \(bool\)0>' -s 18 -c 3 -o gpurun_out/r2s_ktrace_soup \
    python tools/prof_frames.py soup 1920 1080 2 > gpurun_out/r2s_ncu.log 2>&1
tail -2 gpurun_out/r2s_ncu.log; cat gpurun_out/r2s_plain.log
