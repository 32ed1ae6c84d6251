#!/usr/bin/env python
"""How fast the AVI sink takes 600 1080p 24bpp frames (3.73 GB) on this box, by number of payload threads (no GPU).
  python tools/avibench.py [/dev/shm]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from xna_ray_trace_b200 import avi

d = sys.argv[1] if len(sys.argv) > 1 else "/dev/shm"
W, H, N = 1920, 1080, 600
bits = [np.random.randint(0, 255, W * H * 3, dtype=np.uint8) for _ in range(4)]
for threads in (0, 4, 8, 12):
    path = os.path.join(d, "avibench_%d.avi" % threads)
    t0 = time.perf_counter()
    m = avi.AviManager(path, False, io_threads=threads)
    s = m.AddVideoStream(False, 30, 0, W, H, avi.PIXELFORMAT_24BPP_RGB)
    for k in range(N):
        s.AddFrameBits(bits[k & 3])
    m.Close()
    dt = time.perf_counter() - t0
    size = os.path.getsize(path)
    os.unlink(path)
    print("io_threads %2d: %.2f s, %.2f GB/s (%d bytes)" % (threads, dt, size / dt / 1e9, size), flush=True)
