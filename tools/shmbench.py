import os, mmap, time, numpy as np, sys
from concurrent.futures import ThreadPoolExecutor
d = sys.argv[1] if len(sys.argv) > 1 else "/dev/shm"
N, SZ = 200, 6220800
src = [np.random.randint(0, 255, SZ, dtype=np.uint8) for _ in range(4)]
def run(mode, threads):
    path = os.path.join(d, "shmbench.bin")
    if os.path.exists(path): os.unlink(path)
    fd = os.open(path, os.O_RDWR | os.O_CREAT, 0o644)
    t0 = time.perf_counter()
    if mode == "write":
        for i in range(N): os.write(fd, src[i & 3])
    elif mode == "pwrite":
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(lambda i: os.pwrite(fd, src[i & 3], i * (SZ + 8) + 8), range(N)))
    elif mode == "mmap":
        os.ftruncate(fd, N * (SZ + 8) + 8)
        PG = mmap.ALLOCATIONGRANULARITY
        def w(i):
            pos = i * (SZ + 8) + 8
            a = pos // PG * PG
            mm = mmap.mmap(fd, pos - a + SZ, offset=a)
            np.frombuffer(mm, dtype=np.uint8)[pos - a: pos - a + SZ] = src[i & 3]
            mm.close() if False else None
            return mm
        with ThreadPoolExecutor(threads) as ex:
            mms = list(ex.map(w, range(N)))
        del mms
    dt = time.perf_counter() - t0
    os.close(fd); os.unlink(path)
    print("%-7s threads %d: %.2f GB/s" % (mode, threads, N * SZ / dt / 1e9), flush=True)
run("write", 1)
for t in (1, 4, 8): run("pwrite", t)
for t in (1, 4, 8, 16): run("mmap", t)
for t in (8, 16): run("mmap", t)
