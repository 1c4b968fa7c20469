#!/usr/bin/env python
"""Aggregate host memory copy bandwidth of the box (all cores), to put the N=8 host-buffer (e2e) figure in context:
every e2e byte is read from or written to host DRAM by a PCIe device.  P processes (at most 64) each copy a private 256 MiB
buffer back and forth (at most 32 GiB of RAM in total); reports GB/s of read + write traffic.   python tools/host_membw.py [procs]"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np


def work(args):
    reps, barrier_t = args
    a = np.ones(1 << 25, dtype=np.uint64)  # 256 MiB
    b = np.empty_like(a)
    np.copyto(b, a)
    while time.time() < barrier_t:
        pass
    t0 = time.perf_counter()
    for _ in range(reps):
        np.copyto(b, a)
    return time.perf_counter() - t0, 2 * reps * a.nbytes


if __name__ == "__main__":
    for procs in ([min(64, int(sys.argv[1]))] if len(sys.argv) > 1 else [8, 32, 64]):
        with mp.Pool(procs) as pool:
            start = time.time() + 6
            res = pool.map(work, [(24, start)] * procs)
        total = sum(r[1] for r in res)
        print("%3d processes: %.1f GB/s read+write (slowest %.2f s)" % (procs, total / 1e9 / max(r[0] for r in res), max(r[0] for r in res)), flush=True)
