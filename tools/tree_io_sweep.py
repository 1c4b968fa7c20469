#!/usr/bin/env python
"""lb2_strip_tree wall time on a multi-GB /dev/shm tree (copies of the image's torch libraries) for different
numbers of I/O threads, slot and batch sizes.  Each configuration: fresh copy of the tree, fresh context (the
slot ring is sized at first use), one warm-up call on a second copy, then the timed call.

    python tools/tree_io_sweep.py > gpurun_out/tree_io_sweep.txt
"""
import os
import shutil
import sys
import sysconfig
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import copy_tree_parallel  # noqa: E402
from lambdipy_b200 import _native as N  # noqa: E402
from lambdipy_b200 import strip as S  # noqa: E402

sp = sysconfig.get_paths()["purelib"]
base = tempfile.mkdtemp(prefix="lb2_sweep_", dir="/dev/shm")
try:
    master = os.path.join(base, "master")
    for k in range(3):
        shutil.copytree(os.path.join(sp, "torch", "lib"), os.path.join(master, "t%d" % k),
                        ignore=lambda d, names: [n for n in names if not n.endswith(".so")])
    nbytes = sum(os.path.getsize(os.path.join(d, f)) for d, _, fs in os.walk(master) for f in fs)
    print("tree: %.2f GB, nproc %d" % (nbytes / 1e9, os.cpu_count()))
    for threads, slot, batch in [(16, 4, 1024), (32, 4, 1024), (64, 4, 1024), (96, 4, 1024), (64, 8, 1024), (64, 2, 1024), (64, 4, 256), (64, 4, 4096), (32, 16, 1024)]:
        os.environ.update(LB2_IO_THREADS=str(threads), LB2_TREE_SLOT_MB=str(slot), LB2_TREE_BATCH_MB=str(batch))
        with N.Context(0) as ctx:
            ts = []
            for rep in range(3):
                run = os.path.join(base, "run")
                shutil.rmtree(run, ignore_errors=True)
                copy_tree_parallel(master, run)
                t0 = time.perf_counter()
                st = S.strip_tree(run, ctx=ctx)
                ts.append(time.perf_counter() - t0)
            print("threads=%3d slot=%2d MB batch=%4d MB: %s s  best %.2f GB/s  (read_cpu %.2f write_cpu %.2f dma %.2f gpu %.3f, n_failed %d)" %
                  (threads, slot, batch, ["%.3f" % t for t in ts], nbytes / 1e9 / min(ts), st["read_cpu_s"], st["write_cpu_s"], st["dma_wait_s"], st["gpu_s"], st["n_failed"]), flush=True)
finally:
    shutil.rmtree(base, ignore_errors=True)
