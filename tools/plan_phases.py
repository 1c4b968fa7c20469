#!/usr/bin/env python
"""Per-phase cycle counts of the plan kernel (clock64 at the CTA-level sync points and at the end of each
specialised warp's job) for three representative real files, one file per launch.  Needs the diagnostic build:

    python -m lambdipy_b200.build --timing
    LAMBDIPY_B200_LIB=lambdipy_b200/liblambdipy_b200_timing.so python tools/plan_phases.py
"""
import glob
import os
import sys
import sysconfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lambdipy_b200 import _native as N  # noqa: E402
from lambdipy_b200.device import DeviceBatch  # noqa: E402

sp = sysconfig.get_paths()["purelib"]
picks = []
for pat in ("numpy/_core/_simd*.so", "numpy/random/_pcg64*.so", "scipy/fft/_pocketfft/pypocketfft*.so", "scipy/fft/_pocketfft/*.so",
            "torch/lib/libtorch_cuda.so", "torch/lib/libtorch_cpu.so"):
    g = sorted(glob.glob(os.path.join(sp, pat)))
    if g and g[0] not in picks:
        picks.append(g[0])
ctx = N.Context(0)
for p in picks:
    blob = open(p, "rb").read()
    b = DeviceBatch.from_blobs(ctx, [blob])
    for k in range(3):
        if k == 2:
            print("== %s (%d bytes)" % (os.path.relpath(p, sp), len(blob)), flush=True)
            os.environ["LB2_DUMMY"] = "1"
        b.strip_async()
        st = b.results()
    print("   plan_ms=%.4f compact_ms=%.4f tiles=%d" % (st["plan_ms"], st["compact_ms"], st["n_tiles"]), flush=True)
    b.close()
ctx.close()
