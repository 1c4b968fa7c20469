#!/usr/bin/env python
"""Per-file plan-kernel latency on real wheels (one file per launch): finds what the serial planner is slow on."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import elf_fixtures as F
from lambdipy_b200 import _native as N
from lambdipy_b200.device import DeviceBatch
import struct

def info(b):
    shoff, = struct.unpack_from("<Q", b, 0x28); shnum, shstr = struct.unpack_from("<HH", b, 0x3c); phnum, = struct.unpack_from("<H", b, 0x38)
    so, ss = struct.unpack_from("<QQ", b, shoff + shstr * 64 + 24); names = b[so:so + ss]
    notes = 0
    for i in range(shnum):
        n, = struct.unpack_from("<I", b, shoff + i * 64); sz, = struct.unpack_from("<Q", b, shoff + i * 64 + 32)
        if names[n:names.index(b"\0", n)].startswith(b".gnu.build.attributes"): notes += sz
    return shnum, phnum, notes

ctx = N.Context(0)
rows = []
for kind in ("wheels", "torch"):
    for p in F.real_corpus(kind):
        if not p.endswith(".so"): continue
        blob = open(p, "rb").read()
        b = DeviceBatch.from_blobs(ctx, [blob])
        ts = []
        for _ in range(6):
            b.strip_async(); st = b.results(); ts.append(st["plan_ms"])
        shnum, phnum, notes = info(blob)
        rows.append((float(np.median(ts[2:])), len(blob), shnum, phnum, notes, st["n_tiles"], os.path.basename(p)))
        b.close()
rows.sort(reverse=True)
for r in rows[:25]: print("%.3f ms size=%d shnum=%d phnum=%d notes=%d tiles=%d %s" % r)
print("median %.3f ms" % np.median([r[0] for r in rows]))
for r in rows[-5:]: print("%.3f ms size=%d shnum=%d phnum=%d notes=%d tiles=%d %s" % r)
