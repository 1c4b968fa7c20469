#!/usr/bin/env python
"""Differential check of the CPU restatement against the real GNU strip binary.

usage: python oracle/diffcheck.py [--no-merge] [--jobs N] PATH...   (files or directories)
Prints one line per mismatch and a summary.  Test infrastructure (see strip_oracle.c).
"""
import argparse, ctypes, os, subprocess, sys, tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))


def load():
    lib = ctypes.CDLL(os.path.join(HERE, "_build", "libstrip_oracle.so"))
    lib.lbo_strip.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_void_p),
                              ctypes.POINTER(ctypes.c_uint64), ctypes.c_uint]
    lib.lbo_strip.restype = ctypes.c_int
    lib.lbo_free.argtypes = [ctypes.c_void_p]
    return lib


LIB = load()


def oracle_strip(data: bytes, no_merge=False):
    out = ctypes.c_void_p()
    n = ctypes.c_uint64()
    rc = LIB.lbo_strip(data, len(data), ctypes.byref(out), ctypes.byref(n), 1 if no_merge else 0)
    if rc != 0:
        return rc, None
    res = ctypes.string_at(out.value, n.value)
    LIB.lbo_free(out)
    return 0, res


def gnu_strip(path, no_merge=False):
    with tempfile.NamedTemporaryFile(dir="/dev/shm", delete=False) as t:
        tmp = t.name
    try:
        cmd = ["strip", "--strip-unneeded"] + (["--no-merge-notes"] if no_merge else []) + ["-o", tmp, path]
        r = subprocess.run(cmd, capture_output=True)
        if r.returncode != 0:
            return r.returncode, None, r.stderr.decode(errors="replace")
        with open(tmp, "rb") as f:
            return 0, f.read(), r.stderr.decode(errors="replace")
    finally:
        if os.path.exists(tmp):
            os.unlink(tmp)


def first_diff(a, b):
    m = min(len(a), len(b))
    if a[:m] == b[:m]:
        return m
    lo, hi = 0, m
    while hi - lo > 1:
        mid = (lo + hi) // 2
        if a[:mid] == b[:mid]:
            lo = mid
        else:
            hi = mid
    return lo


def check(path, no_merge):
    with open(path, "rb") as f:
        data = f.read()
    if data[:4] != b"\x7fELF":
        return path, "skip-nonelf", ""
    grc, gout, gerr = gnu_strip(path, no_merge)
    orc, oout = oracle_strip(data, no_merge)
    if grc != 0:
        return path, ("ok-both-reject" if orc != 0 else "GNU-rejects"), gerr.strip()[:100]
    if orc != 0:
        return path, "unsupported", "rc=%d" % orc
    if gout == oout:
        return path, "ok", ""
    return path, "MISMATCH", "len gnu=%d oracle=%d first diff @0x%x" % (len(gout), len(oout), first_diff(gout, oout))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-merge", action="store_true")
    ap.add_argument("--jobs", type=int, default=8)
    ap.add_argument("-v", action="store_true")
    ap.add_argument("--all", action="store_true", help="every regular file, not only *.so*")
    ap.add_argument("paths", nargs="+")
    a = ap.parse_args()
    files = []
    for p in a.paths:
        if os.path.isdir(p):
            for d, _, fs in os.walk(p):
                for f in fs:
                    if (a.all or ".so" in f) and not os.path.islink(os.path.join(d, f)):
                        files.append(os.path.join(d, f))
        else:
            files.append(p)
    files.sort()
    counts = {}
    with ThreadPoolExecutor(a.jobs) as ex:
        for path, st, msg in ex.map(lambda f: check(f, a.no_merge), files):
            counts[st] = counts.get(st, 0) + 1
            if st not in ("ok",) or a.v:
                print(st, path, msg)
    print("SUMMARY", counts)
    return 0 if not counts.get("MISMATCH") else 1


if __name__ == "__main__":
    sys.exit(main())
