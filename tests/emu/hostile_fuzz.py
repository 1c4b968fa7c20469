#!/usr/bin/env python
"""Hostile-input fuzz of the PRODUCT's device planner (through the CPU warp emulator) and of the
oracle: random corruption of Ehdr / section-header / program-header / string-table / note bytes with
extreme values.  Properties: neither implementation crashes or hangs; whenever the emulated planner
accepts a file the oracle accepts it too and the bytes are identical; a tile list that does not
cover the output exactly once is an error.

    python tests/emu/hostile_fuzz.py --cases 2000 --seed 1
Each case runs in a worker process so that a crash is caught and reported with its seed."""
import argparse
import multiprocessing as mp
import os
import random
import struct
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

EXTREME = [0, 1, 2, 3, 7, 8, 15, 16, 23, 24, 63, 64, 65, 255, 256, 4095, 4096, 0x7fff, 0x8000, 0xffff, 0x10000, 0x7fffffff, 0x80000000,
           0xfffffff0, 0xfffffffc, 0xffffffff, 0x100000000, 0x7fffffffffffffff, 0x8000000000000000, 0xfffffffffffffff0, 0xffffffffffffffff]


def corrupt(rng, data, log=None):
    b = bytearray(data)
    n = len(b)
    shoff, = struct.unpack_from("<Q", b, 0x28)
    shnum, shstr = struct.unpack_from("<HH", b, 0x3c)
    phnum, = struct.unpack_from("<H", b, 0x38)
    regions = [(0x10, 0x30)]  # Ehdr after e_ident
    if shoff + shnum * 64 <= n:
        regions += [(shoff, shnum * 64)] * 4
        so, ss = struct.unpack_from("<QQ", b, shoff + shstr * 64 + 24) if shstr < shnum else (0, 0)
        if so + ss <= n and ss:
            regions.append((so, ss))
    if phnum and 64 + phnum * 56 <= n:
        regions += [(64, phnum * 56)] * 2
    for _ in range(rng.choice([1, 1, 1, 2, 3, 6])):
        base, ln = rng.choice(regions)
        how = rng.random()
        if how < 0.5:
            w = rng.choice([2, 4, 8])
            off = base + rng.randrange(max(1, ln - w + 1))
            off -= off % w if rng.random() < 0.8 else 0
            v = rng.choice(EXTREME) + rng.choice([0, 0, 0, -1, 1, n, -n, n // 2])
            b[off:off + w] = (v & ((1 << (8 * w)) - 1)).to_bytes(w, "little")
            if log is not None: log.append(("set", base, off - base, w, v & ((1 << (8 * w)) - 1)))
        elif how < 0.8:
            off = base + rng.randrange(ln)
            b[off] ^= 1 << rng.randrange(8)
            if log is not None: log.append(("flip", base, off - base))
        else:
            off = base + rng.randrange(ln)
            b[off] = rng.randrange(256)
            if log is not None: log.append(("byte", base, off - base))
    if rng.random() < 0.1:
        b = b[: rng.randrange(64, n)]
        if log is not None: log.append(("truncate", len(b)))
    return bytes(b)


_state = {}


def _init(gnu=False):
    import emu_lib
    import oracle_lib
    _state["emu"] = emu_lib.load()
    _state["oracle"] = oracle_lib.load()
    _state["gnu"] = gnu


def _one(args):
    seed, k, path = args
    rng = random.Random(seed * 7919 + k)
    with open(path, "rb") as f:
        data = f.read()
    bad = corrupt(rng, data)
    st, got = _state["emu"].strip(bad)
    rc, want = _state["oracle"].strip(bad)
    if rc == 0 and _state.get("gnu"):
        # the parity property on hostile inputs: whatever the oracle accepts, GNU strip accepts and agrees
        import tempfile
        import elf_fixtures as F
        with tempfile.TemporaryDirectory(dir="/dev/shm") as d:
            pth = os.path.join(d, "h.so")
            with open(pth, "wb") as f:
                f.write(bad)
            gnu, err = F.gnu_strip_bytes(pth, d)
        if gnu is None:
            return k, "ORACLE-ACCEPTS-GNU-REFUSES", st, rc
        if gnu != want:
            return k, "ORACLE-DIFFERS-FROM-GNU", st, rc
    if st <= -1000:
        return k, "TILES-DO-NOT-TILE", st, rc
    if st == 0 and (rc != 0 or got != want):
        return k, "EMU-OK-BUT-DIFFERENT", st, rc
    if st != 0 and rc == 0 and st != 8:
        return k, "oracle-accepts-emu-declines", st, rc   # allowed only for planner limits; report
    return k, "ok" if st == 0 else "declined", st, rc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=1000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--jobs", type=int, default=8)
    ap.add_argument("--gnu", action="store_true", help="also run the real GNU strip on everything the oracle accepts")
    a = ap.parse_args()
    import elf_fixtures as F
    import tempfile
    base = tempfile.mkdtemp(prefix="lb2host_")
    v = F.build_variants(os.path.join(base, "fx"))
    seeds = [v[k] for k in sorted(v) if k not in ("c_maxpage_2m", "c_static", "c_static_pie", "c_many_sections")]
    for k, notes in F.note_scenarios().items():
        p = os.path.join(base, "fx", k + ".so")
        if F.with_build_notes(v["c_plain"], p, notes):
            seeds.append(p)
    seeds += [p for p in F.real_corpus("small") if os.path.getsize(p) < 1_000_000][:15]
    rng = random.Random(a.seed)
    jobs = [(a.seed, k, rng.choice(seeds)) for k in range(a.cases)]
    counts = {}
    problems = []
    import emu_lib, oracle_lib
    emu_lib.build(); oracle_lib.build()
    done = set()
    while len(done) < len(jobs):
        todo = [j for j in jobs if j[1] not in done]
        try:
            with mp.get_context("fork").Pool(a.jobs, initializer=_init, initargs=(a.gnu,), maxtasksperchild=200) as pool:
                for k, st, s1, s2 in pool.imap_unordered(_one, todo, chunksize=1):
                    done.add(k)
                    counts[st] = counts.get(st, 0) + 1
                    if st not in ("ok", "declined"):
                        problems.append((k, st, s1, s2))
                        print(st, "case", k, "emu", s1, "oracle", s2, os.path.basename(jobs[k][2]))
        except Exception as e:  # a worker died: find the case by running the rest one by one
            print("worker failure:", type(e).__name__, e)
            break
    print("SUMMARY", dict(sorted(counts.items())), "unfinished", len(jobs) - len(done))
    return 1 if problems or len(done) < len(jobs) else 0


if __name__ == "__main__":
    sys.exit(main())
