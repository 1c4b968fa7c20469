#!/usr/bin/env python
"""Structure fuzzer: mutate real/fixture ELF files (header fields, section flags/alignments/entsizes,
added/removed/renamed sections via objcopy, program-header fields) and compare the CPU restatement
with the real GNU strip on every mutant GNU strip accepts.  A MISMATCH (oracle says ok but bytes
differ) is a bug in the rules; 'unsupported' is fine (the product hands those to host strip).

usage: python oracle/fuzz_vs_gnu.py [--cases N] [--seed S] [--jobs J] [--keep DIR]
Test infrastructure (see strip_oracle.c)."""
import argparse, os, random, shutil, struct, subprocess, sys, tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
import elf_fixtures as F  # noqa: E402
import oracle_lib  # noqa: E402


def shdrs(b):
    shoff, = struct.unpack_from("<Q", b, 0x28)
    shnum, shstr = struct.unpack_from("<HH", b, 0x3c)
    return shoff, shnum, shstr


ONLY_KINDS = None


def mutate(rng, src, dst, tmp):
    """returns a description or None if the mutation could not be applied"""
    kind = rng.choice(ONLY_KINDS or ["align", "entsize", "flags", "objcopy_add", "objcopy_remove", "objcopy_rename", "phdr", "link_info", "type", "addr", "multi",
                       "notes_content", "notes_content"])
    shutil.copy(src, dst)
    with open(dst, "rb") as f:
        b = bytearray(f.read())
    shoff, shnum, shstr = shdrs(b)
    if kind == "notes_content":
        # corrupt / perturb the bytes of a build-attribute note section (sizes, types, ranges, names)
        so, ss = struct.unpack_from("<QQ", b, shoff + shstr * 64 + 24)
        names = bytes(b[so:so + ss])
        cand = []
        for k in range(1, shnum):
            n, t = struct.unpack_from("<II", b, shoff + k * 64)
            nm = names[n:names.index(b"\0", n)]
            off, sz = struct.unpack_from("<QQ", b, shoff + k * 64 + 24)
            if nm.startswith(b".gnu.build.attributes") and t == 7 and sz >= 12:
                cand.append((off, sz))
        if not cand:
            return None
        off, sz = rng.choice(cand)
        for _ in range(rng.choice([1, 1, 2, 4])):
            w = off + 4 * rng.randrange(sz // 4)
            how = rng.random()
            if how < 0.4:
                struct.pack_into("<I", b, w, rng.choice([0, 1, 2, 3, 4, 7, 8, 12, 16, 20, 0x100, 0x101, 0x102, 0xffffffff]))
            elif how < 0.7:
                b[w + rng.randrange(4)] ^= 1 << rng.randrange(8)
            else:
                v, = struct.unpack_from("<I", b, w)
                struct.pack_into("<I", b, w, (v + rng.choice([-16, -1, 1, 5, 16, 0x1000])) & 0xffffffff)
        with open(dst, "wb") as f:
            f.write(b)
        return "notes_content @%#x+%d" % (off, sz)
    phnum, = struct.unpack_from("<H", b, 0x38)
    if shnum < 3:
        return None
    i = rng.randrange(1, shnum)
    o = shoff + i * 64
    desc = kind
    if kind == "align":
        v = rng.choice([0, 1, 2, 3, 4, 8, 16, 24, 32, 64, 128, 4096, 1 << 16, 1 << 21, 6, 12])
        struct.pack_into("<Q", b, o + 48, v); desc += " sec%d=%d" % (i, v)
    elif kind == "entsize":
        v = rng.choice([0, 1, 2, 4, 8, 16, 24, 0x30, 7])
        struct.pack_into("<Q", b, o + 56, v); desc += " sec%d=%d" % (i, v)
    elif kind == "flags":
        fl, = struct.unpack_from("<Q", b, o + 8)
        bit = rng.choice([0x1, 0x4, 0x10, 0x20, 0x40, 0x80, 0x100, 0x200, 0x800, 0x80000000, 0x10000000])
        struct.pack_into("<Q", b, o + 8, fl ^ bit); desc += " sec%d^=%#x" % (i, bit)
    elif kind == "link_info":
        struct.pack_into("<I", b, o + 40 + 4 * rng.randrange(2), rng.choice([0, 1, 2, shnum - 1, shnum, 5, 9])); desc += " sec%d" % i
    elif kind == "type":
        t = rng.choice([1, 7, 8, 14, 15, 16, 0x6ffffff5, 0x70000001, 0x60000000, 6, 5])
        fl, = struct.unpack_from("<Q", b, o + 8)
        if fl & 2 and rng.random() < 0.7:
            return None  # retyping alloc sections mostly just makes BFD refuse the file
        struct.pack_into("<I", b, o + 4, t); desc += " sec%d=%#x" % (i, t)
    elif kind == "addr":
        fl, = struct.unpack_from("<Q", b, o + 8)
        if fl & 2:
            return None
        struct.pack_into("<Q", b, o + 16, rng.choice([0, 0x1000, 0x603020, 0x7, 0x100000001])); desc += " sec%d" % i
    elif kind == "phdr":
        if not phnum:
            return None
        j = rng.randrange(phnum)
        po = 64 + j * 56
        field = rng.choice(["flags", "align", "paddr", "memsz"])
        if field == "flags":
            struct.pack_into("<I", b, po + 4, rng.choice([0, 4, 5, 6, 7]))
        elif field == "align":
            struct.pack_into("<Q", b, po + 48, rng.choice([0, 1, 8, 0x10, 0x1000, 0x10000, 0x200000]))
        elif field == "paddr":
            struct.pack_into("<Q", b, po + 24, rng.choice([0, 0x1000, 0xdead000]))
        else:
            t, = struct.unpack_from("<I", b, po)
            if t == 1:
                return None
            ms, = struct.unpack_from("<Q", b, po + 40)
            struct.pack_into("<Q", b, po + 40, ms + rng.choice([1, 8, 0x40]))
        desc += " ph%d.%s" % (j, field)
    elif kind.startswith("objcopy") or kind == "multi":
        blob = os.path.join(tmp, "blob")
        with open(blob, "wb") as f:
            f.write(bytes(rng.randrange(256) for _ in range(rng.choice([1, 7, 64, 1000, 5000]))))
        names = [".comment", ".note.GNU-stack", ".gnu_debuglink", ".debug_info", ".debug_str", ".symtab", ".strtab", ".gnu.build.attributes",
                 ".note.gnu.build-id", ".gnu.hash", ".eh_frame_hdr", ".data", ".gnu.version", ".note.gnu.property"]
        newn = [".lb2fuzz", ".debug_lb2", ".zdebug_x", ".stabx", ".linefoo", ".gdb_index", ".comment.extra", ".note.lb2", ".gnu.linkonce.wi.x", ".gnu.debuglto_.debug_a",
                ".gnu_debugdata", "lb2noprefix", ".rela.lb2", ".shstrtab2"]
        args = []
        n_ops = 1 if kind != "multi" else rng.randrange(2, 5)
        for _ in range(n_ops):
            op = kind if kind != "multi" else rng.choice(["objcopy_add", "objcopy_remove", "objcopy_rename"])
            if op == "objcopy_add":
                nm = rng.choice(newn)
                args += ["--add-section", "%s=%s" % (nm, blob)]
                if rng.random() < 0.6:
                    args += ["--set-section-flags", "%s=%s" % (nm, rng.choice(["readonly", "debug", "noload,readonly", "data", "contents,readonly", "readonly,merge,strings"]))]
                if rng.random() < 0.5:
                    args += ["--set-section-alignment", "%s=%d" % (nm, rng.choice([1, 2, 4, 8, 16, 64, 4096]))]
            elif op == "objcopy_remove":
                args += ["--remove-section", rng.choice(names)]
            else:
                args += ["--rename-section", "%s=%s" % (rng.choice(names), rng.choice(newn))]
        r = subprocess.run(["objcopy"] + args + [src, dst], capture_output=True)
        if r.returncode != 0:
            return None
        return desc + " " + " ".join(args).replace(blob, "BLOB")
    with open(dst, "wb") as f:
        f.write(b)
    return desc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=500)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--jobs", type=int, default=8)
    ap.add_argument("--keep", default=None)
    ap.add_argument("--kinds", default=None, help="comma-separated subset of mutation kinds")
    a = ap.parse_args()
    global ONLY_KINDS
    if a.kinds:
        ONLY_KINDS = a.kinds.split(",")
    oracle = oracle_lib.load()
    base = tempfile.mkdtemp(prefix="lb2fuzz_", dir="/dev/shm")
    variants = F.build_variants(os.path.join(base, "fx"))
    seeds = [variants[k] for k in sorted(variants) if k not in ("c_maxpage_2m", "c_static", "c_static_pie")]
    seeds += [p for p in F.real_corpus("small") if os.path.getsize(p) < 2_000_000][:25]
    for k, notes in F.note_scenarios().items():
        p = os.path.join(base, "fx", k + ".so")
        if F.with_build_notes(variants["c_plain"], p, notes):
            seeds += [p, p]
    counts = {}

    def one(k):
        rng = random.Random(a.seed * 1000003 + k)
        tmp = os.path.join(base, "w%d" % k)
        os.makedirs(tmp)
        try:
            src = rng.choice(seeds)
            dst = os.path.join(tmp, "m.so")
            desc = mutate(rng, src, dst, tmp)
            if desc is None:
                return "skip", ""
            data = open(dst, "rb").read()
            gnu, err = F.gnu_strip_bytes(dst, tmp)
            rc, out = oracle.strip(data)
            if gnu is None:
                return ("both-reject" if rc != 0 else "GNU-REJECTS-ORACLE-OK"), "%s | %s | %s" % (os.path.basename(src), desc, err.strip()[:80])
            if rc != 0:
                return "unsupported(rc=%d)" % rc, ""
            if out == gnu:
                return "ok", ""
            if a.keep:
                os.makedirs(a.keep, exist_ok=True)
                shutil.copy(dst, os.path.join(a.keep, "case%d.so" % k))
            return "MISMATCH", "case %d | %s | %s" % (k, os.path.basename(src), desc)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)

    with ThreadPoolExecutor(a.jobs) as ex:
        for st, msg in ex.map(one, range(a.cases)):
            counts[st] = counts.get(st, 0) + 1
            if st in ("MISMATCH", "GNU-REJECTS-ORACLE-OK"):
                print(st, msg)
    shutil.rmtree(base, ignore_errors=True)
    print("SUMMARY", dict(sorted(counts.items())))
    return 1 if counts.get("MISMATCH") else 0


if __name__ == "__main__":
    sys.exit(main())
