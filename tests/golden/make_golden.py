#!/usr/bin/env python
"""Regenerates the golden vectors: NAME.in.bin (input ELF) and NAME.gnu.bin (output of this image's
GNU strip 2.42, `strip --strip-unneeded -o`), the pinned parity target of BASELINE.json.

The reference (customink/lambdipy) has no vectors for its strip step; these are
outputs of the real external tool it shells out to (/root/reference/lambdipy/project_build.py:260),
generated in the build container and committed so that the GPU box checks against fixed bytes.
Suffix .bin because the repo's .gitignore excludes *.so.

usage: python tests/golden/make_golden.py
"""
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import elf_fixtures as F  # noqa: E402

PICK = ["c_plain", "c_g", "c_gz", "c_gold", "c_norelro", "c_exec_nopie", "cxx_g", "c_debuglink", "c_extra_sections"]
REAL = ["pillow.libs/libXau-154567c4.so.6.0.0",  # patchelf'd (R10/R12), build-attribute notes (R9), dynsym hoist (R2)
        "numpy.libs/libquadmath-96973f99.so.0.0.0"]


def main():
    tmp = tempfile.mkdtemp()
    v = F.build_variants(tmp)
    items = {k: v[k] for k in PICK if k in v}
    for k, notes in F.note_scenarios().items():
        p = os.path.join(tmp, k + ".so")
        if k in ("notes_gaps", "notes_attrs") and F.with_build_notes(v["c_plain"], p, notes):
            items[k] = p
    sp = F.site_packages()
    for r in REAL:
        p = os.path.join(sp, r)
        if os.path.exists(p) and os.path.getsize(p) < 400_000:
            items["real_" + os.path.basename(r).split("-")[0].split(".")[0]] = p
    for f in os.listdir(HERE):
        if f.endswith(".bin"):
            os.unlink(os.path.join(HERE, f))
    total = 0
    for name, path in sorted(items.items()):
        out, err = F.gnu_strip_bytes(path, tmp)
        assert out is not None, (name, err)
        shutil.copy(path, os.path.join(HERE, name + ".in.bin"))
        os.chmod(os.path.join(HERE, name + ".in.bin"), 0o644)
        with open(os.path.join(HERE, name + ".gnu.bin"), "wb") as f:
            f.write(out)
        total += os.path.getsize(path) + len(out)
        print("%-24s in=%7d  gnu=%7d" % (name, os.path.getsize(path), len(out)))
    print("total bytes", total)
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
