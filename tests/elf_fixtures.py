"""Fixture factory: small ELF shared objects / executables built at test time with the image's
gcc/g++/ld/gold in the variants SURVEY.md 8(c) lists, plus hand-doctored files for the edge
classes (empty, non-ELF, ELF32, no section headers, odd alignments ...).

The reference has no fixtures for this path (its tests never run `strip`), so these
stand in.  Binary outputs live in a temp dir (the repo's .gitignore excludes *.so / *.o); the
small committed golden vectors are under tests/golden/ (see tests/golden/make_golden.py).
"""
import os
import shutil
import struct
import subprocess

C_SRC = r"""
#include <stdio.h>
#include <string.h>
__thread int tls_counter = 7;
__thread char tls_buf[33];
int global_data[64] = {1, 2, 3};
static int hidden_bss[1000];
const char *const names[] = {"alpha", "beta", "gamma"};
__attribute__((constructor)) static void ctor(void) { hidden_bss[0] = 1; }
__attribute__((destructor)) static void dtor(void) { hidden_bss[1] = 2; }
int lb2_fix_sum(const int *v, int n) { int s = tls_counter; for (int i = 0; i < n; i++) s += v[i]; return s + hidden_bss[0]; }
const char *lb2_fix_name(int i) { snprintf(tls_buf, sizeof tls_buf, "%s", names[i % 3]); return tls_buf; }
int main(int argc, char **argv) { (void)argv; return lb2_fix_sum(global_data, argc); }
"""

CXX_SRC = r"""
#include <stdexcept>
#include <string>
#include <vector>
#include <map>
struct Shape { virtual ~Shape() {} virtual double area() const = 0; };
struct Sq : Shape { double s; explicit Sq(double x) : s(x) {} double area() const override { return s * s; } };
thread_local std::vector<int> tl_vec;
static std::map<std::string, int> registry;
extern "C" double lb2_fix_area(double x) {
  if (x < 0) throw std::invalid_argument("negative");
  Sq q(x); tl_vec.push_back((int)x); registry["sq"]++;
  try { if (x > 1e6) throw std::runtime_error("big"); } catch (const std::exception &e) { return -1; }
  return q.area();
}
int main() { return (int)lb2_fix_area(2.0); }
"""

# name -> (language, compiler flags)
VARIANTS = {
    "c_plain": ("c", ["-shared", "-fPIC", "-O1"]),
    "c_g": ("c", ["-shared", "-fPIC", "-O0", "-g"]),
    "c_g3": ("c", ["-shared", "-fPIC", "-O2", "-g3"]),
    "c_gz": ("c", ["-shared", "-fPIC", "-O1", "-g", "-gz"]),
    "c_gc_sections": ("c", ["-shared", "-fPIC", "-O2", "-g", "-ffunction-sections", "-fdata-sections", "-Wl,--gc-sections"]),
    "c_noseparate_code": ("c", ["-shared", "-fPIC", "-O1", "-g", "-Wl,-z,noseparate-code"]),
    "c_maxpage_2m": ("c", ["-shared", "-fPIC", "-O1", "-g", "-Wl,-z,max-page-size=0x200000"]),
    "c_hash_both": ("c", ["-shared", "-fPIC", "-O1", "-g", "-Wl,--hash-style=both"]),
    "c_hash_sysv": ("c", ["-shared", "-fPIC", "-O1", "-Wl,--hash-style=sysv"]),
    "c_no_build_id": ("c", ["-shared", "-fPIC", "-O1", "-g", "-Wl,--build-id=none"]),
    "c_norelro": ("c", ["-shared", "-fPIC", "-O1", "-g", "-Wl,-z,norelro"]),
    "c_now_relro": ("c", ["-shared", "-fPIC", "-O1", "-g", "-Wl,-z,relro,-z,now"]),
    "c_rpath_soname": ("c", ["-shared", "-fPIC", "-O1", "-g", "-Wl,-rpath,$ORIGIN/../lib", "-Wl,-soname,libfixture.so.1"]),
    "c_gold": ("c", ["-shared", "-fPIC", "-O1", "-g", "-fuse-ld=gold"]),
    "c_gold_plain": ("c", ["-shared", "-fPIC", "-O2", "-fuse-ld=gold"]),
    "c_lto": ("c", ["-shared", "-fPIC", "-O2", "-g", "-flto"]),
    "c_prestripped": ("c", ["-shared", "-fPIC", "-O1", "-Wl,-s"]),
    "c_static_pie_like": ("c", ["-pie", "-fPIE", "-O1", "-g"]),
    "c_exec_nopie": ("c", ["-no-pie", "-O1", "-g"]),
    "c_no_comment": ("c", ["-shared", "-fPIC", "-O1", "-fno-ident"]),
    "c_split_dwarf_types": ("c", ["-shared", "-fPIC", "-O1", "-g", "-fdebug-types-section"]),
    "c_dwarf4": ("c", ["-shared", "-fPIC", "-O1", "-gdwarf-4"]),
    "c_Os_unwind": ("c", ["-shared", "-fPIC", "-Os", "-g", "-fasynchronous-unwind-tables"]),
    "c_static": ("c", ["-static", "-O1", "-g"]),
    "c_static_pie": ("c", ["-static-pie", "-O1"]),
    "c_relr": ("c", ["-shared", "-fPIC", "-g", "-Wl,-z,pack-relative-relocs"]),
    "c_ibt_shstk": ("c", ["-shared", "-fPIC", "-g", "-fcf-protection=full", "-Wl,-z,ibt,-z,shstk"]),
    "c_zlib_gnu_debug": ("c", ["-shared", "-fPIC", "-g", "-Wl,--compress-debug-sections=zlib-gnu"]),
    "c_text_segment": ("c", ["-shared", "-fPIC", "-g", "-Wl,-Ttext-segment=0x400000"]),
    "c_sep_loadable": ("c", ["-shared", "-fPIC", "-g", "-Wl,-z,separate-loadable-segments"]),
    "c_bsymbolic_noplt": ("c", ["-shared", "-fPIC", "-g", "-O2", "-fno-plt", "-Wl,-Bsymbolic", "-Wl,-z,nocombreloc"]),
    "cxx_plain": ("cxx", ["-shared", "-fPIC", "-O1"]),
    "cxx_g": ("cxx", ["-shared", "-fPIC", "-O0", "-g"]),
    "cxx_gold_g": ("cxx", ["-shared", "-fPIC", "-O1", "-g", "-fuse-ld=gold"]),
    "cxx_exec_pie": ("cxx", ["-pie", "-fPIE", "-O1", "-g"]),
    "cxx_gz_gc": ("cxx", ["-shared", "-fPIC", "-O2", "-g", "-gz", "-ffunction-sections", "-Wl,--gc-sections"]),
}


def build_variants(outdir):
    """Compile every variant; returns {name: path}.  Variants the toolchain refuses are skipped."""
    os.makedirs(outdir, exist_ok=True)
    csrc, cxxsrc = os.path.join(outdir, "fix.c"), os.path.join(outdir, "fix.cc")
    with open(csrc, "w") as f:
        f.write(C_SRC)
    with open(cxxsrc, "w") as f:
        f.write(CXX_SRC)
    out = {}
    for name, (lang, flags) in VARIANTS.items():
        path = os.path.join(outdir, name + ".so")
        cmd = (["gcc", csrc] if lang == "c" else ["g++", cxxsrc]) + flags + ["-o", path]
        r = subprocess.run(cmd, capture_output=True)
        if r.returncode == 0 and os.path.exists(path):
            out[name] = path
    # post-processed variants
    base = out.get("c_g")
    if base:
        dl = os.path.join(outdir, "c_debuglink.so")
        dbg = os.path.join(outdir, "c_debuglink.debug")
        shutil.copy(base, dl)
        if subprocess.run(["objcopy", "--only-keep-debug", dl, dbg], capture_output=True).returncode == 0 and \
           subprocess.run(["objcopy", "--add-gnu-debuglink=" + dbg, dl], capture_output=True).returncode == 0:
            out["c_debuglink"] = dl
        # already stripped once (idempotence input)
        st = os.path.join(outdir, "c_twice.so")
        if subprocess.run(["strip", "-o", st, base], capture_output=True).returncode == 0:
            out["c_twice"] = st
        # extra non-alloc sections with odd alignment and content
        ex = os.path.join(outdir, "c_extra_sections.so")
        blob = os.path.join(outdir, "blob.bin")
        with open(blob, "wb") as f:
            f.write(bytes(range(251)) * 3)
        if subprocess.run(["objcopy", "--add-section", ".lb2.meta=" + blob, "--set-section-alignment", ".lb2.meta=32",
                           "--add-section", ".stabfoo=" + blob, "--add-section", ".comment2=" + blob, base, ex],
                          capture_output=True).returncode == 0:
            out["c_extra_sections"] = ex
        # more sections than the device planner holds in shared memory (64): planner-limit class
        many = os.path.join(outdir, "c_many_sections.so")
        args = []
        for i in range(70):
            args += ["--add-section", ".lb2.extra%02d=%s" % (i, blob)]
        if subprocess.run(["objcopy"] + args + [base, many], capture_output=True).returncode == 0:
            out["c_many_sections"] = many
    return out


# ---------------------------------------------------------------- build-attribute notes (R9)
GA_VERSION = b"GA$\x013a1\x00"


def _note(name, typ, rng):
    nm = name + b"\0" * ((-len(name)) % 4)
    desc = b"" if rng is None else struct.pack("<QQ", *rng)
    return struct.pack("<III", len(name), len(desc), typ) + nm + desc


def with_build_notes(src, dst, notes):
    """Copy `src` adding a .gnu.build.attributes section made of `notes` = [(name, type, (start,end)|None)]."""
    sec = dst + ".notes"
    with open(sec, "wb") as f:
        f.write(b"".join(_note(*n) for n in notes))
    r = subprocess.run(["objcopy", "--add-section", ".gnu.build.attributes=" + sec,
                        "--set-section-flags", ".gnu.build.attributes=readonly",
                        "--set-section-alignment", ".gnu.build.attributes=4", src, dst], capture_output=True)
    if r.returncode != 0:
        return False
    # objcopy creates it PROGBITS; GNU strip only merges SHT_NOTE sections: patch the type
    patch_section(dst, ".gnu.build.attributes", sh_type=7)
    return True


def note_scenarios():
    O, F = 0x100, 0x101
    V = GA_VERSION
    stack = b"GA*\x02\x03\x00"
    pic = b"GA*\x07\x02\x00"
    fort = b"GA*FORTIFY\x00\x02\x00"
    tool = b"GA$\x05running gcc 13.3.0\x00"
    return {
        "notes_simple": [(V, O, (0x1000, 0x1100)), (V, O, (0x1100, 0x1200)), (V, O, (0x2000, 0x2000)), (V, O, (0x1200, 0x1250))],
        "notes_gaps": [(V, O, (0x1000, 0x1010)), (V, O, (0x1018, 0x1020)), (V, O, (0x1030, 0x10e9)), (V, O, (0x5000, 0x5100)),
                       (V, O, (0, 0)), (V, O, (0x1012, 0x1017)), (V, O, (0x1020, 0x1025))],
        "notes_attrs": [(V, O, (0x1000, 0x1400)), (tool, O, None), (stack, O, None), (pic, O, None), (fort, O, None),
                        (fort, F, (0x1100, 0x1180)), (stack, F, (0x1100, 0x1180)),
                        (V, O, (0x1400, 0x1800)), (tool, O, None), (stack, O, None), (pic, O, None), (fort, O, None),
                        (fort, F, (0x1500, 0x1580)), (V, O, (0x1000, 0x1400)), (pic, O, None)],
        "notes_dups": [(V, O, (0x1000, 0x1400))] * 5 + [(stack, O, None)] * 5 + [(V, O, (0x1000, 0x1400)), (stack, O, (0x1200, 0x1300))],
        "notes_many": [(V, O, (0x1000 + 0x40 * i, 0x1000 + 0x40 * i + (0x20 if i % 3 else 0x40))) for i in range(60)] +
                      [(stack, O, (0x1000 + 0x80 * i, 0x1040 + 0x80 * i)) for i in range(40)] +
                      [(fort, F, (0x3000 - 0x40 * i, 0x3020 - 0x40 * i)) for i in range(30)],
    }


# ---------------------------------------------------------------- ELF patch helpers
def _shdrs(b):
    shoff, = struct.unpack_from("<Q", b, 0x28)
    shnum, shstr = struct.unpack_from("<HH", b, 0x3c)
    so, ss = struct.unpack_from("<QQ", b, shoff + shstr * 64 + 24)
    names = bytes(b[so:so + ss])
    res = []
    for i in range(shnum):
        n, = struct.unpack_from("<I", b, shoff + i * 64)
        res.append((names[n:names.index(b"\0", n)].decode(), shoff + i * 64))
    return res


def patch_section(path, secname, sh_type=None, sh_addralign=None, sh_entsize=None, sh_flags=None):
    with open(path, "rb") as f:
        b = bytearray(f.read())
    for name, o in _shdrs(b):
        if name == secname:
            if sh_type is not None:
                struct.pack_into("<I", b, o + 4, sh_type)
            if sh_flags is not None:
                struct.pack_into("<Q", b, o + 8, sh_flags)
            if sh_addralign is not None:
                struct.pack_into("<Q", b, o + 48, sh_addralign)
            if sh_entsize is not None:
                struct.pack_into("<Q", b, o + 56, sh_entsize)
    with open(path, "wb") as f:
        f.write(b)


def doctored(base_path, outdir):
    """Edge-class inputs derived from one good shared object.  {name: path}."""
    with open(base_path, "rb") as f:
        good = f.read()
    out = {}

    def w(name, data):
        p = os.path.join(outdir, name + ".so")
        with open(p, "wb") as f:
            f.write(data)
        out[name] = p

    w("edge_empty", b"")
    w("edge_text", b"this is not an ELF file\n" * 10)
    w("edge_short_magic", good[:40])
    b = bytearray(good); b[4] = 1; w("edge_elf32_class", bytes(b))
    b = bytearray(good); b[5] = 2; w("edge_big_endian", bytes(b))
    b = bytearray(good); struct.pack_into("<Q", b, 0x28, 0); struct.pack_into("<HH", b, 0x3c, 0, 0); w("edge_no_sections", bytes(b))
    b = bytearray(good); struct.pack_into("<H", b, 0x10, 1); w("edge_et_rel_type", bytes(b))
    b = bytearray(good); struct.pack_into("<Q", b, 0x28, len(good) + 4096); w("edge_shoff_past_eof", bytes(b))
    w("edge_truncated", good[: len(good) // 2])
    # alignment / entsize doctoring (BFD normalises these)
    for i, (secname, al) in enumerate([(".text", 4096), (".comment", 4096), (".dynstr", 3), (".data", 0), (".eh_frame", 24)]):
        p = os.path.join(outdir, "edge_align_%d.so" % i)
        shutil.copy(base_path, p)
        patch_section(p, secname, sh_addralign=al)
        out["edge_align_%d" % i] = p
    for i, (secname, es) in enumerate([(".init_array", 0), (".dynamic", 0), (".gnu.hash", 8), (".text", 0x30), (".comment", 0)]):
        p = os.path.join(outdir, "edge_entsize_%d.so" % i)
        shutil.copy(base_path, p)
        patch_section(p, secname, sh_entsize=es)
        out["edge_entsize_%d" % i] = p
    return out


def gnu_strip(path, out_path, no_merge=False):
    """The parity target: this image's GNU strip (Binutils 2.42).  Returns (rc, stderr)."""
    cmd = ["strip", "--strip-unneeded"] + (["--no-merge-notes"] if no_merge else []) + ["-o", out_path, path]
    r = subprocess.run(cmd, capture_output=True)
    return r.returncode, r.stderr.decode(errors="replace")


def gnu_strip_bytes(path, tmpdir, no_merge=False):
    out = os.path.join(tmpdir, "gnu_strip_out.bin")
    if os.path.exists(out):
        os.unlink(out)
    rc, err = gnu_strip(path, out, no_merge)
    if rc != 0:
        return None, err
    with open(out, "rb") as f:
        return f.read(), err


SITE = None


def site_packages():
    import site
    return site.getsitepackages()[0]


def real_corpus(kind="wheels"):
    """Paths of real shared objects shipped in this image (present on the GPU box too).
    kind: 'wheels' = numpy+scipy+sklearn+PIL(+*.libs) (BASELINE config 2 stand-in, 229 files / 160 MB);
          'torch'  = torch/**/*.so* (config 3 stand-in, 1.5 GB)."""
    sp = site_packages()
    roots = {"wheels": ["numpy", "scipy", "sklearn", "PIL", "numpy.libs", "scipy.libs", "pillow.libs", "scikit_learn.libs"],
             "torch": ["torch"], "small": ["PIL", "pillow.libs", "numpy.libs"]}[kind]
    files = []
    for r in roots:
        for d, _, fs in os.walk(os.path.join(sp, r)):
            for f in fs:
                p = os.path.join(d, f)
                if ".so" in f and not os.path.islink(p):
                    with open(p, "rb") as fh:
                        if fh.read(4) == b"\x7fELF":
                            files.append(p)
    return sorted(files)
