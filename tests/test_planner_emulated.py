"""CPU checks of the PRODUCT's device planner (lambdipy_b200/csrc/plan.cu) without a GPU: the kernel
source is compiled with g++ and one warp is emulated by 32 host threads (tests/emu/plan_emu.cpp);
the emitted tile list is executed with memcpy and must tile the output exactly.  The result is
compared with the oracle and with the real GNU strip.  This is a TEST HARNESS (the package never
loads it); the GPU parity tests remain the gate for the compiled CUDA library."""
import os
import random

import pytest

import elf_fixtures as F

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def emu():
    import emu_lib
    return emu_lib.load()


def _read(p):
    with open(p, "rb") as f:
        return f.read()


def _agree(emu, oracle, data, no_merge=False):
    rc, want = oracle.strip(data, no_merge)
    st, got = emu.strip(data, no_merge)
    if st == 0:
        assert rc == 0 and got == want
    else:
        assert st > -1000, "tile list does not cover the output exactly once"
        assert rc != 0 or st == 8  # 8 = LB2_ST_PLANNER_LIMIT (shared-memory limits of the device planner)
    return st


def test_golden_vectors(emu):
    for f in sorted(os.listdir(GOLDEN)):
        if f.endswith(".in.bin"):
            st, got = emu.strip(_read(os.path.join(GOLDEN, f)))
            assert st == 0 and got == _read(os.path.join(GOLDEN, f.replace(".in.bin", ".gnu.bin"))), f


@pytest.mark.parametrize("no_merge", [False, True])
def test_variants_notes_edges(emu, oracle, variants, note_files, doctored, no_merge):
    n_ok = 0
    for name, p in {**variants, **note_files, **doctored}.items():
        if name == "c_maxpage_2m":
            continue
        st = _agree(emu, oracle, _read(p), no_merge)
        n_ok += st == 0
        if name == "c_many_sections":
            assert st == 8
    assert n_ok >= 45


def test_random_notes_and_structure_fuzz(emu, oracle, variants, fixture_dir, tmp_path):
    import sys
    from test_oracle_vs_gnu_strip import _random_notes, _staircase_notes
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import fuzz_vs_gnu as Z
    rng = random.Random(5)
    for case in range(40):
        p = os.path.join(fixture_dir, "emu_rnd_notes_%d.so" % case)
        assert F.with_build_notes(variants["c_plain"], p, _random_notes(rng, rng.choice([3, 8, 20, 60, 150])))
        assert _agree(emu, oracle, _read(p)) == 0
    c0 = emu.path_counts()
    assert c0[1] > 10, c0          # random ranges nest: the restated merge sort ran
    for case in range(40):
        p = os.path.join(fixture_dir, "emu_stair_notes_%d.so" % case)
        assert F.with_build_notes(variants["c_plain"], p, _staircase_notes(rng, rng.choice([4, 12, 40, 90, 130])))
        assert _agree(emu, oracle, _read(p)) == 0
    for case in range(30):
        p = os.path.join(fixture_dir, "emu_pfx_notes_%d.so" % case)
        assert F.with_build_notes(variants["c_plain"], p, _random_notes(rng, rng.choice([5, 12, 30, 70, 120]), prefix_names=True))
        assert _agree(emu, oracle, _read(p)) == 0
    c1 = emu.path_counts()
    assert c1[2] > 10, c1                                        # prefix-related names: the full merge sort with name compares
    assert c1[0] - c0[0] > 10 and c1[1] - c0[1] > 2, (c0, c1)   # proper orders took the rank sort, the nested ones fell back
    assert c0[3] > 10 and c1[3] - c0[3] > 10, (c0, c1)           # sections with more than 32 notes went to the whole CTA
    seeds = [variants[k] for k in sorted(variants) if k not in ("c_maxpage_2m", "c_static", "c_static_pie")]
    accepted = 0
    for k in range(300):
        d = tmp_path / ("m%d" % k)
        d.mkdir()
        dst = str(d / "m.so")
        if Z.mutate(rng, rng.choice(seeds), dst, str(d)) is None:
            continue
        accepted += _agree(emu, oracle, _read(dst)) == 0
    assert accepted > 80


def test_real_wheels_and_gnu_strip(emu, oracle, tmp_path):
    paths = F.real_corpus("small")
    for p in paths:
        data = _read(p)
        assert _agree(emu, oracle, data) == 0, p
    # and straight against the real binary for a few
    for p in paths[:8]:
        gnu, err = F.gnu_strip_bytes(p, str(tmp_path))
        st, got = emu.strip(_read(p))
        assert st == 0 and got == gnu, p


def test_synthetic_corpus(emu, oracle):
    from lambdipy_b200.corpus import Corpus
    c = Corpus(40, seed=17, max_size=1 << 20)
    for i in range(len(c)):
        assert _agree(emu, oracle, c.materialize(i)) == 0


def test_corrupt_note_sections(emu, oracle, variants, fixture_dir):
    import subprocess
    for sz in (4, 8, 11, 12, 20):
        p = os.path.join(fixture_dir, "emu_tiny_notes_%d.so" % sz)
        sec = p + ".sec"
        with open(sec, "wb") as f:
            f.write(b"\x08\0\0\0" * (sz // 4) + b"\0" * (sz % 4))
        subprocess.run(["objcopy", "--add-section", ".gnu.build.attributes=" + sec, "--set-section-flags",
                        ".gnu.build.attributes=readonly", variants["c_plain"], p], check=True)
        F.patch_section(p, ".gnu.build.attributes", sh_type=7)
        st, out = emu.strip(_read(p))
        rc, _ = oracle.strip(_read(p))
        assert st == 7 and rc == 7, (sz, st, rc)


def test_misplaced_tbss_is_declined_by_planner_and_oracle(emu, oracle, tmp_path):
    """BFD derives the sh_offset it writes for .tbss from the section's address; only the linker's placement (aligned
    end of .tdata) is reproduced, anything else goes to the host strip (gate found by header fuzzing)."""
    import struct
    import subprocess
    src = tmp_path / "t.c"
    src.write_text("__thread int a = 5; __thread int b; __thread char big[100];\nint f(void){ return a + b + big[3]; }\n")
    so = tmp_path / "t.so"
    subprocess.run(["gcc", "-shared", "-fPIC", "-g", "-o", str(so), str(src)], check=True)
    data = bytearray(so.read_bytes())
    shoff, = struct.unpack_from("<Q", data, 0x28)
    shnum, shstr = struct.unpack_from("<HH", data, 0x3c)
    so_, ss_ = struct.unpack_from("<QQ", data, shoff + shstr * 64 + 24)
    names = bytes(data[so_:so_ + ss_])
    tb = [i for i in range(shnum) if names[struct.unpack_from("<I", data, shoff + i * 64)[0]:].split(b"\0")[0] == b".tbss"][0]
    addr, = struct.unpack_from("<Q", data, shoff + tb * 64 + 16)
    assert _agree(emu, oracle, bytes(data)) == 0                      # the natural file takes the device path
    gnu, _ = F.gnu_strip_bytes(str(so), str(tmp_path))
    assert emu.strip(bytes(data))[1] == gnu
    for delta in (4, 16, -4, 256):
        d = bytearray(data)
        struct.pack_into("<Q", d, shoff + tb * 64 + 16, addr + delta)
        assert oracle.strip(bytes(d))[0] == 6 and emu.strip(bytes(d))[0] == 6, delta


def test_huge_extents_with_and_without_room_in_the_expand_list(emu, oracle, monkeypatch):
    """Extents of more than 512 tiles are normally only recorded (the scan launch's extra CTAs write their tiles, in
    parts of 2048); when the list is full the planning CTA writes them itself.  Both routes must tile the output exactly."""
    from lambdipy_b200.corpus import Corpus
    c = Corpus(2, seed=31, min_size=40 << 20, max_size=48 << 20)      # .text / .debug_info extents of 8+ MB
    for i in range(len(c)):
        data = c.materialize(i)
        rc, want = oracle.strip(data)
        assert rc == 0
        st, got = emu.strip(data)
        assert st == 0 and got == want
        monkeypatch.setenv("LB2EMU_BIG_CAP", "0")
        st, got = emu.strip(data)
        assert st == 0 and got == want
        monkeypatch.setenv("LB2EMU_BIG_CAP", "3")                      # room for some files' records only
        st, got = emu.strip(data)
        assert st == 0 and got == want
        monkeypatch.delenv("LB2EMU_BIG_CAP")
