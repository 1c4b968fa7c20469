"""Pins the oracle (oracle/strip_oracle.c) to the real reference tool: this image's GNU strip
(Binutils 2.42), the external binary the reference shells out to
(/root/reference/lambdipy/project_build.py:260).  The reference's own tests hold no vectors for
this path, so the pin is (a) the committed golden outputs of the real binary and
(b) live differential runs against /usr/bin/strip.  CPU only."""
import os
import random
import struct

import pytest

import elf_fixtures as F

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _read(p):
    with open(p, "rb") as f:
        return f.read()


def _diff(oracle, path, tmp, no_merge=False):
    data = _read(path)
    gnu, err = F.gnu_strip_bytes(path, tmp, no_merge)
    rc, out = oracle.strip(data, no_merge)
    return data, gnu, rc, out, err


def test_golden_vectors(oracle):
    ins = sorted(f for f in os.listdir(GOLDEN) if f.endswith(".in.bin"))
    assert len(ins) >= 10
    for f in ins:
        rc, out = oracle.strip(_read(os.path.join(GOLDEN, f)))
        assert rc == 0, f
        assert out == _read(os.path.join(GOLDEN, f.replace(".in.bin", ".gnu.bin"))), f


def test_golden_vectors_match_live_binary(tmp_path):
    """the committed vectors are what this image's strip produces today"""
    for f in sorted(os.listdir(GOLDEN)):
        if f.endswith(".in.bin"):
            p = tmp_path / "x.so"
            p.write_bytes(_read(os.path.join(GOLDEN, f)))
            gnu, err = F.gnu_strip_bytes(str(p), str(tmp_path))
            assert gnu == _read(os.path.join(GOLDEN, f.replace(".in.bin", ".gnu.bin"))), (f, err)


@pytest.mark.parametrize("no_merge", [False, True])
def test_toolchain_variants(oracle, variants, tmp_path, no_merge):
    for name in sorted(variants):
        data, gnu, rc, out, err = _diff(oracle, variants[name], str(tmp_path), no_merge)
        assert gnu is not None, (name, err)
        assert rc == 0 and out == gnu, name
        assert len(out) <= len(data) + 4096


@pytest.mark.parametrize("no_merge", [False, True])
def test_build_attribute_notes(oracle, note_files, tmp_path, no_merge):
    for name in sorted(note_files):
        data, gnu, rc, out, err = _diff(oracle, note_files[name], str(tmp_path), no_merge)
        assert gnu is not None and rc == 0 and out == gnu, name


def _random_notes(rng, n, prefix_names=False):
    O, Fn = 0x100, 0x101
    names = [F.GA_VERSION, b"GA*\x02\x03\x00", b"GA*\x07\x02\x00", b"GA*FORTIFY\x00\x02\x00", b"GA$\x05gcc 13\x00",
             b"GA+stack_clash\x00", b"GA!\x08\x00", b"GA*GOW\x00\x2a\x05\x02\x00", b"GA*GOW\x00\x2a\x05\x00"]
    if prefix_names:  # names that compare EQUAL over the shorter length (memcmp from byte 3): the name order stops being one
        names += [b"GA*GOW\x00\x2a\x05\x00\x00", b"GA+stack_clash\x00\x01\x00", b"GA!\x08\x00\x00", b"GA*\x00"]
    notes = [(F.GA_VERSION, O, (0x1000, 0x1000 + rng.randrange(1, 0x400)))]
    for _ in range(n):
        nm = rng.choice(names)
        typ = O if rng.random() < 0.7 else Fn
        r = rng.random()
        if r < 0.25:
            rng_ = None
        elif r < 0.35:
            rng_ = (0, 0)
        else:
            s = 0x1000 + rng.randrange(0, 0x800) * rng.choice([1, 4, 16])
            e = s + rng.choice([0, 1, 5, 0x10, 0x40, 0x333, 0x1000])
            rng_ = (s, e)
        notes.append((nm, typ, rng_))
        if rng.random() < 0.15:
            notes.append(notes[-1])
    return notes


def _staircase_notes(rng, n):
    """annobin-like sections: per attribute name, ranges whose starts AND ends ascend (disjoint, touching, closer than 16
    bytes, overlapping staircase-wise), equal starts, duplicates, notes without a range -- within one name the sort
    comparator is then a proper order (the device planner's rank-sort path); now and then one nested range breaks it."""
    O, Fn = 0x100, 0x101
    names = [b"GA*\x02\x03\x00", b"GA*\x07\x02\x00", b"GA*FORTIFY\x00\x02\x00", b"GA+stack_clash\x00", b"GA!\x08\x00", b"GA*GOW\x00\x2a\x05\x02\x00"]
    notes = [(F.GA_VERSION, O, (0x1000, 0x1000 + rng.randrange(1, 0x400)))]
    per_name = max(1, n // len(names))
    for nm in rng.sample(names, rng.randrange(1, len(names) + 1)):
        s, e = 0x1000 + rng.randrange(0, 64), 0
        for _ in range(per_name):
            e = max(e + rng.choice([1, 3, 0x20]), s + rng.choice([1, 8, 0x40, 0x200]))
            typ = O if rng.random() < 0.8 else Fn
            notes.append((nm, typ, (s, e) if rng.random() < 0.9 else None))
            if rng.random() < 0.15:
                notes.append(notes[-1])
            if rng.random() < 0.1:
                notes.append((nm, typ, (s, e + rng.choice([1, 0x10]))))      # same start, later end
            s = rng.choice([e, e + 1, e + 8, e + 15, e + 16, e + 17, e + 0x100, max(s + 1, e - rng.choice([1, 4]))])
    if rng.random() < 0.25 and len(notes) > 3:
        nm, typ, r = notes[rng.randrange(1, len(notes))]
        if r:
            notes.insert(rng.randrange(1, len(notes)), (nm, typ, (r[0] + 1, max(r[0] + 1, r[1] - 1))))  # nested: the order stops being one
    rng.shuffle(notes[1:]) if rng.random() < 0.3 else None
    return notes


def test_staircase_note_sections(oracle, variants, fixture_dir, tmp_path):
    rng = random.Random(424242)
    for case in range(40):
        notes = _staircase_notes(rng, rng.choice([4, 12, 40, 90, 200]))
        p = os.path.join(fixture_dir, "stair_notes_%d.so" % case)
        assert F.with_build_notes(variants["c_plain"], p, notes)
        data, gnu, rc, out, err = _diff(oracle, p, str(tmp_path))
        assert gnu is not None, err
        assert rc == 0 and out == gnu, "staircase note case %d" % case


def test_random_note_sections(oracle, variants, fixture_dir, tmp_path):
    """nested, overlapping, adjoining, duplicate and empty ranges over many attribute names: the
    comparator objcopy sorts with is not antisymmetric, the merge-sort sequence matters"""
    rng = random.Random(20260921)
    for case in range(40):
        notes = _random_notes(rng, rng.choice([3, 8, 20, 60, 150]))
        p = os.path.join(fixture_dir, "rnd_notes_%d.so" % case)
        assert F.with_build_notes(variants["c_plain"], p, notes)
        data, gnu, rc, out, err = _diff(oracle, p, str(tmp_path))
        assert gnu is not None, err
        assert rc == 0 and out == gnu, "random note case %d" % case


def test_note_names_that_are_prefixes_of_each_other(oracle, variants, fixture_dir, tmp_path):
    """objcopy compares attribute names over the SHORTER length: a name that continues another compares equal to it but
    not to a third one -- the sort then depends on the comparison sequence even across names"""
    rng = random.Random(9091)
    for case in range(30):
        notes = _random_notes(rng, rng.choice([5, 12, 30, 70, 120]), prefix_names=True)
        p = os.path.join(fixture_dir, "pfx_notes_%d.so" % case)
        assert F.with_build_notes(variants["c_plain"], p, notes)
        data, gnu, rc, out, err = _diff(oracle, p, str(tmp_path))
        assert gnu is not None, err
        assert rc == 0 and out == gnu, "prefix-name note case %d" % case


def test_doctored_edges(oracle, doctored, tmp_path):
    reject = {"edge_empty": 1, "edge_text": 1, "edge_short_magic": 1, "edge_elf32_class": 2, "edge_big_endian": 2,
              "edge_no_sections": 4, "edge_shoff_past_eof": -1, "edge_truncated": -1}
    for name in sorted(doctored):
        data, gnu, rc, out, err = _diff(oracle, doctored[name], str(tmp_path))
        if name in reject:
            assert gnu is None, name            # GNU strip refuses ...
            assert rc == reject[name], (name, rc)  # ... and the oracle classifies the refusal
        elif name == "edge_et_rel_type":
            assert rc == 3                      # ET_REL goes through a different BFD path: out of scope
        else:
            assert gnu is not None and rc == 0 and out == gnu, name


def test_idempotent(oracle, variants):
    for name in ("c_plain", "c_g", "cxx_g", "c_gold", "c_exec_nopie"):
        rc, once = oracle.strip(_read(variants[name]))
        rc2, twice = oracle.strip(once)
        assert rc == 0 and rc2 == 0 and once == twice


def test_real_wheels_sample(oracle, tmp_path):
    """PIL + pillow.libs + numpy.libs of this image: patchelf'd (R10-R12), annobin notes (R9), gold/ld/lld"""
    paths = F.real_corpus("small")
    assert len(paths) >= 20
    for p in paths:
        data, gnu, rc, out, err = _diff(oracle, p, str(tmp_path))
        assert gnu is not None and rc == 0 and out == gnu, p


def test_synthetic_corpus_is_valid_and_matches(oracle, tmp_path):
    from lambdipy_b200.corpus import Corpus
    c = Corpus(60, seed=99, max_size=2 << 20)
    kinds = set()
    for i in range(len(c)):
        data = c.materialize(i)
        kinds.add(c.files[i].kind)
        p = tmp_path / "s.so"
        p.write_bytes(data)
        gnu, err = F.gnu_strip_bytes(str(p), str(tmp_path))
        assert gnu is not None and err.strip() == "", (i, err)
        rc, out = oracle.strip(data)
        assert rc == 0 and out == gnu, i
        assert len(out) < len(data)
    assert kinds == {"tiny", "compact", "sepcode"}


def test_structure_fuzz_against_gnu_strip(oracle, variants, tmp_path):
    """doctored header fields / objcopy-edited sections: the oracle either matches the real binary
    byte for byte or declares the mutant out of contract -- never a silent difference, and never an
    accepted file that GNU strip refuses (see oracle/fuzz_vs_gnu.py, which found the input gate)"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import fuzz_vs_gnu as Z
    seeds = [variants[k] for k in sorted(variants) if k not in ("c_maxpage_2m", "c_static", "c_static_pie")]
    seeds += [p for p in F.real_corpus("small") if os.path.getsize(p) < 1_000_000][:10]
    rng = random.Random(99)
    n_ok = n_unsup = 0
    for k in range(160):
        d = tmp_path / ("m%d" % k)
        d.mkdir()
        dst = str(d / "m.so")
        desc = Z.mutate(rng, rng.choice(seeds), dst, str(d))
        if desc is None:
            continue
        data = _read(dst)
        gnu, err = F.gnu_strip_bytes(dst, str(d))
        rc, out = oracle.strip(data)
        if gnu is None:
            assert rc != 0, ("oracle accepts a file GNU strip refuses", desc, err)
        elif rc == 0:
            assert out == gnu, desc
            n_ok += 1
        else:
            n_unsup += 1
    assert n_ok > 50 and n_unsup > 5


def test_corrupt_and_tiny_note_sections_are_refused(oracle, variants, fixture_dir, tmp_path):
    """objcopy reports 'corrupt GNU build attribute notes' and strip fails: the oracle (and the device
    planner, tests/test_planner_emulated.py) must classify these as bad notes, not pass them through"""
    import subprocess
    for sz in (4, 8, 11, 12, 16, 20):
        p = os.path.join(fixture_dir, "tiny_notes_%d.so" % sz)
        sec = p + ".sec"
        with open(sec, "wb") as f:
            f.write(b"\x08\0\0\0" * (sz // 4) + b"\0" * (sz % 4))
        subprocess.run(["objcopy", "--add-section", ".gnu.build.attributes=" + sec, "--set-section-flags",
                        ".gnu.build.attributes=readonly", variants["c_plain"], p], check=True)
        F.patch_section(p, ".gnu.build.attributes", sh_type=7)
        data, gnu, rc, out, err = _diff(oracle, p, str(tmp_path))
        assert gnu is None and "corrupt GNU build attribute" in err, (sz, err)
        assert rc == 7, (sz, rc)
