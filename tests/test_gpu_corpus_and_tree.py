"""GPU tests at benchmark sizes and through the tree API.

* synthetic corpus (BASELINE config 4): the device generator equals the host generator; a
  stratified sample over the size deciles is bit-exact vs the oracle and GNU strip; at full size the
  size-independent properties hold: outputs are well-formed ELF whose section table ends the file,
  and strip is idempotent (a second, --no-merge-notes pass over the first pass's outputs reproduces them byte for byte).
* lb2_strip_tree on a copy of the real wheels tree == the reference's shell line on another copy.
"""
import ctypes as C
import os
import random
import shutil
import struct
import subprocess

import numpy as np
import pytest

import elf_fixtures as F

pytestmark = pytest.mark.gpu


def test_random_note_sections_gpu(gpu_ctx, oracle, variants, fixture_dir):
    from lambdipy_b200 import strip as S
    from test_oracle_vs_gnu_strip import _random_notes, _staircase_notes
    rng = random.Random(77)
    blobs = []
    for case in range(60):
        p = os.path.join(fixture_dir, "gpu_rnd_notes_%d.so" % case)
        assert F.with_build_notes(variants["c_plain"], p, _random_notes(rng, rng.choice([3, 8, 20, 60, 150])))
        blobs.append(open(p, "rb").read())
    for case in range(60):   # orders that ARE orders: the planner's rank-sort path (and its fallback when one range nests)
        p = os.path.join(fixture_dir, "gpu_stair_notes_%d.so" % case)
        assert F.with_build_notes(variants["c_plain"], p, _staircase_notes(rng, rng.choice([4, 12, 40, 90, 130])))
        blobs.append(open(p, "rb").read())
    for case in range(40):   # attribute names that are prefixes of each other: the slow, fully sequential sort
        p = os.path.join(fixture_dir, "gpu_pfx_notes_%d.so" % case)
        assert F.with_build_notes(variants["c_plain"], p, _random_notes(rng, rng.choice([5, 12, 30, 70, 120]), prefix_names=True))
        blobs.append(open(p, "rb").read())
    outs, status, _ = S.strip_buffers(gpu_ctx, blobs)
    for i, (b, o, st) in enumerate(zip(blobs, outs, status)):
        rc, want = oracle.strip(b)
        assert rc == 0 and st == 0 and o == want, "case %d" % i


def test_synthetic_corpus_device_resident(gpu_ctx, oracle, tmp_path):
    from lambdipy_b200.corpus import Corpus
    from lambdipy_b200.device import DeviceBatch
    corpus = Corpus(400, seed=0xB200)            # ~4.5 GB, sizes 1 KB .. 128 MB
    batch = DeviceBatch.from_corpus(gpu_ctx, corpus)
    try:
        batch.strip_async()
        st = batch.results()
        assert st["n_ok"] == len(corpus) and st["n_unsupported"] == 0 and st["overflow"] == 0
        assert st["in_bytes"] == corpus.total_bytes
        assert st["out_bytes"] == int(batch.out_sizes.sum()) < st["in_bytes"]
        # stratified sample over the size deciles vs oracle and the real binary
        order = np.argsort(batch.sizes)
        sample = sorted(set(int(order[min(len(order) - 1, (d * len(order)) // 10 + k)]) for d in range(10) for k in range(3)))
        for i in sample:
            data = batch.read_input(i)
            assert data == corpus.materialize(i), "device generator != host generator (file %d)" % i
            got = batch.read_output(i)
            rc, want = oracle.strip(data)
            assert rc == 0 and got == want, i
            if len(data) < (32 << 20):
                p = tmp_path / "s.so"
                p.write_bytes(data)
                gnu, err = F.gnu_strip_bytes(str(p), str(tmp_path))
                assert gnu == got, (i, err)
        # full-size structural property: section table ends the file, .shstrtab is last
        out_host = np.empty(int(batch.out_off[-1]), dtype=np.uint8)
        gpu_ctx.d2h(out_host.ctypes.data, batch.d_out, out_host.nbytes)
        for i in range(len(corpus)):
            o, n = int(batch.out_off[i]), int(batch.out_sizes[i])
            hdr = out_host[o:o + 64].tobytes()
            assert hdr[:4] == b"\x7fELF"
            shoff, = struct.unpack_from("<Q", hdr, 0x28)
            shnum, shstrndx = struct.unpack_from("<HH", hdr, 0x3c)
            assert shoff + shnum * 64 == n and shstrndx == shnum - 1
            assert not out_host[o + n:int(batch.out_off[i + 1])].any()  # padding between files untouched/zero
        # idempotence at full size: second pass over the outputs
        second = DeviceBatch(gpu_ctx, batch.out_off, batch.out_sizes[:len(corpus)])
        try:
            gpu_ctx.h2d(second.d_in, out_host.ctypes.data, out_host.nbytes)
            # (GNU strip re-merges already merged build notes differently on a few files -- it is not
            #  idempotent there either -- so the second pass runs like `strip --no-merge-notes`)
            second.strip_async(flags=1)
            st2 = second.results()
            assert st2["n_ok"] == len(corpus)
            assert (second.out_sizes[:len(corpus)] == batch.out_sizes[:len(corpus)]).all()
            out2 = np.empty(int(second.out_off[-1]), dtype=np.uint8)
            gpu_ctx.d2h(out2.ctypes.data, second.d_out, out2.nbytes)
            assert out2.nbytes == out_host.nbytes and (out2 == out_host).all()
        finally:
            second.close()
    finally:
        batch.close()


def test_chunked_device_api_equals_one_batch(gpu_ctx, oracle):
    """lb2_strip_device_chunked (input resident, output streamed through two slots) hands the consumer the same bytes the
    one-batch call produces, chunk after chunk, and reports the same per-file sizes / totals."""
    import ctypes as C
    from lambdipy_b200.corpus import Corpus
    from lambdipy_b200.device import DeviceBatch
    corpus = Corpus(120, seed=0xC0FFEE, max_size=8 << 20)
    whole = DeviceBatch.from_corpus(gpu_ctx, corpus)
    ring = DeviceBatch.from_corpus(gpu_ctx, corpus, chunk_bytes=24 << 20)      # ~8 chunks of whole files
    try:
        whole.strip_async()
        st1 = whole.results()
        want = [whole.read_output(i) for i in range(len(corpus))]
        got, chunks = {}, []

        def on_chunk(chunk, f0, n, d_slot, ooff, osz, status):
            chunks.append((chunk, f0, n))
            assert (status == 0).all()
            for k in range(n):
                buf = C.create_string_buffer(int(osz[k]))
                gpu_ctx.d2h(buf, d_slot + int(ooff[k]), len(buf))
                got[f0 + k] = buf.raw

        st2 = ring.strip_chunked(on_chunk=on_chunk)
        assert len(chunks) >= 4 and [c[0] for c in chunks] == list(range(len(chunks)))
        assert sum(c[2] for c in chunks) == len(corpus) and chunks[0][1] == 0
        assert [got[i] for i in range(len(corpus))] == want
        for k in ("n_ok", "n_unsupported", "in_bytes", "out_bytes", "copy_bytes", "header_bytes", "n_tiles"):
            assert st1[k] == st2[k], k
        assert (ring.out_sizes[:len(corpus)] == whole.out_sizes[:len(corpus)]).all() and not ring.status[:len(corpus)].any()
        for i in (0, 17, 63, 119):
            rc, ob = oracle.strip(whole.read_input(i))
            assert rc == 0 and ob == got[i]
        st3 = ring.strip_chunked()        # no consumer: the benchmark's use
        assert st3["out_bytes"] == st1["out_bytes"] and st3["n_ok"] == len(corpus)
    finally:
        whole.close(); ring.close()


def test_two_batches_in_flight(gpu_ctx, variants):
    """lb2_strip_device_async accepts a second batch before the first is collected (results come back in order); a third
    is refused, and so are the calls that need the workspaces for themselves."""
    from lambdipy_b200 import _native as N
    from lambdipy_b200.device import DeviceBatch
    blobs = [open(variants[k], "rb").read() for k in sorted(variants) if k != "c_many_sections"]
    b = DeviceBatch.from_blobs(gpu_ctx, blobs)
    try:
        b.strip_async()
        ref = b.results()
        want = [b.read_output(i) for i in range(len(blobs))]
        b.strip_async()
        b.strip_async(flags=N.F_NO_MERGE_NOTES)
        with pytest.raises(N.NativeError) as e:
            b.strip_async()
        assert e.value.code == N.LB2_E_STATE
        first = b.results()
        second = b.results()
        assert first["n_ok"] == second["n_ok"] == ref["n_ok"] == len(blobs) and first["out_bytes"] == ref["out_bytes"]
        with pytest.raises(N.NativeError):
            b.results()
        b.strip_async()
        assert b.results()["out_bytes"] == ref["out_bytes"] and [b.read_output(i) for i in range(len(blobs))] == want
    finally:
        b.close()


def test_output_capacity_error_is_reported(gpu_ctx, variants):
    from lambdipy_b200 import _native as N
    from lambdipy_b200.device import DeviceBatch
    blobs = [open(variants[k], "rb").read() for k in sorted(variants) if k != "c_many_sections"]
    b = DeviceBatch.from_blobs(gpu_ctx, blobs)
    try:
        b.out_cap = 4096  # claim a tiny output arena
        b.strip_async()
        with pytest.raises(N.NativeError) as e:
            b.results()
        assert e.value.code == N.LB2_E_CAPACITY
        b.out_cap = b.in_bytes + (16 << 20)
        b.strip_async()
        assert b.results()["n_ok"] == len(blobs)
    finally:
        b.close()


def _copy_tree(dst):
    sp = F.site_packages()
    for r in ("numpy", "PIL", "numpy.libs", "pillow.libs", "sklearn", "scikit_learn.libs"):
        shutil.copytree(os.path.join(sp, r), os.path.join(dst, r), symlinks=True,
                        ignore=shutil.ignore_patterns("*.py", "*.pyc", "*.pyi", "__pycache__", "*.txt", "*.npy", "*.npz"))
    shutil.copy(os.path.join(sp, "numpy.libs", sorted(os.listdir(os.path.join(sp, "numpy.libs")))[0]), os.path.join(dst, "versioned.so.3"))
    os.symlink("versioned.so.3", os.path.join(dst, "libdev.so"))        # link -> file whose own name does not match: target gets stripped
    core = os.path.join(dst, "numpy", "_core")
    so = sorted(f for f in os.listdir(core) if f.endswith(".so"))[0]
    os.symlink(so, os.path.join(core, "alias.so"))                       # a file reached twice: the reference strips it twice


def _snapshot(root):
    out = {}
    for d, dirs, fs in os.walk(root):
        for f in fs:
            p = os.path.join(d, f)
            rel = os.path.relpath(p, root)
            if os.path.islink(p):
                out[rel] = ("link", os.readlink(p))
            else:
                with open(p, "rb") as fh:
                    out[rel] = (os.stat(p).st_mode & 0o7777, fh.read())
    return out


def test_strip_tree_equals_reference_pipeline(gpu_ctx, tmp_path):
    from lambdipy_b200 import strip as S
    a, b = str(tmp_path / "ref"), str(tmp_path / "gpu")
    os.makedirs(a); os.makedirs(b)
    _copy_tree(a); _copy_tree(b)
    os.chmod(os.path.join(a, "versioned.so.3"), 0o640); os.chmod(os.path.join(b, "versioned.so.3"), 0o640)
    rc = subprocess.run(["bash", "-c", 'find %s/ -name "*.so" | xargs strip' % a], capture_output=True)  # the reference's line
    assert rc.returncode == 0, rc.stderr
    st = S.strip_tree(b, ctx=gpu_ctx)
    assert st["n_failed"] == 0 and st["n_fallback"] == 0 and st["n_gpu"] > 80
    assert st["n_skipped"] == 2                      # the two links themselves stay links
    sa, sb = _snapshot(a), _snapshot(b)
    assert set(sa) == set(sb)
    diff = [k for k in sa if sa[k] != sb[k]]
    assert not diff, diff[:5]
    assert not [f for f in os.listdir(os.path.join(b, "numpy.libs")) if ".lb2" in f]  # no temp files left behind


def test_strip_tree_failure_and_tolerance(gpu_ctx, variants, tmp_path):
    from lambdipy_b200 import strip as S
    root = str(tmp_path / "t")
    os.makedirs(root)
    shutil.copy(variants["c_g"], os.path.join(root, "good.so"))
    with open(os.path.join(root, "bogus.so"), "w") as f:
        f.write("not an elf")
    st = S.strip_tree(root, ctx=gpu_ctx)
    assert st["n_gpu"] == 1 and st["n_failed"] == 1          # the reference's script would exit 123 here too
    st = S.strip_tree(root, ctx=gpu_ctx, tolerate_non_elf=True)
    assert st["n_failed"] == 0 and st["n_skipped"] == 1
    # a directory (or dangling link) named *.so makes `strip` fail -> the reference's script exits 123
    os.unlink(os.path.join(root, "bogus.so"))
    os.makedirs(os.path.join(root, "dir.so"))
    os.symlink("missing-target", os.path.join(root, "dangling.so"))
    rc = subprocess.run(["bash", "-c", 'find %s/ -name "*.so" | xargs strip' % root], capture_output=True)
    assert rc.returncode == 123
    st = S.strip_tree(root, ctx=gpu_ctx)
    assert st["n_failed"] == 2
    empty = str(tmp_path / "empty")
    os.makedirs(empty)
    st = S.strip_tree(empty, ctx=gpu_ctx)
    assert st["n_selected"] == 0 and st["n_failed"] == 0     # documented deviation: xargs would run `strip` with no args (rc 123)


def _tree_meta(root):
    out = {}
    for d, dirs, fs in os.walk(root):
        for f in fs:
            p = os.path.join(d, f)
            if not os.path.islink(p):
                st = os.stat(p)
                out[os.path.relpath(p, root)] = (st.st_nlink, st.st_mode & 0o7777, st.st_uid)
    return out


def test_strip_tree_hard_links_keep_the_inode_like_strip_2_42(gpu_ctx, variants, tmp_path):
    """GNU strip 2.42 writes the stripped bytes back into the existing inode: a hard link whose name does not match
    is stripped too, two matching hard links mean the inode is stripped twice, link counts and modes survive."""
    from lambdipy_b200 import strip as S

    def mk(root):
        os.makedirs(os.path.join(root, "pkg"))
        shutil.copy(variants["c_g"], os.path.join(root, "pkg", "liba.so"))
        os.link(os.path.join(root, "pkg", "liba.so"), os.path.join(root, "pkg", "liba.so.1"))        # other name does not match
        shutil.copy(variants["cxx_g"], os.path.join(root, "pkg", "libb.so"))
        os.link(os.path.join(root, "pkg", "libb.so"), os.path.join(root, "libb_alias.so"))           # both names match
        shutil.copy(variants["c_gold"], os.path.join(root, "pkg", "ro.so"))
        os.chmod(os.path.join(root, "pkg", "ro.so"), 0o640)

    a, b = str(tmp_path / "ref"), str(tmp_path / "gpu")
    mk(a); mk(b)
    inodes_before = {k: os.stat(os.path.join(b, k)).st_ino for k in _tree_meta(b)}
    rc = subprocess.run(["bash", "-c", 'find %s/ -name "*.so" | xargs strip' % a], capture_output=True)
    assert rc.returncode == 0, rc.stderr
    st = S.strip_tree(b, ctx=gpu_ctx)
    assert st["n_failed"] == 0 and st["n_fallback"] == 0 and st["n_gpu"] == 3 and st["n_selected"] == 4
    assert _snapshot(a) == _snapshot(b)
    assert _tree_meta(a) == _tree_meta(b)
    assert {k: os.stat(os.path.join(b, k)).st_ino for k in _tree_meta(b)} == inodes_before   # same inodes as before the call
    with open(os.path.join(b, "pkg", "liba.so.1"), "rb") as f1, open(os.path.join(b, "pkg", "liba.so"), "rb") as f2:
        assert f1.read() == f2.read() and os.path.getsize(os.path.join(b, "pkg", "liba.so")) < os.path.getsize(variants["c_g"])


def test_strip_tree_with_cleanup_equals_reference_script(gpu_ctx, variants, tmp_path):
    """LB2_TREE_CLEANUP: the script's rm lines (:256-259) and the strip line (:260) on one walk == the reference's lines."""
    from lambdipy_b200 import strip as S

    def mk(root):
        for sub in ("pkg", "pkg/tests", "pkg/__pycache__", "numpy/tests", "x-1.0.dist-info"):
            os.makedirs(os.path.join(root, sub))
        shutil.copy(variants["c_g"], os.path.join(root, "pkg", "m.so"))
        shutil.copy(variants["cxx_g"], os.path.join(root, "pkg", "tests", "helper.so"))     # removed with its directory, never stripped
        shutil.copy(variants["c_gold"], os.path.join(root, "numpy", "tests", "kept.so"))    # kept by the pattern -> stripped
        with open(os.path.join(root, "pkg", "__pycache__", "m.pyc"), "w") as f:
            f.write("x")

    a, b = str(tmp_path / "ref"), str(tmp_path / "gpu")
    mk(a); mk(b)
    script = "\n".join(["set -ex", "rm -rf %s/*.egg-info" % a, "rm -rf %s/*.dist-info" % a, "find %s/ -name __pycache__ | xargs rm -rf" % a,
                        'find %s/ -name tests | grep -v "numpy" | xargs rm -rf' % a, 'find %s/ -name "*.so" | xargs strip' % a])
    assert subprocess.run(["bash", "-c", script], capture_output=True).returncode == 0
    st = S.strip_tree(b, ctx=gpu_ctx, cleanup=True, keep_tests_regex="numpy")
    assert st["n_removed"] == 3 and st["n_gpu"] == 2 and st["n_failed"] == 0
    assert _snapshot(a) == _snapshot(b)
    assert sorted(os.listdir(b)) == ["numpy", "pkg"] and os.listdir(os.path.join(b, "pkg")) == ["m.so"]


def test_strip_tree_streams_in_batches(gpu_ctx, tmp_path, monkeypatch):
    """Tiny batches and slots force the multi-batch path (upload of batch b+1 overlapping the download of batch b,
    files split over several slices) on the real wheels: the result is still the reference's tree."""
    from lambdipy_b200 import _native as N
    from lambdipy_b200 import strip as S
    monkeypatch.setenv("LB2_TREE_BATCH_MB", "8")
    monkeypatch.setenv("LB2_TREE_SLOT_MB", "1")
    monkeypatch.setenv("LB2_IO_THREADS", "3")
    a, b = str(tmp_path / "ref"), str(tmp_path / "gpu")
    os.makedirs(a); os.makedirs(b)
    _copy_tree(a); _copy_tree(b)
    rc = subprocess.run(["bash", "-c", 'find %s/ -name "*.so" | xargs strip' % a], capture_output=True)
    assert rc.returncode == 0, rc.stderr
    with N.Context(0) as ctx:       # a fresh context: the slot ring is sized at first use
        st = S.strip_tree(b, ctx=ctx)
    assert st["n_failed"] == 0 and st["n_fallback"] == 0
    sa, sb = _snapshot(a), _snapshot(b)
    assert sa == sb
