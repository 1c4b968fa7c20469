"""The clean-up lines of the reference's script (/root/reference/lambdipy/project_build.py:256-259) folded into
the library's directory walk (lb2_tree_cleanup / LB2_TREE_CLEANUP) against the reference's own bash lines run
on an identical tree.  Host logic only: no GPU needed."""
import os
import shutil
import subprocess

import pytest


def _mk(root):
    def d(p):
        os.makedirs(os.path.join(root, p), exist_ok=True)

    def f(p, text="x"):
        d(os.path.dirname(p))
        with open(os.path.join(root, p), "w") as fh:
            fh.write(text)

    d("numpy-1.0.dist-info"); f("numpy-1.0.dist-info/RECORD")
    d("foo.egg-info"); f("foo.egg-info/PKG-INFO")
    f("single.egg-info")                      # a FILE matching the glob goes too
    d(".hidden.dist-info")                    # shell globs skip dot files
    d("pkg/vendored-2.0.dist-info")           # not top level: stays
    f("pkg/__pycache__/a.pyc"); f("pkg/sub/__pycache__/b.pyc"); f("pkg/sub/deep/__pycache__")  # dir, dir, plain file
    f("pkg/tests/test_a.py"); f("pkg/sub/tests/data/x.bin"); f("pkg/sub/tests/tests/y")
    f("numpy/tests/test_n.py"); f("numpy/core/tests/t.py"); f("scipy/linalg/tests/t.py")
    f("other/tests")                          # a plain file called tests
    f("star*dir/tests/t.py")                  # the default pattern "*" keeps paths containing an asterisk
    f("pkg/mod.py"); f("pkg/sub/lib.so", "not elf")
    f("pkg/tests_extra/keep.py"); f("pkg/mytests/keep.py")
    os.symlink("pkg/tests", os.path.join(root, "tests"))  # a symlink named tests: rm -rf removes the link only


def _listing(root):
    out = {}
    for dp, dirs, fs in os.walk(root):
        for n in dirs + fs:
            p = os.path.join(dp, n)
            out[os.path.relpath(p, root)] = "link" if os.path.islink(p) else ("dir" if os.path.isdir(p) else "file")
    return out


def _reference(root, keep_tests):
    pat = "\\|".join(keep_tests) if keep_tests else "*"   # project_build.py:249
    script = "\n".join([
        "set -ex",
        "rm -rf %s/*.egg-info" % root,
        "rm -rf %s/*.dist-info" % root,
        "find %s/ -name __pycache__ | xargs rm -rf" % root,
        'find %s/ -name tests | grep -v "%s" | xargs rm -rf' % (root, pat),
    ])
    r = subprocess.run(["bash", "-c", script], capture_output=True)
    assert r.returncode == 0, r.stderr


@pytest.mark.parametrize("keep", [None, ["numpy", "scipy"], ["sub"]])
def test_cleanup_equals_reference_script(keep, tmp_path):
    from lambdipy_b200 import project_build as mine
    from lambdipy_b200 import strip as S
    a, b = str(tmp_path / "ref" / "build"), str(tmp_path / "b2" / "build")
    _mk(a); _mk(b)
    assert _listing(a) == _listing(b)
    _reference(a, keep)
    removed = S.cleanup_tree(b, mine._keep_pattern(keep))
    assert removed > 0
    la, lb = _listing(a), _listing(b)
    assert la == lb, sorted(set(la.items()) ^ set(lb.items()))
    assert "pkg/vendored-2.0.dist-info" in lb and ".hidden.dist-info" in lb and "numpy-1.0.dist-info" not in lb
    assert ("numpy/tests" in lb) == bool(keep and "numpy" in keep)


def test_cleanup_when_root_path_matches_keep_pattern(tmp_path):
    """grep sees the whole printed path, build directory included (an install_dir containing 'numpy' keeps every tests dir)."""
    from lambdipy_b200 import strip as S
    a, b = str(tmp_path / "numpy-ref" / "build"), str(tmp_path / "numpy-b2" / "build")
    _mk(a); _mk(b)
    _reference(a, ["numpy"])
    S.cleanup_tree(b, "numpy")
    assert _listing(a) == _listing(b)
    assert "pkg/tests" in _listing(b)


def test_script_drops_cleanup_lines_only_for_b200_backend():
    from lambdipy_b200 import project_build as mine
    full = mine._script_lines("/x", "", None)
    assert [l for l in full if l.startswith(("rm -rf", "find"))] and len(full) == 7
    assert mine._script_lines("/x", "", None, cleanup_in_script=False) == full[:3]
