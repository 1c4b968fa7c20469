"""Drop-in check at the CLI level, in the build container only (the reference tree is not on the GPU box):
`lambdipy build --no-docker` of the REFERENCE's own click CLI, with lambdipy_b200.patch applied, runs our
mirror of install_non_resolved_requirements.  The reference's third-party imports that are missing in this
image (docker, requirementslib, PyGithub) are stubbed; everything else is the reference's code."""
import os
import sys
import types

import pytest

REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "lambdipy")), reason="reference tree not present (GPU box)")
def test_reference_cli_build_runs_our_strip_step(tmp_path, monkeypatch, variants):
    click_testing = pytest.importorskip("click.testing")
    for name in ("docker", "docker.errors", "requirementslib", "github", "github.GithubException", "github.GitRelease"):
        m = types.ModuleType(name)
        m.Requirement = object
        m.Github = m.InputGitAuthor = m.GitRelease = object
        m.UnknownObjectException = Exception
        m.BuildError = type("BuildError", (Exception,), {})
        m.from_env = lambda *a, **k: None
        m.__path__ = []
        monkeypatch.setitem(sys.modules, name, m)
    monkeypatch.syspath_prepend(REF)
    for k in [k for k in sys.modules if k == "lambdipy" or k.startswith("lambdipy.")]:
        monkeypatch.delitem(sys.modules, k)
    import lambdipy_b200.patch as patch
    cli = patch.apply()
    import lambdipy.project_build as ref_pb
    from lambdipy_b200 import project_build as mine
    assert ref_pb.install_non_resolved_requirements is mine.install_non_resolved_requirements
    assert cli.install_non_resolved_requirements is mine.install_non_resolved_requirements

    monkeypatch.chdir(tmp_path)
    (tmp_path / "requirements.txt").write_text("")            # nothing to resolve, nothing to pip-install
    monkeypatch.setenv("LAMBDIPY_STRIP_BACKEND", "gnu")       # CPU container: the reference's own line as backend
    monkeypatch.setenv("PYTHON_VERSION", "3.7")
    # an empty tree makes the reference's line fail (xargs runs `strip` without arguments, rc 123):
    r = click_testing.CliRunner().invoke(cli.cli, ["build", "--no-docker"])
    assert r.exit_code == 123, r.output
    # with a shared object in the include path copied first ... the include copy happens AFTER the strip step in
    # the reference (cli.py:69), so instead pre-seed ./build through a patched copy step:
    import shutil
    orig = cli.copy_prepared_releases_to_build_directory

    def seeded(paths, build_directory="./build"):
        orig(paths, build_directory)
        shutil.copy(variants["c_g"], os.path.join(build_directory, "mod.so"))
    monkeypatch.setattr(cli, "copy_prepared_releases_to_build_directory", seeded)
    before = os.path.getsize(variants["c_g"])
    r = click_testing.CliRunner().invoke(cli.cli, ["build", "--no-docker"])
    assert r.exit_code == 0, r.output
    assert "Finalizing the build" in r.output
    assert os.path.getsize(tmp_path / "build" / "mod.so") < before      # stripped by the (gnu) backend of our mirror
    assert not (tmp_path / "build" / "build").exists()
