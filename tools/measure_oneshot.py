#!/usr/bin/env python
"""Cold, one-shot cost of the strip step as `lambdipy build --no-docker` pays it: a FRESH process
imports the mirror, calls install_non_resolved_requirements(..., no_docker=True) on a /dev/shm copy of a real
build tree (BASELINE configs 2 and 3 stand-ins) and exits.  Timed inside that process from before the import to
after the return -- CUDA context creation, library load, pinned ring allocation included -- for
LAMBDIPY_STRIP_BACKEND=gnu (the reference's own line) and =b200, alternating, N repetitions each.
Also: bare `lb2_ctx_create` time in a fresh process.

    python tools/measure_oneshot.py > gpurun_out/oneshot.json
"""
import json
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TREES = {
    "config1_numpy": ["numpy", "numpy.libs"],
    "config2_numpy+scipy+sklearn+PIL": ["numpy", "scipy", "sklearn", "PIL", "numpy.libs", "scipy.libs", "pillow.libs", "scikit_learn.libs"],
    "config3_torch": ["torch"],
}
CHILD = r"""
import sys, time, os, io, contextlib
t0 = time.perf_counter()
sys.path.insert(0, %(root)r)
from lambdipy_b200 import project_build as m
out = io.StringIO()
with contextlib.redirect_stdout(out):
    m.install_non_resolved_requirements({}, [], '3.12', no_docker=True, build_directory=%(bd)r)
print(time.perf_counter() - t0)
"""
CTX = r"""
import sys, time
t0 = time.perf_counter()
sys.path.insert(0, %(root)r)
from lambdipy_b200 import _native as N
t1 = time.perf_counter()
c = N.Context(0)
t2 = time.perf_counter()
c.check(c.lib.lb2_tree_prepare(c.h, 0))
t3 = time.perf_counter()
print(t1 - t0, t2 - t1, t3 - t2)
"""


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    sp = sysconfig.get_paths()["purelib"]
    res = {"nproc": os.cpu_count(), "ctx_create": [], "trees": {}}
    for _ in range(3):
        r = subprocess.run([sys.executable, "-c", CTX % {"root": ROOT}], capture_output=True, text=True)
        if r.returncode == 0:
            a, b, c = [float(x) for x in r.stdout.split()]
            res["ctx_create"].append({"import_s": a, "lb2_ctx_create_s": b, "lb2_tree_prepare_s": c})
        else:
            res["ctx_create"].append({"error": r.stderr[-300:]})
    # the reference executes the script it writes into the build directory: the tree must live on a
    # filesystem mounted exec (/dev/shm is noexec on the GPU boxes)
    base = None
    for cand in ("/dev/shm", "/tmp", ROOT):
        d = tempfile.mkdtemp(prefix="lb2_oneshot_", dir=cand)
        probe = os.path.join(d, "x.sh")
        with open(probe, "w") as f:
            f.write("#!/bin/bash\nexit 0\n")
        os.chmod(probe, 0o755)
        try:
            ok = subprocess.run([probe]).returncode == 0
        except OSError:
            ok = False
        os.unlink(probe)
        if ok:
            base = d
            break
        shutil.rmtree(d, ignore_errors=True)
    res["build_dir_fs"] = os.path.dirname(base)
    try:
        for name, roots in TREES.items():
            roots = [r for r in roots if os.path.isdir(os.path.join(sp, r))]
            if not roots:
                continue
            master = os.path.join(base, "master")
            shutil.rmtree(master, ignore_errors=True)
            for r in roots:
                shutil.copytree(os.path.join(sp, r), os.path.join(master, r), symlinks=True,
                                ignore=lambda d, names: [n for n in names if not (os.path.isdir(os.path.join(d, n)) or ".so" in n)])
            so_bytes = sum(os.path.getsize(os.path.join(d, f)) for d, _, fs in os.walk(master) for f in fs
                           if f.endswith(".so") and not os.path.islink(os.path.join(d, f)))
            t = {"gnu": [], "b200": [], "so_bytes": so_bytes}
            for rep in range(reps):
                for backend in ("gnu", "b200"):
                    bd = os.path.join(base, "build")
                    shutil.rmtree(bd, ignore_errors=True)
                    shutil.copytree(master, bd, symlinks=True)
                    env = dict(os.environ, LAMBDIPY_STRIP_BACKEND=backend)
                    r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "bd": bd}], capture_output=True, text=True, env=env)
                    t[backend].append(float(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else None)
                    if r.returncode != 0:
                        t.setdefault("errors", []).append(r.stderr[-300:])
            res["trees"][name] = t
    finally:
        shutil.rmtree(base, ignore_errors=True)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
