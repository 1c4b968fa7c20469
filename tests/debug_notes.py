import sys, os, tempfile, subprocess
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import elf_fixtures as F, oracle_lib
from lambdipy_b200 import _native as N, strip as S
o = oracle_lib.load()
d = tempfile.mkdtemp()
v = F.build_variants(d)
ctx = N.Context(0)
os.makedirs("gpurun_out/dbg", exist_ok=True)
for rep in range(3):
  for k, n in F.note_scenarios().items():
    p = os.path.join(d, k + ".so")
    assert F.with_build_notes(v["c_g"], p, n)
    data = open(p, "rb").read()
    for nm in (False, True):
        g, err = F.gnu_strip_bytes(p, d, nm)
        rc, ob = o.strip(data, nm)
        outs, st, _ = S.strip_buffers(ctx, [data], flags=1 if nm else 0)
        print(rep, k, nm, "oracle==gnu", g == ob, "gpu==gnu", outs[0] == g, "status", st)
        if rep == 0 and (outs[0] != g or g != ob):
            for tag, b in (("gnu", g), ("oracle", ob), ("gpu", outs[0]), ("in", data)):
                open("gpurun_out/dbg/%s.%d.%s.bin" % (k, nm, tag), "wb").write(b or b"")
