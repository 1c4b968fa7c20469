"""Loader for the CPU restatement under oracle/ (test infrastructure; never imported by the product)."""
import ctypes
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ODIR, "_build", "libstrip_oracle.so")


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        lib.lbo_strip.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_void_p),
                                  ctypes.POINTER(ctypes.c_uint64), ctypes.c_uint]
        lib.lbo_strip.restype = ctypes.c_int
        lib.lbo_free.argtypes = [ctypes.c_void_p]

    def strip(self, data, no_merge=False):
        """-> (rc, bytes|None); rc 0 ok, >0 unsupported class, <0 malformed"""
        out = ctypes.c_void_p()
        n = ctypes.c_uint64()
        rc = self.lib.lbo_strip(data, len(data), ctypes.byref(out), ctypes.byref(n), 1 if no_merge else 0)
        if rc != 0:
            return rc, None
        res = ctypes.string_at(out.value, n.value)
        self.lib.lbo_free(out)
        return 0, res


def build():
    src = os.path.join(ODIR, "strip_oracle.c")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", ODIR, "-s"], check=True)
    return LIB


def load():
    return Oracle(ctypes.CDLL(build()))
