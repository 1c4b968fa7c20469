"""Loader for the CPU warp emulator of the PRODUCT's device planner (tests/emu/plan_emu.cpp).
Test harness only: builds with g++ from lambdipy_b200/csrc/plan.cu, never used by the package."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "emu", "plan_emu.cpp")
LIB = os.path.join(HERE, "emu", "_build_plan_emu.so")
DEPS = [SRC, os.path.join(HERE, "emu", "shim", "cuda_runtime.h"), os.path.join(ROOT, "lambdipy_b200", "csrc", "plan.cu"),
        os.path.join(ROOT, "lambdipy_b200", "csrc", "lb2_common.cuh")]


def build():
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in DEPS):
        subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-pthread", "-shared", "-fPIC", "-I" + os.path.join(HERE, "emu", "shim"),
                        "-o", LIB, SRC], check=True)
    return LIB


class Emu:
    def __init__(self):
        self.lib = ctypes.CDLL(build())
        self.lib.lb2emu_strip.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_uint64)]
        self.lib.lb2emu_strip.restype = ctypes.c_int
        self.lib.lb2emu_free.argtypes = [ctypes.c_void_p]

    def path_counts(self):
        """[rank-sorted, merge-sorted by name rank, merge-sorted with full name compares, merged by the whole CTA] note sections so far"""
        out = (ctypes.c_int * 4)()
        self.lib.lb2emu_path_counts(out)
        return list(out)

    def strip(self, data, no_merge=False):
        """-> (status, bytes|None): status as the device planner reports it (0 ok, >0 class, <0 malformed,
        -1000/-1001: the emitted tiles do not tile the output exactly)"""
        out = ctypes.c_void_p()
        n = ctypes.c_uint64()
        rc = self.lib.lb2emu_strip(data, len(data), 1 if no_merge else 0, ctypes.byref(out), ctypes.byref(n))
        if rc != 0:
            return rc, None
        res = ctypes.string_at(out.value, n.value)
        self.lib.lb2emu_free(out)
        return 0, res


def load():
    return Emu()
