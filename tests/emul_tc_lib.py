"""Builds (g++, host only) and binds tests/emul/libhint_emul_tc.so: the CPU interpreter of the tcgen05 (TF32) kernel's
static program.  Test infrastructure only."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CS = os.path.join(ROOT, "hint_b200", "csrc")
SRC = [os.path.join(HERE, "emul", "emul_tc.cpp"), os.path.join(CS, "plan.cpp"), os.path.join(CS, "plan_tc.cpp")]
DEPS = SRC + [os.path.join(CS, f) for f in ("plan.h", "plan_tc.h")]
LIB = os.path.join(HERE, "emul", "libhint_emul_tc.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not (os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in DEPS)):
            subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", LIB] + SRC, check=True)
        _lib = ctypes.CDLL(LIB)
        _lib.emul_tc_run.restype = ctypes.c_int
        _lib.emul_tc2_run.restype = ctypes.c_int
    return _lib


def _p(a, t=ctypes.c_float):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(t))


def run(d, dc, c_internal, clamp, max_splits, min_split_size, params, x, c=None, rev=False, tf32=False, v2=False):
    B = x.shape[0]
    ci = np.asarray(list(c_internal), dtype=np.int32)
    params = np.ascontiguousarray(params, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    c = None if c is None else np.ascontiguousarray(c, np.float32)
    z = np.full((B, d), np.nan, np.float32)
    J = np.full((B,), np.nan, np.float32)
    info = np.zeros(12, np.int64)
    fn = lib().emul_tc2_run if v2 else lib().emul_tc_run
    rc = fn(ctypes.c_int(d), ctypes.c_int(dc), _p(ci, ctypes.c_int), ctypes.c_int(len(ci)), ctypes.c_double(clamp),
                           ctypes.c_int(max_splits), ctypes.c_int(min_split_size), _p(params), _p(x), _p(c), ctypes.c_longlong(B),
                           ctypes.c_int(1 if rev else 0), ctypes.c_int(1 if tf32 else 0), _p(z), _p(J), _p(info, ctypes.c_longlong))
    return rc, z, J, info
