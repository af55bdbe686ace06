"""Builds (g++, host only) and binds tests/emul/libhint_emul_tc3.so: the CPU interpreter of the tcgen05 TRAINING kernel's
static programs (plan_tc3.h).  Test infrastructure only."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CS = os.path.join(ROOT, "hint_b200", "csrc")
SRC = [os.path.join(HERE, "emul", "emul_tc3.cpp"), os.path.join(CS, "plan.cpp"), os.path.join(CS, "plan_tc3.cpp")]
DEPS = SRC + [os.path.join(CS, f) for f in ("plan.h", "plan_tc3.h")]
LIB = os.path.join(HERE, "emul", "libhint_emul_tc3.so")
_lib = None

INFO = ("ok", "groups", "mma_records", "epi_steps", "chunks", "packed_floats", "smem_bytes", "partial_floats",
        "mma_instr_per_tile", "tensor_cycles_per_tile", "hidden_images", "ring_slots")


def lib():
    global _lib
    if _lib is None:
        if not (os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in DEPS)):
            subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", LIB] + SRC, check=True)
        _lib = ctypes.CDLL(LIB)
        _lib.emul_tc3_backward.restype = ctypes.c_int
        _lib.emul_tc3_transport.restype = ctypes.c_int
    return _lib


def _p(a, t=ctypes.c_float):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(t))


def backward(d, dc, c_internal, clamp, max_splits, min_split_size, params, z, c, dz, dJ, lazy=False, tf32=False):
    """Returns dict(xrec, dx, dc, dparams, info).  Raises LookupError when the block is outside the kernel's envelope."""
    B = z.shape[0]
    ci = np.asarray(list(c_internal), dtype=np.int32)
    f = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
    params, z, c, dz, dJ = f(params), f(z), f(c), f(dz), f(dJ)
    xrec = np.full((B, d), np.nan, np.float32)
    dx = np.full((B, d), np.nan, np.float32)
    dcond = np.full((B, max(dc, 1)), np.nan, np.float32)
    dparams = np.full(params.shape, np.nan, np.float32)
    info = np.zeros(16, np.int64)
    rc = lib().emul_tc3_backward(ctypes.c_int(d), ctypes.c_int(dc), _p(ci, ctypes.c_int), ctypes.c_int(len(ci)), ctypes.c_double(clamp),
                                 ctypes.c_int(max_splits), ctypes.c_int(min_split_size), _p(params), _p(z), _p(c), _p(dz), _p(dJ),
                                 ctypes.c_longlong(B), ctypes.c_int(1 if lazy else 0), ctypes.c_int(1 if tf32 else 0), _p(xrec), _p(dx),
                                 _p(dcond), _p(dparams), _p(info, ctypes.c_longlong))
    inf = dict(zip(INFO, (int(v) for v in info)))
    if rc == 200:
        raise LookupError("outside the tc3 envelope")
    if rc != 0:
        raise RuntimeError(f"emul_tc3_backward failed with code {rc} ({'deadlock in the inferred waits' if rc == 400 else 'internal'})")
    return dict(xrec=xrec, dx=dx, dc=dcond[:, :dc] if dc else None, dparams=dparams, info=inf)


def transport(d, dc, c_internal, clamp, max_splits, min_split_size, params, x, c, rev=False, lazy=False, tf32=False):
    """Forward (rev=False) / inverse transport through the tcgen05 machine's T3K_FORWARD / T3K_INVERSE programs -> dict(z, J, info)."""
    B = x.shape[0]
    ci = np.asarray(list(c_internal), dtype=np.int32)
    f = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
    params, x, c = f(params), f(x), f(c)
    z = np.full((B, d), np.nan, np.float32)
    J = np.full((B,), np.nan, np.float32)
    info = np.zeros(16, np.int64)
    rc = lib().emul_tc3_transport(ctypes.c_int(d), ctypes.c_int(dc), _p(ci, ctypes.c_int), ctypes.c_int(len(ci)), ctypes.c_double(clamp),
                                  ctypes.c_int(max_splits), ctypes.c_int(min_split_size), _p(params), _p(x), _p(c), ctypes.c_longlong(B),
                                  ctypes.c_int(1 if rev else 0), ctypes.c_int(1 if lazy else 0), ctypes.c_int(1 if tf32 else 0), _p(z), _p(J),
                                  _p(info, ctypes.c_longlong))
    inf = dict(zip(INFO, (int(v) for v in info)))
    if rc == 200:
        raise LookupError("outside the tc3 envelope")
    if rc != 0:
        raise RuntimeError(f"emul_tc3_transport failed with code {rc}")
    return dict(z=z, J=J, info=inf)
