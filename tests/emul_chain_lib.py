"""Builds (g++, host only) and binds tests/emul/libhint_emul_chain.so - the fiber-based CPU emulation of the
register-chained warp-MMA kernels (chain_kernels.cuh).  Test infrastructure only; the product never loads it."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "hint_b200", "csrc")
SRC = [os.path.join(HERE, "emul", "emul_chain.cpp"), os.path.join(CSRC, "plan.cpp"), os.path.join(CSRC, "plan_mma.cpp"),
       os.path.join(CSRC, "plan_chain.cpp")]
DEPS = SRC + [os.path.join(HERE, "emul", "emul_mma.cpp")] + [os.path.join(CSRC, f) for f in (
    "plan.h", "plan_mma.h", "plan_chain.h", "mma_kernels.cuh", "chain_kernels.cuh")]
LIB = os.path.join(HERE, "emul", "libhint_emul_chain.so")


def build():
    if os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in DEPS):
        return LIB
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", LIB] + SRC, check=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.emul_chain_run.restype = ctypes.c_int
    return _lib


def _p(a, t=ctypes.c_float):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(t))


def run(d, dc, c_internal, clamp, max_splits, min_split_size, params, x, c=None, rev=False, backward=None, nctas=2, mt=2):
    """backward: None or (dz, dlogdet).  Returns dict(z, J, [xrec, dx, dc, dparams], info); raises LookupError when the
    block is outside the chain kernels' envelope."""
    B = x.shape[0]
    ci = np.asarray(list(c_internal), dtype=np.int32)
    params = np.ascontiguousarray(params, dtype=np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    c = None if c is None else np.ascontiguousarray(c, dtype=np.float32)
    z = np.full((B, d), np.nan, np.float32)
    J = np.full((B,), np.nan, np.float32)
    info = np.zeros(16, np.int64)
    out = dict(z=z, J=J, info=info)
    dz = dl = xrec = dx = dcond = dparams = None
    if backward is not None:
        dz = np.ascontiguousarray(backward[0], np.float32)
        dl = np.ascontiguousarray(backward[1], np.float32)
        xrec = np.full((B, d), np.nan, np.float32)
        dx = np.full((B, d), np.nan, np.float32)
        dcond = np.full((B, dc), np.nan, np.float32) if dc else None
        dparams = np.full(params.shape, np.nan, np.float32)
        out.update(xrec=xrec, dx=dx, dc=dcond, dparams=dparams)
    rc = lib().emul_chain_run(ctypes.c_int(d), ctypes.c_int(dc), _p(ci, ctypes.c_int), ctypes.c_int(len(ci)),
                              ctypes.c_double(clamp), ctypes.c_int(max_splits), ctypes.c_int(min_split_size),
                              _p(params), _p(x), _p(c), ctypes.c_longlong(B), ctypes.c_int(1 if rev else 0),
                              ctypes.c_int(nctas), ctypes.c_int(mt), _p(z), _p(J), _p(dz), _p(dl), _p(xrec),
                              _p(dx), _p(dcond), _p(dparams), _p(info, ctypes.c_longlong))
    if rc == 50:
        raise LookupError("block outside the chain kernels' envelope")
    if rc != 0:
        raise RuntimeError(f"emul_chain_run failed with code {rc}")
    return out
