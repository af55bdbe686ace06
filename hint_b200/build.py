"""In-tree build of the sm_100a shared library (explicit nvcc, no JIT cache).

    python -m hint_b200.build [-v]

Produces hint_b200/libhint_b200.so next to this file.  The .so is git-ignored but travels to the GPU
box with the gpurun snapshot.  Translation units are compiled in parallel (objects under hint_b200/build/)
and only the ones whose dependencies changed are rebuilt.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libhint_b200.so")
HDR_API = os.path.join("..", "..", "include", "hint_b200.h")
# translation unit -> headers it depends on
UNITS = {
    "plan.cpp": ["plan.h", HDR_API],
    "plan_mma.cpp": ["plan.h", "plan_mma.h", HDR_API],
    "plan_chain.cpp": ["plan.h", "plan_chain.h", HDR_API],
    "plan_tc3.cpp": ["plan.h", "plan_tc3.h", HDR_API],
    "tc3_launch.cu": ["launch_count.h", "plan.h", "plan_tc3.h", "tc3_launch.h", "tc3_kernels.cuh", "tcgen05.cuh", HDR_API],
    "chain_launch.cu": ["launch_count.h", "plan.h", "plan_mma.h", "plan_chain.h", "chain_launch.h", "mma_kernels.cuh", "chain_kernels.cuh", HDR_API],
    "mma_launch.cu": ["launch_count.h", "plan.h", "plan_mma.h", "mma_launch.h", "mma_kernels.cuh", HDR_API],
    "train_ops.cu": ["train_ops.h", "launch_count.h", HDR_API],
    "householder.cu": ["householder.h", "mlp_coupling.h", "launch_count.h", HDR_API],
    "mlp_coupling.cu": ["mlp_coupling.h", "launch_count.h", HDR_API],
    "mmd.cu": ["mmd.h", "launch_count.h", HDR_API],
    "capi.cu": ["train_ops.h", "householder.h", "mlp_coupling.h", "mmd.h", "launch_count.h", "plan.h", "plan_mma.h", "mma_launch.h", "plan_chain.h", "chain_launch.h", "plan_tc3.h", "tc3_launch.h", "tcgen05.cuh", 
                "simt_phases.cuh", "simt_kernels.cuh", HDR_API],
}
NVCC_FLAGS = ["-O3", "-std=c++17", "-DHINT_MMA_MINB=2", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _obj(unit):
    return os.path.join(OBJ, os.path.splitext(unit)[0] + ".o")


def _unit_stale(unit):
    o = _obj(unit)
    if not os.path.exists(o):
        return True
    t = os.path.getmtime(o)
    deps = [os.path.join(CSRC, unit)] + [os.path.join(CSRC, h) for h in UNITS[unit]] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.abspath(__file__)]
    for unit, hdrs in UNITS.items():
        deps += [os.path.join(CSRC, unit)] + [os.path.join(CSRC, h) for h in hdrs]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not is_stale():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = nvcc_path()
    todo = [u for u in UNITS if force or _unit_stale(u)]

    def compile_unit(unit):
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, unit), "-o", _obj(unit)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return unit, cmd, res

    with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as ex:
        results = list(ex.map(compile_unit, todo))
    failed = False
    for unit, cmd, res in results:
        if verbose or res.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
        failed |= res.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libhint_b200.so")
    tmp = LIB + f".tmp{os.getpid()}"
    cmd = [nvcc, "-shared", "-o", tmp] + [_obj(u) for u in UNITS]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("linking libhint_b200.so failed")
    os.replace(tmp, LIB)   # atomic: a concurrent importer never maps a half-written library
    return LIB


if __name__ == "__main__":
    build(force="-f" in sys.argv or "--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
