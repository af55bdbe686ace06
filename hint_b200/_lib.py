"""ctypes binding of libhint_b200.so (the C ABI declared in include/hint_b200.h).

There is no CPU or PyTorch fallback: if the shared library cannot be found or built, importing the
product fails loudly, and compute entry points refuse non-CUDA tensors.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhint_b200.so")

HINT_OK, HINT_ERR_INVALID, HINT_ERR_UNSUPPORTED, HINT_ERR_CUDA, HINT_ERR_WORKSPACE = 0, 1, 2, 3, 4
MODE_FP32, MODE_TF32, MODE_TF32X3, MODE_TF32_TCGEN05, MODE_TF32_MMA, MODE_TF32_CHAIN, MODE_TF32_TC3 = 0, 1, 2, 3, 4, 5, 6
WS_FORWARD, WS_BACKWARD = 0, 1

# every symbol include/hint_b200.h declares (tests/test_capi.py checks the header against this list)
EXPORTS = [
    "hint_plan_create", "hint_plan_destroy", "hint_plan_num_nodes", "hint_plan_node", "hint_plan_param_count",
    "hint_plan_param_layout", "hint_plan_flops_per_sample", "hint_plan_tile_rows", "hint_plan_mode_supported",
    "hint_workspace_bytes",
    "hint_forward", "hint_backward", "hint_last_error", "hint_version", "hint_launch_count",
    "hint_backward_nll", "hint_add_noise", "hint_nll_workspace_bytes", "hint_nll_loss", "hint_adam_step",
    "hint_householder_matrix", "hint_householder_matrix_backward", "hint_householder_apply",
    "hint_householder_wgrad_workspace_bytes", "hint_householder_wgrad",
    "hint_mlp_coupling_supported", "hint_mlp_coupling_forward", "hint_mlp_coupling_workspace_bytes", "hint_mlp_coupling_backward",
    "hint_mmd_workspace_bytes", "hint_multi_mmd",
]


class NodeInfo(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("depth", "lo", "hi", "k", "cin", "h", "cout", "leaf", "parent", "upper", "lower")] + \
               [("param_offset", ctypes.c_int64)]


_lib = None


def load():
    """Load (building in-tree first if the .so is missing or stale and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    override = os.environ.get("HINT_B200_LIB")   # developer switch: load a differently-built library (kernel experiments)
    try:
        from . import build as _build
        if not override and _build.is_stale():
            # one builder at a time: under torchrun every rank imports at once, and concurrent nvcc runs would write the same
            # object files / shared library (a rank could dlopen a half-written .so)
            import fcntl
            with open(os.path.join(HERE, ".build.lock"), "w") as lock:
                fcntl.flock(lock, fcntl.LOCK_EX)
                try:
                    if _build.is_stale():
                        _build.build()
                finally:
                    fcntl.flock(lock, fcntl.LOCK_UN)
    except Exception as e:  # no nvcc on this machine: fall through and require the prebuilt .so
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"hint_b200: native library {LIB_PATH} is missing and could not be built ({e}). "
                "Run `python -m hint_b200.build` (needs nvcc, sm_100a). There is no fallback path.") from e
        import warnings
        warnings.warn(f"hint_b200: rebuilding the native library failed ({e}); loading the existing, possibly stale {LIB_PATH}")
    lib = ctypes.CDLL(override or LIB_PATH)
    vp, i32, i64, f32p = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p
    lib.hint_plan_create.restype = ctypes.c_int
    lib.hint_plan_create.argtypes = [i32, i32, ctypes.POINTER(i32), i32, ctypes.c_double, i32, i32, i32, ctypes.POINTER(vp)]
    lib.hint_plan_destroy.restype = None
    lib.hint_plan_destroy.argtypes = [vp]
    lib.hint_plan_num_nodes.restype = i32
    lib.hint_plan_num_nodes.argtypes = [vp]
    lib.hint_plan_node.restype = ctypes.c_int
    lib.hint_plan_node.argtypes = [vp, i32, ctypes.POINTER(NodeInfo)]
    lib.hint_plan_param_count.restype = i64
    lib.hint_plan_param_count.argtypes = [vp]
    lib.hint_plan_param_layout.restype = ctypes.c_int
    lib.hint_plan_param_layout.argtypes = [vp, ctypes.POINTER(i64), i64]
    lib.hint_plan_flops_per_sample.restype = i64
    lib.hint_plan_flops_per_sample.argtypes = [vp]
    lib.hint_plan_tile_rows.restype = i32
    lib.hint_plan_tile_rows.argtypes = [vp, i32]
    lib.hint_plan_mode_supported.restype = i32
    lib.hint_plan_mode_supported.argtypes = [vp, i32]
    lib.hint_workspace_bytes.restype = ctypes.c_size_t
    lib.hint_workspace_bytes.argtypes = [vp, i64, i32]
    lib.hint_forward.restype = ctypes.c_int
    lib.hint_forward.argtypes = [vp, f32p, f32p, f32p, i64, i32, i32, f32p, f32p, vp, ctypes.c_size_t, vp]
    lib.hint_backward.restype = ctypes.c_int
    lib.hint_backward.argtypes = [vp, f32p, f32p, f32p, f32p, f32p, i64, i32, f32p, f32p, f32p, f32p, vp,
                                  ctypes.c_size_t, vp]
    f32 = ctypes.c_float
    lib.hint_backward_nll.restype = ctypes.c_int
    lib.hint_backward_nll.argtypes = [vp, f32p, f32p, f32p, f32p, f32, i64, i32, f32p, f32p, f32p, f32p, vp, ctypes.c_size_t, vp]
    lib.hint_add_noise.restype = ctypes.c_int
    lib.hint_add_noise.argtypes = [f32p, f32p, i64, f32, ctypes.c_uint64, ctypes.c_uint64, vp]
    lib.hint_nll_workspace_bytes.restype = ctypes.c_size_t
    lib.hint_nll_workspace_bytes.argtypes = []
    lib.hint_nll_loss.restype = ctypes.c_int
    lib.hint_nll_loss.argtypes = [f32p, ctypes.POINTER(vp), i32, i64, i32, f32p, vp, ctypes.c_size_t, vp]
    lib.hint_adam_step.restype = ctypes.c_int
    lib.hint_adam_step.argtypes = [i32, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp),
                                   ctypes.POINTER(i64), f32, f32, f32, f32, f32, f32, i64, vp]
    lib.hint_householder_matrix.restype = ctypes.c_int
    lib.hint_householder_matrix.argtypes = [f32p, i32, i32, f32p, vp]
    lib.hint_householder_matrix_backward.restype = ctypes.c_int
    lib.hint_householder_matrix_backward.argtypes = [f32p, f32p, f32p, i32, i32, f32p, vp]
    lib.hint_householder_apply.restype = ctypes.c_int
    lib.hint_householder_apply.argtypes = [f32p, f32p, i64, i32, i32, f32p, vp]
    lib.hint_householder_wgrad_workspace_bytes.restype = ctypes.c_size_t
    lib.hint_householder_wgrad_workspace_bytes.argtypes = [i32]
    lib.hint_householder_wgrad.restype = ctypes.c_int
    lib.hint_householder_wgrad.argtypes = [f32p, f32p, i64, i32, f32p, vp, ctypes.c_size_t, vp]
    lib.hint_mlp_coupling_supported.restype = ctypes.c_int
    lib.hint_mlp_coupling_supported.argtypes = [i32, i32, i32]
    lib.hint_mlp_coupling_forward.restype = ctypes.c_int
    lib.hint_mlp_coupling_forward.argtypes = [f32p, i32, f32p, i32, i32, ctypes.POINTER(vp), f32, i32, i64, f32p, f32p, vp]
    lib.hint_mlp_coupling_workspace_bytes.restype = ctypes.c_size_t
    lib.hint_mlp_coupling_workspace_bytes.argtypes = [i32, i32, i32, i64]
    lib.hint_mlp_coupling_backward.restype = ctypes.c_int
    lib.hint_mlp_coupling_backward.argtypes = [f32p, i32, f32p, i32, i32, ctypes.POINTER(vp), f32, i64, f32p, f32p, f32p, f32p, ctypes.POINTER(vp),
                                               vp, ctypes.c_size_t, vp]
    lib.hint_mmd_workspace_bytes.restype = ctypes.c_size_t
    lib.hint_mmd_workspace_bytes.argtypes = [i64]
    lib.hint_multi_mmd.restype = ctypes.c_int
    lib.hint_multi_mmd.argtypes = [f32p, f32p, i64, i32, ctypes.POINTER(f32), ctypes.POINTER(f32), i32, f32p, vp, ctypes.c_size_t, vp]
    lib.hint_launch_count.restype = ctypes.c_uint64
    lib.hint_launch_count.argtypes = []
    lib.hint_last_error.restype = ctypes.c_char_p
    lib.hint_last_error.argtypes = []
    lib.hint_version.restype = ctypes.c_char_p
    lib.hint_version.argtypes = []
    _lib = lib
    return lib


def last_error():
    return load().hint_last_error().decode()


def check(rc):
    """Map a C status code onto the exception the reference's Python surface would raise."""
    if rc == HINT_OK:
        return
    msg = last_error()
    if rc == HINT_ERR_UNSUPPORTED:
        raise NotImplementedError("hint_b200: " + msg)
    if rc == HINT_ERR_INVALID:
        raise ValueError("hint_b200: " + msg)
    raise RuntimeError(f"hint_b200 (code {rc}): {msg}")


# ---- cheap device guard / stream lookup (torch.cuda.device + torch.cuda.current_stream cost ~20 us per call pair in Python,
# which is what bounds a small-batch step through the autograd wrappers) ----------------------------------------------------------
import contextlib as _contextlib

_NULL_CTX = _contextlib.nullcontext()


def on_device(device):
    """Context manager making ``device`` the current CUDA device; a no-op object when it already is."""
    import torch
    idx = device.index
    if idx is None or idx == torch._C._cuda_getDevice():
        return _NULL_CTX
    return torch.cuda.device(idx)


def stream_of(device):
    """Raw cudaStream_t of torch's current stream on ``device``."""
    import torch
    idx = device.index
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice() if idx is None else idx)
