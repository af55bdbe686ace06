"""CPU tests of the warp-MMA kernels (mma_kernels.cuh): the planner's operand packing / op streams and every index of the
kernels, executed by the fiber-based SIMT emulation (tests/emul/emul_mma.cpp: 256 cooperative fibers per CTA, mma.sync
modelled as a lane exchange with tf32 operand truncation) and compared with the golden vectors of the real reference module.
The GPU parity tests proper are in test_gpu_tf32.py."""
import numpy as np
import pytest

import emul_mma_lib
from conftest import load_golden, plan_kwargs

# fixtures kept small enough that the emulation of the whole suite stays well under a minute
FAST = ["tiny_d2", "tiny_d3_B1", "default_width_d5", "gas_like_d8", "power_like_d6", "single_width_d9", "two_conditions_d10",
        "min_split3_d13_clamp2", "lens_concat_cond_d20_dc2", "plus_ms0_d100", "plus_ms3_d100_narrow"]
SLOW = ["d43_hint8_widths", "plus_concat_cond_d100_dc4"]


def _rel(a, b):
    return float(np.abs(a.astype(np.float64) - b).max() / max(1e-30, np.abs(b).max()))


def _run(name, x3):
    g = load_golden(name)
    pk = plan_kwargs(g["meta"])
    x, c = g["x"], g.get("c")
    B = x.shape[0]
    args = (pk["d"], pk["dc"], pk["c_internal"], pk["clamp"], pk["max_splits"], pk["min_split_size"], g["params"])
    out = emul_mma_lib.run(*args, x, c, rev=False, backward=(g["z64"] / B, np.full(B, -1.0 / B)), x3=x3)
    inv = emul_mma_lib.run(*args, x, c, rev=True, x3=x3)
    return g, out, inv


@pytest.mark.parametrize("name", FAST + SLOW)
def test_emulated_3xtf32_matches_reference(name):
    """3xTF32 mode: fp32-class accuracy (same bounds as the FP32 CUDA-core path's emulation test)."""
    g, out, inv = _run(name, True)
    assert _rel(out["z"], g["z64"]) < 2e-5
    assert np.abs(out["J"] - g["J64"]).max() < 2e-5 * max(1.0, np.abs(g["J64"]).max())
    assert _rel(out["xrec"], g["x"].astype(np.float64)) < 1e-4
    assert _rel(out["dx"], g["dx64"]) < 2e-4
    assert _rel(out["dparams"], g["dparams64"]) < 2e-4
    if g.get("c") is not None:
        assert _rel(out["dc"], g["dc64"]) < 2e-4
    assert _rel(inv["z"], g["xinv64"]) < 2e-5
    assert np.abs(inv["J"] - g["Jinv64"]).max() < 2e-5 * max(1.0, np.abs(g["Jinv64"]).max())


@pytest.mark.parametrize("name", FAST)
def test_emulated_tf32_within_stated_bound(name):
    """Single-pass TF32: z / log-det within 5e-3 (stated bound, test_gpu_tf32.py); gradients in relative L2 norm."""
    g, out, inv = _run(name, False)
    assert _rel(out["z"], g["z64"]) < 5e-3
    assert np.abs(out["J"] - g["J64"]).max() < 5e-3 * max(1.0, np.abs(g["J64"]).max())
    assert _rel(inv["z"], g["xinv64"]) < 5e-3
    for k, ref in (("dx", "dx64"), ("dparams", "dparams64")):
        num = np.linalg.norm(out[k].astype(np.float64) - g[ref])
        assert num / max(1e-30, np.linalg.norm(g[ref])) < 2e-2


@pytest.mark.parametrize("nctas", [1, 2, 5])
def test_partial_gradient_reduction_is_independent_of_cta_count(nctas):
    g = load_golden("gas_like_d8")
    pk = plan_kwargs(g["meta"])
    B = g["x"].shape[0]
    reps = 5   # several tiles per CTA: exercises the store-then-reduce partial-gradient accumulation
    x = np.tile(g["x"], (reps, 1))
    args = (pk["d"], pk["dc"], pk["c_internal"], pk["clamp"], pk["max_splits"], pk["min_split_size"], g["params"])
    z = np.tile(g["z64"], (reps, 1))
    out = emul_mma_lib.run(*args, x, None, backward=(z / B, np.full(B * reps, -1.0 / B)), nctas=nctas, x3=True)
    assert _rel(out["dparams"], reps * g["dparams64"]) < 2e-4
    assert _rel(out["dx"], np.tile(g["dx64"], (reps, 1))) < 2e-4
