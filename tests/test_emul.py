"""CPU tests of the kernels' host logic: the plan / schedule / packed-layout tables and the phase
functions of the FP32 kernels, executed by the host emulation (tests/emul) and compared with the
golden vectors of the real reference module.  (The GPU parity tests proper are in test_gpu_parity.py.)"""
import numpy as np
import pytest

import emul_lib
from conftest import plan_kwargs


def _rel(a, b):
    return float(np.abs(a.astype(np.float64) - b).max() / max(1e-30, np.abs(b).max()))


def test_emulated_kernels_match_reference(golden):
    pk = plan_kwargs(golden["meta"])
    x, c = golden["x"], golden.get("c")
    B = x.shape[0]
    args = (pk["d"], pk["dc"], pk["c_internal"], pk["clamp"], pk["max_splits"], pk["min_split_size"], golden["params"])
    # forward + backward of the NLL loss (train_unconditional.py:128-132): dz = z/B, dJ = -1/B
    z64 = golden["z64"]
    out = emul_lib.run(*args, x, c, rev=False, backward=(z64 / B, np.full(B, -1.0 / B)))
    assert _rel(out["z"], golden["z64"]) < 2e-5
    assert np.abs(out["J"] - golden["J64"]).max() < 2e-5 * max(1.0, np.abs(golden["J64"]).max())
    assert _rel(out["xrec"], golden["x"].astype(np.float64)) < 1e-4
    assert _rel(out["dx"], golden["dx64"]) < 2e-4
    assert _rel(out["dparams"], golden["dparams64"]) < 2e-4
    if c is not None:
        assert _rel(out["dc"], golden["dc64"]) < 2e-4
    # inverse direction
    inv = emul_lib.run(*args, x, c, rev=True)
    assert _rel(inv["z"], golden["xinv64"]) < 2e-5
    assert np.abs(inv["J"] - golden["Jinv64"]).max() < 2e-5 * max(1.0, np.abs(golden["Jinv64"]).max())


@pytest.mark.parametrize("nctas", [1, 2, 5])
def test_partial_gradient_reduction_is_independent_of_cta_count(nctas):
    from conftest import load_golden
    g = load_golden("gas_like_d8")
    pk = plan_kwargs(g["meta"])
    B = g["x"].shape[0]
    # replicate the batch so several tiles exist
    reps = 9
    x = np.tile(g["x"], (reps, 1))
    args = (pk["d"], pk["dc"], pk["c_internal"], pk["clamp"], pk["max_splits"], pk["min_split_size"], g["params"])
    z = np.tile(g["z64"], (reps, 1))
    out = emul_lib.run(*args, x, None, backward=(z / B, np.full(B * reps, -1.0 / B)), nctas=nctas)
    assert _rel(out["dparams"], reps * g["dparams64"]) < 2e-4
    assert _rel(out["dx"], np.tile(g["dx64"], (reps, 1))) < 2e-4
