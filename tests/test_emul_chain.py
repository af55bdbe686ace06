"""CPU tests of the register-chained warp-MMA kernels (chain_kernels.cuh): the planner's super-node grouping, operand
packing and column maps, and every index of the forward / inverse / backward kernels, executed by the fiber-based SIMT
emulation (tests/emul/emul_chain.cpp: cooperative fibers per CTA, mma.sync and ldmatrix modelled as lane exchanges with tf32
operand truncation) and compared with the golden vectors of the real reference module (hint.py:62-101 and its autograd
tape).  The GPU parity tests proper are in test_gpu_tf32.py (mode "tf32_chain")."""
import numpy as np
import pytest

import emul_chain_lib
from conftest import load_golden, plan_kwargs

FIXTURES = ["tiny_d2", "tiny_d3_B1", "default_width_d5", "gas_like_d8", "power_like_d6", "single_width_d9",
            "two_conditions_d10", "min_split3_d13_clamp2", "lens_concat_cond_d20_dc2", "lens_xlane_d20", "d43_hint8_widths",
            "d42_hint8_widths_small_init"]
TF32_TOL = 5e-3   # stated single-pass TF32 bound (test_gpu_tf32.py)


def _rel(a, b):
    return float(np.abs(a.astype(np.float64) - b).max() / max(1.0, np.abs(b).max()))


def _l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(1e-30, np.linalg.norm(b)))


def _args(g):
    pk = plan_kwargs(g["meta"])
    return (pk["d"], pk["dc"], pk["c_internal"], pk["clamp"], pk["max_splits"], pk["min_split_size"], g["params"])


@pytest.mark.parametrize("cfg", [1, 2, 3], ids=["mt1_nw4", "mt2_nw4", "mt1_nw8"])
@pytest.mark.parametrize("name", FIXTURES)
def test_emulated_chain_matches_reference(name, cfg):
    """cfg selects the tile geometry: forward MT = 1 (16 warps) or 2 (8 warps); backward (MT, NW) = (1,4), (2,4), (1,8)."""
    g = load_golden(name)
    x, c = g["x"], g.get("c")
    B = x.shape[0]
    out = emul_chain_lib.run(*_args(g), x, c, rev=False, backward=(g["z64"] / B, np.full(B, -1.0 / B)), mt=cfg, nctas=2)
    inv = emul_chain_lib.run(*_args(g), x, c, rev=True, mt=cfg)
    assert _rel(out["z"], g["z64"]) < TF32_TOL
    assert _rel(out["J"], g["J64"]) < TF32_TOL
    assert _rel(inv["z"], g["xinv64"]) < TF32_TOL
    assert _rel(inv["J"], g["Jinv64"]) < TF32_TOL
    # the backward inverts the TF32 forward with the same TF32 subnets: the reconstruction is much tighter than the bound
    assert _rel(out["xrec"], x.astype(np.float64)) < 1e-4
    assert _l2(out["dx"], g["dx64"]) < 2e-2
    assert _l2(out["dparams"], g["dparams64"]) < 2e-2
    if c is not None:
        assert _l2(out["dc"], g["dc64"]) < 3e-2


def test_super_nodes_group_the_tiny_levels():
    """d=43 `hint_8` tree (31 nodes, hint.py:25-54): the 8 + 16 nodes of the two deepest levels (h = 8) fuse into 4 + 4 super
    nodes, the rest stay single: 15 chain nodes."""
    g = load_golden("d43_hint8_widths")
    out = emul_chain_lib.run(*_args(g), g["x"][:16], None, mt=1)
    assert out["info"][0] == 1 and out["info"][3] == 15


@pytest.mark.parametrize("nctas", [1, 3])
def test_partial_gradients_accumulate_over_tiles_and_ctas(nctas):
    g = load_golden("gas_like_d8")
    B = g["x"].shape[0]
    reps = 5   # 320 samples = 5 tiles of 64: store on the first tile of a CTA, accumulate afterwards
    x = np.tile(g["x"], (reps, 1))
    z = np.tile(g["z64"], (reps, 1))
    out = emul_chain_lib.run(*_args(g), x, None, backward=(z / B, np.full(B * reps, -1.0 / B)), nctas=nctas, mt=1)
    assert _l2(out["dparams"], reps * g["dparams64"]) < 2e-2
    assert _l2(out["dx"], np.tile(g["dx64"], (reps, 1))) < 2e-2


def test_ragged_last_tile():
    g = load_golden("two_conditions_d10")
    x, c = g["x"][:21], g["c"][:21]
    out = emul_chain_lib.run(*_args(g), x, c, mt=2)
    assert _rel(out["z"], g["z64"][:21]) < TF32_TOL and not np.isnan(out["J"]).any()


def test_outside_envelope_is_reported():
    g = load_golden("wide_h_d12")
    with pytest.raises(LookupError):
        emul_chain_lib.run(*_args(g), g["x"], g.get("c"), mt=1)


def _random_cfg(rng):
    d = int(rng.integers(2, 30))
    dc = int(rng.choice([0, 0, 1, 3]))
    n_w = int(rng.integers(1, 5))
    widths = [int(rng.integers(3, 41)) for _ in range(n_w)]
    ms = int(rng.choice([-1, -1, 0, 1, 2, 3]))
    mss = int(rng.choice([2, 2, 3, 4]))
    return d, dc, widths, ms, mss


@pytest.mark.parametrize("seed", range(8))
def test_random_trees_against_the_oracle(seed):
    """Random tree shapes (odd widths, conditions, split limits): planner grouping / packing / column maps of the chain kernels
    against the fp64 oracle (oracle/hint_oracle.py, itself pinned to the real hint.py by the golden vectors)."""
    import torch
    from oracle import hint_oracle as O
    rng = np.random.default_rng(1000 + seed)
    d, dc, widths, ms, mss = _random_cfg(rng)
    plan = O.build_plan(d, dc, widths, ms, mss)
    n = O.param_count(plan)
    params = (0.3 * rng.standard_normal(n)).astype(np.float32)
    B = int(rng.integers(1, 70))
    x = rng.standard_normal((B, d)).astype(np.float32)
    c = rng.standard_normal((B, dc)).astype(np.float32) if dc else None
    p64 = torch.from_numpy(params).double()
    c64 = None if c is None else torch.from_numpy(c).double()
    z_ref, J_ref = O.forward_fast(plan, p64, torch.from_numpy(x).double(), c64)
    dz = torch.from_numpy(rng.standard_normal((B, d))).double() / B
    dJ = torch.from_numpy(rng.standard_normal(B)).double() / B
    _, dx_ref, dc_ref, dp_ref = O.backward_from_output(plan, p64, z_ref, c64, dz, dJ)
    try:
        out = emul_chain_lib.run(d, dc, widths, 4.0, ms, mss, params, x, c, backward=(dz.numpy(), dJ.numpy()), mt=1 + seed % 3, nctas=2)
    except LookupError:
        pytest.skip("outside the chain envelope")
    assert _rel(out["z"], z_ref.numpy()) < TF32_TOL and _rel(out["J"], J_ref.numpy()) < TF32_TOL
    assert _l2(out["dx"], dx_ref.numpy()) < 2e-2 and _l2(out["dparams"], dp_ref.numpy()) < 2e-2
    if dc:
        assert _l2(out["dc"], dc_ref.numpy()) < 3e-2
