"""CPU tests: the oracle against the golden vectors produced by the real reference hint.py
(oracle/gen_golden.py) and against the parameter-count known-answers of the reference configs."""
import numpy as np
import pytest
import torch

from conftest import plan_kwargs
from oracle import hint_oracle as O


def _plan(meta):
    pk = plan_kwargs(meta)
    plan = O.build_plan(pk["d"], pk["dc"], pk["c_internal"], pk["max_splits"], pk["min_split_size"])
    return plan, pk


@pytest.mark.parametrize("kat", O.PARAM_COUNT_KATS, ids=lambda k: k["name"])
def test_param_count_kat(kat):
    plan = O.build_plan(kat["d"], kat["dc"], kat["c_internal"], kat["max_splits"])
    assert O.param_count(plan) == kat["per_block"]
    assert O.param_count(plan) * kat["n_blocks"] == kat["total"]
    e = O.param_entries(plan)
    assert e[-1].offset + int(np.prod(e[-1].shape)) == kat["per_block"]


def test_state_dict_names_and_shapes(golden):
    plan, _ = _plan(golden["meta"])
    mine = [(e.name, list(e.shape)) for e in O.param_entries(plan)]
    assert mine == [(k, s) for k, s in golden["meta"]["state_dict"]]
    assert O.param_count(plan) == golden["params"].size


def test_forward_and_inverse_match_reference(golden):
    plan, pk = _plan(golden["meta"])
    for tag, dt, tol in (("", torch.float32, 2e-5), ("64", torch.float64, 1e-12)):
        flat = torch.from_numpy(golden["params"]).to(dt)
        x = torch.from_numpy(golden["x"]).to(dt)
        c = torch.from_numpy(golden["c"]).to(dt) if "c" in golden else None
        for fn in (O.forward, O.forward_fast):
            z, J = fn(plan, flat, x, c, rev=False, clamp=pk["clamp"])
            scale = max(1.0, float(np.abs(golden["z" + tag]).max()))
            assert np.abs(z.detach().numpy() - golden["z" + tag]).max() <= tol * scale
            assert np.abs(J.detach().numpy() - golden["J" + tag]).max() <= tol * max(1.0, float(np.abs(golden["J" + tag]).max()))
            xi, Ji = fn(plan, flat, x, c, rev=True, clamp=pk["clamp"])
            scale = max(1.0, float(np.abs(golden["xinv" + tag]).max()))
            assert np.abs(xi.detach().numpy() - golden["xinv" + tag]).max() <= tol * scale
            assert np.abs(Ji.detach().numpy() - golden["Jinv" + tag]).max() <= tol * max(1.0, float(np.abs(golden["Jinv" + tag]).max()))


def test_blockwise_timing_ports_match_reference(golden):
    plan, pk = _plan(golden["meta"])
    dt = torch.float64
    flat = torch.from_numpy(golden["params"]).to(dt)
    x = torch.from_numpy(golden["x"]).to(dt)
    c = torch.from_numpy(golden["c"]).to(dt) if "c" in golden else None
    z, J = O.forward_blockwise(plan, flat, x, c, clamp=pk["clamp"])
    assert np.abs(z.numpy() - golden["z64"]).max() <= 1e-12 * max(1.0, np.abs(golden["z64"]).max())
    assert np.abs(J.numpy() - golden["J64"]).max() <= 1e-12 * max(1.0, np.abs(golden["J64"]).max())
    xi, Ji = O.inverse_blockwise(plan, flat, x, c, clamp=pk["clamp"])
    assert np.abs(xi.numpy() - golden["xinv64"]).max() <= 1e-12 * max(1.0, np.abs(golden["xinv64"]).max())
    assert np.abs(Ji.numpy() - golden["Jinv64"]).max() <= 1e-12 * max(1.0, np.abs(golden["Jinv64"]).max())


def _rel(a, b):
    return float(np.abs(a - b).max() / max(1e-30, np.abs(b).max()))


def test_gradients_match_reference(golden):
    """autograd through the oracle AND the hand-written memory-free backward vs reference grads (fp64)."""
    plan, pk = _plan(golden["meta"])
    dt = torch.float64
    flat = torch.from_numpy(golden["params"]).to(dt).requires_grad_(True)
    x = torch.from_numpy(golden["x"]).to(dt).requires_grad_(True)
    c = torch.from_numpy(golden["c"]).to(dt).requires_grad_(True) if "c" in golden else None
    z, J = O.forward(plan, flat, x, c, clamp=pk["clamp"])
    loss = O.nll_loss(z, J)
    loss.backward()
    assert abs(loss.item() - float(golden["loss64"])) <= 1e-10 * max(1.0, abs(float(golden["loss64"])))
    assert _rel(x.grad.numpy(), golden["dx64"]) < 1e-10
    assert _rel(flat.grad.numpy(), golden["dparams64"]) < 1e-10
    if c is not None:
        assert _rel(c.grad.numpy(), golden["dc64"]) < 1e-10
    # hand-written backward from the OUTPUT (the algorithm the CUDA kernel follows)
    B = x.shape[0]
    with torch.no_grad():
        dz = z.detach() / B
        dJ = torch.full((B,), -1.0 / B, dtype=dt)
        xr, dx, dc, dflat = O.backward_from_output(plan, flat.detach(), z.detach(), None if c is None else c.detach(),
                                                   dz, dJ, clamp=pk["clamp"])
    assert _rel(xr.numpy(), golden["x"].astype(np.float64)) < 1e-7
    assert _rel(dx.numpy(), golden["dx64"]) < 1e-7
    assert _rel(dflat.numpy(), golden["dparams64"]) < 1e-7
    if c is not None:
        assert _rel(dc.numpy(), golden["dc64"]) < 1e-7


def test_logdet_is_slogdet_of_autograd_jacobian():
    """Property pin: returned J == log|det d z/d x| on a small block."""
    torch.manual_seed(0)
    plan = O.build_plan(7, 2, [9, 5])
    flat = 0.3 * torch.randn(O.param_count(plan), dtype=torch.float64)
    x = torch.randn(3, 7, dtype=torch.float64)
    c = torch.randn(3, 2, dtype=torch.float64)
    _, J = O.forward(plan, flat, x, c)
    for b in range(3):
        jac = torch.autograd.functional.jacobian(lambda v: O.forward(plan, flat, v[None], c[b:b + 1])[0][0], x[b])
        sign, logabs = torch.linalg.slogdet(jac)
        assert abs(logabs.item() - J[b].item()) < 1e-10


def test_flops_table():
    # SURVEY.md section 8d / BASELINE.md section 4
    for name, per_block in (("plus_hint_4_3", 972792), ("plus_hint_4_full", 967756), ("power_hint_8", 122640),
                            ("gas_hint_8", 121856), ("miniboone_hint_8", 59084), ("d43_hint_8", 59612),
                            ("plus_cond_recursive_4", 1961660), ("lens_xlane_hint_8_full", 52496)):
        kat = next(k for k in O.PARAM_COUNT_KATS if k["name"] == name)
        plan = O.build_plan(kat["d"], kat["dc"], kat["c_internal"], kat["max_splits"])
        assert O.flops_per_sample(plan) == per_block
