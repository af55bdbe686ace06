"""FrEIA-compatible shim: the graph runtime and HouseholderPerm the reference's configs / train scripts need
around the HINT block.  CPU tests: construction through the reference's own call pattern (configs/**: Node(...),
ReversibleGraphNet(...)), the parameter-count known answers of the config comments, HouseholderPerm algebra; the architecture
section of real reference config files is exec'd unmodified when /root/reference is available.  The GPU test runs a two-block
model end to end against the fp64 oracle composition."""
import os
import re

import numpy
import pytest
import torch

from oracle import hint_oracle as O

REF = "/root/reference"


def _build(d, n_blocks, c_internal, max_splits=None, dims_c=0):
    from FrEIA.framework import InputNode, ConditionNode, Node, OutputNode, ReversibleGraphNet
    from FrEIA.modules import HierarchicalAffineCouplingBlock, HouseholderPerm
    lane = [InputNode(d, name="x")]
    cond = ConditionNode(dims_c, name="y") if dims_c else None
    for i in range(n_blocks):
        if i > 0:
            lane.append(Node(lane[-1], HouseholderPerm, {"fixed": True, "n_reflections": d}, name=f"perm_{i}"))
        args = {"c_internal": list(c_internal)}
        if max_splits is not None:
            args["max_splits"] = max_splits
        lane.append(Node(lane[-1], HierarchicalAffineCouplingBlock, args, conditions=cond if cond is not None else [], name=f"hac_{i + 1}"))
    lane.append(OutputNode(lane[-1], name="z"))
    nodes = lane + ([cond] if cond is not None else [])
    if cond is not None:      # a condition must be listed before the nodes that use it
        nodes = [lane[0], cond] + lane[1:]
    return ReversibleGraphNet(nodes, verbose=False)


@pytest.mark.parametrize("d,n_blocks,h,widths,ms,expected", [
    (42, 8, 67, lambda h: [h, h // 2, h // 4, h // 8], None, 250624),       # configs/uci_data/miniboone_hint_8.py:31 "250k"
    (6, 8, 140, lambda h: [h, h // 2, h // 4, h // 8], None, 499632),        # configs/uci_data/power_hint_8.py:30 "500k"
    (8, 8, 128, lambda h: [h, h // 2, h // 4, h // 8], None, 499904),        # gas_hint_8 "500k"
    (100, 4, 314, lambda h: [h, h // 2, h // 4, h // 8], 3, 1967248),        # plus_shape/unconditional_hint_4_3.py:31 "2M"
])
def test_parameter_counts_through_the_shim(d, n_blocks, h, widths, ms, expected):
    model = _build(d, n_blocks, widths(h), ms)
    trainable = list(filter(lambda p: p.requires_grad, model.parameters()))     # the configs' own filter
    assert sum(p.numel() for p in trainable) == expected
    assert [n.name for n in model.node_list][:3] == ["x", "hac_1", "perm_1"]
    assert model.node_list[1].module.__class__.__name__ == "HierarchicalAffineCouplingBlock"


def test_star_imports_leak_numpy_and_torch():
    ns = {}
    exec("from FrEIA.framework import *\nfrom FrEIA.modules import *", ns)
    assert ns["np"] is numpy and ns["torch"] is torch
    for name in ("InputNode", "ConditionNode", "Node", "OutputNode", "ReversibleGraphNet", "HierarchicalAffineCouplingBlock",
                 "HouseholderPerm"):
        assert name in ns


def test_householder_is_orthogonal_and_inverts():
    from FrEIA.modules import HouseholderPerm
    torch.manual_seed(0)
    for fixed in (True, False):
        P = HouseholderPerm([(9,)], n_reflections=9, fixed=fixed).double()
        if fixed:
            P.W = P._matrix(P.Vs.detach())
        x = torch.randn(13, 9, dtype=torch.float64)
        y = P([x])[0]
        assert torch.allclose(P([y], rev=True)[0], x, atol=1e-12)
        assert torch.allclose(y.norm(dim=1), x.norm(dim=1), atol=1e-12)      # |det| = 1: log-jacobian 0
        assert P.jacobian([x]) == 0
        assert P.output_dims([(9,)]) == [(9,)]
        assert any(p.requires_grad for p in P.parameters()) == (not fixed)
    # state_dict round trip restores the same mixing
    A, B = HouseholderPerm([(5,)], n_reflections=5, fixed=True), HouseholderPerm([(5,)], n_reflections=5, fixed=True)
    B.load_state_dict(A.state_dict())
    assert torch.allclose(A.W, B.W)


def test_conditions_reach_the_block_as_dims_c():
    model = _build(20, 2, [68, 34, 17, 17], dims_c=2)
    blk = model.node_list[2].module
    assert blk.plan.dc == 2 and blk.dims_c == [(2,)]


def test_baseline_couplings_are_invertible_with_the_stated_jacobian():
    """AffineCoupling / ExternalAffineCoupling / F_fully_connected (plain PyTorch, off the hot path, parity-unpinned): reverse
    undoes forward, the cached log-det equals slogdet of the autograd Jacobian, and a two-lane graph whose x lane is
    conditioned on an INTERNAL y-lane node reverses (configs/lens_shape/conditional_hint_8_full.py:61-102)."""
    from FrEIA.framework import InputNode, Node, OutputNode, ReversibleGraphNet
    from FrEIA.modules import AffineCoupling, ExternalAffineCoupling, F_fully_connected, HouseholderPerm
    torch.manual_seed(0)
    ac = AffineCoupling([(5,)], F_class=F_fully_connected, F_args={"internal_size": 7}).double()
    x = torch.randn(3, 5, dtype=torch.float64)
    y = ac([x])[0]
    J = ac.jacobian(None)
    assert torch.allclose(ac([y], rev=True)[0], x, atol=1e-10) and torch.allclose(ac.jacobian(None), -J)
    jac = torch.autograd.functional.jacobian(lambda v: ac([v[None]])[0][0], x[0])
    assert abs(float(torch.slogdet(jac)[1]) - float(J[0])) < 1e-9
    ext = ExternalAffineCoupling([(4,)], dims_c=[(2,)], F_class=F_fully_connected, F_args={"internal_size": 6}).double()
    c = torch.randn(3, 2, dtype=torch.float64)
    xe = torch.randn(3, 4, dtype=torch.float64)
    assert torch.allclose(ext([ext([xe], c=[c])[0]], c=[c], rev=True)[0], xe, atol=1e-10)
    # two lanes, x conditioned on an internal y node; the x lane is LISTED FIRST to exercise the dependency ordering
    yl = [InputNode(2, name="y")]
    xl = [InputNode(4, name="x")]
    yl.append(Node(yl[-1], HouseholderPerm, {"fixed": True, "n_reflections": 2}, name="perm_y"))
    xl.append(Node(xl[-1], ExternalAffineCoupling, {"F_class": F_fully_connected, "F_args": {"internal_size": 5}}, conditions=yl[-1], name="y_to_x"))
    yl.append(Node(yl[-1], AffineCoupling, {"F_class": F_fully_connected, "F_args": {"internal_size": 5}}, name="ac_y"))
    yl.append(OutputNode(yl[-1], name="z_y"))
    xl.append(OutputNode(xl[-1], name="z_x"))
    net = ReversibleGraphNet([yl[0]] + xl + yl[1:], verbose=False)
    yy, xx = torch.randn(6, 2), torch.randn(6, 4)
    zx, zy = net([yy, xx])          # outputs in node-list order: z_x is listed before z_y here
    y2, x2 = net([zx, zy], rev=True)
    assert torch.allclose(y2, yy, atol=1e-5) and torch.allclose(x2, xx, atol=1e-5)
    assert net.log_jacobian(run_forward=False).shape == (6,)
    assert any(k.startswith("module_list.2.") for k in net.state_dict())   # keyed by the node's position in the list


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not available")
@pytest.mark.parametrize("cfg,ndim,expected", [
    ("configs/uci_data/miniboone_hint_8.py", 42, 250624),
    ("configs/uci_data/power_hint_8.py", 6, 499632),
])
def test_reference_config_architecture_section_runs_unmodified(cfg, ndim, expected):
    """The MODEL ARCHITECTURE section of the real config (x_lane = [...] ... ReversibleGraphNet(...)) exec'd verbatim against
    the shim; only the data-set dependent dictionary entries are supplied by hand (data.py needs files that are not shipped)."""
    src = open(os.path.join(REF, cfg)).read()
    hidden = int(re.search(r"^\s*'hidden_layer_sizes':\s*(\d+)", src, re.M).group(1))
    n_blocks = int(re.search(r"^\s*'n_blocks':\s*(\d+)", src, re.M).group(1))
    arch = src[src.index("x_lane = [InputNode"):]
    arch = arch[:arch.index("model.to(c['device'])")]
    ns = {"c": {"ndim_x": ndim, "n_blocks": n_blocks, "hidden_layer_sizes": hidden}}
    exec("from FrEIA.framework import *\nfrom FrEIA.modules import *\n" + arch, ns)
    model = ns["model"]
    assert sum(p.numel() for p in model.parameters() if p.requires_grad) == expected


@pytest.mark.gpu
def test_two_block_model_matches_oracle_composition():
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    d, ci, B = 10, [24, 12, 6], 257
    model = _build(d, 2, ci)
    blocks = [n.module for n in model.node_list if n.module is not None and n.module.__class__.__name__.startswith("Hier")]
    perm = [n.module for n in model.node_list if n.module is not None and n.module.__class__.__name__ == "HouseholderPerm"][0]
    with torch.no_grad():
        for b in blocks:
            b.flat.mul_(0.5)
    flats = [b.flat.detach().double().clone() for b in blocks]
    W = perm.W.double().clone()
    model.to(dev)
    x = torch.randn(B, d)
    plan = O.build_plan(d, 0, ci)
    z1, J1 = O.forward_fast(plan, flats[0], x.double())
    z2, J2 = O.forward_fast(plan, flats[1], z1 @ W)
    xg = x.to(dev).requires_grad_(True)
    z = model(xg)
    logj = model.log_jacobian(xg, run_forward=False)
    assert (z.detach().cpu().double() - z2).abs().max() < 1e-5 * max(1.0, z2.abs().max())
    assert (logj.detach().cpu().double() - (J1 + J2)).abs().max() < 1e-5 * max(1.0, (J1 + J2).abs().max())
    loss = 0.5 * torch.sum(z ** 2, dim=1).mean() - logj.mean()        # train_unconditional.py:128-129
    loss.backward()
    assert all(b.flat.grad is not None and torch.isfinite(b.flat.grad).all() for b in blocks)
    with torch.no_grad():
        xr = model(z.detach(), rev=True)
    assert (xr.cpu() - x).abs().max() < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fp32", "tf32"])
def test_training_recipe_through_the_shim_reduces_the_nll(mode):
    """The reference's training recipe (train_unconditional.py:121-144,165-176: 0.005*randn init, noise, NLL, clamp +-5,
    Adam(lr .01, betas (.9,.95), eps 1e-4, weight decay)) on a miniboone-shaped model built through the shim."""
    import hint_b200
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    d, B = 42, 4096
    model = _build(d, 4, [67, 33, 16, 8]).to(dev)
    params = list(filter(lambda p: p.requires_grad, model.parameters()))
    for p in params:
        p.data = 0.005 * torch.randn_like(p.data)
    optim = torch.optim.Adam(params, lr=0.01, betas=(0.9, 0.95), eps=1e-4, weight_decay=1.86e-5)
    means = 2.0 * torch.randn(4, d, device=dev)
    x_all = means[torch.randint(0, 4, (B,), device=dev)] + 0.5 * torch.randn(B, d, device=dev)
    x_all = (x_all - x_all.mean(0)) / x_all.std(0)
    prev = hint_b200.get_precision()
    hint_b200.set_precision(mode)
    try:
        losses = []
        for it in range(40):
            optim.zero_grad()
            x = x_all + 0.01 * torch.randn_like(x_all)
            z = model(x)
            logj = model.log_jacobian(x, run_forward=False)
            loss = 0.5 * torch.sum(z ** 2, dim=1).mean() - logj.mean()
            loss.backward()
            for p in params:
                p.grad.data.clamp_(-5.0, 5.0)
            optim.step()
            losses.append(loss.item())
    finally:
        hint_b200.set_precision(prev)
    assert all(map(lambda v: v == v, losses))
    assert losses[-1] < losses[0] - 1.0
