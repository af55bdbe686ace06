"""Fused baseline couplings (hint_mlp_coupling_* of include/hint_b200.h) against the plain-PyTorch statement of
the same definition in FrEIA/modules/coupling.py, evaluated in float64 on the CPU.  fp32-grade bounds (the kernels use
error-compensated 3 x TF32 products): outputs 2e-5 max-norm relative, gradients 1e-4 relative L2 (looser for the deliberately ill-conditioned 0.3 * randn weights)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _rel(a, b):
    return float(torch.linalg.norm(a.detach().cpu().double() - b.detach()) / max(1e-30, float(torch.linalg.norm(b.detach()))))


def _mk(kind, d_in, d_c, hidden, seed, scale):
    from FrEIA.modules import AffineCoupling, ExternalAffineCoupling, F_fully_connected
    torch.manual_seed(seed)
    cls = AffineCoupling if kind == "affine" else ExternalAffineCoupling
    m = cls([(d_in,)], dims_c=[(d_c,)] if d_c else [], F_class=F_fully_connected, F_args={"internal_size": hidden})
    for p in m.parameters():
        p.data = scale * torch.randn_like(p)
    return m


CASES = [  # kind, d_in, d_c, hidden, B
    ("external", 20, 2, 68, 1000),      # lens conditional_hint_8_full: ac_y_to_x
    ("affine", 2, 0, 17, 1000),         # lens y lane
    ("external", 100, 4, 152, 333),     # plus conditional_hint_8_full
    ("affine", 4, 0, 38, 37),           # plus y lane
    ("affine", 100, 0, 224, 129),       # an inn baseline on the x lane (du = dv = 50), widest hidden layer of the configs
    ("affine", 5, 3, 9, 16),            # odd sizes, conditional AffineCoupling
    ("external", 1, 1, 1, 5),
    ("external", 128, 128, 256, 50),    # envelope corner
]


@pytest.mark.parametrize("kind,d_in,d_c,hidden,B", CASES)
@pytest.mark.parametrize("scale", [0.005, 0.3])
def test_fused_coupling_matches_the_pytorch_definition(dev, kind, d_in, d_c, hidden, B, scale):
    import hint_b200
    m = _mk(kind, d_in, d_c, hidden, 11 + d_in + hidden, scale)
    ref = copy.deepcopy(m).double()
    mg = copy.deepcopy(m).to(dev)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, d_in, generator=g)
    c = torch.randn(B, d_c, generator=g) if d_c else None
    gy = torch.randn(B, d_in, generator=g)
    gj = torch.randn(B, generator=g)

    x64 = x.double().requires_grad_(True)
    c64 = c.double().requires_grad_(True) if d_c else None
    y64 = ref([x64], [c64] if d_c else [])[0]
    j64 = ref.jacobian(None)
    (y64 * gy.double()).sum().add((j64 * gj.double()).sum()).backward()

    xg = x.to(dev).requires_grad_(True)
    cg = c.to(dev).requires_grad_(True) if d_c else None
    n0 = hint_b200._lib.load().hint_launch_count()
    yg = mg([xg], [cg] if d_c else [])[0]
    jg = mg.jacobian(None)
    assert hint_b200._lib.load().hint_launch_count() == n0 + 1           # ONE library launch forward
    (yg * gy.to(dev)).sum().add((jg * gj.to(dev)).sum()).backward()
    assert hint_b200._lib.load().hint_launch_count() - n0 in (3, 4)      # backward: fused kernel + weight gradients (+ split reduction)

    tol = 2e-5 if scale < 0.1 else 2e-4     # 0.3 * randn weights: |t|, |y| reach several hundred, s saturates the clamp
    assert float((yg.cpu().double() - y64).abs().max()) <= tol * max(1.0, float(y64.abs().max()))
    assert float((jg.cpu().double() - j64).abs().max()) <= tol * max(1.0, float(j64.abs().max()))
    gtol = 1e-4 if scale < 0.1 else 5e-3    # 0.3 * randn at hidden = 256: gain ~5 per layer, ReLU masks flip on fp32 rounding
    assert _rel(xg.grad, x64.grad) < gtol
    if d_c:
        assert _rel(cg.grad, c64.grad) < gtol
    for (name, pg), (_, pr) in zip(mg.named_parameters(), ref.named_parameters()):
        assert pg.grad is not None, name
        if float(torch.linalg.norm(pr.grad)) > 1e-12:
            assert _rel(pg.grad, pr.grad) < gtol, (name, _rel(pg.grad, pr.grad))
        else:
            assert float(pg.grad.abs().max()) < 1e-6, name

    # the inverse direction undoes the forward one (no autograd)
    with torch.no_grad():
        back = mg([yg.detach()], [cg.detach()] if d_c else [], rev=True)[0]
        jb = mg.jacobian(None)
    assert float((jb + jg.detach()).abs().max()) <= tol * max(1.0, float(j64.abs().max()))
    if scale < 0.1:      # well-conditioned (with the large weights (y - t) / e(s) cancels digits in any arithmetic)
        assert float((back - xg.detach()).abs().max()) <= 1e-4 * max(1.0, float(x.abs().max()))


def test_fused_coupling_is_deterministic_and_handles_empty_batches(dev):
    m = _mk("external", 20, 2, 68, 3, 0.1).to(dev)
    x = torch.randn(4097, 20, device=dev, requires_grad=True)
    c = torch.randn(4097, 2, device=dev)
    grads = []
    for _ in range(2):
        m.zero_grad()
        y = m([x], [c])[0]
        (y.pow(2).sum() + m.jacobian(None).sum()).backward()
        grads.append([p.grad.clone() for p in m.parameters()])
    for a, b in zip(*grads):
        assert torch.equal(a, b)
    y = m([x[:0]], [c[:0]])[0]
    assert y.shape == (0, 20) and m.jacobian(None).shape == (0,)


def test_paths_the_kernels_do_not_cover_fall_back_to_pytorch(dev):
    """Gradients through rev=True and dropout in training mode use the plain-PyTorch expressions (no library launch)."""
    import hint_b200
    from FrEIA.modules import AffineCoupling, F_fully_connected
    m = _mk("affine", 6, 0, 12, 9, 0.1).to(dev)
    x = torch.randn(50, 6, device=dev, requires_grad=True)
    n0 = hint_b200._lib.load().hint_launch_count()
    y = m([x], rev=True)[0]
    y.sum().backward()
    assert hint_b200._lib.load().hint_launch_count() == n0 and x.grad is not None
    md = AffineCoupling([(6,)], F_class=F_fully_connected, F_args={"internal_size": 12, "dropout": 0.5}).to(dev)
    md([x.detach()])
    assert hint_b200._lib.load().hint_launch_count() == n0
    md.eval()
    md([x.detach()])
    assert hint_b200._lib.load().hint_launch_count() == n0 + 1


def test_two_lane_conditional_model_trains_like_the_pytorch_couplings(dev, monkeypatch):
    """The lens `conditional_hint_*_full` architecture (configs/lens_shape/conditional_hint_8_full.py:61-102, two blocks here)
    through the FrEIA shim, driven like train_conditional.py:119-156 (NLL of both lanes, Adam): the run with the fused coupling
    kernels follows the run with the plain-PyTorch couplings (same seed, fp32 HINT kernels) step for step."""
    import hint_b200
    from FrEIA.framework import InputNode, Node, OutputNode, ReversibleGraphNet
    from FrEIA.modules import (HierarchicalAffineCouplingBlock, HouseholderPerm, AffineCoupling, ExternalAffineCoupling,
                               F_fully_connected)
    from FrEIA.modules import coupling as shim

    def build():
        torch.manual_seed(0)
        y_lane, x_lane = [InputNode(2, name="y")], [InputNode(20, name="x")]
        for i in range(2):
            if i > 0:
                y_lane.append(Node(y_lane[-1], HouseholderPerm, {"fixed": True, "n_reflections": 2}, name=f"perm_y_{i}"))
                x_lane.append(Node(x_lane[-1], HouseholderPerm, {"fixed": True, "n_reflections": 20}, name=f"perm_x_{i}"))
            x_lane.append(Node(x_lane[-1], HierarchicalAffineCouplingBlock, {"c_internal": [68, 34, 17, 17]}, name=f"hac_x_{i+1}"))
            x_lane.append(Node(x_lane[-1], ExternalAffineCoupling, {"F_class": F_fully_connected, "F_args": {"internal_size": 68}},
                               conditions=y_lane[-1], name=f"ac_y_to_x_{i+1}"))
            y_lane.append(Node(y_lane[-1], AffineCoupling, {"F_class": F_fully_connected, "F_args": {"internal_size": 17}}, name=f"ac_y_{i+1}"))
        y_lane.append(OutputNode(y_lane[-1], name="z_y"))
        x_lane.append(OutputNode(x_lane[-1], name="z_x"))
        m = ReversibleGraphNet(y_lane + x_lane, verbose=False)
        for p in m.parameters():
            if p.requires_grad:
                p.data = 0.05 * torch.randn_like(p)
        return m.to(dev)

    def run(m):
        g = torch.Generator().manual_seed(3)
        x = (1.0 + 2.0 * torch.randn(2000, 20, generator=g)).to(dev)
        y = torch.randn(2000, 2, generator=g).to(dev)
        opt = torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=1e-3)
        losses = []
        for _ in range(4):
            opt.zero_grad()
            z_y, z_x = m([y, x])
            J = m.log_jacobian(run_forward=False)
            loss = 0.5 * (z_y.pow(2).sum(1) + z_x.pow(2).sum(1)).mean() - J.mean()
            loss.backward()
            opt.step()
            losses.append(float(loss))
        with torch.no_grad():
            y2, x2 = m(list(m([y, x])), rev=True)
        return losses, float((x2 - x).abs().max()), float((y2 - y).abs().max())

    old = hint_b200.get_precision()
    hint_b200.set_precision("fp32")
    try:
        n0 = hint_b200._lib.load().hint_launch_count()
        fused, ex, ey = run(build())
        n_fused = hint_b200._lib.load().hint_launch_count() - n0
        monkeypatch.setattr(shim._AffineBase, "_fused", lambda self, u, v, rev: None)
        n0 = hint_b200._lib.load().hint_launch_count()
        plain, _, _ = run(build())
        n_plain = hint_b200._lib.load().hint_launch_count() - n0
    finally:
        hint_b200.set_precision(old)
    assert n_fused > n_plain                       # the couplings really ran on the library's kernels
    assert fused[-1] < fused[0]
    for a, b in zip(fused, plain):
        assert abs(a - b) <= 2e-4 * max(1.0, abs(b)), (fused, plain)
    assert ex < 1e-3 and ey < 1e-4
