"""GPU tests of the training edge (train_unconditional.py:121-144,174-176) through the C ABI:
hint_add_noise, hint_nll_loss, hint_backward_nll, hint_adam_step, hint_launch_count and the autograd-free FusedTrainStep
built from them.  Oracles: torch (fp64 where it matters) for the elementwise / reduction kernels, the library's own
hint_backward with a materialised loss gradient for the fused NLL gradient, and the reference-surface step
(module forward + loss.backward() + clamp_ + torch.optim.Adam) for the whole step."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ADAM = dict(lr=0.01, betas=(0.9, 0.95), eps=1e-4, weight_decay=1.86e-5)   # configs/uci_data/miniboone_hint_8.py:38-44


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


def test_fused_clamp_adam_matches_torch(dev):
    """Tolerance 2e-6 relative to max|p| after 5 steps: same formula as torch.optim.Adam, different fp32 association."""
    from hint_b200 import FusedClampAdam
    g = torch.Generator().manual_seed(0)
    sizes = [1, 3, 4, 1021, 4096, 31598]
    ps_a = [torch.randn(n, generator=g).to(dev).requires_grad_(True) for n in sizes]
    ps_b = [p.detach().clone().requires_grad_(True) for p in ps_a]
    opt_a = FusedClampAdam(ps_a, grad_clamp=5.0, **ADAM)
    opt_b = torch.optim.Adam(ps_b, **ADAM)
    for step in range(5):
        for pa, pb in zip(ps_a, ps_b):
            gr = (10.0 * torch.randn(pa.shape, generator=g)).to(dev)   # |g| > 5 occurs: the clamp matters
            pa.grad = gr.clone()
            pb.grad = gr.clone().clamp_(-5.0, 5.0)
        opt_a.step()
        opt_b.step()
    for pa, pb in zip(ps_a, ps_b):
        assert float((pa - pb).detach().abs().max()) <= 2e-6 * max(1.0, float(pb.detach().abs().max()))
    assert opt_a.state[ps_a[0]]["step"] == 5


def test_adam_rejects_cpu_tensors():
    from hint_b200 import FusedClampAdam
    p = torch.zeros(4, requires_grad=True)
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        FusedClampAdam([p]).step()


@pytest.mark.parametrize("B,d,nj", [(1, 1, 1), (77, 43, 3), (4096, 6, 8), (100000, 43, 1)])
def test_nll_loss_matches_fp64(dev, B, d, nj):
    from hint_b200 import nll_loss_fused
    g = torch.Generator().manual_seed(B)
    z = torch.randn(B, d, generator=g)
    js = [torch.randn(B, generator=g) for _ in range(nj)]
    ref0 = 0.5 * (z.double() ** 2).sum(1).mean()
    ref1 = sum(j.double() for j in js).mean()
    out = nll_loss_fused(z.to(dev), [j.to(dev) for j in js]).cpu().double()
    assert abs(float(out[1] - ref0)) <= 1e-6 * max(1.0, abs(float(ref0)))
    assert abs(float(out[2] - ref1)) <= 1e-6 * max(1.0, abs(float(ref1)))
    assert abs(float(out[0] - (ref0 - ref1))) <= 2e-6 * max(1.0, abs(float(ref0)), abs(float(ref1)))


def test_noise_is_standard_normal_and_counter_based(dev):
    from hint_b200 import add_noise
    x = torch.zeros(1 << 20, 3, device=dev)[:-1]   # numel not a multiple of 4: exercises the tail
    a = add_noise(x, 1.0, seed=7, offset=0)
    b = add_noise(x, 1.0, seed=7, offset=0)
    c = add_noise(x, 1.0, seed=7, offset=1)
    d = add_noise(x, 1.0, seed=8, offset=0)
    assert torch.equal(a, b) and not torch.equal(a, c) and not torch.equal(a, d)
    v = a.double().flatten()
    n = v.numel()
    assert abs(float(v.mean())) < 5.0 / np.sqrt(n) and abs(float(v.std()) - 1.0) < 5.0 / np.sqrt(2 * n)
    assert abs(float((v ** 3).mean())) < 0.02 and abs(float((v ** 4).mean()) - 3.0) < 0.05   # skewness 0, kurtosis 3
    assert abs(float((v[:-1] * v[1:]).mean())) < 5.0 / np.sqrt(n)                              # neighbours uncorrelated
    assert float(a.abs().max()) < 7.0 and torch.isfinite(a).all()
    y = torch.randn(1000, 43, device=dev)
    out = add_noise(y, 0.01, seed=1)
    assert 0.008 < float((out - y).std()) < 0.012


CONFIGS = [("d43_chain", 43, 0, [67, 33, 16, 8], "tf32"), ("lens_cond_chain", 20, 2, [68, 34, 17, 17], "tf32"),
           ("gas_tc3", 8, 0, [128, 64, 32, 16], "tf32"), ("power_tc3", 6, 0, [140, 70, 35, 17], "tf32_tc3"),
           ("gas_fp32_materialised", 8, 0, [128, 64, 32, 16], "fp32")]


@pytest.mark.parametrize("name,d,dc,ci,mode", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_backward_nll_equals_backward_with_materialised_gradient(dev, name, d, dc, ci, mode):
    """The loss gradient generated in the tile load (dz = z/B, dlogdet = -1/B) is the same fp32 product torch computes, and the
    kernels are deterministic: results agree to a few ulp (the tcgen05 kernel accumulates its partials with RED, ordered).
    Ragged batch (several tiles, partial last tile)."""
    from hint_b200.block import TreePlan
    tp = TreePlan(d, dc, ci, 4.0, -1, 2, False)
    B = 128 * 5 + 37
    g = torch.Generator().manual_seed(3)
    flat = (0.05 * torch.randn(tp.n_params, generator=g)).to(dev)
    z = torch.randn(B, d, generator=g).to(dev)
    c = torch.randn(B, dc, generator=g).to(dev) if dc else None
    s = 1.0 / B
    dx0, dc0, dp0, _ = tp.backward(z, c, flat, z * s, torch.full((B,), -s, device=dev), mode=mode)
    dx1, dc1, dp1, _ = tp.backward(z, c, flat, None, None, mode=mode, nll_scale=s)
    tol = lambda ref: 1e-6 * max(1e-30, float(ref.abs().max()))
    assert float((dx1 - dx0).abs().max()) <= tol(dx0) and float((dp1 - dp0).abs().max()) <= tol(dp0)
    if dc:
        assert float((dc1 - dc0).abs().max()) <= tol(dc0)
    # earlier blocks: dz from memory, dlogdet generated
    up = torch.randn(B, d, generator=g).to(dev) * s
    dx2, _, dp2, _ = tp.backward(z, c, flat, up, torch.full((B,), -s, device=dev), mode=mode)
    dx3, _, dp3, _ = tp.backward(z, c, flat, up, None, mode=mode, nll_scale=s)
    assert float((dx3 - dx2).abs().max()) <= tol(dx2) and float((dp3 - dp2).abs().max()) <= tol(dp2)


@pytest.mark.parametrize("wl", [dict(d=43, ci=[67, 33, 16, 8], nb=3, mode="tf32"), dict(d=8, ci=[128, 64, 32, 16], nb=2, mode="tf32"),
                                dict(d=20, ci=[68, 34, 17, 17], nb=2, mode="fp32"),
                                dict(d=20, ci=[68, 34, 17, 17], nb=3, mode="fp32", hh="fixed"),
                                dict(d=43, ci=[67, 33, 16, 8], nb=3, mode="fp32", hh="trainable")],
                         ids=["d43_tf32", "gas_tf32", "lens_fp32", "lens_fp32_householder_fixed", "d43_fp32_householder_trainable"])
def test_fused_train_step_matches_the_reference_surface_step(dev, wl):
    """Steps without noise against module forward + nll_loss + backward + clamp_ + torch.optim.Adam.  After ONE step the
    parameters agree to 2e-6 * max|p| in every mode (same kernels, same inputs; only Adam's fp32 association differs).  After
    three steps the bound is 1e-5 in fp32 and 2e-2 in tf32 (measured 1e-3 .. 5e-3): a last-bit difference of a weight can flip its 10-bit tf32 rounding,
    which changes that product by 1e-3 relative and Adam's normalised update m/sqrt(v) carries it into the next step."""
    import hint_b200
    from hint_b200 import HintFlow, nll_loss, FusedClampAdam, FusedTrainStep
    old = hint_b200.get_precision()
    hint_b200.set_precision(wl["mode"])
    try:
        torch.manual_seed(0)
        ma = HintFlow(wl["d"], wl["nb"], wl["ci"], householder=wl.get("hh")).to(dev).init_like_reference_scripts(0.05)
        mb = HintFlow(wl["d"], wl["nb"], wl["ci"], householder=wl.get("hh")).to(dev)
        mb.load_state_dict(ma.state_dict())
        if wl.get("hh") == "trainable":      # reflections of unit scale (0.05 * randn would still be a valid reflection: only the direction counts)
            with torch.no_grad():
                for pa, pb in zip(ma.perms, mb.perms):
                    pa.Vs.normal_()
                    pb.Vs.copy_(pa.Vs)
        B = 1000
        x = torch.randn(B, wl["d"], device=dev)
        opt_a = FusedClampAdam([p for p in ma.parameters() if p.requires_grad], grad_clamp=5.0, **ADAM)
        tr = FusedTrainStep(ma, opt_a, noise=0.0)
        opt_b = torch.optim.Adam([p for p in mb.parameters() if p.requires_grad], **ADAM)
        n0 = hint_b200._lib.load().hint_launch_count()
        for it in range(3):
            la = tr.step(x)
            opt_b.zero_grad()
            z, J = mb(x)
            lb = nll_loss(z, J)
            lb.backward()
            for p in mb.parameters():
                if p.grad is not None:
                    p.grad.clamp_(-5.0, 5.0)
            opt_b.step()
            assert abs(float(la[0]) - float(lb.detach())) <= 1e-4 * max(1.0, abs(float(lb.detach())))
            tol = 2e-6 if it == 0 else (1e-5 if wl["mode"] == "fp32" else 2e-2)
            if wl.get("hh"):
                # with the mixing between the blocks the two paths differ in the last bits of dz (kernel vs autograd order of the
                # same FP32 products) and Adam's normalised update carries that into the parameters; trainable reflections add the
                # gradient through a product of 43 reflections (kernel vs autograd: 1e-4 relative)
                tol *= 10 if wl["hh"] == "fixed" else 100
            with torch.no_grad():
                for pa, pb in zip(ma.parameters(), mb.parameters()):
                    assert float((pa - pb).abs().max()) <= tol * float(pb.abs().max()), (it, float((pa - pb).abs().max()))
        assert hint_b200._lib.load().hint_launch_count() > n0
        # and the loss goes down
        l0 = float(tr.step(x)[0])
        for _ in range(20):
            l1 = float(tr.step(x)[0])
        assert l1 < l0
    finally:
        hint_b200.set_precision(old)


@pytest.mark.parametrize("d,n", [(1, 1), (6, 6), (43, 43), (100, 100), (20, 7), (128, 128)])
def test_householder_kernels_match_the_published_definition(dev, d, n):
    """W = prod (I - 2 v v^T / |v|^2) built, applied and differentiated by the library's kernels against the
    fp64 definition + torch autograd (the same expressions as FrEIA/modules/orthogonal.py).  Bounds: W and x W to 1e-5 relative,
    orthogonality |W W^T - I| <= 2e-5, dVs / dx to 1e-4 relative L2 (fp32 products of up to 128 reflections)."""
    from hint_b200.householder import HouseholderMix, householder_matrix, householder_apply
    g = torch.Generator().manual_seed(100 + d)
    Vs = torch.randn(n, d, generator=g)
    x = torch.randn(777, d, generator=g)

    def matrix64(V):
        W = torch.eye(d, dtype=torch.float64)
        for v in V:
            W = W - 2.0 * torch.outer(W @ v, v) / torch.dot(v, v)
        return W
    V64 = Vs.double().requires_grad_(True)
    x64 = x.double().requires_grad_(True)
    W64 = matrix64(V64)
    W = householder_matrix(Vs.to(dev))
    assert float((W.cpu().double() - W64.detach()).abs().max()) < 1e-5
    assert float((W @ W.t() - torch.eye(d, device=dev)).abs().max()) < 2e-5
    rel = lambda a, b: float(torch.linalg.norm(a.detach().cpu().double() - b.detach()) / max(1e-30, float(torch.linalg.norm(b.detach()))))
    for rev in (False, True):
        y64 = x64 @ (W64.t() if rev else W64)
        gy = torch.randn(777, d, generator=g)
        gV, gx = torch.autograd.grad(y64, (V64, x64), gy.double(), retain_graph=True)
        Vg = Vs.to(dev).requires_grad_(True)
        xg = x.to(dev).requires_grad_(True)
        y = HouseholderMix.apply(xg, Vg, None, rev)
        y.backward(gy.to(dev))
        assert rel(y, y64) < 1e-5 and rel(xg.grad, gx) < 1e-5
        if d == 1:      # W = -1 whatever v is: the true gradient is exactly 0
            assert float(Vg.grad.abs().max()) < 1e-4
        else:
            assert rel(Vg.grad, gV) < 1e-4, (rev, rel(Vg.grad, gV))
        # the inverse direction undoes the mixing
        back = householder_apply(y.detach(), W, transpose=not rev)
        assert float((back - x.to(dev)).abs().max()) < 1e-4


@pytest.mark.parametrize("d", [3, 8, 43, 48, 49, 100])
def test_householder_apply_ragged_unaligned_and_row_isolation(dev, d):
    """`hint_householder_apply` (error-compensated 3 x TF32 MMAs): fp32-grade against fp64 for row counts around the 16/32-row
    ranges, for operands that are not 16-byte aligned (a view starting at row 1), in both directions; a non-finite row stays in its row."""
    from hint_b200.householder import householder_matrix, householder_apply
    g = torch.Generator().manual_seed(7 + d)
    W = householder_matrix(torch.randn(d, d, generator=g).to(dev))
    base = torch.randn(1001, d, generator=g).to(dev)
    for B in (1, 15, 16, 17, 31, 33, 1000):
        for off in (0, 1):
            x = base[off:off + B]
            for tr in (False, True):
                y = householder_apply(x, W, transpose=tr)
                ref = x.double() @ (W.double().t() if tr else W.double())
                assert y.shape == (B, d)
                assert float((y.double() - ref).abs().max()) <= 1e-5 * max(1.0, float(ref.abs().max())), (B, off, tr)
    x = base[:100].clone()
    x[40, d // 2] = float("inf")
    y = householder_apply(x, W)
    ok = torch.ones(100, dtype=torch.bool, device=dev); ok[40] = False
    assert bool(torch.isfinite(y[ok]).all())
    assert float((y[ok].double() - x[ok].double() @ W.double()).abs().max()) <= 1e-5 * 8


def test_freia_householder_perm_uses_the_kernels_on_cuda(dev):
    """The shim's HouseholderPerm (fixed and trainable) gives the same numbers on CUDA (library kernels) as on the CPU (plain
    PyTorch definition), forward, reverse and gradients, and launches library kernels."""
    import hint_b200
    from FrEIA.modules import HouseholderPerm
    torch.manual_seed(0)
    for fixed in (True, False):
        P = HouseholderPerm([(43,)], n_reflections=43, fixed=fixed)
        Pg = HouseholderPerm([(43,)], n_reflections=43, fixed=fixed)
        Pg.load_state_dict(P.state_dict())
        Pg = Pg.to(dev)
        x = torch.randn(300, 43)
        xc, xg = x.clone().requires_grad_(True), x.to(dev).requires_grad_(True)
        n0 = hint_b200._lib.load().hint_launch_count()
        yc, yg = P([xc])[0], Pg([xg])[0]
        assert hint_b200._lib.load().hint_launch_count() > n0
        assert float((yg.cpu() - yc).abs().max()) < 1e-4
        assert float((Pg([yg.detach()], rev=True)[0].cpu() - x).abs().max()) < 1e-4
        w = torch.randn(300, 43)
        (yc * w).sum().backward()
        (yg * w.to(dev)).sum().backward()
        assert float((xg.grad.cpu() - xc.grad).abs().max()) < 1e-4
        if not fixed:
            assert float(torch.linalg.norm(Pg.Vs.grad.cpu() - P.Vs.grad) / torch.linalg.norm(P.Vs.grad)) < 1e-3


def test_graphed_flow_replays_the_same_transport(dev):
    """CUDA-graph replay of forward and inverse at a fixed small batch: bit-identical to the eager calls, follows weight updates,
    and needs fewer host-side launches (the library's kernels are capture-safe)."""
    import hint_b200
    from hint_b200 import HintFlow, GraphedFlow
    old = hint_b200.get_precision()
    hint_b200.set_precision("tf32")
    try:
        torch.manual_seed(0)
        for d, ci, dc in ((20, [68, 34, 17, 17], 0), (8, [128, 64, 32, 16], 0), (20, [68, 34, 17, 17], 2)):
            model = HintFlow(d, 3, ci, dims_c=[(dc,)] if dc else []).to(dev).init_like_reference_scripts(0.05)
            B = 300
            x = torch.randn(B, d, device=dev)
            c = torch.randn(B, dc, device=dev) if dc else None
            gf, gi = GraphedFlow(model, B), GraphedFlow(model, B, rev=True)
            with torch.no_grad():
                z, J = model(x, c)
                zg, Jg = gf(x, c)
                assert torch.equal(z, zg) and torch.equal(J, Jg)
                xr, Jr = gi(z, c)
                xe, Je = model(z, c, rev=True)
                assert torch.equal(xr, xe) and torch.equal(Jr, Je)
                for p in model.parameters():
                    p.mul_(1.1)                      # the graph reads the parameters at replay time
                z2, _ = model(x, c)
                zg2, _ = gf(x, c)
                assert torch.equal(z2, zg2) and not torch.equal(z2, z)
            with pytest.raises(ValueError):
                gf(torch.randn(B + 1, d, device=dev))
    finally:
        hint_b200.set_precision(old)


@pytest.mark.parametrize("n,d", [(4000, 20), (1000, 100), (65, 3), (1, 4), (333, 33)])
def test_fused_multi_mmd_matches_the_scripts_estimator(dev, n, d):
    """hint_multi_mmd against rejection_sampling.py:56-73 restated in float64 (same kernels, same mean over the n x n matrix)."""
    import hint_b200
    g = torch.Generator().manual_seed(n + d)
    x = torch.randn(n, d, generator=g)
    y = 0.3 + 1.2 * torch.randn(n, d, generator=g)
    we = [(0.5, 1), (0.2, 1), (0.2, 0.5)]

    def ref(a, b, we):
        d2 = lambda p, q: torch.cdist(p.double(), q.double()).pow(2)
        k = lambda dd: sum(C ** e * ((C + dd) / e) ** -e for C, e in we)
        return float((k(d2(a, a)) + k(d2(b, b)) - 2 * k(d2(a, b))).mean())
    n0 = hint_b200._lib.load().hint_launch_count()
    got = float(hint_b200.multi_mmd(x.to(dev), y.to(dev), we))
    assert hint_b200._lib.load().hint_launch_count() == n0 + 2
    want = ref(x, y, we)
    assert abs(got - want) <= 2e-5 * max(1e-2, abs(want)), (got, want)
    # identical sets: exactly the estimator's zero (up to fp32 rounding of the kernel values)
    assert abs(float(hint_b200.multi_mmd(x.to(dev), x.to(dev), we))) < 1e-6
    # generic exponents go through powf
    we2 = [(1, 0.5), (0.2, 0.8), (0.2, 0.4)]
    got2, want2 = float(hint_b200.multi_mmd(x.to(dev), y.to(dev), we2)), ref(x, y, we2)
    assert abs(got2 - want2) <= 5e-5 * max(1e-2, abs(want2)), (got2, want2)
    assert float(hint_b200.multi_mmd(x.to(dev), y.to(dev), we)) == got       # deterministic


def test_householder_mixing_at_the_benchmark_size_preserves_norms_and_inverts(dev):
    """Size-independent properties at BASELINE's full batch (2^20 rows, d = 43): the mixing is an isometry (row norms kept to
    fp32 rounding), x W W^T = x, and it is linear ((x + y) W = x W + y W)."""
    from hint_b200.householder import householder_matrix, householder_apply
    torch.manual_seed(0)
    B, d = 1 << 20, 43
    W = householder_matrix(torch.randn(d, d, device=dev))
    x = torch.randn(B, d, device=dev)
    y = householder_apply(x, W)
    nx, ny = x.double().norm(dim=1), y.double().norm(dim=1)
    assert float(((ny - nx).abs() / nx).max()) < 5e-6
    back = householder_apply(y, W, transpose=True)
    assert float((back - x).abs().max()) < 2e-5
    x2 = torch.randn(B, d, device=dev)
    lin = householder_apply(x + x2, W) - y - householder_apply(x2, W)
    assert float(lin.abs().max()) < 2e-5


def test_fused_coupling_round_trip_at_a_large_batch(dev):
    """2^20 samples through the lens y -> x coupling and back: the inverse undoes the forward pass, the log-dets cancel, and the
    rows are independent (a permuted batch gives the permuted result bit for bit)."""
    from FrEIA.modules import ExternalAffineCoupling, F_fully_connected
    torch.manual_seed(1)
    m = ExternalAffineCoupling([(20,)], dims_c=[(2,)], F_class=F_fully_connected, F_args={"internal_size": 68}).to(dev)
    for p in m.parameters():
        p.data = 0.05 * torch.randn_like(p)
    B = 1 << 20
    x, c = torch.randn(B, 20, device=dev), torch.randn(B, 2, device=dev)
    with torch.no_grad():
        y = m([x], [c])[0]
        j = m.jacobian(None)
        back = m([y], [c], rev=True)[0]
        jb = m.jacobian(None)
        perm = torch.randperm(B, device=dev)
        yp = m([x[perm]], [c[perm]])[0]
    assert float((back - x).abs().max()) < 1e-4
    assert float((j + jb).abs().max()) < 1e-5
    assert torch.equal(yp, y[perm])
