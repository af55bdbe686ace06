"""GPU parity tests of the tensor-core modes through the module / C ABI: "tf32" and "tf32x3" (warp-MMA fused-tree kernels:
forward, inverse, backward) and "tf32_tc3" / its alias "tf32_tcgen05" (the tcgen05 / TMEM kernel: transport programs + backward).

Stated bound for single-pass TF32 (10-bit mantissa operands, fp32 accumulate, weights and hidden activations rounded to
nearest, x columns truncated by the tensor core): max-norm error relative to max(1, max|ref|) of
    z, x-reconstruction <= 5e-3,   log-det <= 5e-3
against the fp64 reference on the golden fixtures (Kaiming-scale weights; measured 2e-4 .. 4e-3), and <= 2e-5 on the
reference scripts' own init scale (0.005*randn).  Also checked against the CPU interpreter of the same TF32 program."""
import numpy as np
import pytest
import torch

from conftest import plan_kwargs
from oracle import hint_oracle as O

pytestmark = pytest.mark.gpu

TF32_TOL = 5e-3


def _err(a, ref):
    a = a.detach().double().cpu().numpy()
    return float(np.abs(a - ref).max() / max(1.0, np.abs(ref).max()))


def _split_c(c, dims_c, dev):
    if c is None:
        return []
    out, o = [], 0
    for t in dims_c:
        out.append(torch.from_numpy(c[:, o:o + t[0]].copy()).to(dev))
        o += t[0]
    return out


@pytest.mark.parametrize("mode", ["tf32", "tf32_chain", "tf32_mma", "tf32_tcgen05", "tf32_tc3"])
def test_golden_forward_inverse_tf32(golden, mode):
    import hint_b200
    from hint_b200 import HierarchicalAffineCouplingBlock
    dev = torch.device("cuda:0")
    meta = golden["meta"]
    blk = HierarchicalAffineCouplingBlock([(meta["d"],)], dims_c=[tuple(t) for t in meta["dims_c"]], **meta["kwargs"]).to(dev)
    with torch.no_grad():
        blk.flat.copy_(torch.from_numpy(golden["params"]))
    x = torch.from_numpy(golden["x"]).to(dev)
    cs = _split_c(golden.get("c"), meta["dims_c"], dev)
    cc = torch.cat(cs, dim=1) if cs else None
    try:
        with torch.no_grad():
            z, J = blk.plan.forward(x, cc, blk.flat.detach(), rev=False, mode=mode)
    except NotImplementedError as e:
        assert "envelope" in str(e)
        pytest.skip(str(e))
    tol = 2e-5 if meta["init"] == "randn0.005" else TF32_TOL
    assert _err(z, golden["z64"]) < tol and _err(J, golden["J64"]) < tol
    with torch.no_grad():
        xi, Ji = blk.plan.forward(x, cc, blk.flat.detach(), rev=True, mode=mode)
        assert _err(xi, golden["xinv64"]) < tol and _err(Ji, golden["Jinv64"]) < tol
        xr, Jr = blk.plan.forward(z, cc, blk.flat.detach(), rev=True, mode=mode)
    assert _err(xr, golden["x"].astype(np.float64)) < tol * max(1.0, float(np.abs(golden["z64"]).max()))


CONFIGS = [
    ("d43_hint_8", 43, 0, [67, 33, 16, 8], -1, 3000),
    ("miniboone_hint_8", 42, 0, [67, 33, 16, 8], -1, 1000),
    ("lens_hint_8_full", 20, 0, [68, 34, 17, 17], -1, 2500),
    ("lens_concat_cond", 20, 2, [68, 34, 17, 17], -1, 1000),
    ("gas_like", 8, 0, [64, 32, 16, 8], -1, 4097),
    # the BASELINE configs' REAL widths: wide single nodes, second M tile, ragged K
    ("power_hint_8", 6, 0, [140, 70, 35, 17], -1, 1660),
    ("gas_hint_8", 8, 0, [128, 64, 32, 16], -1, 853),
    ("miniboone_hint_4", 42, 0, [102, 51, 25, 12], -1, 300),
    ("plus_hint_4_3", 100, 0, [314, 157, 78, 39], 3, 500),
    ("plus_hint_4_full", 100, 0, [263, 131, 65, 32, 32], -1, 300),
    ("plus_cond_recursive_4", 100, 4, [267, 133, 66], -1, 300),
    # the hint_4 UCI variants: single nodes of h = 200 / 184 (tcgen05 transport programs up to h = 232; backward: interpreter)
    ("power_hint_4", 6, 0, [200, 100, 50, 25], -1, 700),
    ("gas_hint_4", 8, 0, [184, 92, 46, 23], -1, 700),
]


def _report(tag, **errs):
    """Measured errors go to stdout (pytest -s / the captured log) and to gpurun_out/tf32_errors.txt on the GPU box."""
    line = tag + "  " + "  ".join(f"{k} {v:.2e}" for k, v in errs.items())
    print(line)
    try:
        import os
        from conftest import ROOT
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "tf32_errors.txt"), "a") as f:
            f.write(line + "\n")
    except OSError:
        pass


@pytest.mark.parametrize("mode", ["tf32", "tf32_chain", "tf32_mma", "tf32_tcgen05", "tf32_tc3"])
@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: c[0])
def test_reference_configs_tf32(cfg, mode):
    name, d, dc, ci, ms, B = cfg
    from hint_b200 import HierarchicalAffineCouplingBlock
    dev = torch.device("cuda:0")
    torch.manual_seed(4321)
    blk = HierarchicalAffineCouplingBlock([(d,)], dims_c=[(dc,)] if dc else [], c_internal=list(ci), max_splits=ms)
    flat64 = blk.flat.detach().double().clone()
    blk = blk.to(dev)
    x = torch.randn(B, d)
    c = torch.randn(B, dc) if dc else None
    plan = O.build_plan(d, dc, ci, ms)
    z_ref, J_ref = O.forward_fast(plan, flat64, x.double(), None if c is None else c.double())
    cg = c.to(dev) if dc else None
    with torch.no_grad():
        try:
            z, J = blk.plan.forward(x.to(dev), cg, blk.flat.detach(), mode=mode)
        except NotImplementedError as e:
            assert "envelope" in str(e)
            pytest.skip(str(e))
        z32, J32 = blk.plan.forward(x.to(dev), cg, blk.flat.detach(), mode="fp32")
        xr, Jr = blk.plan.forward(z, cg, blk.flat.detach(), rev=True, mode=mode)
    _report(f"fwd {name:22s} {mode:13s} kaiming", z=_err(z, z_ref.numpy()), J=_err(J, J_ref.numpy()), xrec=_err(xr, x.double().numpy()))
    assert _err(z, z_ref.numpy()) < TF32_TOL and _err(J, J_ref.numpy()) < TF32_TOL
    assert _err(z32, z_ref.numpy()) < 1e-5
    assert _err(xr, x.double().numpy()) < TF32_TOL * max(1.0, float(z_ref.abs().max()))
    assert _err(J + Jr, np.zeros(B)) < TF32_TOL * max(1.0, float(J_ref.abs().max()))


def _l2(a, ref):
    a = a.detach().double().cpu().numpy()
    return float(np.linalg.norm(a - ref) / max(1e-30, np.linalg.norm(ref)))


def test_golden_3xtf32_full_parity(golden):
    """3xTF32 (warp-MMA kernels): same bounds as the FP32 mode - z / x-reconstruction / log-det <= 1e-5 relative to
    max(1, |ref|_inf) against the fp64 reference, gradients <= 2e-4 (robust metric of test_gpu_parity.py not needed on
    these fixtures: measured 2e-7 .. 2e-6)."""
    from hint_b200 import HierarchicalAffineCouplingBlock
    dev = torch.device("cuda:0")
    meta = golden["meta"]
    blk = HierarchicalAffineCouplingBlock([(meta["d"],)], dims_c=[tuple(t) for t in meta["dims_c"]], **meta["kwargs"]).to(dev)
    with torch.no_grad():
        blk.flat.copy_(torch.from_numpy(golden["params"]))
    x = torch.from_numpy(golden["x"]).to(dev)
    B = x.shape[0]
    cs = _split_c(golden.get("c"), meta["dims_c"], dev)
    cc = torch.cat(cs, dim=1) if cs else None
    flat = blk.flat.detach()
    with torch.no_grad():
        z, J = blk.plan.forward(x, cc, flat, mode="tf32x3")
        xi, Ji = blk.plan.forward(x, cc, flat, rev=True, mode="tf32x3")
        z64 = torch.from_numpy(golden["z64"])
        dx, dc, dflat, xrec = blk.plan.backward(z64.float().to(dev), cc, flat, (z64 / B).float().to(dev),
                                                torch.full((B,), -1.0 / B, device=dev), mode="tf32x3", want_xrec=True)
    assert _err(z, golden["z64"]) < 1e-5 and _err(J, golden["J64"]) < 1e-5
    assert _err(xi, golden["xinv64"]) < 1e-5 and _err(Ji, golden["Jinv64"]) < 1e-5
    # inverting amplifies output error by the flow's expansion: bound scaled by |z| as in test_gpu_parity.py
    assert _err(xrec, golden["x"].astype(np.float64)) < 1e-4 * max(1.0, float(np.abs(golden["z64"]).max()))
    # relative L2 error (measured 2e-7 .. 4e-5: ReLU kinks can flip for an isolated sample, see test_gpu_parity.py)
    assert _l2(dx, golden["dx64"]) < 1e-4
    assert _l2(dflat, golden["dparams64"]) < 1e-4
    if cc is not None:
        assert _l2(dc, golden["dc64"]) < 1e-4


@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: c[0])
@pytest.mark.parametrize("mode,tol", [("tf32", 2e-2), ("tf32_mma", 2e-2), ("tf32_tc3", 2e-2), ("tf32x3", 2e-4)])
def test_backward_tensor_core_modes(cfg, mode, tol):
    """Backward of every tensor-core kernel family (tf32 = the default dispatch: register-chained / tcgen05 training kernel /
    interpreter) against the fp64 oracle on the reference configs' real widths (relative L2 error of dx, dc and the flat
    parameter gradient; single-pass TF32 bound 2e-2 on Kaiming-scale weights: ReLU kinks make a few samples flip, see
    test_gpu_parity.py)."""
    name, d, dc, ci, ms, B = cfg
    from hint_b200 import HierarchicalAffineCouplingBlock
    dev = torch.device("cuda:0")
    torch.manual_seed(99)
    blk = HierarchicalAffineCouplingBlock([(d,)], dims_c=[(dc,)] if dc else [], c_internal=list(ci), max_splits=ms)
    if not blk.plan.mode_supported(mode):
        pytest.skip(f"{name} is outside the {mode} envelope")
    with torch.no_grad():
        blk.flat.mul_(0.7)
    flat64 = blk.flat.detach().double().clone()
    blk = blk.to(dev)
    x = torch.randn(B, d)
    c = torch.randn(B, dc) if dc else None
    plan = O.build_plan(d, dc, ci, ms)
    c64 = None if c is None else c.double()
    z_ref, J_ref = O.forward_fast(plan, flat64, x.double(), c64)
    dz = torch.randn(B, d, dtype=torch.float64) / B
    dJ = torch.randn(B, dtype=torch.float64) / B
    _, dx_ref, dc_ref, dflat_ref = O.backward_from_output(plan, flat64, z_ref, c64, dz, dJ)
    cg = c.to(dev) if dc else None
    with torch.no_grad():
        dx, dcc, dflat, xrec = blk.plan.backward(z_ref.float().to(dev), cg, blk.flat.detach(), dz.float().to(dev),
                                                 dJ.float().to(dev), mode=mode, want_xrec=True)
    _report(f"bwd {name:22s} {mode:13s} kaiming*0.7", dx=_l2(dx, dx_ref.numpy()), dparams=_l2(dflat, dflat_ref.numpy()),
            xrec=_l2(xrec, x.double().numpy()))
    assert _l2(dx, dx_ref.numpy()) < tol and _l2(dflat, dflat_ref.numpy()) < tol
    assert _l2(xrec, x.double().numpy()) < (1e-5 if mode == "tf32x3" else 5e-3)
    if dc:   # dc sums the input gradients of EVERY node's subnets (hint.py:76), so it collects the most TF32 rounding noise
        assert _l2(dcc, dc_ref.numpy()) < 1.5 * tol


def test_full_size_properties_chain():
    """BASELINE.json sizes (d=43 `hint_8` widths, batch 256k) through the register-chained kernels: size-independent
    properties instead of the CPU oracle - f^-1(f(x)) = x and J_rev(f(x)) = -J_fwd(x) within the TF32 bound, agreement with
    the FP32 CUDA-core kernels on the same inputs, bit-reproducibility, and exact linearity of the gradients in the upstream
    gradient (scaling by 2 commutes with tf32 rounding)."""
    dev = torch.device("cuda:0")
    from hint_b200 import HierarchicalAffineCouplingBlock
    torch.manual_seed(11)
    B, d = 262144, 43
    blk = HierarchicalAffineCouplingBlock([(d,)], c_internal=[67, 33, 16, 8]).to(dev)
    assert blk.plan.mode_supported("tf32_chain")
    flat = blk.flat.detach()
    x = torch.randn(B, d, device=dev)
    with torch.no_grad():
        z, J = blk.plan.forward(x, None, flat, mode="tf32_chain")
        z2, J2 = blk.plan.forward(x, None, flat, mode="tf32_chain")
        assert torch.equal(z, z2) and torch.equal(J, J2)
        z32, J32 = blk.plan.forward(x, None, flat, mode="fp32")
        scale = max(1.0, z32.abs().max().item())
        assert (z - z32).abs().max().item() < TF32_TOL * scale
        assert (J - J32).abs().max().item() < TF32_TOL * max(1.0, J32.abs().max().item())
        xr, Jr = blk.plan.forward(z, None, flat, rev=True, mode="tf32_chain")
        assert (xr - x).abs().max().item() < TF32_TOL * scale
        assert (J + Jr).abs().max().item() < TF32_TOL * max(1.0, J32.abs().max().item())
        dz = torch.randn(B, d, device=dev) / B
        dJ = torch.randn(B, device=dev) / B
        dx1, _, g1, xrec = blk.plan.backward(z, None, flat, dz, dJ, mode="tf32_chain", want_xrec=True)
        dx2, _, g2, _ = blk.plan.backward(z, None, flat, 2 * dz, 2 * dJ, mode="tf32_chain")
        dx3, _, g3, _ = blk.plan.backward(z, None, flat, dz, dJ, mode="tf32_chain")
        dxf, _, gf, _ = blk.plan.backward(z, None, flat, dz, dJ, mode="fp32")
    assert (xrec - x).abs().max().item() < TF32_TOL * scale
    assert torch.equal(g1, g3) and torch.equal(dx1, dx3)           # per-CTA partial buffers, fixed-order reduction
    assert (g2 - 2 * g1).abs().max().item() <= 1e-5 * g1.abs().max().item() + 1e-7
    assert (dx2 - 2 * dx1).abs().max().item() <= 1e-5 * dx1.abs().max().item() + 1e-9
    rel = lambda a, b: (torch.linalg.norm(a.double() - b.double()) / torch.linalg.norm(b.double())).item()
    assert rel(g1, gf) < 2e-2 and rel(dx1, dxf) < 2e-2


@pytest.mark.parametrize("seed", range(24))
def test_random_trees_chain_against_the_oracle(seed):
    """Random tree shapes through the C ABI in the register-chained mode (forward, inverse, backward) against the fp64 oracle."""
    rng = np.random.default_rng(7000 + seed)
    d = int(rng.integers(2, 48))
    dc = int(rng.choice([0, 0, 1, 3]))
    widths = [int(rng.integers(3, 73)) for _ in range(int(rng.integers(1, 5)))]
    ms = int(rng.choice([-1, -1, 0, 1, 2, 3]))
    mss = int(rng.choice([2, 2, 3, 4]))
    B = int(rng.integers(1, 700))
    from hint_b200 import HierarchicalAffineCouplingBlock
    dev = torch.device("cuda:0")
    torch.manual_seed(seed)
    blk = HierarchicalAffineCouplingBlock([(d,)], dims_c=[(dc,)] if dc else [], c_internal=widths, max_splits=ms, min_split_size=mss)
    if not blk.plan.mode_supported("tf32_chain"):
        pytest.skip("outside the chain envelope")
    with torch.no_grad():
        blk.flat.mul_(0.4)
    flat64 = blk.flat.detach().double().clone()
    blk = blk.to(dev)
    x = torch.randn(B, d)
    c = torch.randn(B, dc) if dc else None
    plan = O.build_plan(d, dc, widths, ms, mss)
    c64 = None if c is None else c.double()
    z_ref, J_ref = O.forward_fast(plan, flat64, x.double(), c64)
    dz = torch.randn(B, d, dtype=torch.float64) / B
    dJ = torch.randn(B, dtype=torch.float64) / B
    _, dx_ref, dc_ref, dflat_ref = O.backward_from_output(plan, flat64, z_ref, c64, dz, dJ)
    cg = c.to(dev) if dc else None
    with torch.no_grad():
        z, J = blk.plan.forward(x.to(dev), cg, blk.flat.detach(), mode="tf32_chain")
        xr, Jr = blk.plan.forward(z, cg, blk.flat.detach(), rev=True, mode="tf32_chain")
        dx, dcc, dflat, _ = blk.plan.backward(z_ref.float().to(dev), cg, blk.flat.detach(), dz.float().to(dev), dJ.float().to(dev),
                                              mode="tf32_chain")
    assert _err(z, z_ref.numpy()) < TF32_TOL and _err(J, J_ref.numpy()) < TF32_TOL
    assert _err(xr, x.double().numpy()) < TF32_TOL * max(1.0, float(z_ref.abs().max()))
    # gradients: TF32 rounding noise grows with the weight scale of wide subnets (ReLU kinks flip for isolated samples), so the
    # bound against the fp64 oracle is loose; the sharp check is agreement with the interpreter warp-MMA kernels, which round at
    # the same points through completely different code
    assert _l2(dx, dx_ref.numpy()) < 4e-2 and _l2(dflat, dflat_ref.numpy()) < 4e-2
    if dc:
        assert _l2(dcc, dc_ref.numpy()) < 4e-2
    if blk.plan.mode_supported("tf32_mma"):
        with torch.no_grad():
            dx2, dc2, dflat2, _ = blk.plan.backward(z_ref.float().to(dev), cg, blk.flat.detach(), dz.float().to(dev), dJ.float().to(dev),
                                                    mode="tf32_mma")
        assert _l2(dx, dx2.double().cpu().numpy()) < 2e-3 and _l2(dflat, dflat2.double().cpu().numpy()) < 2e-3


TRAINED = [
    ("power_hint_8", 6, [140, 70, 35, 17], -1),
    ("gas_hint_8", 8, [128, 64, 32, 16], -1),
    ("miniboone_hint_4", 42, [102, 51, 25, 12], -1),
    ("d43_hint_8", 43, [67, 33, 16, 8], -1),
    ("lens_hint_8_full", 20, [68, 34, 17, 17], -1),
    ("plus_hint_4_3", 100, [314, 157, 78, 39], 3),
]


@pytest.mark.parametrize("cfg", TRAINED, ids=lambda c: c[0])
def test_briefly_trained_weights_meet_the_stated_tf32_bound(cfg):
    """the goldens are init-scale only, so this fixture TRAINS a 2-block flow with the reference recipe
    (train_unconditional.py:121-144: noise 0.01, NLL, clamp 5, Adam lr 0.01 / betas (0.9, 0.95) / eps 1e-4 / wd 1.86e-5, init
    0.005*randn; 300 steps on a fixed batch of 2048 standardised GMM samples: max|w| reaches 1.3 .. 2.5) in the FP32 mode on the
    GPU, then checks every block in the benchmarked mode `tf32` against the fp64 oracle on the activations the flow actually sees.

    Stated bound of the single-pass TF32 mode on trained weights: BASELINE.md section 5's own metric, the relative (L2) error of
    z, <= 2e-3; and in the stricter max-norm metric of these tests, relative to max(1, |ref|_inf):
        z <= 5e-3 (measured 3e-4 .. 4.6e-3), log-det <= 1e-3 (measured 8e-5 .. 9.7e-4), x-reconstruction <= 1e-4,
        gradients (relative L2) dx <= 3e-2, dparams <= 1e-2 (measured 3.6e-3 .. 2.2e-2 / 9e-4 .. 8e-3).
    BASELINE.md section 5 suggested z <= 2e-3 from an emulation on weights with max|w| <= 1.3; this fixture trains harder and the
    error grows with the weight scale, so the bound is restated from measurement.  That the error is the arithmetic's and not
    the kernels' is checked against a yardstick: the same block evaluated by the oracle in float32 ON THE GPU with
    torch.backends.cuda.matmul.allow_tf32 = True - i.e. the reference module as PyTorch itself would run it in TF32 (cuBLAS
    truncates operands, these kernels round to nearest) - whose error against fp64 bounds ours: ours <= 1.5 x PyTorch-TF32's
    (+ 2e-4 floor)."""
    import hint_b200
    from hint_b200 import HintFlow, FusedClampAdam, FusedTrainStep
    name, d, ci, ms = cfg
    dev = torch.device("cuda:0")
    old = hint_b200.get_precision()
    hint_b200.set_precision("fp32")
    try:
        torch.manual_seed(5)
        model = HintFlow(d, 2, ci, max_splits=ms).to(dev).init_like_reference_scripts(0.005)
        g = torch.Generator().manual_seed(17)
        means, stds = 3.0 * torch.randn(8, d, generator=g), 0.3 + torch.rand(8, d, generator=g)
        comp = torch.randint(0, 8, (2048,), generator=g)
        x = means[comp] + stds[comp] * torch.randn(2048, d, generator=g)
        x = ((x - x.mean(0)) / x.std(0)).to(dev)
        opt = FusedClampAdam(list(model.parameters()), grad_clamp=5.0, lr=0.01, betas=(0.9, 0.95), eps=1e-4, weight_decay=1.86e-5)
        tr = FusedTrainStep(model, opt, noise=0.01, seed=3)
        l0 = float(tr.step(x)[0])
        for _ in range(300):
            l1 = float(tr.step(x)[0])
        assert l1 < l0 - 0.5, (l0, l1)     # it did train
    finally:
        hint_b200.set_precision(old)
    plan = O.build_plan(d, 0, ci, ms)
    h = x[:512]
    B = h.shape[0]
    old_tf32 = torch.backends.cuda.matmul.allow_tf32
    for bi, blk in enumerate(model.blocks):
        flat = blk.flat.detach()
        f64 = flat.double().cpu()
        z_ref, J_ref = O.forward_fast(plan, f64, h.double().cpu(), None)
        dz = z_ref / B
        dJ = torch.full((B,), -1.0 / B, dtype=torch.float64)
        _, dx_ref, _, dp_ref = O.backward_from_output(plan, f64, z_ref, None, dz, dJ)
        with torch.no_grad():
            z, J = blk.plan.forward(h, None, flat, mode="tf32")
            xr, _ = blk.plan.forward(z, None, flat, rev=True, mode="tf32")
            dx, _, dflat, _ = blk.plan.backward(z_ref.float().to(dev), None, flat, dz.float().to(dev), dJ.float().to(dev), mode="tf32")
            torch.backends.cuda.matmul.allow_tf32 = True     # yardstick: PyTorch's own TF32 on the same block
            try:
                zt, Jt = O.forward_fast(plan, flat, h, None)
                _, dxt, _, dpt = O.backward_from_output(plan, flat, z_ref.float().to(dev), None, dz.float().to(dev), dJ.float().to(dev))
            finally:
                torch.backends.cuda.matmul.allow_tf32 = old_tf32
        ez, eJ, ex = _err(z, z_ref.numpy()), _err(J, J_ref.numpy()), _err(xr, h.double().cpu().numpy())
        rz = _l2(z, z_ref.numpy())       # BASELINE.md section 5's own metric: relative (L2) error of z
        gdx, gdp = _l2(dx, dx_ref.numpy()), _l2(dflat, dp_ref.numpy())
        tz, tJ = _err(zt, z_ref.numpy()), _err(Jt, J_ref.numpy())
        tdx, tdp = _l2(dxt, dx_ref.numpy()), _l2(dpt, dp_ref.numpy())
        _report(f"trained {name:18s} block {bi} max|w| {float(flat.abs().max()):.2f}", z=ez, z_relL2=rz, J=eJ, xrec=ex, dx=gdx, dparams=gdp,
                torch_tf32_z=tz, torch_tf32_J=tJ, torch_tf32_dx=tdx, torch_tf32_dparams=tdp)
        assert ez <= 5e-3 and eJ <= 1e-3
        assert rz <= 2e-3                # "single-pass TF32 <= 2e-3 relative on z" in its relative-error metric
        assert ex <= 1e-4 * max(1.0, float(z_ref.abs().max()))
        assert gdx <= 3e-2 and gdp <= 1e-2
        assert ez <= 1.5 * tz + 2e-4 and eJ <= 1.5 * tJ + 2e-4
        assert gdx <= 1.5 * tdx + 2e-3 and gdp <= 1.5 * tdp + 2e-3
        h = z_ref.float().to(dev)
