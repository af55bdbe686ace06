"""GPU check + device timing of the warp-MMA (tf32 / tf32x3) kernels against the fp32 CUDA-core kernels and the fp64 oracle.
    python tests/cuda/check_mma.py [B_timing]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hint_b200 import HierarchicalAffineCouplingBlock  # noqa: E402
from oracle import hint_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
CFGS = [("d43", 43, 0, [67, 33, 16, 8], -1, 3001), ("mini42", 42, 0, [67, 33, 16, 8], -1, 777),
        ("lens", 20, 0, [68, 34, 17, 17], -1, 1500), ("lens_c", 20, 2, [68, 34, 17, 17], -1, 1000),
        ("gas", 8, 0, [128, 64, 32, 16], -1, 2049), ("power", 6, 0, [140, 70, 35, 17], -1, 640),
        ("plus_c", 100, 4, [267, 133, 66], -1, 300), ("plus43", 100, 0, [314, 157, 78, 39], 3, 200)]


def rel(a, ref):
    a = a.detach().double().cpu()
    return float((a - ref).abs().max() / max(1e-30, float(ref.abs().max())))


for name, d, dc, ci, ms, B in CFGS:
    torch.manual_seed(7)
    blk = HierarchicalAffineCouplingBlock([(d,)], dims_c=[(dc,)] if dc else [], c_internal=list(ci), max_splits=ms)
    with torch.no_grad():
        blk.flat.mul_(0.7)
    flat64 = blk.flat.detach().double().clone()
    blk = blk.to(dev)
    x = torch.randn(B, d)
    c = torch.randn(B, dc) if dc else None
    plan = O.build_plan(d, dc, ci, ms)
    z_ref, J_ref = O.forward_fast(plan, flat64, x.double(), None if c is None else c.double())
    dz = torch.randn(B, d, dtype=torch.float64) / B
    dJ = torch.randn(B, dtype=torch.float64) / B
    _, dx_ref, dc_ref, dflat_ref = O.backward_from_output(plan, flat64, z_ref, None if c is None else c.double(), dz, dJ)
    xg, cg = x.to(dev), (c.to(dev) if dc else None)
    flat = blk.flat.detach()
    for mode in ("fp32", "tf32", "tf32x3"):
        try:
            with torch.no_grad():
                z, J = blk.plan.forward(xg, cg, flat, mode=mode)
                xi, Ji = blk.plan.forward(z, cg, flat, rev=True, mode=mode)
                dx, dcc, dflat, xrec = blk.plan.backward(z_ref.float().to(dev), cg, flat, dz.float().to(dev), dJ.float().to(dev),
                                                         mode=mode, want_xrec=True)
            torch.cuda.synchronize()
        except NotImplementedError as e:
            print(f"{name:7s} {mode:7s} unsupported: {e}")
            continue
        msg = (f"{name:7s} {mode:7s} z {rel(z, z_ref):.1e} J {rel(J, J_ref):.1e} inv {rel(xi, x.double()):.1e} J+Ji {float((J + Ji).abs().max()):.1e} "
               f"xrec {rel(xrec, x.double()):.1e} dx {rel(dx, dx_ref):.1e} dflat {rel(dflat, dflat_ref):.1e}")
        if dc:
            msg += f" dc {rel(dcc, dc_ref):.1e}"
        print(msg, flush=True)

# ---- timing, d43 hint_8 widths ----
Bt = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
for name, d, dc, ci, ms in [("d43", 43, 0, [67, 33, 16, 8], -1), ("gas", 8, 0, [128, 64, 32, 16], -1), ("lens", 20, 0, [68, 34, 17, 17], -1)]:
    torch.manual_seed(0)
    blk = HierarchicalAffineCouplingBlock([(d,)], c_internal=ci, max_splits=ms).to(dev)
    with torch.no_grad():
        blk.flat.copy_(0.05 * torch.randn_like(blk.flat))
    x = torch.randn(Bt, d, device=dev)
    flat = blk.flat.detach()
    F = blk.plan.flops_per_sample
    for mode in ("fp32", "tf32_tcgen05", "tf32", "tf32x3"):
        try:
            with torch.no_grad():
                z, J = blk.plan.forward(x, None, flat, mode=mode)
                dz = z / Bt
                dJ = torch.full((Bt,), -1.0 / Bt, device=dev)

                def tm(fn, n=5):
                    fn(); torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(n):
                        fn()
                    e1.record(); torch.cuda.synchronize()
                    return e0.elapsed_time(e1) / n
                tf = tm(lambda: blk.plan.forward(x, None, flat, mode=mode))
                ti = tm(lambda: blk.plan.forward(z, None, flat, rev=True, mode=mode))
                tb = tm(lambda: blk.plan.backward(z, None, flat, dz, dJ, mode=mode))
        except NotImplementedError as e:
            print(f"{name} {mode}: unsupported {e}")
            continue
        print(f"{name:5s} {mode:13s} B={Bt}: fwd {tf:7.3f} ms {F * Bt / tf / 1e9:7.2f} TF/s | inv {ti:7.3f} ms | bwd {tb:7.3f} ms {2 * F * Bt / tb / 1e9:7.2f} TF/s "
              f"(algorithmic) | fwd {Bt / tf / 1e3:6.1f} M samples/s/block", flush=True)
