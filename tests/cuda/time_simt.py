"""Device timing of one block's fp32 forward and backward (CUDA events)."""
import sys, os
import torch
from hint_b200 import HierarchicalAffineCouplingBlock
cfgs = {"d43": (43, 0, [67, 33, 16, 8], -1), "power": (6, 0, [140, 70, 35, 17], -1)}
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 19
dev = torch.device("cuda:0")
for name, (d, dc, ci, ms) in cfgs.items():
    torch.manual_seed(0)
    blk = HierarchicalAffineCouplingBlock([(d,)], c_internal=ci, max_splits=ms).to(dev)
    Bc = B if d < 100 else B // 8
    x = torch.randn(Bc, d, device=dev)
    flat = blk.flat.detach()
    with torch.no_grad():
        z, J = blk.plan.forward(x, None, flat, mode="fp32")
        dz = z / Bc; dJ = torch.full((Bc,), -1.0 / Bc, device=dev)
        def tm(fn, n=5):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n): fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        tf = tm(lambda: blk.plan.forward(x, None, flat, mode="fp32"))
        tb = tm(lambda: blk.plan.backward(z, None, flat, dz, dJ))
    F = blk.plan.flops_per_sample
    print(f"{name:7s} TM {blk.plan.tile_rows(0)}/{blk.plan.tile_rows(1)}: fwd {tf:8.3f} ms {F*Bc/tf/1e9:6.2f} TF/s | bwd {tb:8.3f} ms {2*F*Bc/tb/1e9:6.2f} TF/s (algorithmic)")
