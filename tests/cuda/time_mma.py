"""Device timing of the warp-MMA kernels on one d=43 hint_8-width block.  python tests/cuda/time_mma.py [B] [modes...]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hint_b200 import HierarchicalAffineCouplingBlock
Bt = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
modes = sys.argv[2:] or ["tf32"]
dev = torch.device("cuda:0")
for name, d, dc, ci, ms in [("d43", 43, 0, [67, 33, 16, 8], -1)]:
    torch.manual_seed(0)
    blk = HierarchicalAffineCouplingBlock([(d,)], c_internal=ci, max_splits=ms).to(dev)
    with torch.no_grad():
        blk.flat.copy_(0.05 * torch.randn_like(blk.flat))
    x = torch.randn(Bt, d, device=dev)
    flat = blk.flat.detach()
    F = blk.plan.flops_per_sample
    for mode in modes:
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
            tb = tm(lambda: blk.plan.backward(z, None, flat, dz, dJ, mode=mode))
        print(f"{name:5s} {mode:13s} B={Bt}: fwd {tf:7.3f} ms {F * Bt / tf / 1e9:7.2f} TF/s | bwd {tb:7.3f} ms {2 * F * Bt / tb / 1e9:7.2f} TF/s", flush=True)
