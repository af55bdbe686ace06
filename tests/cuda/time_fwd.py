"""Quick device timing of one block's forward in each mode (CUDA events, inputs larger than L2)."""
import sys
import torch
from hint_b200 import HierarchicalAffineCouplingBlock

cfgs = {"d43": (43, 0, [67, 33, 16, 8], -1), "lens": (20, 0, [68, 34, 17, 17], -1), "gas64": (8, 0, [64, 32, 16, 8], -1)}
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
dev = torch.device("cuda:0")
for name, (d, dc, ci, ms) in cfgs.items():
    torch.manual_seed(0)
    blk = HierarchicalAffineCouplingBlock([(d,)], c_internal=ci, max_splits=ms).to(dev)
    x = torch.randn(B, d, device=dev)
    for mode in ("fp32", "tf32"):
        with torch.no_grad():
            for _ in range(3):
                z, J = blk.plan.forward(x, None, blk.flat.detach(), mode=mode)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                z, J = blk.plan.forward(x, None, blk.flat.detach(), mode=mode)
            e1.record()
            torch.cuda.synchronize()
        ms_ = e0.elapsed_time(e1) / 10
        print(f"{name:6s} {mode}: {ms_:8.3f} ms/block  {B / ms_ / 1e3:8.1f} Msamples/s/block  {blk.plan.flops_per_sample * B / ms_ / 1e9:7.2f} TFLOP/s")
