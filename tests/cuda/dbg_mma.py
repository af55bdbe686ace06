"""Per-barrier cycle breakdown of the warp-MMA kernels (run with HINT_B200_MMA_DEBUG=1)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hint_b200 import HierarchicalAffineCouplingBlock
B = 148 * 2 * 64 * 6
dev = torch.device("cuda:0")
torch.manual_seed(0)
blk = HierarchicalAffineCouplingBlock([(43,)], c_internal=[67, 33, 16, 8]).to(dev)
x = torch.randn(B, 43, device=dev)
mode = sys.argv[1] if len(sys.argv) > 1 else "tf32_mma"
with torch.no_grad():
    for _ in range(2):
        z, J = blk.plan.forward(x, None, blk.flat.detach(), mode=mode)
        out = blk.plan.backward(z, None, blk.flat.detach(), z / B, torch.full((B,), -1.0 / B, device=dev), mode=mode)
torch.cuda.synchronize()
