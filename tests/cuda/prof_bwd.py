"""ncu driver: forward + backward of one d=43 hint_8-width block in fp32 mode."""
import sys
import torch
from hint_b200 import HierarchicalAffineCouplingBlock
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 64 * 8
dev = torch.device("cuda:0")
torch.manual_seed(0)
blk = HierarchicalAffineCouplingBlock([(43,)], c_internal=[67, 33, 16, 8]).to(dev)
x = torch.randn(B, 43, device=dev)
with torch.no_grad():
    for _ in range(3):
        z, J = blk.plan.forward(x, None, blk.flat.detach(), mode="fp32")
        out = blk.plan.backward(z, None, blk.flat.detach(), z / B, torch.full((B,), -1.0 / B, device=dev))
torch.cuda.synchronize()
print("ok")
