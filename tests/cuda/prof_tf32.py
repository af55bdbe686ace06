"""Small driver for ncu: a few forward launches of one d=43 hint_8-width block in a given mode."""
import sys
import torch
from hint_b200 import HierarchicalAffineCouplingBlock

mode = sys.argv[1] if len(sys.argv) > 1 else "tf32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 128 * 8
dev = torch.device("cuda:0")
torch.manual_seed(0)
blk = HierarchicalAffineCouplingBlock([(43,)], c_internal=[67, 33, 16, 8]).to(dev)
x = torch.randn(B, 43, device=dev)
with torch.no_grad():
    for _ in range(4):
        z, J = blk.plan.forward(x, None, blk.flat.detach(), mode=mode)
torch.cuda.synchronize()
print("ok", float(z.abs().max()))
