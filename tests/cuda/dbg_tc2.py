"""Cycle breakdown of the TF32 v2 kernel (HINT_B200_TC_DEBUG=1), one block, B samples."""
import os, sys
os.environ["HINT_B200_TC_DEBUG"] = "1"
import torch
from hint_b200 import HierarchicalAffineCouplingBlock
cfgs = {"d43": (43, 0, [67, 33, 16, 8], -1), "lens": (20, 0, [68, 34, 17, 17], -1), "gas64": (8, 0, [64, 32, 16, 8], -1)}
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 8
dev = torch.device("cuda:0")
for name, (d, dc, ci, ms) in cfgs.items():
    torch.manual_seed(0)
    blk = HierarchicalAffineCouplingBlock([(d,)], c_internal=ci, max_splits=ms).to(dev)
    x = torch.randn(B, d, device=dev)
    with torch.no_grad():
        for rev in (False, True):
            for _ in range(2):
                print(name, "rev" if rev else "fwd", flush=True)
                z, J = blk.plan.forward(x, None, blk.flat.detach(), rev=rev, mode="tf32")
                torch.cuda.synchronize()
