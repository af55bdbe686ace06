"""Forward + inverse launches per config and mode (for ncu captures): python run_fwd_once.py <cfg> <mode> <B>."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hint_b200.block import TreePlan
CFG = {"gas": (8, 0, [128, 64, 32, 16]), "power": (6, 0, [140, 70, 35, 17]), "d43": (43, 0, [67, 33, 16, 8]), "lens": (20, 0, [68, 34, 17, 17])}
name = sys.argv[1] if len(sys.argv) > 1 else "gas"
mode = sys.argv[2] if len(sys.argv) > 2 else "tf32"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 18
d, dc, ci = CFG[name]
dev = torch.device("cuda:0")
tp = TreePlan(d, dc, ci, 4.0, -1, 2, False)
flat = (0.05 * torch.randn(tp.n_params)).to(dev)
x = torch.randn(B, d, device=dev)
for _ in range(3):
    z, J = tp.forward(x, None, flat, False, mode=mode)
    tp.forward(z, None, flat, True, mode=mode)
torch.cuda.synchronize()
print("done", name, mode, B)
