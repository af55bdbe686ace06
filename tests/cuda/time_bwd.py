"""Developer aid: backward launch time per config and mode: python time_bwd.py <cfg> <B> <mode> [<mode> ...]."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hint_b200.block import TreePlan
CFG = {"plus43": (100, 0, [314, 157, 78, 39], 3), "plus4f": (100, 0, [263, 131, 65, 32, 32]), "gas": (8, 0, [128, 64, 32, 16]), "power": (6, 0, [140, 70, 35, 17]), "d43": (43, 0, [67, 33, 16, 8]), "lens": (20, 0, [68, 34, 17, 17])}
name = sys.argv[1]; B = int(sys.argv[2]); modes = sys.argv[3:] or ["tf32"]
d, dc, ci = CFG[name][:3]
ms = CFG[name][3] if len(CFG[name]) > 3 else -1
dev = torch.device("cuda:0")
tp = TreePlan(d, dc, ci, 4.0, ms, 2, False)
print('tile rows fwd/bwd', tp.tile_rows(0), tp.tile_rows(1))
flat = (0.05 * torch.randn(tp.n_params)).to(dev)
z = torch.randn(B, d, device=dev); dz = torch.randn(B, d, device=dev) / B; dJ = torch.full((B,), -1.0 / B, device=dev)
for mode in modes:
    for _ in range(3): tp.backward(z, None, flat, dz, dJ, mode=mode)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    n = 10; e0.record()
    for _ in range(n): tp.backward(z, None, flat, dz, dJ, mode=mode)
    e1.record(); torch.cuda.synchronize()
    print(f"{name} B={B} {mode}: backward {e0.elapsed_time(e1) / n:.3f} ms")
