"""Backward timing of the plus hint_4_3 block vs batch size (interpreter warp-MMA kernel): python time_bwd_plus.py"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hint_b200.block import TreePlan
dev = torch.device("cuda:0")
tp = TreePlan(100, 0, [314, 157, 78, 39], 4.0, 3, 2, False)
flat = (0.02 * torch.randn(tp.n_params)).to(dev)
print("tile rows fwd/bwd", tp.tile_rows(0), tp.tile_rows(1))
for B in (500, 2000, 10000, 20000, 40000, 160000):
    z = torch.randn(B, 100, device=dev); dz = torch.randn(B, 100, device=dev) / B; dJ = torch.full((B,), -1.0 / B, device=dev)
    for _ in range(2):
        tp.backward(z, None, flat, dz, dJ, mode="tf32")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        tp.backward(z, None, flat, dz, dJ, mode="tf32")
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"B={B:7d}: bwd {ms:8.3f} ms -> {2 * tp.flops_per_sample * B / ms / 1e9:6.1f} TFLOP/s algorithmic, {B / ms / 1e3:7.3f} M samples/s", flush=True)
