"""Timing of the Householder mixing kernels: python time_hh.py"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hint_b200.householder import householder_apply, householder_matrix, _wgrad
dev = torch.device("cuda:0")
for d, B in ((43, 1 << 20), (100, 1 << 18), (8, 1 << 20), (20, 1 << 20)):
    W = householder_matrix(torch.randn(d, d, device=dev))
    x = torch.randn(B, d, device=dev)
    n = min(B, 1 << 16); yr = x[:n].double() @ W.double(); ya = householder_apply(x, W)[:n].double(); yt = (x[:n] @ W).double()
    print(f"d={d:3d}: apply max|err| vs fp64 {float((ya - yr).abs().max()):.2e} (torch fp32 matmul: {float((yt - yr).abs().max()):.2e}), max|y| {float(yr.abs().max()):.2f}", flush=True)
    for name, fn in (("apply", lambda: householder_apply(x, W)), ("wgrad", lambda: _wgrad(x, x))):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"d={d:3d} B={B}: {name} {ms:.3f} ms  ({2 * B * d * 4 / ms / 1e6:.0f} GB/s streamed, {2 * B * d * d / ms / 1e9:.1f} TFLOP/s)", flush=True)
