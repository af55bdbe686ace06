"""Fused multi-kernel MMD vs the reference script's PyTorch expression (rejection_sampling.py:56-73): python time_mmd.py"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import hint_b200
dev = torch.device("cuda:0")
def torch_mmd(x, y, we=((0.5, 1), (0.2, 1), (0.2, 0.5))):
    xx, yy, xy = torch.mm(x, x.t()), torch.mm(y, y.t()), torch.mm(x, y.t())
    rx = xx.diag().unsqueeze(0).expand_as(xx); ry = yy.diag().unsqueeze(0).expand_as(yy)
    dxx = torch.clamp(rx.t() + rx - 2. * xx, 0, float("inf")); dyy = torch.clamp(ry.t() + ry - 2. * yy, 0, float("inf")); dxy = torch.clamp(rx.t() + ry - 2. * xy, 0, float("inf"))
    XX, YY, XY = torch.zeros_like(xx), torch.zeros_like(xx), torch.zeros_like(xx)
    for C, a in we:
        XX += C ** a * ((C + dxx) / a) ** -a; YY += C ** a * ((C + dyy) / a) ** -a; XY += C ** a * ((C + dxy) / a) ** -a
    return torch.mean(XX + YY - 2. * XY)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for n, d in ((4000, 20), (4000, 100), (16384, 20)):
    x = torch.randn(n, d, device=dev); y = 0.2 + torch.randn(n, d, device=dev)
    a, b = float(hint_b200.multi_mmd(x, y)), float(torch_mmd(x, y))
    print(f"n={n} d={d}: fused {t(lambda: hint_b200.multi_mmd(x, y)):.3f} ms, PyTorch expression {t(lambda: torch_mmd(x, y)):.3f} ms; values {a:.6f} / {b:.6f}", flush=True)
