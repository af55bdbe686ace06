"""Timing of the Householder matrix build / backward kernels and the plus-shape small-batch step pieces: python time_hh_matrix.py"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hint_b200.householder import householder_matrix, householder_vs_grad, _wgrad
dev = torch.device("cuda:0")
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for d, B in ((100, 500), (100, 10000), (43, 1 << 20), (20, 10000)):
    Vs = torch.randn(d, d, device=dev); W = householder_matrix(Vs)
    x = torch.randn(B, d, device=dev); dy = torch.randn(B, d, device=dev)
    print(f"d={d} B={B}: matrix {t(lambda: householder_matrix(Vs)):.3f} ms, wgrad {t(lambda: _wgrad(x, dy)):.3f} ms, vs_grad (wgrad + matrix backward) {t(lambda: householder_vs_grad(x, dy, Vs, W)):.3f} ms", flush=True)
