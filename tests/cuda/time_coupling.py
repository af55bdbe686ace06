"""Fused baseline coupling kernels vs the plain-PyTorch modules: python time_coupling.py"""
import sys, os, copy, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hint_b200 import coupling as K
from FrEIA.modules import ExternalAffineCoupling, AffineCoupling, F_fully_connected
dev = torch.device("cuda:0")
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, du, dv, H, B in (("lens ext", 2, 20, 68, 10000), ("lens y", 1, 1, 17, 10000), ("plus ext 8", 4, 100, 152, 10000), ("plus ext 4", 4, 100, 224, 10000),
                           ("lens ext", 2, 20, 68, 500), ("plus ext 4", 4, 100, 224, 500), ("lens ext 1M", 2, 20, 68, 1 << 20)):
    m = ExternalAffineCoupling([(dv,)], dims_c=[(du,)], F_class=F_fully_connected, F_args={"internal_size": H}).to(dev)
    params = [p.detach() for p in K.subnet_params(m.s, m.t)]
    u = torch.randn(B, du, device=dev); v = torch.randn(B, dv, device=dev); dy = torch.randn(B, dv, device=dev); dj = torch.randn(B, device=dev)
    f = t(lambda: K.forward(u, v, params, 5.0))
    b = t(lambda: K.backward(u, v, params, 5.0, dy, dj))
    flops = 2 * 2 * (du * H + 2 * H * H + H * dv) * B
    print(f"{name:12s} du={du} dv={dv} H={H} B={B}: fused forward {f*1e3:.1f} us ({flops/f/1e9:.1f} TFLOP/s), backward (2 launches) {b*1e3:.1f} us ({2*flops/b/1e9:.1f} TFLOP/s)", flush=True)
