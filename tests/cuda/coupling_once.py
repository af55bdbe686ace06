"""One forward + backward of the fused lens y -> x coupling at B = 10000 (for ncu captures)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hint_b200 import coupling as K
from FrEIA.modules import ExternalAffineCoupling, F_fully_connected
dev = torch.device("cuda:0")
du, dv, H, B = 2, 20, 68, 10000
m = ExternalAffineCoupling([(dv,)], dims_c=[(du,)], F_class=F_fully_connected, F_args={"internal_size": H}).to(dev)
params = [p.detach() for p in K.subnet_params(m.s, m.t)]
u = torch.randn(B, du, device=dev); v = torch.randn(B, dv, device=dev); dy = torch.randn(B, dv, device=dev); dj = torch.randn(B, device=dev)
for _ in range(3):
    K.forward(u, v, params, 5.0); K.backward(u, v, params, 5.0, dy, dj)
torch.cuda.synchronize()
