import sys, os, torch
sys.path.insert(0, "/root/repo")
from hint_b200.householder import householder_apply, householder_matrix
dev = torch.device("cuda:0")
d, B = 43, 1 << 20
W = householder_matrix(torch.randn(d, d, device=dev)); x = torch.randn(B, d, device=dev)
for _ in range(2): householder_apply(x, W)
torch.cuda.synchronize()
