"""Host-side (Python) overhead of one block's autograd forward + backward at a small batch: python prof_host.py"""
import sys, os, time, cProfile, pstats, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import hint_b200
from hint_b200 import HierarchicalAffineCouplingBlock
from hint_b200.householder import HouseholderPerm
hint_b200.set_precision("tf32")
dev = torch.device("cuda:0")
blk = HierarchicalAffineCouplingBlock([(20,)], c_internal=[68, 34, 17, 17]).to(dev)
perm = HouseholderPerm([(20,)], n_reflections=20, fixed=True).to(dev)
x = torch.randn(1000, 20, device=dev, requires_grad=True)
def step():
    z = blk([perm([x])[0]])[0]
    J = blk.jacobian(None)
    (0.5 * z.pow(2).sum(1).mean() - J.mean()).backward()
for _ in range(20): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): step()
torch.cuda.synchronize()
print(f"block + perm fwd+bwd, B=1000: {(time.perf_counter() - t0) / 200 * 1e3:.3f} ms per step (host-bound)")
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
