"""Forward / inverse launch time of the chain kernels (d43, lens): python time_fwd_chain.py"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hint_b200.block import TreePlan
CFG = {"d43": (43, 0, [67, 33, 16, 8]), "lens": (20, 0, [68, 34, 17, 17]), "d42": (42, 0, [67, 33, 16, 8])}
dev = torch.device("cuda:0")
for name in sys.argv[1:] or ["d43", "lens"]:
    d, dc, ci = CFG[name]
    B = 1 << 20
    tp = TreePlan(d, dc, ci, 4.0, -1, 2, False)
    flat = (0.05 * torch.randn(tp.n_params)).to(dev)
    x = torch.randn(B, d, device=dev)
    for rev in (False, True):
        for _ in range(3): tp.forward(x, None, flat, rev, mode="tf32")
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): tp.forward(x, None, flat, rev, mode="tf32")
        e1.record(); torch.cuda.synchronize()
        print(f"{name} B={B} rev={rev} cfg={os.environ.get('HINT_B200_CHAIN_FWD', 'default')}: {e0.elapsed_time(e1) / 10:.3f} ms", flush=True)
