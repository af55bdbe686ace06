"""Forward / inverse timing per config and mode: python time_fwd.py [B]."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hint_b200.block import TreePlan
CFG = {"plus43": (100, 0, [314, 157, 78, 39], 3), "plus4f": (100, 0, [263, 131, 65, 32, 32]), "pluscond": (100, 4, [267, 133, 66]),
       "gas": (8, 0, [128, 64, 32, 16]), "power": (6, 0, [140, 70, 35, 17]), "mini4": (42, 0, [102, 51, 25, 12]), "power4": (6, 0, [200, 100, 50, 25]),
       "d43": (43, 0, [67, 33, 16, 8]), "lens": (20, 0, [68, 34, 17, 17])}
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
dev = torch.device("cuda:0")
for name, cfg in CFG.items():
    d, dc, ci = cfg[:3]
    tp = TreePlan(d, dc, ci, 4.0, cfg[3] if len(cfg) > 3 else -1, 2, False)
    flat = (0.02 * torch.randn(tp.n_params)).to(dev)
    x = torch.randn(B, d, device=dev)
    c = torch.randn(B, dc, device=dev) if dc else None
    for mode in ("tf32", "tf32_tc3", "tf32_mma"):
        if not tp.mode_supported(mode):
            continue
        for rev in (False, True):
            for _ in range(2):
                tp.forward(x, c, flat, rev, mode=mode)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                tp.forward(x, c, flat, rev, mode=mode)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"{name:7s} {mode:9s} {'inv' if rev else 'fwd'} B={B}: {ms:.3f} ms -> {tp.flops_per_sample * B / ms / 1e9:.1f} TFLOP/s", flush=True)
