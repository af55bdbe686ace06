"""Developer aid: times one block's forward / inverse / backward launches (CUDA events) for a workload and mode.
    python tools/time_block.py [--workload d43_hint_8] [--mode tf32] [--batch 1048576] [--what fwd,inv,bwd]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS, synthetic_batch  # noqa: E402
from hint_b200 import HierarchicalAffineCouplingBlock  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="d43_hint_8")
    ap.add_argument("--mode", default="tf32")
    ap.add_argument("--batch", type=int, default=1 << 20)
    ap.add_argument("--what", default="fwd,inv,bwd")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    w = WORKLOADS[a.workload]
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    blk = HierarchicalAffineCouplingBlock([(w["d"],)], dims_c=[(w["dc"],)] if w["dc"] else [], c_internal=w["c_internal"],
                                          max_splits=w["max_splits"]).to(dev)
    with torch.no_grad():
        blk.flat.copy_(0.05 * torch.randn_like(blk.flat))
    x, c = synthetic_batch(torch, a.batch, w["d"], w["dc"], dev, 1)
    flat = blk.flat.detach()
    B = a.batch
    with torch.no_grad():
        z, J = blk.plan.forward(x, c, flat, mode=a.mode)
        z32, J32 = blk.plan.forward(x, c, flat, mode="fp32")
        print(f"mode {a.mode}: max|z - z_fp32| = {(z - z32).abs().max().item():.3e}  max|J - J_fp32| = {(J - J32).abs().max().item():.3e}")
        dz = z / B
        dJ = torch.full((B,), -1.0 / B, device=dev)
        fns = {"fwd": lambda: blk.plan.forward(x, c, flat, mode=a.mode),
               "inv": lambda: blk.plan.forward(z, c, flat, rev=True, mode=a.mode),
               "bwd": lambda: blk.plan.backward(z, c, flat, dz, dJ, mode=a.mode)}
        for name in a.what.split(","):
            fn = fns[name]
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.reps
            F = blk.plan.flops_per_sample * (2 if name == "bwd" else 1)
            print(f"{name}: {ms:.3f} ms per launch  ({B / ms * 1e-3:.1f} M samples/s, {F * B / ms * 1e-9:.1f} TFLOP/s algorithmic)")


if __name__ == "__main__":
    main()
