"""Developer check of the tcgen05 training kernel (mode tf32_tc3) on the GPU: parity of the backward against the fp64 oracle on
the reference configs' real widths (several tiles incl. a ragged one), then timing against the other backward kernels."""
import sys, os, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hint_b200.block import TreePlan
from oracle import hint_oracle as O

CONFIGS = [
    ("tiny", 2, 0, [5], -1, 77, 0.3),
    ("gas_hint_8", 8, 0, [128, 64, 32, 16], -1, 300, 0.1),
    ("power_hint_8", 6, 0, [140, 70, 35, 17], -1, 300, 0.1),
    ("d43_hint_8", 43, 0, [67, 33, 16, 8], -1, 300, 0.15),
    ("lens_hint_8_full", 20, 0, [68, 34, 17, 17], -1, 300, 0.15),
    ("lens_concat_cond", 20, 2, [68, 34, 17, 17], -1, 300, 0.15),
]

def l2(a, ref):
    a = a.detach().double().cpu().numpy(); ref = ref.numpy()
    return float(np.linalg.norm(a - ref) / max(1e-30, np.linalg.norm(ref)))

def main():
    dev = torch.device("cuda:0")
    only = sys.argv[1:] 
    for name, d, dc, ci, ms, B, scale in CONFIGS:
        if only and name not in only and "time" not in only: continue
        tp = TreePlan(d, dc, ci, 4.0, ms, 2, False)
        if not tp.mode_supported("tf32_tc3"):
            print(name, "outside tc3 envelope"); continue
        plan = O.build_plan(d, dc, ci, ms)
        g = torch.Generator().manual_seed(7)
        flat = (scale * torch.randn(tp.n_params, generator=g)).float()
        x = torch.randn(B, d, generator=g)
        c = torch.randn(B, dc, generator=g) if dc else None
        f64 = flat.double()
        z_ref, J_ref = O.forward_fast(plan, f64, x.double(), None if c is None else c.double())
        dz = torch.randn(B, d, generator=g).double() / B
        dJ = torch.randn(B, generator=g).double() / B
        xr_ref, dx_ref, dc_ref, dp_ref = O.backward_from_output(plan, f64, z_ref, None if c is None else c.double(), dz, dJ)
        zg, cg = z_ref.float().to(dev), (c.to(dev) if dc else None)
        res = {}
        for mode in ("tf32_tc3", "tf32_mma"):
            try:
                dx, dcg, dflat, xrec = tp.backward(zg, cg, flat.to(dev), dz.float().to(dev), dJ.float().to(dev), mode=mode, want_xrec=True)
                torch.cuda.synchronize()
            except Exception as e:
                print(name, mode, "FAILED:", e); continue
            res[mode] = (l2(dx, dx_ref), l2(dflat, dp_ref), float((xrec.cpu().double() - x.double()).abs().max()), l2(dcg, dc_ref) if dc else 0.0)
            print(f"{name:18s} {mode:9s} dx {res[mode][0]:.2e} dparams {res[mode][1]:.2e} xrec {res[mode][2]:.2e} dc {res[mode][3]:.2e}", flush=True)
            if mode == "tf32_tc3" and res[mode][1] > 3e-2:
                # which parameter tensors are off
                got = dflat.cpu().double().numpy(); ref = dp_ref.numpy()
                for nm, off, shape in tp.entries:
                    n = int(np.prod(shape)); a, b = got[off:off + n], ref[off:off + n]
                    err = np.linalg.norm(a - b) / max(1e-30, np.linalg.norm(b))
                    if err > 3e-2: print(f"    {nm:28s} {shape} rel-l2 {err:.2e} |got| {np.linalg.norm(a):.2e} |ref| {np.linalg.norm(b):.2e}")
    if "time" in only or not only:
        for name, d, dc, ci, ms, B in [("gas_hint_8", 8, 0, [128, 64, 32, 16], -1, 262144), ("power_hint_8", 6, 0, [140, 70, 35, 17], -1, 65536 * 4),
                                     ("d43_hint_8", 43, 0, [67, 33, 16, 8], -1, 262144), ("lens_hint_8_full", 20, 0, [68, 34, 17, 17], -1, 262144)]:
            tp = TreePlan(d, dc, ci, 4.0, ms, 2, False)
            flat = (0.05 * torch.randn(tp.n_params)).to(dev)
            z = torch.randn(B, d, device=dev); dz = torch.randn(B, d, device=dev) / B; dJ = torch.full((B,), -1.0 / B, device=dev)
            for mode in ("tf32_tc3", "tf32", "tf32_mma"):
                if not tp.mode_supported(mode): continue
                try:
                    for _ in range(2): tp.backward(z, None, flat, dz, dJ, mode=mode)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(5): tp.backward(z, None, flat, dz, dJ, mode=mode)
                    e1.record(); torch.cuda.synchronize()
                    ms_ = e0.elapsed_time(e1) / 5
                    fl = 2 * tp.flops_per_sample * B
                    print(f"time {name:18s} {mode:9s} B={B}: {ms_:.3f} ms  -> {fl / ms_ / 1e9:.1f} TFLOP/s algorithmic (bwd = 2x fwd flops)", flush=True)
                except Exception as e:
                    print("time", name, mode, "FAILED:", e)

if __name__ == "__main__":
    main()
