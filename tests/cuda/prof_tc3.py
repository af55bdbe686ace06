"""Developer aid: timeline of one steady-state tile of the tcgen05 training kernel (CTA 0, second tile): when every MMA record was
issued, when every epilogue step passed its wait and when it finished.  Prints a per-step table and totals per step type."""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hint_b200.block import TreePlan
from hint_b200 import _lib
TYPES = ["IN", "HID", "OUTS", "CPL", "DS", "DHID", "DA", "FLUSH"]
CFG = {"gas": (8, 0, [128, 64, 32, 16]), "power": (6, 0, [140, 70, 35, 17]), "d43": (43, 0, [67, 33, 16, 8]), "lens": (20, 0, [68, 34, 17, 17])}

def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "gas"
    verbose = len(sys.argv) > 2
    d, dc, ci = CFG[name]
    B = 148 * 128 * 3
    dev = torch.device("cuda:0")
    tp = TreePlan(d, dc, ci, 4.0, -1, 2, False)
    flat = (0.05 * torch.randn(tp.n_params)).to(dev)
    z = torch.randn(B, d, device=dev); dz = torch.randn(B, d, device=dev) / B; dJ = torch.full((B,), -1.0 / B, device=dev)
    lib = _lib.load()
    nbytes = lib.hint_workspace_bytes(tp._h, B, _lib.WS_BACKWARD)
    ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    prof = torch.zeros(8192, dtype=torch.int64, device=dev)
    dx = torch.zeros(B, d, device=dev); dflat = torch.zeros_like(flat)
    info = (ctypes.c_int32 * 4096)()
    vp = ctypes.c_void_p
    lib.hint_dev_tc3_profile.restype = ctypes.c_int
    for _ in range(2):
        rc = lib.hint_dev_tc3_profile(tp._h, vp(z.data_ptr()), None, vp(flat.data_ptr()), vp(dz.data_ptr()), vp(dJ.data_ptr()), ctypes.c_int64(B),
                                      vp(dx.data_ptr()), None, vp(dflat.data_ptr()), vp(prof.data_ptr()), info, vp(ws.data_ptr()),
                                      vp(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, _lib.last_error()
        torch.cuda.synchronize()
    p = prof.cpu().numpy()
    nm, ne, ninstr, tcyc = info[0], info[1], info[2], info[3]
    iss = p[:nm]; arr0 = p[1024:1024 + 2 * ne:2]; arr1 = p[1025:1025 + 2 * ne:2]; wdone = p[2048:2048 + ne]
    t0 = min(iss.min(), wdone.min())
    end = max(arr0.max(), arr1.max())
    print(f"{name}: tile {end - t0} cycles; {nm} records ({ninstr} MMA instructions, tensor floor {tcyc} cycles), {ne} epilogue steps")
    tot_work = {}; tot_wait = {}
    prev_end = t0
    for si in range(ne):
        ty = TYPES[info[16 + si] & 0xFF]
        a = max(arr0[si], arr1[si])
        work = a - wdone[si]; wait = wdone[si] - prev_end
        tot_work[ty] = tot_work.get(ty, 0) + work; tot_wait[ty] = tot_wait.get(ty, 0) + wait
        sub = p[2560 + 3 * si: 2563 + 3 * si]
        if verbose: print(f"  step {si:3d} {ty:5s} wait_mma {info[16 + si] >> 8:3d}: waited {wait:6d}  work {work:6d} (wg0 {arr0[si] - wdone[si]:6d} wg1 {arr1[si] - wdone[si]:6d}) body {sub[0] - wdone[si]:5d} wait_st {sub[1] - sub[0]:5d} fence.proxy {sub[2] - sub[1]:5d} publish {arr0[si] - sub[2]:5d}  t={a - t0}")
        prev_end = a
    print("  epilogue by type: " + "  ".join(f"{k}: work {tot_work[k]} wait {tot_wait[k]}" for k in tot_work))
    print(f"  epilogue total work {sum(tot_work.values())}  total wait {sum(tot_wait.values())}")
    gaps = np.diff(iss)
    ss = np.array([(info[1024 + i] >> 8) & 1 for i in range(nm)])
    nk = np.array([info[1024 + i] & 0xFF for i in range(nm)])
    if verbose:
        for i in range(nm): print(f"  rec {i:3d} {'SS' if ss[i] else 'TS'} N={info[1024 + i] >> 16:3d} nk={nk[i]:2d} flags={(info[1024 + i] >> 8) & 0xFF:02x} issued t={iss[i] - t0:7d}  (+{0 if i == 0 else gaps[i - 1]})  fetch+waits {p[4096 + 3 * i] - iss[i] if i < 512 else -1:5d}  mma issue {p[4097 + 3 * i] - p[4096 + 3 * i] if i < 512 else -1:5d}  commit {p[4098 + 3 * i] - p[4097 + 3 * i] if i < 512 else -1:5d}")
    print(f"  issuer: first record t={iss[0] - t0}, last t={iss[-1] - t0}; median gap per record {int(np.median(gaps))}; records with gap > 1000: {(gaps > 1000).sum()} (sum {gaps[gaps > 1000].sum()})")

if __name__ == "__main__":
    main()
