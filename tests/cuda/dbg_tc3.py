"""Developer aid: step-limited run of the tcgen05 training kernel against the CPU interpreter of the same program.
For n = 1..n_steps: run the first n epilogue steps of ONE tile on the GPU (hint_dev_tc3_debug) and in tests/emul/emul_tc3.cpp,
compare TMEM, the operand images and the per-sample state; report the first step whose results differ."""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from hint_b200.block import TreePlan
from hint_b200 import _lib
import emul_tc3_lib

TYPES = ["IN", "HID", "OUTS", "CPL", "DS", "DHID", "DA", "FLUSH"]

def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    cfgs = {"tiny": (2, 0, [5], -1, 0.3), "gas": (8, 0, [128, 64, 32, 16], -1, 0.1), "power": (6, 0, [140, 70, 35, 17], -1, 0.1),
            "d43": (43, 0, [67, 33, 16, 8], -1, 0.15), "lens": (20, 0, [68, 34, 17, 17], -1, 0.15)}
    d, dc, ci, ms, scale = cfgs[name]
    B = 100
    dev = torch.device("cuda:0")
    tp = TreePlan(d, dc, ci, 4.0, ms, 2, False)
    g = torch.Generator().manual_seed(3)
    flat = (scale * torch.randn(tp.n_params, generator=g)).float()
    craft = os.environ.get('CRAFT')
    if craft and name == 'tiny':
        flat.zero_(); flat[5:10] = torch.tensor([1., 2., 3., 4., 5.]); flat[46 + 5:46 + 10] = torch.tensor([1., 2., 3., 4., 5.])
        if craft == 'eye': flat[10:35] = torch.eye(5).flatten()
        if craft == 'bias': flat[35:40] = torch.tensor([10., 20., 30., 40., 50.])
        if craft == 'both': flat[10:35] = torch.eye(5).flatten(); flat[35:40] = torch.tensor([10., 20., 30., 40., 50.])
        if craft == 'full': flat[10:35] = torch.arange(25.).float() + 1
        rnd = (0.3 * torch.randn(92, generator=torch.Generator().manual_seed(11))).float()
        if craft in ('r1', 'r12'): flat[0:10] = rnd[0:10]
        if craft == 'r1': flat[10:35] = torch.eye(5).flatten(); flat[35:40] = torch.tensor([10., 20., 30., 40., 50.])
        if craft in ('r2', 'r12'): flat[10:40] = rnd[10:40]
        if craft == 'r1p': flat[0:5] = rnd[0:5].abs() * 0.1; flat[10:35] = torch.eye(5).flatten()
        if craft == 'r2i': flat[10:35] = torch.round(rnd[10:35] * 20)
        if craft == 'r2h': flat[10:35] = torch.round(rnd[10:35] * 20) / 1024
    z = torch.randn(B, d, generator=g); dz = torch.randn(B, d, generator=g) / B; dJ = torch.randn(B, generator=g) / B
    lib = _lib.load()
    lib.hint_dev_tc3_debug.restype = ctypes.c_int
    nbytes = lib.hint_workspace_bytes(tp._h, B, _lib.WS_BACKWARD)
    ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    dump = torch.zeros(131072 + 4096, dtype=torch.float32, device=dev)
    zg, dzg, dJg, fg = z.to(dev), dz.to(dev), dJ.to(dev), flat.to(dev)
    xrec = torch.zeros(B, d, device=dev); dx = torch.zeros(B, d, device=dev)
    layout = (ctypes.c_int32 * 32)()
    elib = emul_tc3_lib.lib()
    ci_arr = np.asarray(ci, np.int32)
    P = lambda a, t=ctypes.c_float: a.ctypes.data_as(ctypes.POINTER(t))
    first_bad = None
    n = 1
    n_steps = None
    vp = ctypes.c_void_p
    while True:
        dump.zero_()
        rc = lib.hint_dev_tc3_debug(tp._h, ctypes.c_int32(n), vp(zg.data_ptr()), None, vp(fg.data_ptr()), vp(dzg.data_ptr()), vp(dJg.data_ptr()),
                                    ctypes.c_int64(B), vp(xrec.data_ptr()), vp(dx.data_ptr()), None, vp(dump.data_ptr()), layout, vp(ws.data_ptr()),
                                    vp(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, _lib.last_error()
        torch.cuda.synchronize()
        L = list(layout)
        sm_xs, sm_gs, sm_os, xp, op = L[0:5]; sm_img = L[5:10]; rows = L[10:15]; n_steps = L[15]
        dmp = dump.cpu().numpy()
        tm_g = dmp[:65536].reshape(128, 512); sm_g = dmp[65536:]
        tm_e = np.zeros((128, 512), np.float32); img_e = np.zeros(sum(rows) * 128, np.float32)
        xs_e = np.zeros(128 * xp, np.float32); gs_e = np.zeros(128 * xp, np.float32); os_e = np.zeros(128 * op, np.float32)
        info = np.zeros(16, np.int64)
        zc, dzc, dJc, fc = (np.ascontiguousarray(a.numpy(), np.float32) for a in (z, dz, dJ, flat))
        rc = elib.emul_tc3_debug(ctypes.c_int(d), ctypes.c_int(dc), P(ci_arr, ctypes.c_int), ctypes.c_int(len(ci)), ctypes.c_double(4.0), ctypes.c_int(ms),
                                 ctypes.c_int(2), P(fc), P(zc), None, P(dzc), P(dJc), ctypes.c_longlong(B), ctypes.c_int(1), ctypes.c_int(n), P(tm_e), P(img_e),
                                 P(xs_e), P(gs_e), P(os_e), P(info, ctypes.c_longlong))
        assert rc == 0, rc
        def cmp(a, b, what):
            m = np.isfinite(b)
            if not m.any(): return 0.0
            return float(np.abs(np.where(m, a, 0) - np.where(m, b, 0)).max())
        errs = {"tmem": cmp(tm_g, tm_e, "tmem"), "xs": cmp(sm_g[sm_xs // 4: sm_xs // 4 + 128 * xp], xs_e, "xs"),
                "gs": cmp(sm_g[sm_gs // 4: sm_gs // 4 + 128 * xp], gs_e, "gs"), "os": cmp(sm_g[sm_os // 4: sm_os // 4 + 128 * op], os_e, "os")}
        o = 0
        for i in range(5):
            if rows[i] and sm_img[i] // 4 + rows[i] * 128 <= sm_g.size:
                errs[f"img{i}"] = cmp(sm_g[sm_img[i] // 4: sm_img[i] // 4 + rows[i] * 128], img_e[o:o + rows[i] * 128], "img")
            o += rows[i] * 128
        worst = max(errs.values())
        if os.environ.get('TIMES') and n <= 4:
            ts = dump[131072:].cpu().numpy().view(np.int64)
            t0 = ts[ts > 0].min()
            print('  issue clocks (after wait) of records:', (ts[:6] - t0).tolist()); print('  arrive clocks of epi steps x warps:', [(ts[64 + 8 * s: 64 + 8 * s + 8] - t0).tolist() for s in range(min(n, 4))])
        if craft and n in (2, 3, 4):
            np.set_printoptions(precision=3, suppress=True, linewidth=200)
            print('gpu Q lane0', tm_g[0, 24:32], 'lane5', tm_g[5, 24:32]); print('emu Q lane0', tm_e[0, 24:32], 'lane5', tm_e[5, 24:32]); print('gpu P lane0', tm_g[0, 0:24]); print('gpu OUT lane0', tm_g[0, 96:100], 'emu', tm_e[0, 96:100])
        bad_cols = ""
        if errs["tmem"] > 1e-2:
            m = np.isfinite(tm_e)
            dcol = np.abs(np.where(m, tm_g, 0) - np.where(m, tm_e, 0)).max(axis=0)
            bad = np.nonzero(dcol > 1e-2)[0]
            bad_cols = f" bad tmem cols {bad[:12]}..{bad[-1]} (n={bad.size})"
        print(f"n={n:3d} {TYPES[int(info[12])]:5s} wait_mma={int(info[13]):3d} nrec={int(info[14]):3d} a={int(info[15]):3d} " + " ".join(f"{k} {v:.1e}" for k, v in errs.items()) + bad_cols, flush=True)
        if worst > 1e-2 and first_bad is None:
            first_bad = n
            np.save(os.path.join(ROOT, "gpurun_out", f"dbg_tc3_{name}_gpu_tmem.npy"), tm_g)
            np.save(os.path.join(ROOT, "gpurun_out", f"dbg_tc3_{name}_emu_tmem.npy"), tm_e)
            if n > 6 + 1 and not craft: break
        if first_bad is not None and n >= first_bad + 2 and not craft: break
        if craft and n >= 4: break
        n += 1
        if n > n_steps: break
    print("first bad step:", first_bad, "of", n_steps)

if __name__ == "__main__":
    main()
