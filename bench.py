"""bench.py — the reference's headline metric on B200: samples/s of the HINT coupling-block hot path.

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--mode fp32|tf32|tf32x3]
    python bench.py --impl reference ...      # the reference algorithm's CPU path (oracle port) on the host cores

One "step" = one full training step of the workload's model on one batch of synthetic data:
x += 0.01*randn -> forward (all blocks) -> NLL loss -> backward -> [NCCL grad all-reduce] -> clamp +-5 -> Adam
(train_unconditional.py:121-144,174-176).  `value` = samples/s of that step with the batch resident in HBM;
`e2e` = the same through host buffers (pinned H2D of the batch and D2H of the loss inside the timed region);
forward+logdet and inverse throughputs are reported under "extra".  Default workload: the d=43 `hint_8` model of
BASELINE.json's weak-scaling sweep at 1,048,576 samples per GPU (the only listed config that is both a training
workload and defined for 1/2/4/8 GPUs); the other configs are parity-test cases and optional --workload values.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: d, dc, n_blocks, c_internal, max_splits, batch per GPU          
    "d43_hint_8": dict(d=43, dc=0, n_blocks=8, c_internal=[67, 33, 16, 8], max_splits=-1, batch=1 << 20),
    "miniboone_hint_8": dict(d=42, dc=0, n_blocks=8, c_internal=[67, 33, 16, 8], max_splits=-1, batch=1 << 18),
    "gas_hint_8": dict(d=8, dc=0, n_blocks=8, c_internal=[128, 64, 32, 16], max_splits=-1, batch=1 << 18),
    "power_hint_8": dict(d=6, dc=0, n_blocks=8, c_internal=[140, 70, 35, 17], max_splits=-1, batch=1 << 16),
    "lens_hint_8_full": dict(d=20, dc=0, n_blocks=8, c_internal=[68, 34, 17, 17], max_splits=-1, batch=10000),
    "lens_concat_cond": dict(d=20, dc=2, n_blocks=8, c_internal=[68, 34, 17, 17], max_splits=-1, batch=10000),
    "plus_hint_4_3": dict(d=100, dc=0, n_blocks=4, c_internal=[314, 157, 78, 39], max_splits=3, batch=10000),
    "plus_hint_4_full": dict(d=100, dc=0, n_blocks=4, c_internal=[263, 131, 65, 32, 32], max_splits=-1, batch=10000),
    "plus_cond_recursive_4": dict(d=100, dc=4, n_blocks=4, c_internal=[267, 133, 66], max_splits=-1, batch=10000),
}
METRIC = "samples/sec (train step: fwd+logdet, NLL, backward, clamp, Adam)"
ADAM = dict(lr=0.01, betas=(0.9, 0.95), eps=1e-4, weight_decay=1.86e-5)  # miniboone_hint_8.py:38-44, train_unconditional.py:174-176


def workload_config(name, w, batch, n_gpus):
    return {"workload": f"{name}: {w['n_blocks']} HINT blocks d={w['d']} dc={w['dc']} c_internal={w['c_internal']} "
                        f"max_splits={w['max_splits']}, full training step",
            "batch_per_gpu": batch, "global_batch": batch * n_gpus, "parallelism": f"dp{n_gpus}",
            "l2_hygiene": "inputs larger than L2" if batch * w["d"] * 4 > 126e6 else "L2 flushed between timed iterations",
            "weights": "0.005*randn (train_unconditional.py:165-167)", "data": "8-component diagonal GMM, standardised"}


def synthetic_batch(torch, B, d, dc, device, seed):
    """K=8 diagonal Gaussian mixture, standardised per feature (the recipe data.py:335-344 applies to UCI data)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    means = 3.0 * torch.randn(8, d, generator=g)
    stds = 0.3 + torch.rand(8, d, generator=g)
    gd = torch.Generator(device=device).manual_seed(seed + 1)
    comp = torch.randint(0, 8, (B,), generator=gd, device=device)
    x = means.to(device)[comp] + stds.to(device)[comp] * torch.randn(B, d, generator=gd, device=device)
    x = (x - x.mean(0)) / x.std(0)
    c = torch.randn(B, dc, generator=gd, device=device) if dc else None
    return x.contiguous(), c


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference algorithm on the host cores
# ------------------------------------------------------------------------------------------------------------
def cpu_train_steps(w, sample, steps, warmup, threads):
    import torch
    from oracle import hint_oracle as O
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    plan = O.build_plan(w["d"], w["dc"], w["c_internal"], w["max_splits"])
    flats = [(0.005 * torch.randn(O.param_count(plan))).requires_grad_(True) for _ in range(w["n_blocks"])]
    opt = torch.optim.Adam(flats, **ADAM)
    x, c = synthetic_batch(torch, sample, w["d"], w["dc"], "cpu", 1)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        xn = x + 0.01 * torch.randn_like(x)
        J = 0
        for f in flats:
            xn, Jb = O.forward_blockwise(plan, f, xn, c)
            J = J + Jb
        loss = O.nll_loss(xn, J)
        loss.item()
        loss.backward()
        for f in flats:
            f.grad.data.clamp_(-5.0, 5.0)
        opt.step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return times


def cpu_infer(w, sample, threads, rev, reps=3):
    import torch
    from oracle import hint_oracle as O
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    plan = O.build_plan(w["d"], w["dc"], w["c_internal"], w["max_splits"])
    flats = [0.005 * torch.randn(O.param_count(plan)) for _ in range(w["n_blocks"])]
    x, c = synthetic_batch(torch, sample, w["d"], w["dc"], "cpu", 1)
    ts = []
    with torch.no_grad():
        for it in range(reps + 1):
            t0 = time.perf_counter()
            v = x
            for f in (reversed(flats) if rev else flats):
                v, _ = (O.inverse_blockwise if rev else O.forward_blockwise)(plan, f, v, c)
            if it:
                ts.append(time.perf_counter() - t0)
    return statistics.median(ts)


def run_reference(args, w, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = min(w["batch"], args.cpu_sample)
    times = cpu_train_steps(w, sample, args.steps, args.warmup, threads)
    total = sum(times)
    value = sample * len(times) / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": workload_config(name, w, args.batch or w["batch"], args.gpus),
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port",
                             "sample": f"{sample} samples per step, {len(times)} steps (oracle/hint_oracle.py forward_blockwise "
                                       "+ torch autograd + Adam; the Python reference cannot travel to the GPU box)"},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def run_ours(args, w, name):
    import torch
    import torch.distributed as dist
    import hint_b200
    from hint_b200 import HintFlow, nll_loss, BucketedGradAllReduce, broadcast_parameters

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    hint_b200.set_precision(args.mode)
    B = args.batch or w["batch"]
    torch.manual_seed(0)
    model = HintFlow(w["d"], w["n_blocks"], w["c_internal"], dims_c=[(w["dc"],)] if w["dc"] else [],
                     max_splits=w["max_splits"]).to(dev)
    model.init_like_reference_scripts(0.005)
    broadcast_parameters(model)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, fused=True, **ADAM)
    reducer = BucketedGradAllReduce(model)
    x, c = synthetic_batch(torch, B, w["d"], w["dc"], dev, 1 + rank)
    loss_acc = torch.zeros((), device=dev)

    def train_step(xb, cb):
        opt.zero_grad(set_to_none=True)
        xn = xb + 0.01 * torch.randn_like(xb)
        z, J = model(xn, cb)
        loss = nll_loss(z, J)
        loss.backward()
        reducer.finish()
        for p in params:
            p.grad.clamp_(-5.0, 5.0)
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K calls bracketed by barrier+sync, CUDA events on the current stream, max over ranks (ms total)."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    flush = None
    if B * w["d"] * 4 <= 126e6:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed_avg(fn, steps):
        """ms per call; when the inputs fit in L2 every call is timed alone with an L2 flush before it."""
        if flush is None:
            return timed(fn, steps) / steps
        total = 0.0
        for _ in range(steps):
            flush.zero_()
            total += timed(fn, 1)
        return total / steps

    # ---- training step, batch resident in HBM ----------------------------------------------------------------
    def step_resident():
        loss_acc.add_(train_step(x, c).detach())

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_step = timed_avg(step_resident, args.steps)
    clocks = sampler.stop() if sampler else None
    value = B * world / (ms_step * 1e-3)

    # ---- end to end: host batch (pinned) -> H2D -> step -> D2H of the loss, every step ------------------------
    xh = x.cpu().pin_memory()
    ch = c.cpu().pin_memory() if c is not None else None
    h2d = xh.numel() * 4 + (ch.numel() * 4 if ch is not None else 0)

    def step_e2e():
        xb = xh.to(dev, non_blocking=True)
        cb = ch.to(dev, non_blocking=True) if ch is not None else None
        train_step(xb, cb).item()

    step_e2e()
    ms_e2e = timed_avg(step_e2e, args.steps)
    e2e = {"value": B * world / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
           "ms_per_step": ms_e2e}

    # ---- forward+logdet and inverse (no grad, no collective) --------------------------------------------------
    extra = {}
    with torch.no_grad():
        z0, _ = model(x, c)
        for nm, fn in (("fwd_logdet", lambda: model(x, c)), ("inverse", lambda: model(z0, c, rev=True))):
            for _ in range(2):
                fn()
            ms = timed_avg(fn, args.steps)
            extra[nm + "_samples_per_s"] = B * world / (ms * 1e-3)
            extra[nm + "_ms"] = ms
        xr, _ = model(z0, c, rev=True)
        extra["invertibility_max_abs_err"] = float((xr - x).abs().max().item())
    F = model.flops_per_sample
    extra["flops_per_sample_fwd"] = F
    extra["train_tflops_algorithmic"] = 3 * F * value / 1e12
    extra["fwd_tflops_algorithmic"] = F * extra["fwd_logdet_samples_per_s"] / 1e12

    # ---- roofline of the dominant kernel: the fused backward of one block -------------------------------------
    blk = model.blocks[0]
    with torch.no_grad():
        zb = blk([x], c=[] if c is None else [c])[0]
        dz = zb / B
        dJ = torch.full((B,), -1.0 / B, device=dev)
        flat = blk.flat.detach()
        run_b = lambda: blk.plan.backward(zb, c, flat, dz, dJ)
        run_b()
        ms_b = timed_avg(run_b, args.steps)
        run_f = lambda: blk.plan.forward(x, c, flat)
        run_f()
        ms_f = timed_avg(run_f, args.steps)
    peaks, peak_src = measured_peaks()
    tf32_peak = peaks["bf16_tflops_sustained"] / 2.0
    Fb = blk.plan.flops_per_sample
    achieved = 2 * Fb * B / (ms_b * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(name, {}).get("bwd_dram_bytes_per_launch")
    kname = {"fp32": "hint_bwd_fp32_kernel", "tf32_tcgen05": "hint_bwd_fp32_kernel", "tf32x3": "hint_bwd_mma_kernel<TM,3xTF32>"}.get(
        args.mode, "hint_bwd_mma_kernel<TM,TF32>")
    chain = args.mode in ("tf32", "tf32_chain") and blk.plan.mode_supported("tf32_chain") and os.environ.get("HINT_B200_TF32_BWD") != "mma"
    if chain:
        kname = "hint_bwd_chain_kernel<MT=1,NW=4> (register-chained warp-MMA)"
    # second denominator: what mma.sync.m16n8k8 tf32 itself sustains on this chip (profiles/ubench5_r01_mma_sync_tf32.txt);
    # the kernels of this path issue warp-level MMAs, tcgen05 does not fit the 8..72-wide layers
    mma_sync_peak = 270.0
    roofline = {"bound": "tensor", "kernel": f"{kname} (one block, B={B})", "achieved": achieved,
                "peak": tf32_peak, "unit": "TFLOP/s", "frac": achieved / tf32_peak, "traffic": traffic,
                "peak_source": f"TF32 dense = bf16_tflops_sustained/2 of {peak_src}",
                "flops_per_launch": 2 * Fb * B, "ms_per_launch": ms_b,
                "frac_of_mma_sync_tf32_peak": achieved / mma_sync_peak, "mma_sync_tf32_peak_tflops": mma_sync_peak,
                "fwd_kernel": {"achieved": Fb * B / (ms_f * 1e-3) / 1e12, "ms_per_launch": ms_f},
                "hbm_gbs_fwd_streaming": (2 * w["d"] + 1 + w["dc"]) * 4 * B / (ms_f * 1e-3) / 1e9,
                "note": "launch = pack + fused kernel (+ partial-gradient reduce), timed with CUDA events on the launch stream"}

    line = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.mode,
            "data": "synthetic", "config": workload_config(name, w, B, world), "roofline": roofline, "e2e": e2e,
            "gpu_launches": args.steps * w["n_blocks"] * 5, "clocks": clocks, "extra": extra,
            "final_loss": float(loss_acc.item()) / max(1, args.steps + args.warmup)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sample = min(B, args.cpu_sample)
        ts = cpu_train_steps(w, sample, 3, 1, threads)
        v = sample / statistics.median(ts)
        line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": threads, "kind": "port",
                                "sample": f"{sample} samples per step, median of 3 steps after 1 warm-up (oracle port of the "
                                          "reference algorithm, torch CPU fp32, all host threads)",
                                "fwd_logdet_samples_per_s": sample / cpu_infer(w, sample, threads, False),
                                "inverse_samples_per_s": sample / cpu_infer(w, sample, threads, True)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="d43_hint_8", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="samples per GPU (default: the workload's)")
    ap.add_argument("--mode", default=os.environ.get("HINT_B200_MODE", "tf32"), choices=["fp32", "tf32", "tf32x3", "tf32_mma", "tf32_tcgen05"])
    ap.add_argument("--cpu-sample", type=int, default=32768, help="samples per CPU step (bounded sample of the workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w, args.workload)
    else:
        run_ours(args, w, args.workload)


if __name__ == "__main__":
    main()
