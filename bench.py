"""bench.py — the reference's headline metric on B200: samples/s of the HINT coupling-block hot path.

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--mode fp32|tf32|tf32x3]
    python bench.py --impl reference ...      # the reference algorithm's CPU path (oracle port) on the host cores

One "step" = one full training step of the workload's model on one batch of synthetic data:
x += 0.01*randn -> forward (all blocks) -> NLL loss -> backward -> [NCCL grad all-reduce] -> clamp +-5 -> Adam
(train_unconditional.py:121-144,174-176), run through hint_b200.FusedTrainStep: every kernel of the step is this library's
(`gpu_launches` is counted by the library).  `value` = samples/s of that step with the batch resident in HBM;
`e2e` = the same through host buffers (pinned H2D of the batch and D2H of the loss inside the timed region, the copy
prefetched one step ahead on a copy stream); forward+logdet and inverse throughputs are under "extra" and, as fractions of the
TF32 peak MEASURED in the same run, under roofline.phases; "modes" repeats the step in fp32 / 3xTF32 and "configs" runs every
BASELINE workload at its stated batch (1 GPU).  Default workload: the d=43 `hint_8` model of
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
# every BASELINE.json workload at its stated batch: plus_shape at 500 (BASELINE text) and 10 000 (the configs' own)
CONFIG_SWEEP = [("plus_hint_4_3", 500), ("plus_hint_4_3", 10000), ("plus_hint_4_full", 10000), ("plus_cond_recursive_4", 10000),
                ("lens_hint_8_full", 10000), ("lens_concat_cond", 10000), ("power_hint_8", 1 << 16), ("gas_hint_8", 1 << 18),
                ("miniboone_hint_8", 1 << 18), ("d43_hint_8", 1 << 20)]
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


def measure_tf32_peak(torch, dev):
    """TF32 dense tensor-pipe peak of THIS box: torch.matmul (cuBLAS, allow_tf32) 8192^3, same method as MEASURED_PEAKS.json's
    bf16 figures: best single launch of 10 (burst) and back-to-back launches for ~1.5 s (sustained).  A yardstick only."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        out = torch.empty(n, n, device=dev)
        for _ in range(3):
            torch.matmul(a, b, out=out)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b, out=out); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        reps = max(20, int(1500.0 / best))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            torch.matmul(a, b, out=out)
        e1.record()
        torch.cuda.synchronize()
        fl = 2.0 * n ** 3
        return {"burst_tflops": fl / (best * 1e-3) / 1e12, "sustained_tflops": fl * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12,
                "how": f"torch.matmul fp32 inputs, allow_tf32, {n}^3: best of 10 single launches (burst), {reps} back-to-back (sustained)"}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
        del a, b, out


def bwd_kernel_name(plan, mode):
    if mode == "fp32":
        return "hint_bwd_fp32_kernel"
    if mode == "tf32x3":
        return "hint_bwd_mma_kernel<TM,3xTF32>"
    if mode in ("tf32", "tf32_chain") and plan.mode_supported("tf32_chain"):
        return "hint_bwd_chain_kernel<MT=1,NW=4> (register-chained warp-MMA)"
    if mode in ("tf32", "tf32_tc3", "tf32_tcgen05") and plan.mode_supported("tf32_tc3"):
        return "hint_tc3_bwd_kernel (tcgen05 / TMEM)"
    return "hint_bwd_mma_kernel<TM,TF32>"


def run_ours(args, w, name):
    import torch
    import torch.distributed as dist
    import hint_b200
    from hint_b200 import HintFlow, nll_loss, BucketedGradAllReduce, broadcast_parameters, FusedClampAdam, FusedTrainStep
    from hint_b200 import _lib as hlib

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
    lib = hlib.load()
    launches = lambda: int(lib.hint_launch_count())

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

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def make_timed_avg(in_l2):
        def timed_avg(fn, steps):
            """ms per call; when the inputs fit in L2 every call is timed alone with an L2 flush before it."""
            if not in_l2:
                return timed(fn, steps) / steps
            total = 0.0
            for _ in range(steps):
                flush_buf.zero_()
                total += timed(fn, 1)
            return total / steps
        return timed_avg

    def build(wl, B, mode, seed_rank=True, householder=None):
        hint_b200.set_precision(mode)
        torch.manual_seed(0)
        model = HintFlow(wl["d"], wl["n_blocks"], wl["c_internal"], dims_c=[(wl["dc"],)] if wl["dc"] else [],
                         max_splits=wl["max_splits"], householder=householder).to(dev)
        model.init_like_reference_scripts(0.005)
        broadcast_parameters(model)
        params = [p for p in model.parameters() if p.requires_grad]
        opt = FusedClampAdam(params, grad_clamp=5.0, **ADAM)
        trainer = FusedTrainStep(model, opt, noise=0.01, seed=1234 + rank)
        x, c = synthetic_batch(torch, B, wl["d"], wl["dc"], dev, 1 + (rank if seed_rank else 0))
        return model, params, opt, trainer, x, c

    def phases(wl, B, mode, steps, warmup, householder=None):
        """train step / fwd+logdet / inverse of one workload: samples/s (whole job) and the launch count of one train step."""
        model, params, opt, trainer, x, c = build(wl, B, mode, householder=householder)
        tavg = make_timed_avg(B * wl["d"] * 4 <= 126e6)
        for _ in range(warmup):
            trainer.step(x, c)
        n0 = launches()
        ms_t = tavg(lambda: trainer.step(x, c), steps)
        per_step = (launches() - n0) // steps
        out = {"train_samples_per_s": B * world / (ms_t * 1e-3), "train_ms": ms_t, "launches_per_train_step": per_step}
        with torch.no_grad():
            z0, _ = model(x, c)
            for nm, fn in (("fwd_logdet", lambda: model(x, c)), ("inverse", lambda: model(z0, c, rev=True))):
                fn()
                ms = tavg(fn, steps)
                out[nm + "_samples_per_s"] = B * world / (ms * 1e-3)
                out[nm + "_ms"] = ms
            if B <= 65536:      # the configs' own batch sizes are launch-bound: CUDA-graph replay of the same launches
                try:
                    from hint_b200 import GraphedFlow
                    gf = GraphedFlow(model, B)
                    gf(x, c)
                    ms = tavg(lambda: gf(x, c), steps)
                    out["fwd_logdet_graph_samples_per_s"] = B * world / (ms * 1e-3)
                    del gf
                except Exception as e:
                    out["fwd_logdet_graph_error"] = str(e)[:120]
        out["flops_per_sample_fwd"] = model.flops_per_sample
        return out, (model, params, opt, trainer, x, c)

    B = args.batch or w["batch"]
    tf32_peak = measure_tf32_peak(torch, dev) if (rank == 0 and not args.skip_peak) else None
    hint_b200.set_precision(args.mode)
    model, params, opt, trainer, x, c = build(w, B, args.mode)
    timed_avg = make_timed_avg(B * w["d"] * 4 <= 126e6)

    # ---- training step, batch resident in HBM: the library's fused step (noise, forward, NLL, backward, [all-reduce], clamp+Adam) ----
    last_loss = [None]

    def step_resident():
        last_loss[0] = trainer.step(x, c)

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    n0 = launches()
    ms_step = timed_avg(step_resident, args.steps)
    gpu_launches = launches() - n0
    clocks = sampler.stop() if sampler else None
    value = B * world / (ms_step * 1e-3)
    final_loss = float(last_loss[0][0].item())

    # ---- end to end: pinned host batch -> H2D (copy stream, double-buffered prefetch) -> step -> D2H of the loss, every step ----------
    xh = x.cpu().pin_memory()
    ch = c.cpu().pin_memory() if c is not None else None
    h2d = xh.numel() * 4 + (ch.numel() * 4 if ch is not None else 0)
    copy_stream = torch.cuda.Stream(device=dev)
    xbuf = [torch.empty_like(x) for _ in range(2)]
    cbuf = [torch.empty_like(c) for _ in range(2)] if c is not None else [None, None]
    loss_host = torch.zeros(2, 3).pin_memory()
    ready = [torch.cuda.Event() for _ in range(2)]      # H2D of slot i finished
    consumed = [torch.cuda.Event() for _ in range(2)]   # the step that read slot i finished
    loss_done = [torch.cuda.Event() for _ in range(2)]

    def prefetch(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            xbuf[slot].copy_(xh, non_blocking=True)
            if ch is not None:
                cbuf[slot].copy_(ch, non_blocking=True)
            ready[slot].record(copy_stream)

    def run_e2e(steps):
        """Every step: its own H2D copy of the batch (prefetched one step ahead on the copy stream) and a D2H read of its loss
        (read on the host one step later, so the host never drains the GPU queue).  All copies are inside the timed region."""
        cur = torch.cuda.current_stream()
        for sl in range(2):
            consumed[sl].record(cur)
        prefetch(0)
        for i in range(steps):
            sl = i & 1
            if i + 1 < steps:
                prefetch(sl ^ 1)
            cur.wait_event(ready[sl])
            l3 = trainer.step(xbuf[sl], cbuf[sl])
            consumed[sl].record(cur)
            loss_host[sl].copy_(l3, non_blocking=True)
            loss_done[sl].record(cur)
            if i > 0:
                loss_done[sl ^ 1].synchronize()
                float(loss_host[sl ^ 1][0])
        loss_done[(steps - 1) & 1].synchronize()
        float(loss_host[(steps - 1) & 1][0])

    run_e2e(2)
    ms_e2e = timed(lambda: run_e2e(args.steps), 1) / args.steps
    e2e = {"value": B * world / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12,
           "ms_per_step": ms_e2e, "how": "FusedTrainStep.step on a batch copied from pinned host memory every step (copy stream, one "
                                         "step of prefetch) + D2H of the 3-float loss every step"}

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

    # ---- the reference-surface path (nn.Module forward + torch autograd + clamp_ + torch Adam), for comparison ------------------
    opt_t = torch.optim.Adam(params, fused=True, **ADAM)
    reducer = BucketedGradAllReduce(model)

    def autograd_step():
        opt_t.zero_grad(set_to_none=True)
        z, J = model(x + 0.01 * torch.randn_like(x), c)
        loss = nll_loss(z, J)
        loss.backward()
        reducer.finish()
        for p in params:
            p.grad.clamp_(-5.0, 5.0)
        opt_t.step()

    autograd_step()
    ms_ag = timed_avg(autograd_step, max(2, args.steps // 2))
    reducer.remove()
    extra["autograd_path_train_samples_per_s"] = B * world / (ms_ag * 1e-3)
    extra["autograd_path_note"] = "same step through HierarchicalAffineCouplingBlock.forward + loss.backward() + clamp_ + torch.optim.Adam(fused)"
    for p in params:
        p.grad = torch.zeros_like(p)   # the fused trainer writes gradients in place

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
    if tf32_peak is None:
        tf32_peak = {"burst_tflops": peaks["bf16_tflops"] / 2.0, "sustained_tflops": peaks["bf16_tflops_sustained"] / 2.0,
                     "how": "bf16 figures of MEASURED_PEAKS.json / 2"}
    peak_alone = tf32_peak["burst_tflops"]          # a kernel timed alone
    peak_step = tf32_peak["sustained_tflops"]       # work timed inside a long step
    Fb = blk.plan.flops_per_sample
    achieved = 2 * Fb * B / (ms_b * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(name, {}).get("bwd_dram_bytes_per_launch")
    mma_sync_peak = 270.0   # profiles/ubench5_r01_mma_sync_tf32.txt: what mma.sync.m16n8k8 tf32 sustains on this chip
    kname = bwd_kernel_name(blk.plan, args.mode)
    fp32_mode = args.mode == "fp32"
    roofline = {"bound": "tensor", "kernel": f"{kname} (one block, B={B})", "achieved": achieved,
                "peak": peak_alone, "unit": "TFLOP/s", "frac": achieved / peak_alone, "traffic": traffic,
                "peak_source": "TF32 dense, measured on this box in this run (burst: the kernel is timed alone): " + tf32_peak["how"],
                "tf32_peak_measured": tf32_peak,
                "flops_per_launch": 2 * Fb * B, "ms_per_launch": ms_b,
                "algorithmic_bytes_per_launch": 4 * (3 * w["d"] + 1 + w["dc"]) * B,
                "frac_of_mma_sync_tf32_peak": achieved / mma_sync_peak, "mma_sync_tf32_peak_tflops": mma_sync_peak,
                "phases": {
                    "train_step": {"tflops": 3 * F * value / world / 1e12, "frac": 3 * F * value / world / 1e12 / peak_step},
                    "fwd_logdet": {"tflops": F * extra["fwd_logdet_samples_per_s"] / world / 1e12,
                                   "frac": F * extra["fwd_logdet_samples_per_s"] / world / 1e12 / peak_step},
                    "inverse": {"tflops": F * extra["inverse_samples_per_s"] / world / 1e12,
                                "frac": F * extra["inverse_samples_per_s"] / world / 1e12 / peak_step},
                    "peak": peak_step, "peak_kind": "TF32 sustained (phases are timed inside long steps), per GPU"},
                "fwd_kernel": {"achieved": Fb * B / (ms_f * 1e-3) / 1e12, "ms_per_launch": ms_f, "frac": Fb * B / (ms_f * 1e-3) / 1e12 / peak_alone},
                "hbm_gbs_fwd_streaming": (2 * w["d"] + 1 + w["dc"]) * 4 * B / (ms_f * 1e-3) / 1e9,
                "hbm_peak_gbs": peaks.get("hbm_gbs"),
                "note": ("launch = pack + fused kernel (+ partial-gradient reduce), timed with CUDA events on the launch stream"
                         + ("; fp32 mode runs on CUDA cores: the tensor peak is not its bound" if fp32_mode else ""))}

    line = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.mode,
            "data": "synthetic", "config": workload_config(name, w, B, world), "roofline": roofline, "e2e": e2e,
            "gpu_launches": gpu_launches, "gpu_launches_how": "hint_launch_count() (C ABI) after - before the timed region: every kernel "
            "this library launched; the step launches no torch kernels" + (" (NCCL all-reduces excluded)" if world > 1 else ""),
            "clocks": clocks, "extra": extra, "final_loss": final_loss}

    # ---- the other arithmetic modes on the same workload, and every BASELINE workload at its stated batch (1 GPU only) ----------
    if world == 1 and not args.quick:
        del model, params, opt, trainer, opt_t
        modes = {}
        for md in ("fp32", "tf32x3"):
            if md == args.mode:
                continue
            try:
                ph, keep = phases(w, B, md, 3, 3)
                del keep
                modes[md] = {k: ph[k] for k in ("train_samples_per_s", "train_ms", "fwd_logdet_samples_per_s", "inverse_samples_per_s")}
            except Exception as e:   # a mode outside a kernel family's envelope is reported, not fatal
                modes[md] = {"error": str(e)[:200]}
        line["modes"] = modes
        # the configs' full x-lane: the same blocks with the inter-block HouseholderPerm mixing between them (SURVEY.md 8f-2;
        # uci configs: fixed reflections, plus_shape hint_4_3: trainable) - kernels of this library (householder.cu), not cuBLAS
        hh = {}
        for kind in ("fixed", "trainable"):
            try:
                ph, keep = phases(w, B, args.mode, 3, 3, householder=kind)
                del keep
                hh[kind] = {k: ph[k] for k in ("train_samples_per_s", "train_ms", "launches_per_train_step", "fwd_logdet_samples_per_s",
                                               "inverse_samples_per_s")}
            except Exception as e:
                hh[kind] = {"error": str(e)[:200]}
        line["with_householder"] = hh
        cfgs = []
        for wn, Bw in CONFIG_SWEEP:
            if wn == name and Bw == B:
                continue
            wl = WORKLOADS[wn]
            try:
                ph, keep = phases(wl, Bw, args.mode, 3, 3)
                plan0 = keep[0].blocks[0].plan
                entry = {"workload": wn, "batch": Bw, "bwd_kernel": bwd_kernel_name(plan0, args.mode)}
                del keep
                entry.update(ph)
                entry["train_tflops_algorithmic"] = 3 * ph["flops_per_sample_fwd"] * ph["train_samples_per_s"] / 1e12
                entry["train_frac_of_tf32_peak"] = entry["train_tflops_algorithmic"] / peak_step
                entry["fwd_frac_of_tf32_peak"] = ph["flops_per_sample_fwd"] * ph["fwd_logdet_samples_per_s"] / 1e12 / peak_step
                # the same model with the config's own inter-block HouseholderPerm (plus_shape hint_4_3: trainable reflections,
                # configs/plus_shape/unconditional_hint_4_3.py:60-71; uci / lens hint_8: fixed)
                kind = "trainable" if wn.startswith("plus") else "fixed"
                ph, keep = phases(wl, Bw, args.mode, 3, 3, householder=kind)
                del keep
                entry["with_householder"] = {"kind": kind, "train_samples_per_s": ph["train_samples_per_s"], "train_ms": ph["train_ms"],
                                             "launches_per_train_step": ph["launches_per_train_step"]}
                cfgs.append(entry)
            except Exception as e:
                cfgs.append({"workload": wn, "batch": Bw, "error": str(e)[:200]})
        line["configs"] = cfgs
        # the 2-lane conditional model of configs/lens_shape/conditional_hint_8_full.py:61-102 through the FrEIA shim exactly as
        # train_conditional.py:119-156 drives it (autograd, both lanes, fused couplings + Householder + HINT kernels): SURVEY.md 8f-4
        try:
            line["conditional_2lane"] = conditional_2lane(torch, dev, launches, 10000)
        except Exception as e:
            line["conditional_2lane"] = {"error": str(e)[:200]}
        hint_b200.set_precision(args.mode)
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



def conditional_2lane(torch, dev, launches, B, ndim_x=20, ndim_y=2, n_blocks=8, h=68):
    """fwd+bwd of the lens `conditional_hint_8_full` architecture (x lane: HouseholderPerm, HINT block, ExternalAffineCoupling;
    y lane: HouseholderPerm, AffineCoupling) built through the FrEIA shim; ms per step and library launches per step."""
    from FrEIA.framework import InputNode, Node, OutputNode, ReversibleGraphNet
    from FrEIA.modules import (HierarchicalAffineCouplingBlock, HouseholderPerm, AffineCoupling, ExternalAffineCoupling,
                               F_fully_connected)
    y_lane, x_lane = [InputNode(ndim_y, name="y")], [InputNode(ndim_x, name="x")]
    for i in range(n_blocks):
        if i > 0:
            y_lane.append(Node(y_lane[-1], HouseholderPerm, {"fixed": True, "n_reflections": ndim_y}, name=f"perm_y_{i}"))
            x_lane.append(Node(x_lane[-1], HouseholderPerm, {"fixed": True, "n_reflections": ndim_x}, name=f"perm_x_{i}"))
        x_lane.append(Node(x_lane[-1], HierarchicalAffineCouplingBlock, {"c_internal": [h, h // 2, h // 4, h // 4]}, name=f"hac_x_{i+1}"))
        x_lane.append(Node(x_lane[-1], ExternalAffineCoupling, {"F_class": F_fully_connected, "F_args": {"internal_size": h}},
                           conditions=y_lane[-1], name=f"ac_y_to_x_{i+1}"))
        y_lane.append(Node(y_lane[-1], AffineCoupling, {"F_class": F_fully_connected, "F_args": {"internal_size": h // 4}}, name=f"ac_y_{i+1}"))
    y_lane.append(OutputNode(y_lane[-1], name="z_y"))
    x_lane.append(OutputNode(x_lane[-1], name="z_x"))
    m = ReversibleGraphNet(y_lane + x_lane, verbose=False).to(dev)
    for p in m.parameters():
        if p.requires_grad:
            p.data = 0.005 * torch.randn_like(p)
    x, y = torch.randn(B, ndim_x, device=dev), torch.randn(B, ndim_y, device=dev)

    def step():
        m.zero_grad(set_to_none=True)
        z_y, z_x = m([y, x])
        J = m.log_jacobian(run_forward=False)
        (0.5 * (z_y.pow(2).sum(1) + z_x.pow(2).sum(1)).mean() - J.mean()).backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    n0 = launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    return {"workload": "lens conditional_hint_8_full (2 lanes, 8 blocks), forward + backward through the FrEIA shim", "batch": B,
            "step_ms": ms, "samples_per_s": B / (ms * 1e-3), "library_launches_per_step": (launches() - n0) // 10,
            "note": "host-bound: Python graph runtime + autograd; the couplings, mixings and HINT blocks are library kernels"}

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="d43_hint_8", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="samples per GPU (default: the workload's)")
    ap.add_argument("--mode", default=os.environ.get("HINT_B200_MODE", "tf32"), choices=["fp32", "tf32", "tf32x3", "tf32_mma", "tf32_tcgen05", "tf32_chain", "tf32_tc3"])
    ap.add_argument("--cpu-sample", type=int, default=32768, help="samples per CPU step (bounded sample of the workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-peak", action="store_true", help="do not measure the TF32 peak with cuBLAS in this run (profiler launch lists); "
                                                             "the roofline then uses MEASURED_PEAKS.json's bf16 figures / 2")
    ap.add_argument("--quick", action="store_true", help="skip the extra modes and the sweep over the other BASELINE workloads")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w, args.workload)
    else:
        run_ours(args, w, args.workload)


if __name__ == "__main__":
    main()
