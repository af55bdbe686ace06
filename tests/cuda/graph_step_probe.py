"""Would replaying the fused train step as ONE CUDA graph help at the configs' own batch sizes?  (Probe only: the captured step
re-uses the captured noise offset and Adam step count.)  python graph_step_probe.py"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import hint_b200
from hint_b200 import HintFlow, FusedClampAdam, FusedTrainStep
hint_b200.set_precision("tf32")
dev = torch.device("cuda:0")
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, d, nb, ci, ms, B in (("lens", 20, 8, [68, 34, 17, 17], -1, 10000), ("plus43", 100, 4, [314, 157, 78, 39], 3, 500), ("plus43", 100, 4, [314, 157, 78, 39], 3, 10000),
                               ("power", 6, 8, [140, 70, 35, 17], -1, 65536)):
    model = HintFlow(d, nb, ci, max_splits=ms).to(dev).init_like_reference_scripts(0.005)
    opt = FusedClampAdam(list(model.parameters()), grad_clamp=5.0, lr=0.01, betas=(0.9, 0.95), eps=1e-4, weight_decay=1.86e-5)
    tr = FusedTrainStep(model, opt, noise=0.01, seed=0)
    x = torch.randn(B, d, device=dev)
    eager = t(lambda: tr.step(x))
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3): tr.step(x)
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = tr.step(x)
        graphed = t(lambda: g.replay())
        print(f"{name} B={B}: eager {eager:.3f} ms, graph replay {graphed:.3f} ms", flush=True)
    except Exception as e:
        print(f"{name} B={B}: eager {eager:.3f} ms, capture failed: {str(e)[:200]}", flush=True)
