"""Does CUDA-graph capture of the library's launches pay off at the configs' own (small) batch sizes?"""
import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import hint_b200
from hint_b200 import HintFlow
hint_b200.set_precision("tf32")
dev = torch.device("cuda:0")
for name, d, ci, ms, B in [("lens_hint_8_full", 20, [68, 34, 17, 17], -1, 10000), ("plus_hint_4_3", 100, [314, 157, 78, 39], 3, 10000),
                           ("plus_hint_4_3", 100, [314, 157, 78, 39], 3, 500), ("miniboone", 42, [67, 33, 16, 8], -1, 300)]:
    nb = 4 if d == 100 else 8
    model = HintFlow(d, nb, ci, max_splits=ms).to(dev).init_like_reference_scripts(0.005)
    x = torch.randn(B, d, device=dev)
    with torch.no_grad():
        for _ in range(3):
            z, J = model(x)
        torch.cuda.synchronize()
        def timeit(fn, n=20):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(n): fn()
            torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
        eager = timeit(lambda: model(x))
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2): model(x)
        torch.cuda.current_stream().wait_stream(s)
        try:
            with torch.cuda.graph(g):
                zg, Jg = model(x)
            graph = timeit(lambda: g.replay())
            ok = float((zg - z).abs().max())
        except Exception as e:
            graph, ok = float("nan"), str(e)[:100]
    print(f"{name:18s} B={B:6d}: forward eager {eager:.3f} ms  graph replay {graph:.3f} ms  (max |z_graph - z| {ok})", flush=True)
