"""Generate golden vectors by running the REAL reference module.  TEST INFRASTRUCTURE ONLY.

Run in the build container (where /root/reference exists):

    python oracle/gen_golden.py

It imports ``/root/reference/hint.py`` unmodified.  ``hint.py:6`` imports
``FrEIA.modules.orthogonal.HouseholderPerm`` at module level; FrEIA is absent, so a
placeholder package is injected into ``sys.modules`` (the class is only instantiated when
``reshuffle=True``, hint.py:36-39, which no fixture uses).  No reference source is copied: only
numeric inputs/outputs are stored, as ``tests/golden/<case>.npz``:

    x, c (opt), params (flat, ``parameters()`` order)           inputs
    z, J            forward(x)                     fp32 (+ z64, J64: same module cast to fp64)
    xinv, Jinv      forward(x, rev=True)           fp32 (+ fp64)
    dx, dc, dparams grads of 0.5*sum(z^2,1).mean() - J.mean()  (train_unconditional.py:128-132), fp32 + fp64
    meta            json: ctor kwargs, state_dict names/shapes in order

The GPU box has no /root/reference; tests only read the .npz files.
"""
import json
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def import_reference_hint():
    freia = types.ModuleType("FrEIA")
    modules = types.ModuleType("FrEIA.modules")
    orth = types.ModuleType("FrEIA.modules.orthogonal")

    class HouseholderPerm:  # placeholder; never instantiated by the fixtures
        def __init__(self, *a, **k):
            raise NotImplementedError("FrEIA is not available; reshuffle=True is unpinned")

    orth.HouseholderPerm = HouseholderPerm
    freia.modules = modules
    modules.orthogonal = orth
    sys.modules.setdefault("FrEIA", freia)
    sys.modules.setdefault("FrEIA.modules", modules)
    sys.modules.setdefault("FrEIA.modules.orthogonal", orth)
    sys.path.insert(0, REF)
    import hint  # noqa: E402  (the reference's hint.py)
    assert os.path.abspath(hint.__file__).startswith(REF), hint.__file__
    return hint


CASES = [
    # name, d, dims_c, kwargs, B, weight recipe
    dict(name="power_like_d6", d=6, dims_c=[], kw=dict(c_internal=[20, 10, 5, 2]), B=37, init="default"),
    dict(name="gas_like_d8", d=8, dims_c=[], kw=dict(c_internal=[16, 8, 4, 2]), B=64, init="default"),
    dict(name="d43_hint8_widths", d=43, dims_c=[], kw=dict(c_internal=[67, 33, 16, 8]), B=130, init="default"),
    dict(name="d42_hint8_widths_small_init", d=42, dims_c=[], kw=dict(c_internal=[67, 33, 16, 8]), B=65, init="randn0.005"),
    dict(name="lens_xlane_d20", d=20, dims_c=[], kw=dict(c_internal=[68, 34, 17, 17]), B=100, init="default"),
    dict(name="plus_full_d100_narrow", d=100, dims_c=[], kw=dict(c_internal=[24, 12, 6, 3, 3]), B=50, init="default"),
    dict(name="plus_ms3_d100_narrow", d=100, dims_c=[], kw=dict(c_internal=[32, 16, 8, 4], max_splits=3), B=33, init="default"),
    dict(name="plus_ms0_d100", d=100, dims_c=[], kw=dict(c_internal=[16, 8], max_splits=0), B=40, init="default"),
    dict(name="lens_concat_cond_d20_dc2", d=20, dims_c=[(2,)], kw=dict(c_internal=[17, 8, 4]), B=70, init="default"),
    dict(name="plus_concat_cond_d100_dc4", d=100, dims_c=[(4,)], kw=dict(c_internal=[20, 10, 5]), B=48, init="default"),
    dict(name="two_conditions_d10", d=10, dims_c=[(2,), (3,)], kw=dict(c_internal=[12, 6]), B=29, init="default"),
    dict(name="default_width_d5", d=5, dims_c=[], kw=dict(), B=31, init="default"),
    dict(name="single_width_d9", d=9, dims_c=[], kw=dict(c_internal=[7]), B=32, init="default"),
    dict(name="min_split3_d13_clamp2", d=13, dims_c=[], kw=dict(c_internal=[10, 5], min_split_size=3, clamp=2.0), B=45, init="default"),
    dict(name="tiny_d2", d=2, dims_c=[], kw=dict(c_internal=[4]), B=17, init="default"),
    dict(name="tiny_d3_B1", d=3, dims_c=[], kw=dict(c_internal=[5, 3]), B=1, init="default"),
    dict(name="wide_h_d12", d=12, dims_c=[], kw=dict(c_internal=[150, 75, 37]), B=129, init="randn0.05"),
]


def run_case(hint, case, seed):
    torch.manual_seed(seed)
    d, dims_c = case["d"], case["dims_c"]
    kw = {k: (list(v) if isinstance(v, list) else v) for k, v in case["kw"].items()}
    blk = hint.HierarchicalAffineCouplingBlock([(d,)], dims_c=dims_c, **kw)
    if case["init"].startswith("randn"):
        scale = float(case["init"][5:])
        for p in blk.parameters():
            p.data = scale * torch.randn_like(p.data)
    B = case["B"]
    x = torch.randn(B, d)
    cs = [torch.randn(B, dc[0]) for dc in dims_c]

    out = {}
    names = [(k, list(v.shape)) for k, v in blk.state_dict().items()]
    assert [k for k, _ in names] == [k for k, _ in blk.named_parameters()]
    out["params"] = torch.cat([p.detach().reshape(-1) for p in blk.parameters()]).numpy()
    out["x"] = x.numpy()
    if cs:
        out["c"] = torch.cat(cs, dim=1).numpy()

    for tag, dt in (("", torch.float32), ("64", torch.float64)):
        m = blk.to(dt)
        xx = x.to(dt).clone().requires_grad_(True)
        cc = [ci.to(dt).clone().requires_grad_(True) for ci in cs]
        for p in m.parameters():
            p.grad = None
        z = m([xx], c=cc, rev=False)[0]
        J = m.jacobian([xx], c=cc, rev=False)
        loss = 0.5 * torch.sum(z ** 2, dim=1).mean() - J.mean()
        loss.backward()
        out["z" + tag] = z.detach().numpy()
        out["J" + tag] = J.detach().numpy()
        out["loss" + tag] = np.asarray(loss.item())
        out["dx" + tag] = xx.grad.numpy()
        if cc:
            out["dc" + tag] = torch.cat([ci.grad for ci in cc], dim=1).numpy()
        out["dparams" + tag] = torch.cat([p.grad.reshape(-1) for p in m.parameters()]).numpy()
        with torch.no_grad():
            xi = m([x.to(dt)], c=[ci.to(dt) for ci in cs], rev=True)[0]
            Ji = m.jacobian(None)
            out["xinv" + tag] = xi.numpy()
            out["Jinv" + tag] = Ji.numpy()
            # round trip through the reference itself: f^-1(f(x))
            xr = m([z.detach()], c=[ci.to(dt) for ci in cs], rev=True)[0]
            out["xrec" + tag] = xr.numpy()
        blk = m
    blk.to(torch.float32)
    meta = dict(name=case["name"], d=d, dims_c=[list(t) for t in dims_c], kwargs=case["kw"], B=B,
                init=case["init"], seed=seed, state_dict=names, torch=torch.__version__,
                source="/root/reference/hint.py HierarchicalAffineCouplingBlock, run on CPU")
    out["meta"] = np.asarray(json.dumps(meta))
    return out


def main():
    hint = import_reference_hint()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)  # deterministic reduction order
    for i, case in enumerate(CASES):
        out = run_case(hint, case, seed=1000 + i)
        path = os.path.join(OUT, case["name"] + ".npz")
        np.savez_compressed(path, **out)
        print(f"{case['name']:36s} params={out['params'].size:7d}  B={case['B']:4d}  "
              f"|z|max={np.abs(out['z']).max():.3f}  J range=[{out['J'].min():.3f},{out['J'].max():.3f}]  "
              f"inv err={np.abs(out['xrec'] - out['x']).max():.2e}  {os.path.getsize(path) // 1024} KiB")


if __name__ == "__main__":
    main()
