"""Golden vectors for ``reshuffle=True`` (hint.py:36-39,64-65,93-94) from the REAL reference module.  TEST INFRASTRUCTURE ONLY.

    python oracle/gen_golden_reshuffle.py        (build container: /root/reference must exist)

hint.py instantiates ``FrEIA.modules.orthogonal.HouseholderPerm([shape], dims_c=..., n_reflections=d_node, fixed=True)`` per tree
node and calls ``perm([x])[0]`` / ``perm([x], rev=True)[0]``.  FrEIA is not part of the reference, so the stand-in injected here
follows the PUBLISHED definition (W = prod_i (I - 2 v_i v_i^T / |v_i|^2), forward x W, reverse x W^T) - the same one
hint_b200.householder.HouseholderPerm states.  What these vectors pin is therefore the reference's RECURSION (where the mixings
sit relative to the splits, children and couplings), i.e. hint_b200's claim that they compose into one matrix in front of the
un-shuffled tree; the definition of W itself stays parity-unpinned.  Stored per case: x, c, trainable params (flat, parameters()
order without the fixed reflections), the reflections of every node keyed by the reference's state_dict name, z / J / xinv /
Jinv and the NLL gradients in fp64."""
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


class HouseholderPerm(nn.Module):
    def __init__(self, dims_in, dims_c=[], n_reflections=1, fixed=False):
        super().__init__()
        self.width = dims_in[0][0]
        self.Vs = nn.Parameter(torch.randn(n_reflections, self.width), requires_grad=not fixed)

    def matrix(self):
        W = torch.eye(self.width, dtype=self.Vs.dtype)
        for v in self.Vs:
            W = W - 2.0 * torch.outer(W @ v, v) / torch.dot(v, v)
        return W

    def forward(self, x, c=[], rev=False):
        W = self.matrix()
        return [x[0] @ (W.t() if rev else W)]


def import_reference_hint():
    freia, modules, orth = types.ModuleType("FrEIA"), types.ModuleType("FrEIA.modules"), types.ModuleType("FrEIA.modules.orthogonal")
    orth.HouseholderPerm = HouseholderPerm
    freia.modules, modules.orthogonal = modules, orth
    sys.modules["FrEIA"], sys.modules["FrEIA.modules"], sys.modules["FrEIA.modules.orthogonal"] = freia, modules, orth
    sys.path.insert(0, REF)
    import hint
    assert os.path.abspath(hint.__file__).startswith(REF), hint.__file__
    return hint


CASES = [dict(name="reshuffle_d13", d=13, dims_c=[], kw=dict(c_internal=[10, 5], reshuffle=True), B=40),
         dict(name="reshuffle_cond_d10_ms1", d=10, dims_c=[(3,)], kw=dict(c_internal=[12, 6], max_splits=1, reshuffle=True), B=25)]


def main():
    hint = import_reference_hint()
    torch.set_num_threads(1)
    for i, case in enumerate(CASES):
        torch.manual_seed(4000 + i)
        d, dims_c = case["d"], case["dims_c"]
        blk = hint.HierarchicalAffineCouplingBlock([(d,)], dims_c=dims_c, **{k: (list(v) if isinstance(v, list) else v) for k, v in case["kw"].items()}).double()
        B = case["B"]
        x = torch.randn(B, d, dtype=torch.float64, requires_grad=True)
        cs = [torch.randn(B, t[0], dtype=torch.float64, requires_grad=True) for t in dims_c]
        out = {"x": x.detach().numpy()}
        if cs:
            out["c"] = torch.cat([c.detach() for c in cs], dim=1).numpy()
        trainable = [p for p in blk.parameters() if p.requires_grad]
        out["params"] = torch.cat([p.detach().reshape(-1) for p in trainable]).numpy()
        for k, v in blk.state_dict().items():
            if k.endswith("perm.Vs"):
                out["vs:" + k] = v.numpy()
        z = blk([x], c=cs)[0]
        J = blk.jacobian([x], c=cs)
        (0.5 * torch.sum(z ** 2, dim=1).mean() - J.mean()).backward()
        out.update(z64=z.detach().numpy(), J64=J.detach().numpy(), dx64=x.grad.numpy(),
                   dparams64=torch.cat([p.grad.reshape(-1) for p in trainable]).numpy())
        if cs:
            out["dc64"] = torch.cat([c.grad for c in cs], dim=1).numpy()
        with torch.no_grad():
            xi = blk([x.detach()], c=[c.detach() for c in cs], rev=True)[0]
            out.update(xinv64=xi.numpy(), Jinv64=blk.jacobian(None).numpy())
        out["meta"] = np.asarray(json.dumps(dict(name=case["name"], d=d, dims_c=[list(t) for t in dims_c], kwargs=case["kw"], B=B,
                                                 source="/root/reference/hint.py with the published HouseholderPerm definition injected")))
        np.savez_compressed(os.path.join(OUT, case["name"] + ".npz"), **out)
        print(case["name"], out["params"].size, "params;", sum(1 for k in out if k.startswith("vs:")), "perms")


if __name__ == "__main__":
    main()
