"""CPU oracle for HINT's recursive affine coupling block.  TEST INFRASTRUCTURE ONLY.

This file is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it.  Nothing under ``hint_b200/`` imports it, and the product
path raises if its CUDA library is missing (there is no CPU fallback).

It restates, from scratch and in a different shape (flattened pre-order node
table + level-synchronous sweeps instead of Python recursion), the algorithm of
the reference file ``hint.py``:

* tree construction rules ............ hint.py:25-54
* 3-layer ReLU MLP subnets ........... hint.py:10-13
* soft clamp e(s)=exp(clamp*0.636*atan s)  hint.py:56-60   (literal 0.636, not 2/pi)
* forward / inverse order, log-det ... hint.py:62-101
* block wrapper, default clamp 4.0 ... hint.py:108-129
* NLL training loss .................. train_unconditional.py:128-132

Parity status: PINNED.  ``oracle/gen_golden.py`` imports the real
``/root/reference/hint.py`` (with a stub for the absent FrEIA import) and stores
its outputs under ``tests/golden/``; ``tests/test_oracle.py`` checks this oracle
against every one of those fixtures and against the parameter-count
known-answers written in the reference configs.
``reshuffle=True`` (FrEIA HouseholderPerm, source absent) is NOT covered:
parity unpinned for that option, and the product rejects it.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch

SOFT_CLAMP_CONST = 0.636  # hint.py:57,60


@dataclass
class Node:
    idx: int            # pre-order index (root = 0, then the upper subtree, then the lower subtree)
    depth: int
    lo: int             # first column of the block input owned by this node
    hi: int             # one past the last column
    k: int              # split index: upper = [lo, lo+k), lower = [lo+k, hi)      hint.py:41
    cin: int            # subnet input width  = k + dc                               hint.py:44
    h: int              # subnet hidden width = c_internal[min(depth, len-1)]        hint.py:31-34,50
    cout: int           # subnet output width = (hi-lo) - k                          hint.py:44
    leaf: bool
    parent: int = -1
    upper: int = -1
    lower: int = -1
    path: str = "tree"  # state_dict prefix, e.g. "tree.upper.lower"


def build_plan(d: int, dc: int = 0, c_internal: Sequence[int] = (), max_splits: int = -1,
               min_split_size: int = 2) -> List[Node]:
    """Flatten the recursion of hint.py:25-54 into a pre-order node table.

    Width rule (hint.py:29-34,50-52): empty list -> [d_root]; the list is consumed one entry per
    level and its last entry repeats forever.  A node has children iff
    ``width >= 2*min_split_size and max_splits != 0`` (hint.py:47); children get ``max_splits-1``.
    """
    widths = list(c_internal) if len(c_internal) > 0 else [d]
    nodes: List[Node] = []

    def rec(lo: int, hi: int, depth: int, splits_left: int, parent: int, path: str) -> int:
        w = hi - lo
        k = w // 2
        idx = len(nodes)
        internal = (w >= 2 * min_split_size) and (splits_left != 0)
        n = Node(idx=idx, depth=depth, lo=lo, hi=hi, k=k, cin=k + dc,
                 h=int(widths[min(depth, len(widths) - 1)]), cout=w - k,
                 leaf=not internal, parent=parent, path=path)
        nodes.append(n)
        if internal:
            n.upper = rec(lo, lo + k, depth + 1, splits_left - 1, idx, path + ".upper")
            n.lower = rec(lo + k, hi, depth + 1, splits_left - 1, idx, path + ".lower")
        return idx

    rec(0, d, 0, max_splits, -1, "tree")
    return nodes


@dataclass
class ParamEntry:
    name: str
    node: int
    net: str       # "s" or "t"
    layer: int     # 0, 1, 2  (state_dict indices 0, 2, 4)
    kind: str      # "weight" or "bias"
    shape: Tuple[int, ...]
    offset: int


def param_entries(plan: List[Node]) -> List[ParamEntry]:
    """Canonical flat parameter order == ``parameters()`` order of the reference module:
    pre-order over nodes; per node s.{0,2,4}.{weight,bias} then t.{...} (registration order in
    hint.py:44-45,49-52; nn.Linear weight is [out, in])."""
    out: List[ParamEntry] = []
    off = 0
    for n in plan:
        dims = [(n.h, n.cin), (n.h, n.h), (n.cout, n.h)]
        for net in ("s", "t"):
            for layer, (o, i) in enumerate(dims):
                for kind, shape in (("weight", (o, i)), ("bias", (o,))):
                    numel = o * i if kind == "weight" else o
                    out.append(ParamEntry(f"{n.path}.{net}.{2 * layer}.{kind}", n.idx, net, layer,
                                          kind, shape, off))
                    off += numel
    return out


def param_count(plan: List[Node]) -> int:
    return sum(2 * (n.h * n.cin + n.h + n.h * n.h + n.h + n.cout * n.h + n.cout) for n in plan)


def flops_per_sample(plan: List[Node]) -> int:
    """Algorithmic forward FLOPs per sample per block: true dims, GEMMs only."""
    return sum(2 * 2 * (n.cin * n.h + n.h * n.h + n.h * n.cout) for n in plan)


def _views(plan: List[Node], flat: torch.Tensor):
    """Per node: {(net, layer): (W, b)} as views into the flat parameter vector."""
    table = [dict() for _ in plan]
    ws = {}
    for e in param_entries(plan):
        numel = 1
        for s in e.shape:
            numel *= s
        v = flat[e.offset:e.offset + numel].view(*e.shape)
        ws[(e.node, e.net, e.layer, e.kind)] = v
    for n in plan:
        for net in ("s", "t"):
            for layer in range(3):
                table[n.idx][(net, layer)] = (ws[(n.idx, net, layer, "weight")],
                                              ws[(n.idx, net, layer, "bias")])
    return table


def _mlp(a: torch.Tensor, tab, net: str, keep: bool = False):
    """W3 relu(W2 relu(W1 a + b1) + b2) + b3   (hint.py:10-13)."""
    W1, b1 = tab[(net, 0)]
    W2, b2 = tab[(net, 1)]
    W3, b3 = tab[(net, 2)]
    h1 = torch.relu(torch.addmm(b1, a, W1.t()))
    h2 = torch.relu(torch.addmm(b2, h1, W2.t()))
    o = torch.addmm(b3, h2, W3.t())
    return (o, h1, h2) if keep else o


def _levels(plan: List[Node]) -> List[List[Node]]:
    depth = max(n.depth for n in plan)
    lv: List[List[Node]] = [[] for _ in range(depth + 1)]
    for n in plan:
        lv[n.depth].append(n)
    return lv


def forward(plan: List[Node], flat: torch.Tensor, x: torch.Tensor, c: Optional[torch.Tensor] = None,
            rev: bool = False, clamp: float = 4.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Transport + log|det J| of one block (hint.py:62-101), level-synchronous.

    forward: deepest level first, root last (children before own coupling, hint.py:70-73);
    inverse: root first (hint.py:85-88).  Every node owns the contiguous columns [lo, hi).
    Works under autograd (no in-place writes on graph tensors).
    """
    tab = _views(plan, flat)
    alpha = clamp * SOFT_CLAMP_CONST
    cols = [x[:, j:j + 1] for j in range(x.shape[1])]   # column list -> functional updates
    J = torch.zeros(x.shape[0], dtype=x.dtype, device=x.device)
    levels = _levels(plan)
    order = levels if rev else levels[::-1]
    for level in order:
        for n in level:
            xu = torch.cat(cols[n.lo:n.lo + n.k], dim=1)
            xl = torch.cat(cols[n.lo + n.k:n.hi], dim=1)
            a = torch.cat([xu, c], dim=1) if c is not None and c.shape[1] > 0 else xu
            s = _mlp(a, tab[n.idx], "s")
            t = _mlp(a, tab[n.idx], "t")
            la = alpha * torch.atan(s)
            if not rev:
                xl = torch.exp(la) * xl + t
                J = J + la.sum(dim=1)
            else:
                xl = (xl - t) / torch.exp(la)
                J = J - la.sum(dim=1)
            for j in range(n.cout):
                cols[n.lo + n.k + j] = xl[:, j:j + 1]
    return torch.cat(cols, dim=1), J


def forward_fast(plan: List[Node], flat: torch.Tensor, x: torch.Tensor, c: Optional[torch.Tensor] = None,
                 rev: bool = False, clamp: float = 4.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Same arithmetic as :func:`forward`, written with in-place slice updates (no autograd).
    Used for CPU-baseline timing, where the column-list bookkeeping above would dominate."""
    tab = _views(plan, flat)
    alpha = clamp * SOFT_CLAMP_CONST
    X = x.clone()
    J = torch.zeros(x.shape[0], dtype=x.dtype, device=x.device)
    levels = _levels(plan)
    with torch.no_grad():
        for level in (levels if rev else levels[::-1]):
            for n in level:
                xu = X[:, n.lo:n.lo + n.k]
                a = torch.cat([xu, c], dim=1) if c is not None and c.shape[1] > 0 else xu
                la = alpha * torch.atan(_mlp(a, tab[n.idx], "s"))
                t = _mlp(a, tab[n.idx], "t")
                xl = X[:, n.lo + n.k:n.hi]
                if not rev:
                    X[:, n.lo + n.k:n.hi] = torch.exp(la) * xl + t
                    J += la.sum(dim=1)
                else:
                    X[:, n.lo + n.k:n.hi] = (xl - t) / torch.exp(la)
                    J -= la.sum(dim=1)
    return X, J


def forward_blockwise(plan: List[Node], flat: torch.Tensor, x: torch.Tensor, c: Optional[torch.Tensor] = None,
                      clamp: float = 4.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Forward direction with the reference's own op mix per node (views for the split, 6 addmm, one cat of
    the node's columns - hint.py:68-90), autograd-friendly.  This is the CPU "port" that bench.py times as the
    reference arm / cpu_baseline: same GEMM shapes, same elementwise passes and the same per-node concat copies."""
    tab = _views(plan, flat)
    alpha = clamp * SOFT_CLAMP_CONST
    has_c = c is not None and c.shape[1] > 0
    out = {}
    Js = {}
    for level in _levels(plan)[::-1]:
        for n in level:
            if n.leaf:
                xu, xl = x[:, n.lo:n.lo + n.k], x[:, n.lo + n.k:n.hi]
                Jc = None
            else:
                xu, xl = out.pop(n.upper), out.pop(n.lower)
                Jc = Js.pop(n.upper) + Js.pop(n.lower)
            a = torch.cat([xu, c], dim=1) if has_c else xu
            s = _mlp(a, tab[n.idx], "s")
            t = _mlp(a, tab[n.idx], "t")
            xl = torch.exp(alpha * torch.atan(s)) * xl + t
            Jn = torch.sum(alpha * torch.atan(s), dim=1)   # the reference evaluates atan twice (hint.py:57,60)
            out[n.idx] = torch.cat([xu, xl], dim=1)
            Js[n.idx] = Jn if Jc is None else Jc + Jn
    return out[0], Js[0]


def inverse_blockwise(plan: List[Node], flat: torch.Tensor, z: torch.Tensor, c: Optional[torch.Tensor] = None,
                      clamp: float = 4.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """rev=True direction with the reference's op mix (root coupling first, hint.py:79-94).  CPU timing port."""
    tab = _views(plan, flat)
    alpha = clamp * SOFT_CLAMP_CONST
    has_c = c is not None and c.shape[1] > 0
    inp = {0: z}
    done = {}
    J = torch.zeros(z.shape[0], dtype=z.dtype, device=z.device)
    for level in _levels(plan):
        for n in level:
            xin = inp.pop(n.idx)
            xu, xl = xin[:, :n.k], xin[:, n.k:]
            a = torch.cat([xu, c], dim=1) if has_c else xu
            s = _mlp(a, tab[n.idx], "s")
            t = _mlp(a, tab[n.idx], "t")
            xl = (xl - t) / torch.exp(alpha * torch.atan(s))
            J = J - torch.sum(alpha * torch.atan(s), dim=1)
            if n.leaf:
                done[n.idx] = torch.cat([xu, xl], dim=1)
            else:
                inp[n.upper], inp[n.lower] = xu, xl
    for level in _levels(plan)[::-1]:
        for n in level:
            if not n.leaf:
                done[n.idx] = torch.cat([done.pop(n.upper), done.pop(n.lower)], dim=1)
    return done[0], J


def backward_from_output(plan: List[Node], flat: torch.Tensor, z: torch.Tensor, c: Optional[torch.Tensor],
                         dz: torch.Tensor, dJ: torch.Tensor, clamp: float = 4.0):
    """Memory-free backward of the *forward* direction, restated by hand (no autograd).

    Sweeps root-first (the inverse order), at every node recomputing s,t from the already
    available x_upper', recovering x_lower' = (z_lower - t)/e and emitting gradients
    (formulas: SURVEY.md section 8a, derived from hint.py:76-80):
        dt = dz_l ; dg = dz_l*x_l'*e + dJ ; ds = dg*alpha/(1+s^2) ; dx_l' = dz_l*e
        MLP backward -> da ; dx_u' = dz_u + da[:k] ; dc += da[k:]
    Returns (x_reconstructed, dx, dc, dflat).  This is the algorithm the CUDA backward follows.
    """
    tab = _views(plan, flat)
    dflat = torch.zeros_like(flat)
    dtab = _views(plan, dflat)
    alpha = clamp * SOFT_CLAMP_CONST
    Z = z.clone()
    D = dz.clone()
    has_c = c is not None and c.shape[1] > 0
    dC = torch.zeros_like(c) if has_c else None
    for level in _levels(plan):
        for n in level:
            zu = Z[:, n.lo:n.lo + n.k]
            a = torch.cat([zu, c], dim=1) if has_c else zu
            s, h1s, h2s = _mlp(a, tab[n.idx], "s", keep=True)
            t, h1t, h2t = _mlp(a, tab[n.idx], "t", keep=True)
            la = alpha * torch.atan(s)
            e = torch.exp(la)
            zl = Z[:, n.lo + n.k:n.hi]
            dzl = D[:, n.lo + n.k:n.hi]
            xl = (zl - t) / e
            d_t = dzl
            d_s = (dzl * xl * e + dJ[:, None]) * alpha / (1.0 + s * s)
            da = torch.zeros_like(a)
            for net, dout, h1, h2 in (("s", d_s, h1s, h2s), ("t", d_t, h1t, h2t)):
                W1, _ = tab[n.idx][(net, 0)]
                W2, _ = tab[n.idx][(net, 1)]
                W3, _ = tab[n.idx][(net, 2)]
                dW1, db1 = dtab[n.idx][(net, 0)]
                dW2, db2 = dtab[n.idx][(net, 1)]
                dW3, db3 = dtab[n.idx][(net, 2)]
                dW3 += dout.t() @ h2
                db3 += dout.sum(0)
                dh2 = (dout @ W3) * (h2 > 0).to(h2.dtype)
                dW2 += dh2.t() @ h1
                db2 += dh2.sum(0)
                dh1 = (dh2 @ W2) * (h1 > 0).to(h1.dtype)
                dW1 += dh1.t() @ a
                db1 += dh1.sum(0)
                da += dh1 @ W1
            Z[:, n.lo + n.k:n.hi] = xl
            D[:, n.lo + n.k:n.hi] = dzl * e
            D[:, n.lo:n.lo + n.k] = D[:, n.lo:n.lo + n.k] + da[:, :n.k]
            if has_c:
                dC += da[:, n.k:]
    return Z, D, dC, dflat


def nll_loss(z: torch.Tensor, J: torch.Tensor) -> torch.Tensor:
    """0.5*sum(z^2,1).mean() - J.mean()   (train_unconditional.py:128-132)."""
    return 0.5 * torch.sum(z ** 2, dim=1).mean() - J.mean()


# Known-answer table: parameter budgets written as comments in the reference configs,
# reproduced from the construction rules.  (d, dc, c_internal, max_splits, n_blocks, total)
def _ci(h, n):
    return [h // (2 ** i) for i in range(n)]


PARAM_COUNT_KATS = [
    # configs/plus_shape/unconditional_hint_4_3.py:31,69-70   "2M"
    dict(name="plus_hint_4_3", d=100, dc=0, c_internal=_ci(314, 4), max_splits=3, n_blocks=4, per_block=491812, total=1967248),
    # configs/plus_shape/unconditional_hint_1_full.py:31      "200k"
    dict(name="plus_hint_1_full", d=100, dc=0, c_internal=_ci(110, 3), max_splits=-1, n_blocks=1, per_block=199788, total=199788),
    # configs/plus_shape/unconditional_hint_4_full.py         "2M"
    dict(name="plus_hint_4_full", d=100, dc=0, c_internal=_ci(263, 4) + [263 // 8], max_splits=-1, n_blocks=4, per_block=495866, total=1983464),
    # configs/plus_shape/unconditional_hint_8_full.py         "2M"
    dict(name="plus_hint_8_full", d=100, dc=0, c_internal=_ci(176, 4) + [176 // 8], max_splits=-1, n_blocks=8, per_block=254272, total=2034176),
    # configs/uci_data/power_hint_8.py:30,66                  "500k"
    dict(name="power_hint_8", d=6, dc=0, c_internal=_ci(140, 4), max_splits=-1, n_blocks=8, per_block=62454, total=499632),
    dict(name="power_hint_4", d=6, dc=0, c_internal=_ci(200, 4), max_splits=-1, n_blocks=4, per_block=125214, total=500856),
    # configs/uci_data/gas_hint_8.py                          "500k"
    dict(name="gas_hint_8", d=8, dc=0, c_internal=_ci(128, 4), max_splits=-1, n_blocks=8, per_block=62488, total=499904),
    dict(name="gas_hint_4", d=8, dc=0, c_internal=_ci(184, 4), max_splits=-1, n_blocks=4, per_block=125880, total=503520),
    # configs/uci_data/miniboone_hint_8.py:31,67              "250k"
    dict(name="miniboone_hint_8", d=42, dc=0, c_internal=_ci(67, 4), max_splits=-1, n_blocks=8, per_block=31328, total=250624),
    dict(name="miniboone_hint_4", d=42, dc=0, c_internal=_ci(102, 4), max_splits=-1, n_blocks=4, per_block=62690, total=250760),
    # configs/lens_shape/unconditional_hint_1_full.py         "100k"
    dict(name="lens_hint_1_full", d=20, dc=0, c_internal=_ci(139, 3), max_splits=-1, n_blocks=1, per_block=99298, total=99298),
    # configs/plus_shape/conditional_recursive_cinn_4.py:31,68   "4M"
    dict(name="plus_cond_recursive_4", d=100, dc=4, c_internal=_ci(267, 3), max_splits=-1, n_blocks=4, per_block=1001570, total=4006280),
    # lens x-lane of configs/lens_shape/conditional_hint_8_full.py:74
    dict(name="lens_xlane_hint_8_full", d=20, dc=0, c_internal=_ci(68, 3) + [68 // 4], max_splits=-1, n_blocks=8, per_block=27696, total=221568),
    # BASELINE.json d=43 variant with miniboone hint_8 widths
    dict(name="d43_hint_8", d=43, dc=0, c_internal=_ci(67, 4), max_splits=-1, n_blocks=8, per_block=31598, total=252784),
]
