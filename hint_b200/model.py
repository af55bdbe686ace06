"""A chain of HINT coupling blocks: the x-lane of the reference configs (e.g. configs/uci_data/miniboone_hint_8.py:55-71),
with the blocks applied back to back or - ``householder='fixed' | 'trainable'`` - with the inter-block HouseholderPerm mixing
of the configs between them (FrEIA code, published definition, parity-unpinned).  Used by bench.py, the DP trainer and the smoke
test; the reference's own scripts build the same chain through the FrEIA graph shim instead."""
import torch
import torch.nn as nn

from .block import HierarchicalAffineCouplingBlock
from .householder import HouseholderPerm


class HintFlow(nn.Module):
    def __init__(self, d, n_blocks, c_internal, dims_c=(), clamp=4.0, max_splits=-1, min_split_size=2, householder=None):
        super().__init__()
        self.d = int(d)
        self.blocks = nn.ModuleList([
            HierarchicalAffineCouplingBlock([(d,)], dims_c=list(dims_c), c_internal=list(c_internal), clamp=clamp,
                                            max_splits=max_splits, min_split_size=min_split_size)
            for _ in range(n_blocks)])
        if householder not in (None, "fixed", "trainable"):
            raise ValueError("householder must be None, 'fixed' or 'trainable'")
        # perms[i] mixes the output of block i before block i + 1 ({'fixed': ..., 'n_reflections': d} in the configs)
        self.perms = nn.ModuleList([HouseholderPerm([(d,)], n_reflections=d, fixed=householder == "fixed")
                                    for _ in range(n_blocks - 1)]) if householder else None

    @property
    def flops_per_sample(self):
        return sum(b.plan.flops_per_sample for b in self.blocks)

    def forward(self, x, c=None, rev=False):
        """-> (z, logdet).  rev=True applies the inverse blocks in reverse order."""
        cs = [] if c is None else [c]
        J = None
        n = len(self.blocks)
        for i in (range(n - 1, -1, -1) if rev else range(n)):
            if self.perms is not None and not rev and i > 0:
                x = self.perms[i - 1]([x])[0]
            blk = self.blocks[i]
            x = blk([x], c=cs, rev=rev)[0]
            J = blk.jac if J is None else J + blk.jac
            if self.perms is not None and rev and i > 0:
                x = self.perms[i - 1]([x], rev=True)[0]
        return x, J

    def init_like_reference_scripts(self, init_scale=0.005, generator=None):
        """p = init_scale * randn for every trainable parameter (train_unconditional.py:165-167)."""
        with torch.no_grad():
            for p in self.parameters():
                if p.requires_grad:
                    p.copy_(init_scale * torch.randn(p.shape, generator=generator, device=p.device, dtype=p.dtype))
        return self


def nll_loss(z, logdet):
    """train_unconditional.py:128-132"""
    return 0.5 * torch.sum(z ** 2, dim=1).mean() - logdet.mean()


class GraphedFlow:
    """CUDA-graph replay of ``HintFlow.forward`` / the inverse for ONE fixed batch size.

    At the reference configs' own batch sizes (300 ... 10 000 samples) sampling and density evaluation are launch-bound: an 8-block
    model is 16 small launches (pack + fused tree kernel per block) plus the host-side tensor bookkeeping around each.  Capturing
    them once and replaying removes the host from the loop (measured on B200, lens `hint_8_full`, B = 10 000: 0.41 -> 0.25 ms per
    forward).  The library's launches are capture-safe: every kernel goes to the caller's stream, nothing synchronises or
    allocates device memory after the first (warm-up) call.  Parameters are read by the pack kernels at replay time, so weight
    updates between calls are seen.  The returned tensors are static buffers that the next call overwrites."""

    def __init__(self, model, batch, rev=False, warmup=2):
        self.model, self.rev = model, bool(rev)
        p = next(model.parameters())
        dc = model.blocks[0].plan.dc
        self.x = torch.zeros(int(batch), model.d, dtype=torch.float32, device=p.device)
        self.c = torch.zeros(int(batch), dc, dtype=torch.float32, device=p.device) if dc else None
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad():
            side = torch.cuda.Stream(device=p.device)
            side.wait_stream(torch.cuda.current_stream(p.device))
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):      # plans' device tables, workspace sizes: everything lazy happens here
                    model(self.x, self.c, rev=self.rev)
            torch.cuda.current_stream(p.device).wait_stream(side)
            with torch.cuda.graph(self.graph):
                self.z, self.J = model(self.x, self.c, rev=self.rev)

    @torch.no_grad()
    def __call__(self, x, c=None):
        if tuple(x.shape) != tuple(self.x.shape):
            raise ValueError(f"GraphedFlow was captured for batch shape {tuple(self.x.shape)}, got {tuple(x.shape)}")
        self.x.copy_(x, non_blocking=True)
        if self.c is not None:
            self.c.copy_(c, non_blocking=True)
        self.graph.replay()
        return self.z, self.J
