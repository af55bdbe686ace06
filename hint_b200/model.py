"""A chain of HINT coupling blocks: the x-lane of the reference configs with the blocks applied back to back
(e.g. configs/uci_data/miniboone_hint_8.py:55-71 minus the inter-block HouseholderPerm nodes, which are FrEIA code
and not part of hint.py).  Used by bench.py, the DP trainer and the smoke test; the reference's own scripts build
the same chain through the FrEIA graph shim instead."""
import torch
import torch.nn as nn

from .block import HierarchicalAffineCouplingBlock


class HintFlow(nn.Module):
    def __init__(self, d, n_blocks, c_internal, dims_c=(), clamp=4.0, max_splits=-1, min_split_size=2):
        super().__init__()
        self.d = int(d)
        self.blocks = nn.ModuleList([
            HierarchicalAffineCouplingBlock([(d,)], dims_c=list(dims_c), c_internal=list(c_internal), clamp=clamp,
                                            max_splits=max_splits, min_split_size=min_split_size)
            for _ in range(n_blocks)])

    @property
    def flops_per_sample(self):
        return sum(b.plan.flops_per_sample for b in self.blocks)

    def forward(self, x, c=None, rev=False):
        """-> (z, logdet).  rev=True applies the inverse blocks in reverse order."""
        cs = [] if c is None else [c]
        J = None
        for blk in (reversed(self.blocks) if rev else self.blocks):
            x = blk([x], c=cs, rev=rev)[0]
            J = blk.jac if J is None else J + blk.jac
        return x, J

    def init_like_reference_scripts(self, init_scale=0.005, generator=None):
        """p = init_scale * randn for every trainable parameter (train_unconditional.py:165-167)."""
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(init_scale * torch.randn(p.shape, generator=generator, device=p.device, dtype=p.dtype))
        return self


def nll_loss(z, logdet):
    """train_unconditional.py:128-132"""
    return 0.5 * torch.sum(z ** 2, dim=1).mean() - logdet.mean()
