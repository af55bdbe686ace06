"""Data-parallel training of a chain of HINT blocks: one process per GPU, weights replicated, the batch sharded.

The only exchange is the gradient all-reduce (average).  Every block keeps its parameters in ONE flat tensor, so each
block is one NCCL bucket: a post-accumulate-grad hook launches the block's all-reduce as soon as its fused backward
kernel has produced the gradient (the last block's bucket goes first), overlapping the transfer with the backward of
the earlier blocks.  ``finish()`` waits for the outstanding buckets; gradient clamp and the optimizer step come after
it, so single-process semantics (clamp AFTER averaging) are kept.  Inference needs no collective at all.
"""
import torch
import torch.distributed as dist


class BucketedGradAllReduce:
    def __init__(self, module, process_group=None):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.handles = []
        self.params = [p for p in module.parameters() if p.requires_grad]
        self._hooks = []
        if self.world > 1:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _on_grad(self, p):
        # average over ranks: the loss means are over the global batch
        if p.is_cuda:
            self.handles.append(dist.all_reduce(p.grad, op=dist.ReduceOp.AVG, group=self.group, async_op=True))
        else:  # gloo (CPU tests): no AVG
            h = dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self.handles.append((h, p))

    def finish(self):
        for h in self.handles:
            if isinstance(h, tuple):
                h[0].wait()
                h[1].grad.div_(self.world)
            else:
                h.wait()
        self.handles.clear()

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks.clear()


def broadcast_parameters(module, src=0, process_group=None):
    if dist.is_initialized() and dist.get_world_size(process_group) > 1:
        for p in module.parameters():
            dist.broadcast(p.data, src=src, group=process_group)


def shard_rows(n_rows, rank, world):
    """Contiguous row shard [lo, hi) of a batch for rank (sampling / density evaluation: no collective)."""
    per = (n_rows + world - 1) // world
    lo = min(n_rows, rank * per)
    return lo, min(n_rows, lo + per)
