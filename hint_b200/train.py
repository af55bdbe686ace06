"""The training step around the coupling blocks, as this library's own kernels.

The reference's step (train_unconditional.py:121-144) is
    x += noise*randn ; z, J = model(x) ; loss = 0.5*mean|z|^2 - mean J ; loss.backward() ;
    for p: p.grad.clamp_(-5, 5) ; Adam.step()
``FusedTrainStep`` runs the same arithmetic without the autograd tape: Philox noise kernel -> the blocks' fused forward kernels
-> one NLL reduction kernel -> the blocks' memory-free backward kernels in reverse order, the last block generating the loss
gradient (dz = z/B, dlogdet = -1/B) inside its tile load -> [NCCL all-reduce of each block's flat gradient as soon as it
exists] -> ONE clamp+Adam launch over all flat parameter tensors.  No `.item()` on the critical path: the loss stays a device
tensor.  ``FusedClampAdam`` is usable on its own as a drop-in for `clamp_` + `torch.optim.Adam` over any fp32 CUDA parameters.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib
from .block import get_precision
from .householder import householder_apply, householder_matrix, householder_vs_grad


def _stream():
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())   # inside _lib.on_device(...)


def add_noise(x, sigma, seed, offset=0, out=None):
    """out = x + sigma * N(0, 1) (train_unconditional.py:121-123); counter-based Philox stream (seed, offset)."""
    if not x.is_cuda or x.dtype != torch.float32:
        raise RuntimeError("hint_b200.add_noise: x must be a float32 CUDA tensor (there is no CPU path)")
    x = x.contiguous()
    if x.data_ptr() % 16:
        x = x.clone()
    if out is None:
        out = torch.empty_like(x)
    with _lib.on_device(x.device):
        _lib.check(_lib.load().hint_add_noise(x.data_ptr(), out.data_ptr(), x.numel(), float(sigma), int(seed) & (2 ** 64 - 1),
                                              int(offset) & (2 ** 64 - 1), _stream()))
    return out


def nll_loss_fused(z, logdet):
    """-> float32[3] device tensor: [0.5*mean_b|z_b|^2 - mean_b logdet_b, first term, second term] (train_unconditional.py:128-132).
    ``logdet``: one [B] tensor or a list of per-block [B] tensors (summed inside the reduction kernel)."""
    if not z.is_cuda:
        raise RuntimeError("hint_b200.nll_loss_fused: CUDA tensors only (there is no CPU path)")
    lib = _lib.load()
    z = z.contiguous()
    if z.data_ptr() % 16:
        z = z.clone()
    js = [j.contiguous() for j in (logdet if isinstance(logdet, (list, tuple)) else [logdet])]
    B, d = z.shape
    if any(tuple(j.shape) != (B,) or j.dtype != torch.float32 or not j.is_cuda for j in js):
        raise ValueError("hint_b200.nll_loss_fused: every logdet must be a float32 CUDA tensor of shape [B]")
    jp = (ctypes.c_void_p * len(js))(*[j.data_ptr() for j in js])
    with _lib.on_device(z.device):
        out = torch.empty(3, dtype=torch.float32, device=z.device)
        nbytes = lib.hint_nll_workspace_bytes()
        ws = torch.empty(nbytes, dtype=torch.uint8, device=z.device)
        _lib.check(lib.hint_nll_loss(z.data_ptr(), jp, len(js), B, d, out.data_ptr(), ws.data_ptr(), nbytes, _stream()))
    return out


class FusedClampAdam(torch.optim.Optimizer):
    """`p.grad.clamp_(-c, c)` for every p, then `torch.optim.Adam(lr, betas, eps, weight_decay).step()` - one kernel launch for
    all parameter tensors (hint_adam_step).  Same state names as torch.optim.Adam (`step`, `exp_avg`, `exp_avg_sq`)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_clamp=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, grad_clamp=grad_clamp))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        lib = _lib.load()
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("FusedClampAdam: parameters and gradients must be contiguous float32 CUDA tensors")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
            step = self.state[ps[0]]["step"]
            if any(self.state[p]["step"] != step for p in ps):
                raise RuntimeError("FusedClampAdam: parameters of one group must share the step count")
            n = len(ps)
            arr = lambda ts: (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts])
            sizes = (ctypes.c_int64 * n)(*[p.numel() for p in ps])
            b1, b2 = group["betas"]
            with _lib.on_device(ps[0].device):
                _lib.check(lib.hint_adam_step(n, arr(ps), arr([p.grad for p in ps]), arr([self.state[p]["exp_avg"] for p in ps]),
                                              arr([self.state[p]["exp_avg_sq"] for p in ps]), sizes, float(group["lr"]), float(b1),
                                              float(b2), float(group["eps"]), float(group["weight_decay"]),
                                              float(group["grad_clamp"]), int(step), _stream()))
        return loss


class FusedTrainStep:
    """One reference training step of a ``HintFlow`` without autograd (see the module docstring).

    step(x, c=None) -> float32[3] device tensor (loss, 0.5*mean|z|^2, mean logdet) of the batch BEFORE the update.
    Data parallel: pass ``world_size`` > 1 (torch.distributed initialised); every block's flat gradient is all-reduced (AVG)
    right after its backward kernel, overlapping the earlier blocks' backward; the loss gradient is scaled by 1/B_local and the
    average over ranks makes it the global-batch mean, as in hint_b200.parallel.BucketedGradAllReduce."""

    def __init__(self, model, optimizer, noise=0.01, seed=0, process_group=None):
        self.model, self.opt, self.noise, self.seed = model, optimizer, float(noise), int(seed)
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.calls = 0
        for blk in model.blocks:
            if blk.flat.grad is None:
                blk.flat.grad = torch.zeros_like(blk.flat)
        self.perms = getattr(model, "perms", None)
        if self.perms is not None:
            for pm in self.perms:
                if not pm.fixed and pm.Vs.grad is None:
                    pm.Vs.grad = torch.zeros_like(pm.Vs)

    @torch.no_grad()
    def step(self, x, c=None):
        model = self.model
        mode = get_precision()
        B = x.shape[0]
        h = add_noise(x, self.noise, self.seed, self.calls) if self.noise else x
        self.calls += 1
        perms = self.perms
        Ws = None
        if perms is not None:      # the mixing matrices of this step: constants, or rebuilt from the trainable reflections (one launch each)
            Ws = [pm.W if pm.fixed else householder_matrix(pm.Vs) for pm in perms]
        zs, Js = [], []
        for i, blk in enumerate(model.blocks):
            if Ws is not None and i > 0:
                h = householder_apply(h, Ws[i - 1])
            h, Jb = blk.plan.forward(h, c, blk.flat.detach(), False, mode=mode)
            zs.append(h)
            Js.append(Jb)
        loss = nll_loss_fused(h, Js)
        handles = []
        dz = None
        inv_b = 1.0 / B
        for i in range(len(model.blocks) - 1, -1, -1):
            blk = model.blocks[i]
            dz, _, _, _ = blk.plan.backward(zs[i], c, blk.flat.detach(), dz, None, mode=mode, want_dc=False, nll_scale=inv_b,
                                            out=blk.flat.grad)
            zs[i] = None
            if self.world > 1:
                handles.append(dist.all_reduce(blk.flat.grad, op=dist.ReduceOp.AVG, group=self.group, async_op=True))
            if Ws is not None and i > 0:      # through the mixing in front of block i: y = z_{i-1} W
                pm = perms[i - 1]
                if not pm.fixed:
                    householder_vs_grad(zs[i - 1], dz, pm.Vs, Ws[i - 1], out=pm.Vs.grad)
                    if self.world > 1:
                        handles.append(dist.all_reduce(pm.Vs.grad, op=dist.ReduceOp.AVG, group=self.group, async_op=True))
                dz = householder_apply(dz, Ws[i - 1], transpose=True)
        for hd in handles:
            hd.wait()
        self.opt.step()
        return loss


def multi_mmd(x, y, widths_exponents=((0.5, 1), (0.2, 1), (0.2, 0.5))):
    """The sampling scripts' evaluation metric (rejection_sampling.py:56-73): multi-kernel MMD^2 of two sets of n samples with
    inverse-multiquadric kernels, as ONE fused pair-tile kernel (nothing n x n is materialised) -> 0-dim device tensor.  No autograd
    (the reference only evaluates it)."""
    if not (x.is_cuda and y.is_cuda) or x.dtype != torch.float32 or y.dtype != torch.float32:
        raise RuntimeError("hint_b200.multi_mmd: float32 CUDA tensors required (there is no CPU path)")
    if x.dim() != 2 or x.shape != y.shape:
        raise ValueError("hint_b200.multi_mmd: x and y must both be [n, d] (the reference averages XX + YY - 2 XY elementwise)")
    x, y = x.detach().contiguous(), y.detach().contiguous()
    n, d = x.shape
    k = len(widths_exponents)
    C = (ctypes.c_float * k)(*[float(w) for w, _ in widths_exponents])
    a = (ctypes.c_float * k)(*[float(e) for _, e in widths_exponents])
    lib = _lib.load()
    with _lib.on_device(x.device):
        out = torch.empty((), dtype=torch.float32, device=x.device)
        nbytes = lib.hint_mmd_workspace_bytes(n)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        _lib.check(lib.hint_multi_mmd(x.data_ptr(), y.data_ptr(), n, d, C, a, k, out.data_ptr(), ws.data_ptr(), nbytes, _stream()))
    return out
