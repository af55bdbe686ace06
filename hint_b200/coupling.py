"""Host side of the fused baseline couplings (C ABI: include/hint_b200.h, hint_mlp_coupling_*): FrEIA's
``AffineCoupling`` / ``ExternalAffineCoupling`` with ``F_fully_connected`` subnets as the reference's 2-lane conditional configs
use them (configs/lens_shape/conditional_hint_8_full.py:78-89).  One forward launch; backward = one fused launch + one
weight-gradient launch, instead of ~75 PyTorch kernels per coupling and step.  CUDA float32 tensors only - there is no CPU
path here (FrEIA/modules/coupling.py keeps the plain-PyTorch statement of the definition, which is also the checker)."""
import ctypes

import torch

from . import _lib

PARAM_ORDER = ("fc1", "fc2", "fc2b", "fc3")


def supported(du, dv, hidden):
    return bool(_lib.load().hint_mlp_coupling_supported(int(du), int(dv), int(hidden)))


def _ptrs(ts):
    return (ctypes.c_void_p * 16)(*[t.data_ptr() for t in ts])


def _check(u, v, params):
    if not (u.is_cuda and v.is_cuda) or u.dtype != torch.float32 or v.dtype != torch.float32:
        raise RuntimeError("hint_b200.coupling: float32 CUDA tensors required (there is no CPU path)")
    if len(params) != 16:
        raise ValueError("hint_b200.coupling: 16 parameter tensors expected ([s, t] x [fc1, fc2, fc2b, fc3] x [weight, bias])")
    H, du = params[0].shape
    dv = params[6].shape[0]
    if u.shape[1] != du or v.shape[1] != dv or u.shape[0] != v.shape[0]:
        raise ValueError(f"hint_b200.coupling: u {tuple(u.shape)} / v {tuple(v.shape)} do not match the subnets ({du} -> {H} -> {dv})")
    return int(du), int(dv), int(H)


def forward(u, v, params, clamp, rev=False):
    """y [B, dv], logdet [B] of y = e(s(u)) v + t(u) (rev: (v - t(u)) / e(s(u))); no autograd."""
    du, dv, H = _check(u, v, params)
    u, v = u.contiguous(), v.contiguous()
    params = [p.contiguous() for p in params]
    B = u.shape[0]
    with _lib.on_device(u.device):
        y = torch.empty_like(v)
        jac = torch.empty(B, dtype=torch.float32, device=u.device)
        _lib.check(_lib.load().hint_mlp_coupling_forward(u.data_ptr(), du, v.data_ptr(), dv, H, _ptrs(params), float(clamp), 1 if rev else 0,
                                                         B, y.data_ptr(), jac.data_ptr(), _lib.stream_of(u.device)))
    return y, jac


def backward(u, v, params, clamp, dy, djac):
    """Gradients of the rev=False direction: (du, dv, [16 parameter gradients])."""
    du, dv, H = _check(u, v, params)
    u, v, dy = u.contiguous(), v.contiguous(), dy.contiguous()
    djac = None if djac is None else djac.contiguous()
    params = [p.contiguous() for p in params]
    B = u.shape[0]
    lib = _lib.load()
    with _lib.on_device(u.device):
        gu, gv = torch.empty_like(u), torch.empty_like(v)
        gp = [torch.empty_like(p) for p in params]     # (16 cached-allocator hits: cheaper on the host than views of one buffer)
        nbytes = lib.hint_mlp_coupling_workspace_bytes(du, dv, H, B)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=u.device)
        _lib.check(lib.hint_mlp_coupling_backward(u.data_ptr(), du, v.data_ptr(), dv, H, _ptrs(params), float(clamp), B, dy.data_ptr(),
                                                  djac.data_ptr() if djac is not None else None, gu.data_ptr(), gv.data_ptr(), _ptrs(gp),
                                                  ws.data_ptr(), nbytes, _lib.stream_of(u.device)))
    return gu, gv, gp


class MlpCouplingFn(torch.autograd.Function):
    """(y, logdet) = coupling(u, v; params), differentiable in the rev=False direction."""

    @staticmethod
    def forward(ctx, u, v, clamp, rev, *params):
        y, jac = forward(u, v, params, clamp, rev)
        ctx.save_for_backward(u, v, *params)
        ctx.clamp, ctx.rev = float(clamp), bool(rev)
        return y, jac

    @staticmethod
    def backward(ctx, dy, djac):
        if ctx.rev:
            raise NotImplementedError("hint_b200.coupling: gradients of the rev=True direction are not implemented by the fused kernels")
        u, v, *params = ctx.saved_tensors
        if dy is None:
            dy = torch.zeros_like(v)
        gu, gv, gp = backward(u, v, params, ctx.clamp, dy, djac)
        return (gu, gv, None, None, *gp)


def subnet_params(s, t):
    """The 16 tensors of two F_fully_connected-shaped modules in the C ABI's order."""
    out = []
    for net in (s, t):
        for name in PARAM_ORDER:
            lin = getattr(net, name)
            out += [lin.weight, lin.bias]
    return out
