"""Host side of the inter-block Householder mixing (FrEIA ``HouseholderPerm`` as the reference's configs use it,
e.g. configs/plus_shape/unconditional_hint_4_3.py:60-71).  Thin autograd wrapper over the C ABI (include/hint_b200.h:
hint_householder_*): W is rebuilt from the reflections by one kernel per call when they are trainable, applied by an error-compensated 3 x TF32 tensor-core
kernel, and differentiated without stored intermediates.  CUDA tensors only - there is no CPU path."""
import torch
import torch.nn as nn

from . import _lib


def _stream():
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())   # inside _lib.on_device(...)


def _dense(t):
    t = t.contiguous()
    return t.clone() if t.data_ptr() % 16 else t


def householder_matrix(Vs):
    """W [d, d] = prod_i (I - 2 v_i v_i^T / |v_i|^2), v_i = Vs[i] (no autograd; see HouseholderMix for the differentiable form)."""
    if not Vs.is_cuda or Vs.dtype != torch.float32:
        raise RuntimeError("hint_b200.householder_matrix: float32 CUDA tensor required (there is no CPU path)")
    Vs = _dense(Vs.detach())
    n, d = Vs.shape
    with _lib.on_device(Vs.device):
        W = torch.empty(d, d, dtype=torch.float32, device=Vs.device)
        _lib.check(_lib.load().hint_householder_matrix(Vs.data_ptr(), n, d, W.data_ptr(), _stream()))
    return W


def householder_apply(x, W, transpose=False):
    """y = x W (or x W^T), fp32-grade (error-compensated 3 x TF32 tensor-core products, fp32 accumulation)."""
    if not x.is_cuda or x.dtype != torch.float32:
        raise RuntimeError("hint_b200.householder_apply: float32 CUDA tensors required (there is no CPU path)")
    x, W = x.contiguous(), _dense(W)       # the kernel takes x / y at any 4-byte alignment
    B, d = x.shape
    with _lib.on_device(x.device):
        y = torch.empty_like(x)
        _lib.check(_lib.load().hint_householder_apply(x.data_ptr(), W.data_ptr(), B, d, 1 if transpose else 0, y.data_ptr(), _stream()))
    return y


def _wgrad(x, dz):
    lib = _lib.load()
    x, dz = _dense(x), _dense(dz)
    B, d = x.shape
    with _lib.on_device(x.device):
        dW = torch.empty(d, d, dtype=torch.float32, device=x.device)
        nbytes = lib.hint_householder_wgrad_workspace_bytes(d)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        _lib.check(lib.hint_householder_wgrad(x.data_ptr(), dz.data_ptr(), B, d, dW.data_ptr(), ws.data_ptr(), nbytes, _stream()))
    return dW


def householder_vs_grad(x, dy, Vs, W, out=None):
    """Gradient of the reflections for y = x W(Vs): dW = x^T dy (deterministic two-stage reduction), then the chain of
    reflections walked backwards without stored intermediates.  ``out`` receives the result when given."""
    dW = _wgrad(x, dy)
    V = _dense(Vs.detach())
    n, d = V.shape
    with _lib.on_device(x.device):
        dVs = out if (out is not None and out.is_contiguous() and out.data_ptr() % 16 == 0) else torch.empty_like(V)
        _lib.check(_lib.load().hint_householder_matrix_backward(V.data_ptr(), _dense(W).data_ptr(), dW.data_ptr(), n, d, dVs.data_ptr(), _stream()))
        if out is not None and dVs is not out:
            out.copy_(dVs)
            dVs = out
    return dVs


class HouseholderMix(torch.autograd.Function):
    """y = x W(Vs) (rev=False) or x W(Vs)^T (rev=True).  ``W`` may be passed pre-built (fixed reflections)."""

    @staticmethod
    def forward(ctx, x, Vs, W, rev):
        if W is None:
            W = householder_matrix(Vs)
        ctx.rev = bool(rev)
        ctx.save_for_backward(x, Vs, W)
        return householder_apply(x, W, transpose=rev)

    @staticmethod
    def backward(ctx, dy):
        x, Vs, W = ctx.saved_tensors
        dy = dy.contiguous()
        dx = householder_apply(dy, W, transpose=not ctx.rev) if ctx.needs_input_grad[0] else None
        dVs = None
        if ctx.needs_input_grad[1]:
            dW = _wgrad(dy, x) if ctx.rev else _wgrad(x, dy)      # y = x W^T: dW = dy^T x;  y = x W: dW = x^T dy
            lib = _lib.load()
            V = _dense(Vs.detach())
            n, d = V.shape
            with _lib.on_device(x.device):
                dVs = torch.empty_like(V)
                _lib.check(lib.hint_householder_matrix_backward(V.data_ptr(), W.data_ptr(), dW.data_ptr(), n, d, dVs.data_ptr(), _stream()))
        return dx, dVs, None, None


class HouseholderPerm(nn.Module):
    """FrEIA ``HouseholderPerm`` (published definition, parity-unpinned): the fixed / learned orthogonal mixing the reference
    inserts between HINT blocks.  CUDA float32 inputs run on the library kernels above; CPU tensors use the plain PyTorch
    expressions, which define the semantics (and serve the shim's CPU tests)."""

    def __init__(self, dims_in, dims_c=[], n_reflections=1, fixed=False):
        super().__init__()
        assert len(dims_in) == 1 and len(dims_in[0]) == 1, "HouseholderPerm mixes flat feature vectors"
        self.width = int(dims_in[0][0])
        self.n_reflections = int(n_reflections)
        self.fixed = bool(fixed)
        self.conditional = len(dims_c) > 0          # accepted and ignored: the mixing does not depend on the condition
        self.Vs = nn.Parameter(torch.randn(self.n_reflections, self.width), requires_grad=not self.fixed)
        if self.fixed:
            self.register_buffer("W", self._matrix(self.Vs.detach()), persistent=False)

    @staticmethod
    def _matrix(Vs):
        W = torch.eye(Vs.shape[1], dtype=Vs.dtype, device=Vs.device)
        for v in Vs:
            W = W - 2.0 * torch.outer(W @ v, v) / torch.dot(v, v)      # W (I - 2 v v^T / |v|^2)
        return W

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        if self.fixed:
            self.W = self._matrix(self.Vs.detach())

    def forward(self, x, c=[], rev=False):
        if x[0].is_cuda and x[0].dtype == torch.float32 and x[0].dim() == 2 and self.width <= 128:
            return [HouseholderMix.apply(x[0], self.Vs, self.W if self.fixed else None, bool(rev))]
        W = self.W if self.fixed else self._matrix(self.Vs)
        return [x[0] @ (W.t() if rev else W)]

    def jacobian(self, x, c=[], rev=False):
        return 0

    def output_dims(self, input_dims):
        assert len(input_dims) == 1, "Can only use one input."
        return input_dims
