"""The baseline couplings the 2-lane conditional HINT configs put around the HINT block:
``ExternalAffineCoupling`` (y -> x lane) and ``AffineCoupling`` (y lane) with ``F_fully_connected`` subnets, e.g.
configs/lens_shape/conditional_hint_8_full.py:78-89; the `*_inn_*` / `*_cinn_*` baseline configs use the same classes.

FrEIA's source is not part of the reference and no version is pinned, so these follow the published FrEIA definitions of
that period - **parity-unpinned** (DESIGN.md section 2): a four-layer fully connected subnet, a single affine transformation
of the lower half (AffineCoupling) or of the whole input driven by the condition alone (ExternalAffineCoupling), with the
same soft clamp hint.py:56-60 uses, e(s) = exp(clamp * 0.636 * atan(s)).

On CUDA float32 tensors both couplings run on the library's fused kernels (hint_b200/csrc/mlp_coupling.cu through
hint_b200.coupling: one launch forward, two backward); the plain-PyTorch expressions below are the definition, the CPU path of
this test shim, the checker of the kernels, and the fallback for what the kernels do not cover (custom subnets, dropout in
training mode, gradients through rev=True, widths outside the envelope)."""
import torch
import torch.nn as nn


class F_fully_connected(nn.Module):
    """Fully connected subnet: Linear-ReLU x3 + Linear (fc1, fc2, fc2b, fc3), optional dropout; internal_size defaults to 2*size."""

    def __init__(self, size_in, size, internal_size=None, dropout=0.0):
        super().__init__()
        if not internal_size:
            internal_size = 2 * size
        self.d1, self.d2, self.d2b = nn.Dropout(p=dropout), nn.Dropout(p=dropout), nn.Dropout(p=dropout)
        self.fc1 = nn.Linear(size_in, internal_size)
        self.fc2 = nn.Linear(internal_size, internal_size)
        self.fc2b = nn.Linear(internal_size, internal_size)
        self.fc3 = nn.Linear(internal_size, size)
        self.nl1, self.nl2, self.nl2b = nn.ReLU(), nn.ReLU(), nn.ReLU()

    def forward(self, x):
        out = self.nl1(self.d1(self.fc1(x)))
        out = self.nl2(self.d2(self.fc2(out)))
        out = self.nl2b(self.d2b(self.fc2b(out)))
        return self.fc3(out)


class _AffineBase(nn.Module):
    def __init__(self, clamp):
        super().__init__()
        self.clamp = float(clamp)
        self.jac = None

    def log_e(self, s):
        return self.clamp * 0.636 * torch.atan(s)

    def e(self, s):
        return torch.exp(self.log_e(s))

    def _fused(self, u, v, rev):
        """(y, logdet) from the fused kernels, or None when this call is outside what they cover."""
        if not (u.is_cuda and v.is_cuda and u.dtype == torch.float32 and v.dtype == torch.float32):
            return None
        s, t = self.s, self.t
        if type(s) is not F_fully_connected or type(t) is not F_fully_connected:
            return None
        if self.training and (s.d1.p > 0 or t.d1.p > 0):
            return None
        from hint_b200 import coupling as _k
        params = _k.subnet_params(s, t)
        if rev and torch.is_grad_enabled() and (u.requires_grad or v.requires_grad or any(p.requires_grad for p in params)):
            return None
        if not _k.supported(u.shape[1], v.shape[1], s.fc1.out_features):
            return None
        return _k.MlpCouplingFn.apply(u, v, self.clamp, bool(rev), *params)

    def jacobian(self, x, c=[], rev=False):
        """Cached log|det J| of the last call (as hint.py:128-129 does), so ``jacobian(None)`` works (train_conditional.py:50-55)."""
        return self.jac

    def output_dims(self, input_dims):
        assert len(input_dims) == 1, "Can only use one input."
        return input_dims


class AffineCoupling(_AffineBase):
    """x = [x1 | x2] (split at d//2); y2 = e(s(x1, c)) * x2 + t(x1, c), x1 unchanged."""

    def __init__(self, dims_in, dims_c=[], F_class=F_fully_connected, F_args={}, clamp=5.0):
        super().__init__(clamp)
        channels = dims_in[0][0]
        self.split_len1 = channels // 2
        self.split_len2 = channels - channels // 2
        assert all(tuple(dims_c[i][1:]) == tuple(dims_in[0][1:]) for i in range(len(dims_c))), \
            "Dimensions of input and one or more conditions don't agree."
        self.conditional = len(dims_c) > 0
        condition_length = sum(dims_c[i][0] for i in range(len(dims_c)))
        self.s = F_class(self.split_len1 + condition_length, self.split_len2, **F_args)
        self.t = F_class(self.split_len1 + condition_length, self.split_len2, **F_args)

    def forward(self, x, c=[], rev=False):
        x1, x2 = x[0].narrow(1, 0, self.split_len1), x[0].narrow(1, self.split_len1, self.split_len2)
        x1_c = torch.cat([x1, *c], 1) if self.conditional else x1
        fused = self._fused(x1_c, x2, rev) if x[0].dim() == 2 else None
        if fused is not None:
            y2, self.jac = fused
            return [torch.cat((x1, y2), 1)]
        s, t = self.s(x1_c), self.t(x1_c)
        if not rev:
            y2 = self.e(s) * x2 + t
            self.jac = torch.sum(self.log_e(s), dim=tuple(range(1, s.dim())))
        else:
            y2 = (x2 - t) / self.e(s)
            self.jac = -torch.sum(self.log_e(s), dim=tuple(range(1, s.dim())))
        return [torch.cat((x1, y2), 1)]


class ExternalAffineCoupling(_AffineBase):
    """y = e(s(c)) * x + t(c): the whole input is transformed, scale and shift come from the condition alone."""

    def __init__(self, dims_in, dims_c=[], F_class=F_fully_connected, F_args={}, clamp=5.0):
        super().__init__(clamp)
        assert len(dims_c) > 0, "ExternalAffineCoupling needs a condition"
        channels = dims_in[0][0]
        condition_length = sum(dims_c[i][0] for i in range(len(dims_c)))
        self.s = F_class(condition_length, channels, **F_args)
        self.t = F_class(condition_length, channels, **F_args)

    def forward(self, x, c=[], rev=False):
        cc = torch.cat(list(c), 1) if len(c) > 1 else c[0]
        fused = self._fused(cc, x[0], rev) if x[0].dim() == 2 else None
        if fused is not None:
            y, self.jac = fused
            return [y]
        s, t = self.s(cc), self.t(cc)
        if not rev:
            y = self.e(s) * x[0] + t
            self.jac = torch.sum(self.log_e(s), dim=tuple(range(1, s.dim())))
        else:
            y = (x[0] - t) / self.e(s)
            self.jac = -torch.sum(self.log_e(s), dim=tuple(range(1, s.dim())))
        return [y]
