"""Host-side mirror of the reference module surface (/root/reference/hint.py:104-133).

``HierarchicalAffineCouplingBlock`` keeps the FrEIA-module protocol of the reference
(ctor kwargs, ``forward(x_list, c=[], rev=False) -> [z]``, ``jacobian()``, ``output_dims``) and its
``state_dict`` key names (``tree.upper.s.0.weight`` ...), but holds ALL parameters of the block in one
flat fp32 ``nn.Parameter`` in the reference's ``parameters()`` order, and runs the whole coupling tree
in one fused CUDA kernel through the C ABI (include/hint_b200.h).  Backward is memory-free: autograd
keeps only the block output.

No CPU fallback: non-CUDA tensors raise.
"""
import ctypes
import math
import os

import torch
import torch.nn as nn

from . import _lib

_MODES = {"fp32": _lib.MODE_FP32, "tf32": _lib.MODE_TF32, "tf32x3": _lib.MODE_TF32X3, "3xtf32": _lib.MODE_TF32X3,
          "tf32_tcgen05": _lib.MODE_TF32_TCGEN05, "tf32_mma": _lib.MODE_TF32_MMA, "tf32_chain": _lib.MODE_TF32_CHAIN,
          "tf32_tc3": _lib.MODE_TF32_TC3}
_mode = os.environ.get("HINT_B200_MODE", "fp32").lower()


def set_precision(mode: str):
    """Select the GEMM arithmetic of all blocks without touching the configs: 'fp32' | 'tf32' | 'tf32x3'."""
    global _mode
    if mode.lower() not in _MODES:
        raise ValueError(f"unknown precision mode {mode!r}; choose from {sorted(_MODES)}")
    _mode = mode.lower()


def get_precision() -> str:
    return _mode


# The backward is memory-free: autograd keeps only the block OUTPUT and the kernel re-derives every node's input as
# x_l = (z_l - t) / e(s).  With clamp = 4 the scale e lies in [e^-4, e^4]; the division and the cancellation in z_l - t amplify
# rounding by up to ~55x per nesting level, so a strongly trained / saturated coupling can lose accuracy in the reconstruction
# (and with it in the gradients) where the reference, which stores its activations, does not.  `set_backward_check(tol)` makes
# every forward also keep the block INPUT and every backward compare it with the kernel's reconstruction:
#   max|x_rec - x| <= tol * max(1, max|x|)   else RuntimeError (or a warning with raise_error=False).
_bwd_check = None


def set_backward_check(tol=1e-3, raise_error=True):
    """Enable (tol > 0) or disable (tol = None / 0) the reconstruction check of the memory-free backward.  Costs one extra
    [B, d] tensor per block kept for backward and one write + one read of it."""
    global _bwd_check
    _bwd_check = (float(tol), bool(raise_error)) if tol else None


def get_backward_check():
    return _bwd_check


def linear_subnet_constructor(c_in, c_out, c_internal):
    """Shape contract of the only subnet the fused kernels implement (hint.py:10-13):
    Linear(c_in, h) - ReLU - Linear(h, h) - ReLU - Linear(h, c_out).  Passing this function (or None) as
    ``subnet_constructor`` selects the fused path; any other callable is rejected."""
    return nn.Sequential(nn.Linear(c_in, c_internal), nn.ReLU(),
                         nn.Linear(c_internal, c_internal), nn.ReLU(),
                         nn.Linear(c_internal, c_out))


class TreePlan:
    """Python handle of a ``hint_plan_t`` (tree of hint.py:25-54 flattened by the C++ planner)."""

    def __init__(self, d, dc, c_internal, clamp, max_splits, min_split_size, reshuffle):
        lib = _lib.load()
        self._lib = lib
        ci = (ctypes.c_int32 * len(c_internal))(*[int(v) for v in c_internal])
        handle = ctypes.c_void_p()
        _lib.check(lib.hint_plan_create(int(d), int(dc), ci, len(c_internal), float(clamp), int(max_splits),
                                        int(min_split_size), 1 if reshuffle else 0, ctypes.byref(handle)))
        self._h = handle
        self._ws_cache = {}
        self.d, self.dc, self.clamp = int(d), int(dc), float(clamp)
        self.c_internal, self.max_splits, self.min_split_size = [int(v) for v in c_internal], int(max_splits), int(min_split_size)
        n = lib.hint_plan_num_nodes(handle)
        self.nodes = []
        for i in range(n):
            info = _lib.NodeInfo()
            _lib.check(lib.hint_plan_node(handle, i, ctypes.byref(info)))
            self.nodes.append({f: getattr(info, f) for f, _ in _lib.NodeInfo._fields_})
        self.n_params = int(lib.hint_plan_param_count(handle))
        self.flops_per_sample = int(lib.hint_plan_flops_per_sample(handle))
        offs = (ctypes.c_int64 * (n * 12))()
        _lib.check(lib.hint_plan_param_layout(handle, offs, n * 12))
        paths = {}
        for i, nd in enumerate(self.nodes):
            if nd["parent"] < 0:
                paths[i] = "tree"
            else:
                par = self.nodes[nd["parent"]]
                paths[i] = paths[nd["parent"]] + (".upper" if par["upper"] == i else ".lower")
        self.paths = paths
        # (name, offset, shape) in the reference's parameters() order
        self.entries = []
        for i, nd in enumerate(self.nodes):
            dims = [(nd["h"], nd["cin"]), (nd["h"], nd["h"]), (nd["cout"], nd["h"])]
            for net, netname in enumerate("st"):
                for layer, (o, k) in enumerate(dims):
                    base = i * 12 + net * 6 + layer * 2
                    self.entries.append((f"{paths[i]}.{netname}.{2 * layer}.weight", int(offs[base]), (o, k)))
                    self.entries.append((f"{paths[i]}.{netname}.{2 * layer}.bias", int(offs[base + 1]), (o,)))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.hint_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def tile_rows(self, which=_lib.WS_FORWARD):
        return int(self._lib.hint_plan_tile_rows(self._h, which))

    def mode_supported(self, mode: str) -> bool:
        """True when this block fits the kernel family behind `mode` (every family has a shape / shared-memory envelope)."""
        return bool(self._lib.hint_plan_mode_supported(self._h, _MODES[mode]))

    # -- launches ------------------------------------------------------------------------------
    def _ws_bytes(self, B, kind):
        key = (B, kind)
        n = self._ws_cache.get(key)
        if n is None:
            if len(self._ws_cache) > 64:
                self._ws_cache.clear()
            n = self._ws_cache[key] = self._lib.hint_workspace_bytes(self._h, B, kind)
        return n

    @staticmethod
    def _check(t, name, shape=None):
        if t is None:
            return
        if not t.is_cuda:
            raise RuntimeError(f"hint_b200: {name} must be a CUDA tensor (there is no CPU path)")
        if t.dtype != torch.float32:
            raise TypeError(f"hint_b200: {name} must be float32, got {t.dtype}")
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError(f"hint_b200: {name} has shape {tuple(t.shape)}, expected {tuple(shape)}")

    @staticmethod
    def _dense16(t):
        """Contiguous and 16-byte aligned (the C ABI's contract).  `.contiguous()` is a no-op for a row slice like x[1:], whose
        storage offset can leave the base pointer misaligned when d*4 is not a multiple of 16 - the reference accepts any tensor,
        so such views are copied here instead of being rejected by the library."""
        if t is None:
            return None
        t = t.contiguous()
        return t.clone() if t.data_ptr() % 16 else t

    def forward(self, x, c, flat, rev=False, mode=None):
        """z, logdet = f(x; c) (rev=False) or f^-1(x; c) (rev=True).  hint.py:62-101."""
        B = x.shape[0]
        self._check(x, "x", (B, self.d))
        self._check(flat, "params", (self.n_params,))
        if self.dc:
            if c is None:
                raise ValueError("hint_b200: block was built with dims_c but no condition was passed")
            self._check(c, "c", (B, self.dc))
            c = self._dense16(c)
        x = self._dense16(x)
        flat = self._dense16(flat)
        with _lib.on_device(x.device):
            z = torch.empty_like(x)
            J = torch.empty(B, dtype=torch.float32, device=x.device)
            nbytes = self._ws_bytes(B, _lib.WS_FORWARD)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
            _lib.check(self._lib.hint_forward(self._h, x.data_ptr(), c.data_ptr() if self.dc else None, flat.data_ptr(), B,
                                              1 if rev else 0, _MODES[mode or _mode], z.data_ptr(), J.data_ptr(),
                                              ws.data_ptr(), nbytes, _lib.stream_of(x.device)))
        return z, J

    def backward(self, z, c, flat, dz, dJ, mode=None, want_xrec=False, want_dc=True, nll_scale=None, out=None):
        """Gradients of the rev=False direction from the block OUTPUT z (memory-free backward).

        ``nll_scale`` (dJ = None): the upstream gradient is the NLL loss's (train_unconditional.py:128-132): dlogdet = -nll_scale
        and, when dz is None too (last block), dz = nll_scale * z - both generated inside the kernel (hint_backward_nll) where
        the kernel family supports it and materialised here otherwise.  ``out``: optional tensor receiving the parameter gradient."""
        B = z.shape[0]
        self._check(z, "z", (B, self.d))
        if nll_scale is not None:
            if dz is not None:
                self._check(dz, "dz", (B, self.d))
            try:
                return self._backward_launch(z, c, flat, dz, None, mode, want_xrec, want_dc, float(nll_scale), out)
            except NotImplementedError:   # kernel family without the fused form: materialise the two gradients
                dz = z * float(nll_scale) if dz is None else dz
                dJ = torch.full((B,), -float(nll_scale), dtype=torch.float32, device=z.device)
        self._check(dz, "dz", (B, self.d))
        self._check(dJ, "dlogdet", (B,))
        return self._backward_launch(z, c, flat, dz, dJ, mode, want_xrec, want_dc, None, out)

    def _backward_launch(self, z, c, flat, dz, dJ, mode, want_xrec, want_dc, nll_scale, out):
        B = z.shape[0]
        self._check(flat, "params", (self.n_params,))
        if self.dc:
            self._check(c, "c", (B, self.dc))
            c = self._dense16(c)
        z, dz, dJ, flat = self._dense16(z), self._dense16(dz), self._dense16(dJ), self._dense16(flat)
        with _lib.on_device(z.device):
            dx = torch.empty_like(z)
            dc = torch.empty(B, self.dc, dtype=torch.float32, device=z.device) if (self.dc and want_dc) else None
            xrec = torch.empty_like(z) if want_xrec else None
            dflat = out if (out is not None and out.is_contiguous() and out.data_ptr() % 16 == 0 and out.shape == flat.shape) else torch.empty_like(flat)
            nbytes = self._ws_bytes(B, _lib.WS_BACKWARD)
            if nbytes == 0:
                _lib.check(_lib.HINT_ERR_CUDA)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=z.device)
            tail = (xrec.data_ptr() if want_xrec else None, dx.data_ptr(), dc.data_ptr() if dc is not None else None,
                    dflat.data_ptr(), ws.data_ptr(), nbytes, _lib.stream_of(z.device))
            if nll_scale is not None:
                _lib.check(self._lib.hint_backward_nll(self._h, z.data_ptr(), c.data_ptr() if self.dc else None, flat.data_ptr(),
                                                       dz.data_ptr() if dz is not None else None, nll_scale, B, _MODES[mode or _mode], *tail))
            else:
                _lib.check(self._lib.hint_backward(self._h, z.data_ptr(), c.data_ptr() if self.dc else None, flat.data_ptr(),
                                                   dz.data_ptr(), dJ.data_ptr(), B, _MODES[mode or _mode], *tail))
            if out is not None and dflat is not out:
                out.copy_(dflat)
                dflat = out
        return dx, dc, dflat, xrec


def _inverse_graph(plan, x, c, flat):
    """The inverse transport of one block written with differentiable PyTorch ops (hint.py:82-96, root coupling first, then the
    children), on the tensors' own (CUDA) device.  Used ONLY to differentiate the rev=True direction, which the reference never
    does in its scripts (model_inverse always runs under no_grad / detach) but its autograd would allow: the fused kernels cover
    the forward-direction backward, this keeps the rev direction differentiable instead of raising.  Not a compute path of
    forward / inverse / training."""
    alpha = plan.clamp * 0.636
    views = {name: flat[off:off + math.prod(shape)].view(shape) for name, off, shape in plan.entries}
    cols = list(x.unbind(dim=1))
    J = x.new_zeros(x.shape[0])
    order = sorted(range(len(plan.nodes)), key=lambda i: plan.nodes[i]["depth"])     # root level first (hint.py:85-88)
    for i in order:
        nd, path = plan.nodes[i], plan.paths[i]
        lo, k, hi = nd["lo"], nd["k"], nd["hi"]
        a = torch.stack(cols[lo:lo + k], dim=1)
        if plan.dc:
            a = torch.cat([a, c], dim=1)

        def mlp(net):
            h = torch.relu(torch.addmm(views[f"{path}.{net}.0.bias"], a, views[f"{path}.{net}.0.weight"].t()))
            h = torch.relu(torch.addmm(views[f"{path}.{net}.2.bias"], h, views[f"{path}.{net}.2.weight"].t()))
            return torch.addmm(views[f"{path}.{net}.4.bias"], h, views[f"{path}.{net}.4.weight"].t())
        la = alpha * torch.atan(mlp("s"))
        xl = (torch.stack(cols[lo + k:hi], dim=1) - mlp("t")) * torch.exp(-la)
        J = J - la.sum(dim=1)
        for j in range(hi - lo - k):
            cols[lo + k + j] = xl[:, j]
    return torch.stack(cols, dim=1), J


class _CouplingFn(torch.autograd.Function):
    """z, logdet = block(x, c).  Saves only the OUTPUT z (+ c, params); backward re-derives the rest."""

    @staticmethod
    def forward(ctx, x, c, flat, plan, rev):
        mode = _mode   # the precision mode is fixed per call: backward uses the kernels of the mode the forward ran in
        z, J = plan.forward(x, c, flat, rev, mode=mode)
        ctx.plan, ctx.rev, ctx.mode = plan, rev, mode
        ctx.check = _bwd_check if not rev else None
        ctx.save_for_backward(z, c if c is not None else x.new_empty(0), flat, x.detach() if (ctx.check or rev) else x.new_empty(0))
        return z, J

    @staticmethod
    def backward(ctx, dz, dJ):
        z, c, flat, x_in = ctx.saved_tensors
        plan = ctx.plan
        if ctx.rev:
            # gradients through the inverse: differentiate the PyTorch restatement at the saved INPUT of the rev call
            with torch.enable_grad():
                xi = x_in.detach().requires_grad_(True)
                ci = c.detach().requires_grad_(True) if plan.dc else None
                fi = flat.detach().requires_grad_(True)
                y, Jr = _inverse_graph(plan, xi, ci, fi)
                outs, grads = [], []
                if dz is not None:
                    outs.append(y); grads.append(dz)
                if dJ is not None:
                    outs.append(Jr); grads.append(dJ)
                ins = [xi, fi] + ([ci] if plan.dc else [])
                g = torch.autograd.grad(outs, ins, grads, allow_unused=True)
            gx, gf = g[0], g[1]
            gc = g[2] if plan.dc else None
            return (gx if ctx.needs_input_grad[0] else None, gc if ctx.needs_input_grad[1] else None,
                    gf if ctx.needs_input_grad[2] else None, None, None)
        if dz is None:
            dz = torch.zeros_like(z)
        if dJ is None:
            dJ = torch.zeros(z.shape[0], dtype=z.dtype, device=z.device)
        dx, dc, dflat, xrec = plan.backward(z, c if plan.dc else None, flat, dz, dJ, mode=ctx.mode, want_dc=ctx.needs_input_grad[1],
                                            want_xrec=ctx.check is not None)
        if ctx.check is not None:
            tol, fatal = ctx.check
            err = float((xrec - x_in).abs().max())
            scale = max(1.0, float(x_in.abs().max()))
            if not (err <= tol * scale):
                msg = (f"hint_b200: the memory-free backward reconstructed the block input with max abs error {err:.3e} "
                       f"(> {tol:g} * {scale:.3g}) in mode '{ctx.mode}': the coupling is too ill-conditioned for "
                       "recomputation by inversion; gradients of this step are unreliable (use mode 'fp32' / 'tf32x3' or a smaller clamp)")
                if fatal:
                    raise RuntimeError(msg)
                import warnings
                warnings.warn(msg)
        return (dx if ctx.needs_input_grad[0] else None, dc if ctx.needs_input_grad[1] else None,
                dflat if ctx.needs_input_grad[2] else None, None, None)


class _LinearView:
    def __init__(self, weight, bias):
        self.weight, self.bias = weight, bias


class _SubnetView:
    """Read/write views of one subnet's tensors with the reference's indices: net[0], net[2], net[4]."""

    def __init__(self, layers):
        self._layers = layers

    def __getitem__(self, i):
        return self._layers[i]


class HierarchicalAffineCouplingTree:
    """Structural view of one tree node (reference: hint.py:21-101).  Not an nn.Module: it owns no storage;
    ``.s[0].weight`` etc. are views into the owning block's flat parameter."""

    def __init__(self, block, idx):
        self._block, self._idx = block, idx
        nd = block.plan.nodes[idx]
        self.data_shape = (nd["hi"] - nd["lo"],)
        self.clamp = block.plan.clamp
        self.split_idx = nd["k"]
        self.conditional = block.plan.dc > 0
        self.leaf = bool(nd["leaf"])
        self.perm = None

    def _net(self, netname):
        b = self._block
        path = b.plan.paths[self._idx]
        views = b.named_views()
        return _SubnetView({2 * l: _LinearView(views[f"{path}.{netname}.{2 * l}.weight"], views[f"{path}.{netname}.{2 * l}.bias"])
                            for l in range(3)})

    @property
    def s(self):
        return self._net("s")

    @property
    def t(self):
        return self._net("t")

    @property
    def upper(self):
        nd = self._block.plan.nodes[self._idx]
        return None if nd["leaf"] else HierarchicalAffineCouplingTree(self._block, nd["upper"])

    @property
    def lower(self):
        nd = self._block.plan.nodes[self._idx]
        return None if nd["leaf"] else HierarchicalAffineCouplingTree(self._block, nd["lower"])


class HierarchicalAffineCouplingBlock(nn.Module):
    """Drop-in for hint.py:104-133 (same ctor signature, forward/jacobian/output_dims protocol)."""

    def __init__(self, dims_in, dims_c=[], conv=False, subnet_constructor=None, c_internal=[], clamp=4.,
                 max_splits=-1, min_split_size=2, reshuffle=False):
        super().__init__()
        assert all([tuple(dims_c[i][1:]) == tuple(dims_in[0][1:]) for i in range(len(dims_c))]), \
            "Dimensions of input and one or more conditions don't agree."
        if conv or len(dims_in[0]) != 1:
            raise NotImplementedError("hint_b200: conv=True / image-shaped inputs are outside the fused path "
                                      "(no reference config uses them)")
        if subnet_constructor is not None and subnet_constructor is not linear_subnet_constructor:
            raise NotImplementedError("hint_b200: only the default 3-layer ReLU MLP subnet (hint.py:10-13) is fused; "
                                      "custom subnet_constructor is not supported and there is no fallback")
        d = int(dims_in[0][0])
        dc = int(sum(dims_c[i][0] for i in range(len(dims_c))))
        self.dims_c = [tuple(t) for t in dims_c]
        # reshuffle=True (hint.py:36-39,64-65,93-94): every tree node mixes its inputs with a FIXED orthogonal matrix before the
        # split.  The mixings are applied top-down while the recursion descends and the couplings bottom-up while it returns, so
        # all of them compose into ONE d x d orthogonal matrix M = D_0 D_1 ... (D_l = block-diagonal of the depth-l nodes' W) in
        # front of the un-shuffled tree: forward x M -> tree, inverse tree^-1 -> . M^T, log|det| unchanged.  The tree runs in the
        # fused kernels as always, the mixing in hint_householder_apply (FP32).  W follows the published FrEIA definition
        # (product of d_node random Householder reflections, parity-unpinned: FrEIA's source is not part of the reference).
        self.plan = TreePlan(d, dc, list(c_internal), clamp, max_splits, min_split_size, False)
        self.flat = nn.Parameter(torch.empty(self.plan.n_params, dtype=torch.float32))
        self.reshuffle = bool(reshuffle)
        if self.reshuffle:
            if d > 128:
                raise NotImplementedError("hint_b200: reshuffle=True is implemented for d <= 128")
            for i, nd in enumerate(self.plan.nodes):      # fixed reflections: buffers (not trainable), saved under the reference's names
                self.register_buffer(f"_perm_vs_{i}", torch.randn(nd["hi"] - nd["lo"], nd["hi"] - nd["lo"]), persistent=False)
            self.register_buffer("perm_M", self._compose_perm(), persistent=False)
            self.register_buffer("_no_vs", torch.zeros(1), persistent=False)
        self.reset_parameters()
        self.jac = None

    @property
    def perm_vs(self):
        return [getattr(self, f"_perm_vs_{i}") for i in range(len(self.plan.nodes))]

    def _compose_perm(self):
        d = self.plan.d
        depth = max(nd["depth"] for nd in self.plan.nodes)
        M = torch.eye(d, dtype=torch.float64)
        for lv in range(depth + 1):
            D = torch.eye(d, dtype=torch.float64)
            for nd, vs in zip(self.plan.nodes, self.perm_vs):
                if nd["depth"] != lv:
                    continue
                n = nd["hi"] - nd["lo"]
                W = torch.eye(n, dtype=torch.float64)
                for v in vs.detach().double().cpu():
                    W = W - 2.0 * torch.outer(W @ v, v) / torch.dot(v, v)
                D[nd["lo"]:nd["hi"], nd["lo"]:nd["hi"]] = W
            M = M @ D
        return M.float().to(self.flat.device)

    # -- parameters ------------------------------------------------------------------------------
    def reset_parameters(self):
        """nn.Linear's default init per layer (what the reference gets from torch): U(-1/sqrt(fan_in), +)."""
        with torch.no_grad():
            fan_in = 0
            for name, off, shape in self.plan.entries:
                if len(shape) == 2:  # a bias uses the bound of the weight that precedes it
                    fan_in = shape[1]
                bound = 1.0 / math.sqrt(fan_in) if fan_in > 0 else 0.0
                self.flat.data[off:off + math.prod(shape)].uniform_(-bound, bound)

    def named_views(self):
        """{reference state_dict name: view into the flat parameter}."""
        out = {}
        data = self.flat.data
        for name, off, shape in self.plan.entries:
            n = 1
            for s in shape:
                n *= s
            out[name] = data[off:off + n].view(*shape)
        return out

    @property
    def tree(self):
        return HierarchicalAffineCouplingTree(self, 0)

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        for name, v in self.named_views().items():
            destination[prefix + name] = v if keep_vars else v.detach().clone()
        if self.reshuffle:      # the reference keeps one HouseholderPerm per tree node: <path>.perm.Vs
            for i, vs in enumerate(self.perm_vs):
                destination[prefix + self.plan.paths[i] + ".perm.Vs"] = vs if keep_vars else vs.detach().clone()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        views = self.named_views()
        seen = set()
        for name, v in views.items():
            key = prefix + name
            if key not in state_dict:
                missing_keys.append(key)
                continue
            seen.add(key)
            src = state_dict[key]
            if tuple(src.shape) != tuple(v.shape):
                error_msgs.append(f"size mismatch for {key}: copying a param with shape {tuple(src.shape)} from "
                                  f"checkpoint, the shape in current model is {tuple(v.shape)}.")
                continue
            with torch.no_grad():
                v.copy_(src)
        if self.reshuffle:
            for i, vs in enumerate(self.perm_vs):
                key = prefix + self.plan.paths[i] + ".perm.Vs"
                if key not in state_dict:
                    missing_keys.append(key)
                    continue
                seen.add(key)
                with torch.no_grad():
                    vs.copy_(state_dict[key])
            self.perm_M = self._compose_perm()
        if strict:
            for key in state_dict.keys():
                if key.startswith(prefix) and key not in seen:
                    unexpected_keys.append(key)

    # -- FrEIA module protocol -------------------------------------------------------------------
    def forward(self, x, c=[], rev=False):
        x0 = x[0]
        cc = None
        if self.plan.dc:
            cc = c[0] if len(c) == 1 else torch.cat(list(c), dim=1)
        if self.reshuffle and not rev:
            from .householder import HouseholderMix
            x0 = HouseholderMix.apply(x0, self._no_vs, self.perm_M, False)
        if torch.is_grad_enabled() and (x0.requires_grad or self.flat.requires_grad or (cc is not None and cc.requires_grad)):
            z, self.jac = _CouplingFn.apply(x0, cc, self.flat, self.plan, bool(rev))
        else:
            z, self.jac = self.plan.forward(x0, cc, self.flat.detach(), bool(rev))
        if self.reshuffle and rev:
            from .householder import HouseholderMix
            z = HouseholderMix.apply(z, self._no_vs, self.perm_M, True)
        return [z]

    def jacobian(self, x, c=[], rev=False):
        return self.jac

    def output_dims(self, input_dims):
        assert len(input_dims) == 1, "Can only use one input."
        return input_dims

    def extra_repr(self):
        p = self.plan
        return (f"d={p.d}, dc={p.dc}, nodes={len(p.nodes)}, params={p.n_params}, clamp={p.clamp}, "
                f"tile_rows(fwd/bwd)={p.tile_rows(0)}/{p.tile_rows(1)}, mode={_mode}")
