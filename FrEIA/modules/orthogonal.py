"""``HouseholderPerm`` of the FrEIA shim: the fixed / learned orthogonal mixing the reference inserts between HINT blocks
(configs/uci_data/miniboone_hint_8.py:60-63: ``{'fixed': True, 'n_reflections': ndim_x}``) and, with ``reshuffle=True``,
inside the tree (hint.py:36-39).  FrEIA's source is not part of the reference, so this follows the published definition
W = prod_i (I - 2 v_i v_i^T / |v_i|^2), forward x W, reverse x W^T, log|det| = 0  -- **parity-unpinned** (SURVEY.md 8c).

On CUDA tensors the work runs in hint_b200's own kernels (hint_householder_*: W rebuilt from the reflections in one launch per
call when they are trainable instead of a 100-iteration Python loop, FP32 FFMA application, backward without stored
intermediates); the plain PyTorch expressions below remain for CPU tensors - the shim's CPU tests - and define the semantics."""
import torch
import torch.nn as nn


class HouseholderPerm(nn.Module):
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
            from hint_b200.householder import HouseholderMix
            return [HouseholderMix.apply(x[0], self.Vs, self.W if self.fixed else None, bool(rev))]
        W = self.W if self.fixed else self._matrix(self.Vs)
        return [x[0] @ (W.t() if rev else W)]

    def jacobian(self, x, c=[], rev=False):
        return 0

    def output_dims(self, input_dims):
        assert len(input_dims) == 1, "Can only use one input."
        return input_dims
