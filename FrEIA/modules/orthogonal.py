"""``HouseholderPerm`` of the FrEIA shim: the fixed / learned orthogonal mixing the reference inserts between HINT blocks
(configs/uci_data/miniboone_hint_8.py:60-63: ``{'fixed': True, 'n_reflections': ndim_x}``) and, with ``reshuffle=True``,
inside the tree (hint.py:36-39).  FrEIA's source is not part of the reference, so this follows the published definition
W = prod_i (I - 2 v_i v_i^T / |v_i|^2), forward x W, reverse x W^T, log|det| = 0  -- **parity-unpinned** (SURVEY.md 8c).
The class lives in hint_b200.householder (library kernels on CUDA, plain PyTorch on CPU tensors)."""
from hint_b200.householder import HouseholderPerm  # noqa: F401
