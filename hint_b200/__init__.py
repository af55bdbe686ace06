"""hint_b200 — B200-native implementation of HINT's recursive affine coupling block (hint.py of vislearn/HINT).

Importing this package loads (building in-tree if needed) the sm_100a shared library; there is no CPU path.
"""
from . import _lib
from .block import (HierarchicalAffineCouplingBlock, HierarchicalAffineCouplingTree, TreePlan,
                    linear_subnet_constructor, set_precision, get_precision, set_backward_check, get_backward_check)
from .model import HintFlow, GraphedFlow, nll_loss
from .parallel import BucketedGradAllReduce, broadcast_parameters, shard_rows
from .train import FusedClampAdam, FusedTrainStep, add_noise, nll_loss_fused, multi_mmd

__version__ = _lib.load().hint_version().decode()

__all__ = ["HierarchicalAffineCouplingBlock", "HierarchicalAffineCouplingTree", "TreePlan",
           "linear_subnet_constructor", "set_precision", "get_precision", "set_backward_check", "get_backward_check", "HintFlow", "GraphedFlow", "nll_loss",
           "BucketedGradAllReduce", "broadcast_parameters", "shard_rows",
           "FusedClampAdam", "FusedTrainStep", "add_noise", "nll_loss_fused", "multi_mmd"]
