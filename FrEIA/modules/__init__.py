"""Modules of the FrEIA shim: the B200-native HINT block under its reference name, ``HouseholderPerm``, and the baseline
couplings of the 2-lane / `*_inn_*` / `*_cinn_*` configs (plain PyTorch, off the hot path).  ``np`` / ``torch``
leak through the star-import on purpose."""
import numpy as np  # noqa: F401
import torch  # noqa: F401

from hint_b200 import HierarchicalAffineCouplingBlock, HierarchicalAffineCouplingTree, linear_subnet_constructor  # noqa: F401
from .orthogonal import HouseholderPerm  # noqa: F401
from .coupling import AffineCoupling, ExternalAffineCoupling, F_fully_connected  # noqa: F401

__all__ = ["np", "torch", "HierarchicalAffineCouplingBlock", "HierarchicalAffineCouplingTree", "linear_subnet_constructor",
           "HouseholderPerm", "AffineCoupling", "ExternalAffineCoupling", "F_fully_connected"]
