"""Modules of the FrEIA shim: the B200-native HINT block under its reference name, ``HouseholderPerm``, and loud stubs for the
baseline couplings (outside the hot path).  ``np`` / ``torch`` leak through the star-import on purpose."""
import numpy as np  # noqa: F401
import torch  # noqa: F401

from hint_b200 import HierarchicalAffineCouplingBlock, HierarchicalAffineCouplingTree, linear_subnet_constructor  # noqa: F401
from .orthogonal import HouseholderPerm  # noqa: F401

__all__ = ["np", "torch", "HierarchicalAffineCouplingBlock", "HierarchicalAffineCouplingTree", "linear_subnet_constructor",
           "HouseholderPerm", "AffineCoupling", "ExternalAffineCoupling", "F_fully_connected"]


def _unpinned(name):
    class _Stub:
        def __init__(self, *a, **k):
            raise NotImplementedError(f"FrEIA.modules.{name}: the baseline couplings of the *_inn_* / *_cinn_* configs are outside "
                                      "hint_b200's hot path and FrEIA's source is not part of the reference (parity unpinned)")
    _Stub.__name__ = name
    return _Stub


AffineCoupling = _unpinned("AffineCoupling")
ExternalAffineCoupling = _unpinned("ExternalAffineCoupling")
F_fully_connected = _unpinned("F_fully_connected")
