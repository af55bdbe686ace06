"""Import shim: lets ``import hint`` / ``from hint import HierarchicalAffineCouplingBlock`` resolve to the
B200-native implementation, with the public names of the reference module of the same name."""
from hint_b200 import (HierarchicalAffineCouplingBlock, HierarchicalAffineCouplingTree,  # noqa: F401
                       linear_subnet_constructor)
