"""Minimal FrEIA-compatible shim so that the reference's configs and training scripts, which do
``from FrEIA.framework import *`` / ``from FrEIA.modules import *`` (e.g. configs/uci_data/miniboone_hint_8.py:4-5,
train_unconditional.py:4-5), build and run their HINT models on the B200-native coupling block.

Scope: the graph runtime (InputNode / ConditionNode / Node / OutputNode / ReversibleGraphNet), the HINT block (the hot path,
``hint_b200``) and the inter-block ``HouseholderPerm``.  The reference pins NO FrEIA version and ships none of its sources, so
everything in this package except the HINT block follows the published FrEIA definitions and is **parity-unpinned**
(DESIGN.md section 2); the baseline couplings of the 2-lane / `*_inn_*` / `*_cinn_*` configs (AffineCoupling,
ExternalAffineCoupling, F_fully_connected) are plain PyTorch modules in FrEIA.modules.coupling."""
from . import framework, modules  # noqa: F401
