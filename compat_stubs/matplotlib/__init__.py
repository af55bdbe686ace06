"""Stand-in for matplotlib (reference: monitoring.py:2-3, data.py:8; plot_data.py and eval_shapes.py are out of scope):
every pyplot call is accepted and draws nothing."""
from . import colors, pyplot  # noqa: F401

__version__ = "0.0-hint_b200-stub"


def use(*args, **kwargs):
    return None
