import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def golden_names():
    """Fixtures of the un-shuffled block (oracle/gen_golden.py); the reshuffle_* fixtures (oracle/gen_golden_reshuffle.py) have
    their own schema and tests."""
    return sorted(n for n in (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
                  if not n.startswith("reshuffle"))


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    g["meta"] = json.loads(str(g["meta"]))
    return g


def plan_kwargs(meta):
    """ctor kwargs of a fixture -> (d, dc, c_internal, max_splits, min_split_size, clamp)."""
    kw = meta["kwargs"]
    return dict(d=meta["d"], dc=sum(t[0] for t in meta["dims_c"]), c_internal=list(kw.get("c_internal", [])),
                max_splits=kw.get("max_splits", -1), min_split_size=kw.get("min_split_size", 2),
                clamp=kw.get("clamp", 4.0))


@pytest.fixture(params=golden_names())
def golden(request):
    return load_golden(request.param)
