"""Run one of the reference's own training scripts UNMODIFIED against the B200-native block.

    python tools/run_reference.py --reference /path/to/vislearn-HINT --config configs.uci_data.miniboone_hint_8 \\
        [--script train_unconditional] [--epochs 1] [--batches 5] [--mode tf32] [--workdir DIR]

What it does, and nothing else:
  * puts this repo (the `FrEIA` shim, `hint`, `hint_b200`) and the reference checkout on sys.path - the reference's hint.py is
    shadowed by this repo's `hint` module, its scripts / configs / data.py / monitoring.py are imported as they are;
  * appends compat_stubs/ for those of matplotlib / visdom / shapely whose real import fails;
  * chdir's into a work directory that holds the data files the configs load at import (synthetic ones from
    tools/make_synthetic_data.py unless the directory already has them);
  * imports the config module, shortens the run through the config NAMEDTUPLE (`n_epochs`, `max_batches_per_epoch` - the
    reference's own knobs) and calls the script's `main(c)`.
The scripts say `'device': 'cuda'`; a CUDA device is required (hint_b200 has no CPU path)."""
import argparse
import importlib
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def prepare_imports(reference):
    for p in (reference, ROOT):                       # ROOT first: `import hint` / `import FrEIA` resolve to this repo
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    stubs = os.path.join(ROOT, "compat_stubs")
    missing = []
    for pkg in ("matplotlib", "visdom", "shapely"):
        try:
            importlib.import_module(pkg)
        except ImportError:
            missing.append(pkg)
    if missing and stubs not in sys.path:
        sys.path.append(stubs)                        # appended: a real package always wins
    return missing


def run(reference, config, script="train_unconditional", epochs=1, batches=5, mode=None, workdir=None, rows=20000):
    missing = prepare_imports(os.path.abspath(reference))
    if workdir is None:
        workdir = tempfile.mkdtemp(prefix="hint_b200_ref_")
    if not os.path.exists(os.path.join(workdir, "uci_data", "miniboone", "data.npy")):
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import make_synthetic_data
        make_synthetic_data.main(workdir, rows)
    os.chdir(workdir)
    if mode:
        import hint_b200
        hint_b200.set_precision(mode)
    cfg = importlib.import_module(config)
    c = cfg.c
    repl = {}
    if epochs is not None:
        repl["n_epochs"] = int(epochs)
        if hasattr(c, "pre_low_lr"):
            repl["pre_low_lr"] = min(c.pre_low_lr, int(epochs))
    if batches is not None:
        repl["max_batches_per_epoch"] = int(batches)
    c = c._replace(**repl)
    mod = importlib.import_module(script)
    result = mod.main(c)
    return result, c, missing


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True)
    ap.add_argument("--config", required=True, help="e.g. configs.uci_data.miniboone_hint_8")
    ap.add_argument("--script", default="train_unconditional", choices=["train_unconditional", "train_conditional"])
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--batches", type=int, default=5)
    ap.add_argument("--mode", default=None)
    ap.add_argument("--workdir", default=None)
    a = ap.parse_args()
    res, c, missing = run(a.reference, a.config, a.script, a.epochs, a.batches, a.mode, a.workdir)
    print(f"final test loss {float(res):.4f}  (config {c.suffix}; stubbed imports: {missing or 'none'})")
