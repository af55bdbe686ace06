"""the reference's OWN scripts and configs, unmodified, against this repo's block.

These tests import the real files from /root/reference (train_unconditional.py, train_conditional.py, data.py, monitoring.py,
configs/**) through tools/run_reference.py: the `FrEIA` shim and `hint` module of this repo on sys.path, compat_stubs/ for the
absent matplotlib / visdom / shapely, synthetic data files of the shapes the loaders expect.  They run only where the
reference checkout exists (this container); the GPU box has no /root/reference, and there the same recipe is covered by
tests/test_freia_shim.py and tests/test_gpu_train.py.

There is no GPU here and the product has no CPU path, so for these tests only:
  * a 'cuda' device request is mapped to the CPU (the configs hard-code `'device': 'cuda'`), and
  * TreePlan.forward / TreePlan.backward are routed to tests/emul (libhint_emul.so): the FP32 kernels' own phase functions
    (simt_phases.cuh) compiled for the host - the same code the GPU runs, test infrastructure, not a product fallback.
What is verified: every import resolves, the configs build their graphs through the shim with the parameter counts the config
comments state, `main(c)` runs epochs end to end (training + test pass + the monitoring calls + model_inverse with the x lane
conditioned on internal y-lane nodes), the loss is finite and decreases."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

REFERENCE = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REFERENCE, "train_unconditional.py")),
                                reason="the reference checkout is not present on this machine")


def _to_cpu_args(args, kwargs):
    def fix(v):
        if isinstance(v, str) and v.startswith("cuda"):
            return "cpu"
        if isinstance(v, torch.device) and v.type == "cuda":
            return torch.device("cpu")
        return v
    return tuple(fix(a) for a in args), {k: fix(v) for k, v in kwargs.items()}


@pytest.fixture
def reference_env(monkeypatch, tmp_path):
    import emul_lib
    from hint_b200.block import TreePlan
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_synthetic_data
    import run_reference

    t_to, m_to = torch.Tensor.to, torch.nn.Module.to

    def tensor_to(self, *a, **k):
        a, k = _to_cpu_args(a, k)
        return t_to(self, *a, **k)

    def module_to(self, *a, **k):
        a, k = _to_cpu_args(a, k)
        return m_to(self, *a, **k)

    monkeypatch.setattr(torch.Tensor, "to", tensor_to)
    monkeypatch.setattr(torch.nn.Module, "to", module_to)

    def args_of(plan):
        return (plan.d, plan.dc, plan.c_internal, plan.clamp, plan.max_splits, plan.min_split_size)

    def forward(self, x, c, flat, rev=False, mode=None):
        out = emul_lib.run(*args_of(self), flat.detach().numpy(), x.detach().numpy(), None if c is None else c.detach().numpy(), rev=rev)
        return torch.from_numpy(out["z"]), torch.from_numpy(out["J"])

    def backward(self, z, c, flat, dz, dJ, mode=None, want_xrec=False, want_dc=True, nll_scale=None, out=None):
        cn = None if c is None else c.detach().numpy()
        x = emul_lib.run(*args_of(self), flat.detach().numpy(), z.detach().numpy(), cn, rev=True)["z"]
        r = emul_lib.run(*args_of(self), flat.detach().numpy(), x, cn, backward=(dz.detach().numpy(), dJ.detach().numpy()))
        dc = torch.from_numpy(r["dc"]) if (self.dc and want_dc) else None
        return torch.from_numpy(r["dx"]), dc, torch.from_numpy(r["dparams"]), torch.from_numpy(r["xrec"])

    monkeypatch.setattr(TreePlan, "forward", forward)
    monkeypatch.setattr(TreePlan, "backward", backward)
    work = str(tmp_path)
    make_synthetic_data.main(work, rows=4000)
    cwd = os.getcwd()
    path = list(sys.path)
    mods = set(sys.modules)
    yield run_reference, work
    os.chdir(cwd)
    sys.path[:] = path
    stubs = os.path.join(ROOT, "compat_stubs")
    for m in set(sys.modules) - mods:      # configs / scripts / data / monitoring / stubs: re-imported fresh by the next test
        f = getattr(sys.modules[m], "__file__", None) or ""
        if f.startswith(REFERENCE) or f.startswith(stubs):
            del sys.modules[m]


def _small_loaders(c, n_train, n_test, batch):
    """The loaders are configuration data (fields of the config namedtuple, like n_epochs): same tensors, fewer rows, so that
    the host emulation finishes in seconds.  The configs' own loaders use batches of 300 .. 10 000."""
    from torch.utils.data import DataLoader, TensorDataset
    tr, te = c.train_loader.dataset.tensors, c.test_loader.dataset.tensors
    return c._replace(train_loader=DataLoader(TensorDataset(*(t[:n_train] for t in tr)), batch_size=batch, shuffle=True, drop_last=True),
                      test_loader=DataLoader(TensorDataset(*(t[:n_test] for t in te)), batch_size=n_test, shuffle=True, drop_last=True))


def test_train_unconditional_runs_unmodified_on_the_miniboone_config(reference_env, capsys):
    run_reference, work = reference_env
    run_reference.prepare_imports(REFERENCE)
    os.chdir(work)
    cfg = importlib.import_module("configs.uci_data.miniboone_hint_8")
    tu = importlib.import_module("train_unconditional")
    c = _small_loaders(cfg.c, 600, 300, 300)._replace(n_epochs=1, pre_low_lr=0, max_batches_per_epoch=2)
    first = float(tu.main(c))
    out = capsys.readouterr().out
    assert "250,624 trainable parameters" in out          # configs/uci_data/miniboone_hint_8.py:31 ("250k")
    assert type(c.model).__module__ == "FrEIA.framework"
    import hint_b200
    hacs = [n.module for n in c.model.node_list if n.name.startswith("hac_")]
    assert len(hacs) == 8 and all(isinstance(m, hint_b200.HierarchicalAffineCouplingBlock) for m in hacs)
    # more epochs of the same loop lower the test loss (init_scale 0: main() keeps the parameters, train_unconditional.py:164-167)
    last = float(tu.main(c._replace(n_epochs=3, init_scale=0)))
    assert np.isfinite(first) and np.isfinite(last) and last < first, (first, last)
    # sampling through the graph's reverse pass (train_unconditional.py:58-62: c.model_inverse)
    with torch.no_grad():
        x = c.model_inverse(torch.randn(7, c.ndim_x))
        z = c.model(x)
        assert x.shape == (7, 42) and torch.isfinite(x).all()
        assert float((c.model_inverse(z) - x).abs().max()) < 1e-3


def test_train_conditional_runs_unmodified_on_the_two_lane_lens_config(reference_env, capsys):
    """configs/lens_shape/conditional_hint_8_full.py: y lane (AffineCoupling) + x lane (HINT block, ExternalAffineCoupling
    conditioned on internal y-lane nodes); train_conditional.py calls node.module.jacobian(None) and model_inverse."""
    run_reference, work = reference_env
    run_reference.prepare_imports(REFERENCE)
    os.chdir(work)
    cfg = importlib.import_module("configs.lens_shape.conditional_hint_8_full")
    tc = importlib.import_module("train_conditional")
    c = _small_loaders(cfg.c, 600, 300, 300)._replace(n_epochs=1, pre_low_lr=0, max_batches_per_epoch=2)
    loss = float(tc.main(c))
    assert np.isfinite(loss)
    names = [n.name for n in c.model.node_list]
    assert "ac_y_to_x_1" in names and "hac_x_8" in names and "perm_y_7" in names
    with torch.no_grad():
        y = torch.randn(5, c.ndim_y)
        x = c.model_inverse(y, torch.randn(5, c.ndim_x))
        z_y, z_x = c.model([y, x])
        y2, x2 = c.model([z_y, z_x], rev=True)
    assert x.shape == (5, 20) and float((x2 - x).abs().max()) < 1e-3 and float((y2 - y).abs().max()) < 1e-4


def test_rejection_sampling_imports_unmodified_and_its_mmd_runs(reference_env, monkeypatch):
    """rejection_sampling.py (the evaluation script north_star names): imports under the environment stubs (data.py, scipy, tqdm,
    matplotlib) with nothing of this repo on its path but the shim; its `multi_mmd` (rejection_sampling.py:56-73) is run on CPU
    tensors (`.cuda()` is the identity here) and compared with a direct evaluation of the same inverse-multiquadric estimator."""
    run_reference, work = reference_env
    run_reference.prepare_imports(REFERENCE)
    os.chdir(work)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    rs = importlib.import_module("rejection_sampling")
    g = torch.Generator().manual_seed(0)
    x, y = torch.randn(64, 5, generator=g), 0.5 + torch.randn(64, 5, generator=g)
    got = float(rs.multi_mmd(x, y))
    d2 = lambda a, b: torch.cdist(a.double(), b.double()).pow(2)
    k = lambda dd: sum(C ** a * ((C + dd) / a) ** -a for C, a in [(0.5, 1), (0.2, 1), (0.2, 0.5)])
    want = float((k(d2(x, x)) + k(d2(y, y)) - 2 * k(d2(x, y))).mean())
    assert abs(got - want) < 1e-4 * max(1.0, abs(want))
    assert float(rs.multi_mmd(x, x)) < 1e-5 < got


CONFIG_DIRS = ["configs/uci_data", "configs/lens_shape", "configs/plus_shape"]


def test_every_config_of_the_reference_builds_through_the_shim(reference_env):
    """All config files import and build their model; parameter counts of the documented ones match the config comments
    (SURVEY.md 8c).  Stale configs that import names the reference itself no longer defines are skipped, and counted."""
    run_reference, work = reference_env
    run_reference.prepare_imports(REFERENCE)
    os.chdir(work)
    expected = {"configs.plus_shape.unconditional_hint_4_3": 1967248, "configs.plus_shape.unconditional_hint_1_full": 199788,
                "configs.plus_shape.unconditional_hint_4_full": 1983464, "configs.plus_shape.unconditional_hint_8_full": 2034176,
                "configs.uci_data.power_hint_8": 499632, "configs.uci_data.power_hint_4": 500856,
                "configs.uci_data.gas_hint_8": 499904, "configs.uci_data.gas_hint_4": 503520,
                "configs.uci_data.miniboone_hint_8": 250624, "configs.uci_data.miniboone_hint_4": 250760,
                "configs.lens_shape.unconditional_hint_1_full": 99298, "configs.plus_shape.conditional_recursive_cinn_4": 4006280}
    built, stale = 0, []
    for d in CONFIG_DIRS:
        for f in sorted(os.listdir(os.path.join(REFERENCE, d))):
            if not f.endswith(".py") or f.startswith("__"):
                continue
            name = d.replace("/", ".") + "." + f[:-3]
            try:
                m = importlib.import_module(name)
            except (ImportError, TypeError, AttributeError) as e:
                # the reference's own stale files (they import the abstract FourierCurveModel) - not the shim's
                if "FrEIA" in str(e) or "hint" in str(e).lower():
                    raise
                stale.append((name, str(e)[:80]))
                continue
            hint_params = [p for n in m.model.node_list if n.name.startswith(("hac", "hint")) and n.module is not None
                           for p in n.module.parameters()]
            if name in expected:
                n_hint = sum(p.numel() for p in hint_params)
                n_all = sum(p.numel() for p in m.model.params_trainable)
                assert expected[name] in (n_hint, n_all), (name, n_hint, n_all)
            built += 1
    assert "configs.plus_shape.unconditional_hint_4_3_reshuffle" in sys.modules      # reshuffle=True builds too
    assert built >= 60, (built, stale)
