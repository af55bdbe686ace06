"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol the header
declares, and its planner (tree / parameter layout, integer work -> bit-exact) agrees with the oracle and with
the reference's state_dict layout stored in the golden fixtures.  No compute call is made here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, plan_kwargs
from oracle import hint_oracle as O

import hint_b200
from hint_b200 import _lib, HierarchicalAffineCouplingBlock, TreePlan


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "hint_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(hint_[a-z_0-9]+)\s*\(", header)))
    assert declared == sorted(_lib.EXPORTS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in _lib.load().hint_version()


@pytest.mark.parametrize("kat", O.PARAM_COUNT_KATS, ids=lambda k: k["name"])
def test_planner_matches_oracle_and_param_kats(kat):
    plan = TreePlan(kat["d"], kat["dc"], kat["c_internal"], 4.0, kat["max_splits"], 2, False)
    assert plan.n_params == kat["per_block"]
    oplan = O.build_plan(kat["d"], kat["dc"], kat["c_internal"], kat["max_splits"])
    assert len(plan.nodes) == len(oplan)
    for a, b in zip(plan.nodes, oplan):
        for f in ("depth", "lo", "hi", "k", "cin", "h", "cout", "parent", "upper", "lower"):
            assert a[f] == getattr(b, f), (f, a, b)
        assert bool(a["leaf"]) == b.leaf
    assert plan.flops_per_sample == O.flops_per_sample(oplan)
    mine = [(n, o, tuple(s)) for n, o, s in plan.entries]
    ref = [(e.name, e.offset, tuple(e.shape)) for e in O.param_entries(oplan)]
    assert mine == ref


def test_state_dict_layout_matches_reference(golden):
    meta = golden["meta"]
    pk = plan_kwargs(meta)
    blk = HierarchicalAffineCouplingBlock([(meta["d"],)], dims_c=[tuple(t) for t in meta["dims_c"]], **meta["kwargs"])
    sd = blk.state_dict()
    assert [(k, list(v.shape)) for k, v in sd.items()] == [(k, s) for k, s in meta["state_dict"]]
    assert blk.plan.n_params == golden["params"].size
    # loading a reference-layout state_dict fills the flat parameter in parameters() order
    ref_sd, off = {}, 0
    for k, s in meta["state_dict"]:
        n = int(np.prod(s))
        ref_sd[k] = torch.from_numpy(golden["params"][off:off + n].reshape(s).copy())
        off += n
    blk.load_state_dict(ref_sd)
    assert np.array_equal(blk.flat.detach().numpy(), golden["params"])
    assert [p.requires_grad for p in blk.parameters()] == [True]
    t = blk.tree
    assert t.s[0].weight.shape == tuple(meta["state_dict"][0][1])
    with pytest.raises(RuntimeError):
        blk.load_state_dict({k: v for k, v in list(ref_sd.items())[1:]})


def test_ctor_errors_mirror_reference_surface():
    with pytest.raises(AssertionError):
        HierarchicalAffineCouplingBlock([(8, 4, 4)], dims_c=[(2, 3, 3)])
    with pytest.raises(NotImplementedError):
        HierarchicalAffineCouplingBlock([(8,)], conv=True)
    with pytest.raises(NotImplementedError):
        HierarchicalAffineCouplingBlock([(8,)], subnet_constructor=lambda a, b, c: None)
    with pytest.raises(NotImplementedError):
        HierarchicalAffineCouplingBlock([(200,)], c_internal=[8], reshuffle=True)   # the composed mixing kernel covers d <= 128
    assert HierarchicalAffineCouplingBlock([(8,)], reshuffle=True).perm_M.shape == (8, 8)
    with pytest.raises(ValueError):
        HierarchicalAffineCouplingBlock([(1,)])
    blk = HierarchicalAffineCouplingBlock([(8,)], subnet_constructor=hint_b200.linear_subnet_constructor)
    assert blk.output_dims([(8,)]) == [(8,)]
    with pytest.raises(AssertionError):
        blk.output_dims([(8,), (8,)])
    with pytest.raises(RuntimeError, match="CUDA"):
        blk([torch.randn(3, 8)])   # no CPU fallback


def test_default_init_distribution_matches_nn_linear():
    torch.manual_seed(0)
    blk = HierarchicalAffineCouplingBlock([(100,)], c_internal=[64, 32])
    v = blk.named_views()
    w = v["tree.s.2.weight"]
    bound = 1 / np.sqrt(64)
    assert w.abs().max() <= bound and w.abs().max() > 0.9 * bound
    assert abs(float(w.std()) - bound / np.sqrt(3)) < 0.05 * bound


def test_import_shim_names():
    import hint
    assert hint.HierarchicalAffineCouplingBlock is HierarchicalAffineCouplingBlock
    assert callable(hint.linear_subnet_constructor)


@pytest.mark.parametrize("name", ["reshuffle_d13", "reshuffle_cond_d10_ms1"])
def test_reshuffle_composes_into_one_matrix_in_front_of_the_tree(name):
    """hint.py:36-39,64-65,93-94 with reshuffle=True, golden vectors from the REAL module (oracle/gen_golden_reshuffle.py, the
    published HouseholderPerm definition injected): the block's composed matrix perm_M followed by the UN-shuffled tree (oracle)
    reproduces the reference's forward, log-det and inverse to 1e-12 (fp64), and the reference's state_dict names round-trip."""
    from conftest import load_golden
    g = load_golden(name)
    meta = g["meta"]
    kw = dict(meta["kwargs"])
    blk = HierarchicalAffineCouplingBlock([(meta["d"],)], dims_c=[tuple(t) for t in meta["dims_c"]], **kw)
    sd = blk.state_dict()
    perm_keys = sorted(k for k in sd if k.endswith("perm.Vs"))
    assert perm_keys == sorted(k[3:] for k in g if k.startswith("vs:"))
    for k in perm_keys:
        sd[k] = torch.from_numpy(g["vs:" + k]).float()
    blk.load_state_dict(sd)
    assert blk.plan.n_params == g["params"].size       # the fixed reflections are not trainable parameters
    # compose in fp64 from the loaded reflections (perm_M itself is the fp32 copy the kernels use)
    M = torch.eye(meta["d"], dtype=torch.float64)
    depth = max(nd["depth"] for nd in blk.plan.nodes)
    for lv in range(depth + 1):
        D = torch.eye(meta["d"], dtype=torch.float64)
        for nd, path in zip(blk.plan.nodes, [blk.plan.paths[i] for i in range(len(blk.plan.nodes))]):
            if nd["depth"] != lv:
                continue
            n = nd["hi"] - nd["lo"]
            W = torch.eye(n, dtype=torch.float64)
            for v in torch.from_numpy(g["vs:" + path + ".perm.Vs"]):
                W = W - 2.0 * torch.outer(W @ v, v) / torch.dot(v, v)
            D[nd["lo"]:nd["hi"], nd["lo"]:nd["hi"]] = W
        M = M @ D
    assert float((blk.perm_M.double() - M).abs().max()) < 1e-6
    pk = plan_kwargs(meta)
    plan = O.build_plan(pk["d"], pk["dc"], pk["c_internal"], pk["max_splits"], pk["min_split_size"])
    f64 = torch.from_numpy(g["params"])
    x = torch.from_numpy(g["x"])
    c = torch.from_numpy(g["c"]) if "c" in g else None
    z, J = O.forward(plan, f64, x @ M, c, clamp=pk["clamp"])
    assert float((z - torch.from_numpy(g["z64"])).abs().max()) < 1e-12 and float((J - torch.from_numpy(g["J64"])).abs().max()) < 1e-12
    xi, Ji = O.forward(plan, f64, x, c, rev=True, clamp=pk["clamp"])
    assert float((xi @ M.t() - torch.from_numpy(g["xinv64"])).abs().max()) < 1e-12
    assert float((Ji - torch.from_numpy(g["Jinv64"])).abs().max()) < 1e-12


def test_widening_entry_points_validate_their_arguments_without_a_gpu():
    """Argument checks of the training-edge / Householder / coupling / MMD entry points return the documented status codes before
    any CUDA call (no device needed), and the size queries answer on the host."""
    import ctypes
    lib = _lib.load()
    vp16 = (ctypes.c_void_p * 16)()
    one = ctypes.c_void_p(16)      # a non-null placeholder pointer: rejected calls never dereference it
    assert lib.hint_mlp_coupling_supported(2, 20, 68) == 1
    assert lib.hint_mlp_coupling_supported(0, 20, 68) == 0 and lib.hint_mlp_coupling_supported(2, 129, 68) == 0
    assert lib.hint_mlp_coupling_supported(2, 20, 257) == 0
    assert lib.hint_mlp_coupling_forward(one, 2, one, 200, 68, vp16, 5.0, 0, 10, one, one, None) == _lib.HINT_ERR_UNSUPPORTED
    assert lib.hint_mlp_coupling_forward(one, 2, one, 20, 68, vp16, 5.0, 0, 10, one, one, None) == _lib.HINT_ERR_INVALID   # null parameters
    assert "null parameter" in _lib.last_error()
    assert lib.hint_mlp_coupling_forward(one, 2, one, 20, 68, vp16, 5.0, 0, -1, one, one, None) == _lib.HINT_ERR_INVALID
    assert lib.hint_mlp_coupling_workspace_bytes(2, 20, 68, 1000) >= 4 * 2 * 1000 * (6 * 68 + 20)
    assert lib.hint_mlp_coupling_workspace_bytes(2, 200, 68, 1000) == 0
    full16 = (ctypes.c_void_p * 16)(*[16] * 16)
    assert lib.hint_mlp_coupling_backward(one, 2, one, 20, 68, full16, 5.0, 10, one, None, one, one, full16, one, 8, None) == _lib.HINT_ERR_WORKSPACE
    f3 = (ctypes.c_float * 3)(0.5, 0.2, 0.2)
    bad = (ctypes.c_float * 3)(0.5, -0.2, 0.2)
    assert lib.hint_mmd_workspace_bytes(4000) >= 8 * 3 * 63 * 63
    assert lib.hint_multi_mmd(one, one, 0, 20, f3, f3, 3, one, one, 1 << 20, None) == _lib.HINT_ERR_INVALID
    assert lib.hint_multi_mmd(one, one, 100, 20, f3, bad, 3, one, one, 1 << 20, None) == _lib.HINT_ERR_INVALID
    assert lib.hint_multi_mmd(one, one, 100, 20, f3, f3, 9, one, one, 1 << 20, None) == _lib.HINT_ERR_INVALID
    assert lib.hint_multi_mmd(one, one, 4000, 20, f3, f3, 3, one, one, 8, None) == _lib.HINT_ERR_WORKSPACE
    assert lib.hint_householder_apply(one, one, 10, 129, 0, one, None) == _lib.HINT_ERR_INVALID
    assert lib.hint_householder_apply(one, one, 10, 20, 0, one, None) == _lib.HINT_ERR_INVALID        # y aliases x
    assert lib.hint_householder_wgrad_workspace_bytes(43) > 0 and lib.hint_householder_wgrad_workspace_bytes(0) == 0
