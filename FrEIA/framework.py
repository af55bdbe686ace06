"""Graph runtime of the FrEIA shim: the node types and ``ReversibleGraphNet`` with the call surface the reference uses
(configs/**: ``Node(inputs, module_type, module_args, conditions=, name=)``, ``ReversibleGraphNet(node_list, verbose=)``;
train_unconditional.py:124-125: ``model(x)``, ``model.log_jacobian(x, run_forward=False)``; train_conditional.py:50-55:
``node.module.jacobian(None)``; model_inverse: ``model(z, rev=True)``).  Pure host logic - the arithmetic lives in the modules.

The star-import deliberately leaks ``np`` and ``torch``: the reference scripts use ``np`` without importing it
(train_unconditional.py:95,158,198)."""
import numpy as np  # noqa: F401  (leaked on purpose)
import torch
import torch.nn as nn

__all__ = ["np", "torch", "nn", "Node", "InputNode", "ConditionNode", "OutputNode", "ReversibleGraphNet"]


def _as_list(v):
    if v is None:
        return []
    return list(v) if isinstance(v, (list, tuple)) and not (len(v) == 2 and isinstance(v[1], int) and isinstance(v[0], Node)) else [v]


class Node:
    """One invertible module of the graph.  ``inputs``: a Node, a ``(Node, output_index)`` pair, or a list of those;
    ``conditions``: a Node / ConditionNode or a list of them (their first output is passed as ``c=[...]``)."""

    def __init__(self, inputs, module_type, module_args, conditions=[], name=None):
        self.inputs = [(i, 0) if isinstance(i, Node) else (i[0], int(i[1])) for i in _as_list(inputs)]
        self.conditions = [c if isinstance(c, Node) else c[0] for c in _as_list(conditions)]
        self.module_type, self.module_args = module_type, dict(module_args or {})
        self.name = name if name is not None else f"node_{id(self) & 0xffff:04x}"
        self.module = None
        self.input_dims, self.output_dims = None, None
        for k in range(8):                       # FrEIA exposes node.out0, node.out1, ... as (node, index) handles
            setattr(self, f"out{k}", (self, k))

    def build_modules(self):
        self.input_dims = [n.output_dims[i] for n, i in self.inputs]
        if self.conditions:
            cond_dims = [c.output_dims[0] for c in self.conditions]
            self.module = self.module_type(self.input_dims, dims_c=cond_dims, **self.module_args)
        else:
            self.module = self.module_type(self.input_dims, **self.module_args)
        self.output_dims = self.module.output_dims(self.input_dims)
        return self.module


class InputNode(Node):
    def __init__(self, *dims, name="node"):
        super().__init__([], None, {}, name=name)
        self.output_dims = [tuple(int(d) for d in dims)]

    def build_modules(self):
        return None


class ConditionNode(InputNode):
    pass


class OutputNode(Node):
    def __init__(self, inputs, name="node"):
        super().__init__(inputs, None, {}, name=name)

    def build_modules(self):
        self.input_dims = [n.output_dims[i] for n, i in self.inputs]
        self.output_dims = list(self.input_dims)
        return None


class ReversibleGraphNet(nn.Module):
    """Executes the node list (given in a valid topological order, as every reference config writes it) forward, or backward
    with every module called with ``rev=True``.  Inputs / outputs / conditions are ordered as their nodes appear in the list."""

    def __init__(self, node_list, ind_in=None, ind_out=None, verbose=True):
        super().__init__()
        self.node_list = list(node_list)
        self.in_nodes = [n for n in self.node_list if isinstance(n, InputNode) and not isinstance(n, ConditionNode)]
        self.cond_nodes = [n for n in self.node_list if isinstance(n, ConditionNode)]
        self.out_nodes = [n for n in self.node_list if isinstance(n, OutputNode)]
        # Build in dependency order: the reference's 2-lane configs list the whole y lane before the x lane, but nothing stops a
        # config from listing a node before its condition lane, so the order is resolved here instead of being assumed.
        built, order = set(), []
        pending = list(self.node_list)
        while pending:
            progress = False
            for n in list(pending):
                if all(src in built for src, _ in n.inputs) and all(src in built for src in n.conditions):
                    n.build_modules()
                    built.add(n); order.append(n); pending.remove(n); progress = True
            assert progress, "cyclic or dangling node references: " + ", ".join(n.name for n in pending)
        self._order = order
        # one entry per node, None for the input / output nodes: state_dict keys are module_list.<position in node_list>.*
        self.module_list = nn.ModuleList([n.module for n in self.node_list])
        self._values = None
        self._rev = False
        if verbose:
            for n in self.node_list:
                print(f"{n.name}: {n.input_dims} -> {n.output_dims}")

    @staticmethod
    def _tensors(v):
        if v is None:
            return []
        return list(v) if isinstance(v, (list, tuple)) else [v]

    def forward(self, x, c=None, rev=False, intermediate_outputs=False):
        xs, cs = self._tensors(x), self._tensors(c)
        vals = {}
        assert len(cs) == len(self.cond_nodes), f"expected {len(self.cond_nodes)} condition tensors, got {len(cs)}"
        for n, t in zip(self.cond_nodes, cs):
            vals[(n, 0)] = t
        if not rev:
            assert len(xs) == len(self.in_nodes), f"expected {len(self.in_nodes)} input tensors, got {len(xs)}"
            for n, t in zip(self.in_nodes, xs):
                vals[(n, 0)] = t
            for n in self._order:
                if isinstance(n, InputNode):
                    continue
                ins = [vals[k] for k in n.inputs]
                if isinstance(n, OutputNode):
                    vals[(n, 0)] = ins[0]
                    continue
                kw = {"c": [vals[(cn, 0)] for cn in n.conditions]} if n.conditions else {}
                outs = n.module(ins, rev=False, **kw)
                for k, o in enumerate(outs):
                    vals[(n, k)] = o
            result = [vals[(n, 0)] for n in self.out_nodes]
        else:
            assert len(xs) == len(self.out_nodes), f"expected {len(self.out_nodes)} output tensors, got {len(xs)}"
            for n, t in zip(self.out_nodes, xs):
                vals[n.inputs[0]] = t
            # A node runs backwards once all its outputs AND all its conditions are known.  A condition may be an internal node of
            # another lane (configs/lens_shape/conditional_hint_8_full.py:78-83: the x lane is conditioned on the y lane), whose
            # value only appears when that lane has been reversed far enough - so this is a worklist, not a fixed order.
            pending = [n for n in reversed(self._order) if not isinstance(n, (InputNode, OutputNode))]
            while pending:
                progress = False
                for n in list(pending):
                    if all((n, k) in vals for k in range(len(n.output_dims))) and all((cn, 0) in vals for cn in n.conditions):
                        outs = [vals[(n, k)] for k in range(len(n.output_dims))]
                        kw = {"c": [vals[(cn, 0)] for cn in n.conditions]} if n.conditions else {}
                        ins = n.module(outs, rev=True, **kw)
                        for key, t in zip(n.inputs, ins):
                            vals[key] = t
                        pending.remove(n); progress = True
                assert progress, "reverse pass is stuck on: " + ", ".join(n.name for n in pending)
            result = [vals[(n, 0)] for n in self.in_nodes]
        self._values, self._rev = vals, bool(rev)
        if intermediate_outputs:
            return {(n.name, k): t for (n, k), t in vals.items()}
        return result[0] if len(result) == 1 else tuple(result)

    def log_jacobian(self, x=None, c=None, rev=False, run_forward=True, intermediate_outputs=False):
        """Sum of the modules' log|det J| (each module caches the value of its last call, hint.py:124-129)."""
        if run_forward or self._values is None:
            self.forward(x, c, rev=rev)
        total = 0
        per_node = {}
        for n in self.node_list:
            if n.module is None:
                continue
            kw = {"c": [self._values[(cn, 0)] for cn in n.conditions]} if n.conditions else {}
            j = n.module.jacobian([self._values.get(k) for k in n.inputs], rev=self._rev, **kw)
            per_node[n.name] = j
            total = total + j
        return per_node if intermediate_outputs else total
