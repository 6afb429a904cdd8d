"""Expression wrappers — the callers the drop-in stays behind.

``Expression`` mirrors /root/reference/src/Expression.jl:435-520 (``eval_tree_array(ex, X)``,
``eval_grad_tree_array(ex, X)``, ``ex(X)``, ``ex'(X)``); ``ParametricExpression`` mirrors
/root/reference/src/ParametricExpression.jl:361-390: per-sample parameters
``parameters[:, classes[j]]``.  The device path gathers them at operand-fetch time instead
of materialising ``parameters[:, classes]`` and the ``vcat`` onto X (:380-385).
"""
from __future__ import annotations

import numpy as np

from . import device as D
from .evaluate import (EvalContext, _context, _device_of, _is_torch, _prep, _resolve_dtype, _to_host,
                       eval_grad_tree_array, eval_tree_array, validate_input)
from .node import Node
from .operators import OperatorEnum


class Expression:
    """``Expression(tree; operators, variable_names)``"""

    def __init__(self, tree: Node, *, operators: OperatorEnum, variable_names=None):
        self.tree = tree
        self.operators = operators
        self.variable_names = variable_names

    def get_tree(self):
        return self.tree

    def _ops(self, operators):
        return operators if operators is not None else self.operators

    def eval_tree_array(self, X, operators=None, **kws):
        validate_input(self.tree, X)
        return eval_tree_array(self.tree, X, self._ops(operators), **kws)

    def eval_grad_tree_array(self, X, operators=None, **kws):
        validate_input(self.tree, X)
        return eval_grad_tree_array(self.tree, X, self._ops(operators), **kws)

    def __call__(self, X, operators=None, **kws):
        out, ok = self.eval_tree_array(X, operators, **kws)
        if not ok:
            out[...] = float("nan")
        return out

    def gradient(self, X, operators=None, *, variable=True):
        """``ex'(X; variable)`` (src/Expression.jl:483-513)."""
        _, g, ok = self.eval_grad_tree_array(X, operators, variable=variable)
        if not ok:
            g[...] = float("nan")
        return g


def ParametricNode(*args, **kws):
    """Node constructor accepting ``parameter=p`` leaves
    (/root/reference/src/ParametricExpression.jl:52-74)."""
    return Node(*args, **kws)


class ParametricExpression:
    """``ParametricExpression(tree; operators, variable_names, parameters, parameter_names)``;
    ``parameters`` has shape (n_params, n_classes)."""

    def __init__(self, tree: Node, *, operators: OperatorEnum, parameters, variable_names=None,
                 parameter_names=None):
        self.tree = tree
        self.operators = operators
        self.parameters = np.asarray(parameters)
        if self.parameters.ndim != 2:
            raise ValueError("parameters must be (n_params, n_classes)")
        self.variable_names = variable_names
        self.parameter_names = parameter_names

    def get_tree(self):
        return self.tree

    def eval_tree_array(self, X, classes=None, operators=None, *, eval_context=None, **kws):
        if classes is None:
            raise RuntimeError("Incorrect call. You must pass the `classes::Vector` argument "
                               "when calling `eval_tree_array`.")  # :358-360
        out, ok = eval_parametric_trees_array([self], X, classes, operators, eval_context=eval_context,
                                              **kws)
        return out[0], bool(ok[0])

    def __call__(self, X, classes, operators=None, **kws):
        out, ok = self.eval_tree_array(X, classes, operators, **kws)
        if not ok:
            out[...] = float("nan")
        return out

    def eval_grad_tree_array(self, X, classes, operators=None, *, variable=False):
        """``eval_grad_tree_array(convert(Node, ex), vcat(parameters[:, classes], X), operators;
        variable)`` — how the reference differentiates a ParametricExpression
        (/root/reference/src/ParametricExpression.jl:305-350, /root/reference/src/ChainRules.jl:56-77):
        the per-sample parameter rows are the first ``n_params`` feature directions, the rows of X
        follow, then the constants.  Returns (y[N], grad[G, N], complete)."""
        from .evaluate import _mode_of
        X = _prep(X)
        ops = operators if operators is not None else self.operators
        cl = np.asarray(classes.cpu() if _is_torch(classes) else classes)
        assert len(cl) == X.shape[1] and (cl.size == 0 or (cl.min() >= 1 and cl.max() <= self.parameters.shape[1]))
        dt = _resolve_dtype([self.tree], X)
        pop = D.Population([self.tree], ops, dt, ctx=D.Context.get(_device_of(X)), n_params=self.parameters.shape[0])
        out, grad, off, ok = pop.eval_grad_parametric(X, np.asarray(self.parameters, dtype=dt)[None], cl.astype(np.int64) - 1,
                                                      _mode_of(variable))
        N = out.shape[1]
        G = int(off[1]) // N if N else 0
        g = grad.view(N, G).T
        host = not _is_torch(X)
        return _to_host(out[0], host), (g.cpu().numpy() if host else g), bool(ok[0])

    def parameter_gradient(self, X, classes, dY, operators=None):
        """d (sum_j dY[j] * y[j]) / d parameters, shape (n_params, n_classes): the scatter-add of the
        per-sample parameter-row gradients over the samples of each class (what Zygote's pullback
        through ``parameters[:, classes]`` produces)."""
        y, g, ok = self.eval_grad_tree_array(X, classes, operators, variable=True)
        n_params, n_classes = self.parameters.shape
        g = np.asarray(g.cpu() if _is_torch(g) else g)[:n_params]
        cl = np.asarray(classes.cpu() if _is_torch(classes) else classes).astype(np.int64) - 1
        out = np.zeros((n_params, n_classes), dtype=g.dtype)
        np.add.at(out.T, cl, (g * np.asarray(dY)[None, :]).T)
        if not ok:
            out[...] = float("nan")
        return out


def eval_parametric_trees_array(exprs, X, classes, operators=None, *, eval_context=None, **kws):
    """Batched ParametricExpression evaluation: every expression has its own
    ``parameters`` (same shape); ``classes`` is 1-based like the reference.
    Returns (out[P, N], ok[P])."""
    ctx = _context(eval_context, kws)
    X = _prep(X)
    exprs = list(exprs)
    ops = operators if operators is not None else exprs[0].operators
    classes = np.asarray(classes.cpu() if _is_torch(classes) else classes)
    assert len(classes) == X.shape[1]                                   # :378
    n_classes = exprs[0].parameters.shape[1]
    assert classes.size == 0 or classes.max() <= n_classes              # :379
    assert classes.size == 0 or classes.min() >= 1
    dt = _resolve_dtype([e.tree for e in exprs], X)
    pop = D.Population([e.tree for e in exprs], ops, dt, ctx=D.Context.get(_device_of(X)),
                       bumper=ctx.bumper, use_fused=ctx.use_fused, n_params=exprs[0].parameters.shape[0])
    params = np.stack([np.asarray(e.parameters, dtype=dt) for e in exprs])
    out, ok = pop.eval_parametric(X, params, classes.astype(np.int64) - 1, early_exit=ctx.early_exit)
    host = not _is_torch(X)
    return _to_host(out, host), (_to_host(ok, host).astype(bool) if host else ok.bool())


__all__ = ["Expression", "ParametricExpression", "ParametricNode", "eval_parametric_trees_array",
           "EvalContext"]
