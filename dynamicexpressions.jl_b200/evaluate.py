"""The reference's evaluation API, served by the B200 kernels.

Same names, argument meaning and return conventions as
/root/reference/src/Evaluate.jl:279-309 (``eval_tree_array``),
/root/reference/src/EvaluateDerivative.jl:40-53, 193-228 (``eval_diff_tree_array``,
``eval_grad_tree_array``) and /root/reference/src/EvaluationHelpers.jl:29-33, 56-62
(``tree(X, operators)``, ``tree'(X, operators)``), with 1-based indices where Julia has
them.  ``X`` has shape (nfeatures, nsamples).  numpy in -> numpy out; torch CUDA tensor
in -> torch CUDA tensors out.

Each single-tree call packs a population of one; the batched forms
(``eval_trees_array`` etc., or :class:`dexb200.device.Population` directly) are what a
caller with many trees should use — that is the point of the device path.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import device as D
from .node import Node, count_constant_nodes, max_feature, tree_dtype
from .operators import OperatorEnum


@dataclass(frozen=True)
class EvalContext:
    """Mirror of ``EvalContext`` (/root/reference/src/Evaluate.jl:156-181).

    ``turbo`` is accepted for signature compatibility and ignored: it selects the
    LoopVectorization CPU kernels in the reference and has no meaning on the device.
    ``bumper=True`` selects the *semantics* of the Bumper evaluator (unfused validity
    checks, /root/reference/ext/DynamicExpressionsBumperExt.jl:11-89).  ``buffer`` is not
    needed: temporaries never leave the SM."""
    turbo: bool = False
    bumper: bool = False
    early_exit: bool = True
    use_fused: bool = True
    buffer: object = None

    def __post_init__(self):
        if self.bumper and self.buffer is not None:
            raise AssertionError("bumper and buffer are mutually exclusive")  # :176-178


EvalOptions = EvalContext  # deprecated alias (:183)


def _context(eval_context, kws):
    bad = set(kws) - {"turbo", "bumper", "eval_options"}
    if bad:
        raise ValueError(f"Invalid keyword argument(s): {sorted(bad)}")  # :206-208
    if "eval_options" in kws:
        assert eval_context is None, "Cannot use both `eval_context` and deprecated `eval_options`."
        eval_context = kws["eval_options"]
    if kws.get("turbo") is not None or kws.get("bumper") is not None:
        assert eval_context is None, \
            "Cannot use both `eval_context` and deprecated flags `turbo` and `bumper`."
    if eval_context is not None:
        return eval_context
    return EvalContext(turbo=bool(kws.get("turbo") or False), bumper=bool(kws.get("bumper") or False))


def _is_torch(x):
    return not isinstance(x, (np.ndarray, list, tuple))


def _resolve_dtype(trees, X):
    """Promotion rule of src/Evaluate.jl:317-327: promote_type(tree T, X T)."""
    xdt = np.dtype(str(X.dtype).replace("torch.", "")).type
    tdt = None
    for t in trees:
        tdt = tree_dtype(t)
        if tdt is not None:
            break
    if tdt is None:
        tdt = xdt
    if xdt not in (np.float32, np.float64):
        raise TypeError(f"the device path evaluates Float32/Float64 only, got X of {xdt}")
    return np.promote_types(tdt, xdt).type


def _prep(X):
    if isinstance(X, (list, tuple)):
        X = np.asarray(X)
    if isinstance(X, np.ndarray) and X.ndim == 1:
        X = X.reshape(-1, 1)  # vector overload, :311-315
    elif _is_torch(X) and X.dim() == 1:
        X = X.reshape(-1, 1)
    return X


def _device_of(X):
    if _is_torch(X) and X.is_cuda:
        return X.device.index
    return None


def _to_host(t, like_numpy):
    return t.cpu().numpy() if like_numpy else t


def eval_trees_array(trees, X, operators: OperatorEnum, *, eval_context=None, **kws):
    """Batched ``[eval_tree_array(t, X, operators) for t in trees]`` in one launch:
    returns (out[P, N], ok[P])."""
    ctx = _context(eval_context, kws)
    X = _prep(X)
    trees = list(trees)
    dt = _resolve_dtype(trees, X)
    pop = D.Population(trees, operators, dt, ctx=D.Context.get(_device_of(X)), bumper=ctx.bumper,
                       use_fused=ctx.use_fused)
    out, ok = pop.eval(X, early_exit=ctx.early_exit)
    host = not _is_torch(X)
    return _to_host(out, host), (_to_host(ok, host).astype(bool) if host else ok.bool())


def eval_tree_array(tree: Node, X, operators: OperatorEnum, *, eval_context=None, **kws):
    """``(output, complete) = eval_tree_array(tree, cX, operators; eval_context)``
    (/root/reference/src/Evaluate.jl:279-309)."""
    out, ok = eval_trees_array([tree], X, operators, eval_context=eval_context, **kws)
    return out[0], bool(ok[0])


def _mode_of(variable):
    if variable is True or variable == "features":
        return D.GRAD_FEATURES
    if variable is False or variable == "constants":
        return D.GRAD_CONSTANTS
    if variable in ("both", ":both"):
        return D.GRAD_BOTH
    raise ValueError("variable must be True, False or 'both'")


def eval_grad_trees_array(trees, X, operators: OperatorEnum, *, variable=False, turbo=False):
    """Batched ``eval_grad_tree_array``: (out[P, N], [grad_t (G_t, N)], ok[P])."""
    X = _prep(X)
    trees = list(trees)
    dt = _resolve_dtype(trees, X)
    pop = D.Population(trees, operators, dt, ctx=D.Context.get(_device_of(X)))
    mode = _mode_of(variable)
    out, grad, off, ok = pop.eval_grad(X, mode)
    N = out.shape[1]
    grads = []
    for t in range(pop.n_trees):
        n = int(off[t + 1] - off[t])
        G = n // N if N else 0
        grads.append(grad[int(off[t]):int(off[t + 1])].view(N, G).T)
    host = not _is_torch(X)
    if host:
        grads = [g.cpu().numpy() for g in grads]
    return _to_host(out, host), grads, (_to_host(ok, host).astype(bool) if host else ok.bool())


def eval_grad_tree_array(tree: Node, X, operators: OperatorEnum, *, variable=False, turbo=False):
    """``(evaluation, gradient, complete)``; gradient is (G, N) with G = nfeatures
    (``variable=True``), count_constant_nodes(tree) (``variable=False``) or both, features
    first (/root/reference/src/EvaluateDerivative.jl:193-228)."""
    out, grads, ok = eval_grad_trees_array([tree], X, operators, variable=variable, turbo=turbo)
    return out[0], grads[0], bool(ok[0])


def eval_diff_tree_array(tree: Node, X, operators: OperatorEnum, direction: int, *, turbo=False):
    """``(evaluation, derivative, complete)`` along feature ``direction`` (1-based)
    (/root/reference/src/EvaluateDerivative.jl:40-53); never reports failure."""
    X = _prep(X)
    dt = _resolve_dtype([tree], X)
    pop = D.Population([tree], operators, dt, ctx=D.Context.get(_device_of(X)))
    out, dout, ok = pop.eval_diff(X, int(direction) - 1)
    host = not _is_torch(X)
    return _to_host(out[0], host), _to_host(dout[0], host), bool(ok[0])


def call_tree(tree: Node, X, operators: OperatorEnum, **kws):
    """``tree(X, operators)``: NaN-filled on failure
    (/root/reference/src/EvaluationHelpers.jl:29-33)."""
    out, ok = eval_tree_array(tree, X, operators, **kws)
    if not ok:
        out[...] = float("nan")  # set_nan!, src/Utils.jl:73-76
    return out


def grad_tree(tree: Node, X, operators: OperatorEnum, *, variable=True, **kws):
    """``tree'(X, operators; variable)``: NaN-filled gradient on failure
    (/root/reference/src/EvaluationHelpers.jl:56-62, 90-91)."""
    _, grad, ok = eval_grad_tree_array(tree, X, operators, variable=variable)
    if not ok:
        grad[...] = float("nan")
    return grad


def validate_input(tree_or_trees, X):
    """``_validate_input`` (/root/reference/src/Expression.jl:401-409)."""
    X = _prep(X)
    if len(X.shape) != 2:
        raise AssertionError("X must be a matrix")
    trees = [tree_or_trees] if isinstance(tree_or_trees, Node) else list(tree_or_trees)
    mf = max((max_feature(t) for t in trees), default=0)
    if mf > X.shape[0]:
        raise AssertionError(f"expression uses feature x{mf} but X has {X.shape[0]} rows")


__all__ = ["EvalContext", "EvalOptions", "eval_tree_array", "eval_trees_array",
           "eval_grad_tree_array", "eval_grad_trees_array", "eval_diff_tree_array", "call_tree",
           "grad_tree", "validate_input", "count_constant_nodes"]
