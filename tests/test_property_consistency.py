"""Property tests after /root/reference/test/test_supposition_consistency.jl:16-104 (generator:
test/supposition_utils.jl:14-62): random trees over ((abs, cos, exp), (+, -, *, /), (fma, clamp, +, max)),
any finite constants, X 5 x (1..16) of any finite values; whenever the evaluation is `complete`
it must equal an INDEPENDENT evaluation of the same expression.

  * CPU: the oracle against a plain numpy recursion over the tree (the reference evaluates the
    printed expression with Julia itself) — one more pin of the oracle, now on ternary operators,
    deep nesting and extreme magnitudes;
  * GPU: the device against the oracle on populations of such trees (flags exactly, values in the
    classes of tests/parity_util.py), Float32 and Float64, ragged tiny sample counts.
"""
import math

import numpy as np
import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import HealthCheck, given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402

import dexb200  # noqa: E402

N_FEATURES = 5
OPS = {1: ("abs", "cos", "exp"), 2: ("+", "-", "*", "/"), 3: ("fma", "clamp", "+", "max")}

def _wrap(child):
    return st.one_of(
        st.tuples(st.just(1), st.integers(1, 3), child),
        st.tuples(st.just(2), st.integers(1, 4), child, child),
        st.tuples(st.just(3), st.integers(1, 4), child, child, child),
    )


def strategies(width):
    """(trees, matrices) over the finite floats of the given width (Data.Floats{T}(nans = false, infs = false)).
    A tree is nested tuples: ("c", value) | ("x", feature) | (degree, op index (1-based), children...)."""
    finite = st.floats(allow_nan=False, allow_infinity=False, width=width)
    leaves = st.one_of(finite.map(lambda v: ("c", v)), st.integers(1, N_FEATURES).map(lambda i: ("x", i)))
    tr = st.recursive(leaves, _wrap, max_leaves=24)
    mats = st.integers(1, 16).flatmap(
        lambda bs: st.lists(finite, min_size=N_FEATURES * bs, max_size=N_FEATURES * bs).map(
            lambda v: np.asarray(v, dtype=np.float64).reshape(N_FEATURES, bs)))
    return tr, mats


trees, matrices = strategies(64)


def to_node(t, dtype):
    N_ = dexb200.Node
    if t[0] == "c":
        return N_(val=float(np.asarray(t[1], dtype=dtype)), T=dtype)
    if t[0] == "x":
        return N_(feature=t[1], T=dtype)
    kids = tuple(to_node(c, dtype) for c in t[2:])
    return N_(op=t[1], children=kids) if len(kids) == 3 else N_(t[1], *kids)


def numpy_eval(t, X):
    """The expression evaluated by plain numpy broadcasting in float64 (Julia semantics of the
    operators: fma exactly rounded, clamp(x, lo, hi) = x > hi ? hi : (x < lo ? lo : x), NaN-propagating max)."""
    with np.errstate(all="ignore"):
        if t[0] == "c":
            return np.full(X.shape[1], t[1], dtype=np.float64)
        if t[0] == "x":
            return X[t[1] - 1].copy()
        a = [numpy_eval(c, X) for c in t[2:]]
        d, op = t[0], t[1]
        if d == 1:
            return (np.abs, np.cos, np.exp)[op - 1](a[0])
        if d == 2:
            return (np.add, np.subtract, np.multiply, np.divide)[op - 1](a[0], a[1])
        x, y, z = a
        if op == 1:     # fma: one rounding (long double carries the exact product of two doubles' leading bits)
            return np.array([_fma(p, q, r) for p, q, r in zip(x, y, z)])
        if op == 2:
            return np.where(x > z, z, np.where(x < y, y, x))
        if op == 3:
            return x + y + z
        m = np.maximum(np.maximum(x, y), z)
        return m


def _fma(a, b, c):
    if hasattr(math, "fma"):
        try:
            return math.fma(a, b, c)
        except (OverflowError, ValueError):
            pass
    with np.errstate(all="ignore"):
        return float(np.longdouble(a) * np.longdouble(b) + np.longdouble(c))


def approx(a, b):
    """Julia's `≈` for vectors: norm(a - b) <= sqrt(eps) * max(norm(a), norm(b))."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    s = max(float(np.max(np.abs(a), initial=0.0)), float(np.max(np.abs(b), initial=0.0)), 1e-300)
    return np.linalg.norm(a / s - b / s) <= math.sqrt(np.finfo(np.float64).eps) * max(np.linalg.norm(a / s), np.linalg.norm(b / s))


@settings(max_examples=400, deadline=None, suppress_health_check=list(HealthCheck))
@given(tree=trees, X=matrices)
def test_oracle_equals_plain_numpy_evaluation(tree, X, oracle):
    ops = dexb200.OperatorEnum(OPS)
    wire = dexb200.to_wire(to_node(tree, np.float64))
    y, ok = oracle.eval_tree_array(wire, ops.opcodes, X)
    if not ok:      # the reference filters on `complete` as well
        return
    want = numpy_eval(tree, X)
    assert np.isfinite(want).all(), "complete, but the plain evaluation is not finite"
    assert approx(y, want), (tree, y, want)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_device_equals_oracle_on_generated_populations(dtype, oracle):
    from tests.test_gpu_parity import _check_population
    ops = dexb200.OperatorEnum(OPS)
    seen = {"trees": 0, "complete": 0}
    tr, mats = strategies(32 if dtype == np.float32 else 64)

    @settings(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
    @given(pop=st.lists(tr, min_size=8, max_size=24), X=mats)
    def run(pop, X):
        nodes = [to_node(t, dtype) for t in pop]
        wires, offsets = dexb200.to_wire_population(nodes)
        Xd = X.astype(dtype)
        errs, ok = _check_population(oracle, wires, offsets, ops, Xd, dtype, label="", min_strict=0.0)
        seen["trees"] += len(pop)
        seen["complete"] += int(np.sum(ok))

    run()
    assert seen["trees"] > 200 and seen["complete"] > 50, seen
