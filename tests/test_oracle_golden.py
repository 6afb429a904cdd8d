"""The CPU oracle (oracle/) against the reference's own known-answer tests
(tests/golden/reference_known_answers.json) and against a dumb numpy recursion.
Runs without a GPU."""
import numpy as np
import pytest

import dexb200
from dexb200 import to_wire
from tests.golden_util import (CONTEXTS, expected_grad, expected_y, load_cases, make_matrix,
                               make_operators, make_tree)

CASES = load_cases()


def _flags(o, ctx):
    f = 0
    if ctx.get("early_exit", True):
        f |= o.EARLY_EXIT
    if ctx.get("use_fused", True):
        f |= o.USE_FUSED
    if ctx.get("bumper", False):
        f |= o.BUMPER
    return f


def _tol(case, dtype):
    e = case["expect"]
    if dtype == np.float32:
        return e.get("rtol32", 1e-4), e.get("atol", 1e-5)
    return e.get("rtol", 1e-6), e.get("atol", 1e-9)


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_oracle_matches_reference_known_answers(case, oracle):
    for dt in case["dtypes"]:
        dtype = np.dtype(dt).type
        ops = make_operators(case)
        tree = make_tree(case["tree"], ops, dtype)
        wire = to_wire(tree)
        X = make_matrix(case["X"], dtype)
        P = cls0 = None
        if "parameters" in case:
            P = make_matrix(case["parameters"], dtype)
            cls0 = np.array(case["classes"], dtype=np.int32) - 1
        for cname in case.get("contexts", ["default"]):
            flags = _flags(oracle, CONTEXTS[cname])
            if P is not None:
                y, ok = oracle.eval_parametric(wire, ops.opcodes, X, P, cls0, flags)
            else:
                y, ok = oracle.eval_tree_array(wire, ops.opcodes, X, flags)
            e = case["expect"]
            assert ok == e["ok"], (case["id"], dt, cname)
            want = expected_y(case, X, P, cls0)
            if ok and want is not None:
                rtol, atol = _tol(case, dtype)
                for j, w in enumerate(want):
                    if w is None:
                        assert not np.isfinite(y[j])
                    elif np.isfinite(w):
                        assert abs(y[j] - w) <= atol + rtol * abs(w), (case["id"], dt, cname, j, y[j], w)
                    else:
                        assert (np.isnan(w) and np.isnan(y[j])) or y[j] == w
            # second opinion: numpy recursion (elementwise, no early exit)
            if ok and want is not None and P is None:
                y2 = oracle.numpy_eval(wire, ops.opcodes, X, dexb200.OPCODE_INFO)
                fin = np.isfinite(y2)
                rtol, atol = _tol(case, dtype)
                np.testing.assert_allclose(y[fin], y2[fin], rtol=rtol * 10, atol=atol * 10)


GRAD_CASES = [c for c in CASES if "grad_mode" in c]


@pytest.mark.parametrize("case", GRAD_CASES, ids=[c["id"] for c in GRAD_CASES])
def test_oracle_gradients_match_reference_known_answers(case, oracle):
    mode = {"features": oracle.GRAD_FEATURES, "constants": oracle.GRAD_CONSTANTS,
            "both": oracle.GRAD_BOTH}[case["grad_mode"]]
    for dt in case["dtypes"]:
        dtype = np.dtype(dt).type
        ops = make_operators(case)
        tree = make_tree(case["tree"], ops, dtype)
        wire = to_wire(tree)
        X = make_matrix(case["X"], dtype)
        y, g, ok = oracle.eval_grad_tree_array(wire, ops.opcodes, X, mode)
        e = case["expect"]
        assert ok == e["ok"]
        if e.get("grad_all_nan"):
            # tree'(X) NaN-fills when !complete (src/EvaluationHelpers.jl:56-62)
            assert not ok
            continue
        want = expected_grad(case, X)
        rtol = e.get("grad_rtol", 1e-4 if dtype == np.float32 else 1e-6)
        atol = e.get("grad_atol", 1e-5 if dtype == np.float32 else 1e-9)
        assert g.shape == want.shape
        np.testing.assert_allclose(g, want, rtol=rtol, atol=atol)
        # eval_diff == rows of eval_grad (test/test_derivatives.jl:96-117)
        if mode == oracle.GRAD_FEATURES:
            for k in range(X.shape[0]):
                y1, d1, ok1 = oracle.eval_diff_tree_array(wire, ops.opcodes, X, k)
                assert ok1
                np.testing.assert_allclose(d1, g[k], rtol=1e-6, atol=1e-7)
                np.testing.assert_allclose(y1, y, rtol=1e-6, atol=1e-7)


def test_constant_numbering_is_leaf_order(oracle):
    """index_constant_nodes order == get_scalar_constants order
    (/root/reference/test/test_derivatives.jl:146-170)."""
    ops = dexb200.OperatorEnum({1: ("cos",), 2: ("+", "*", "-")})
    N = dexb200.Node
    tree = N(2, N(1, N(val=1.5), N(3, N(feature=1), N(val=2.5))), N(1, N(1, N(val=3.5), N(val=4.5))))
    consts, _ = dexb200.get_scalar_constants(tree)
    assert list(consts) == [1.5, 2.5, 3.5, 4.5]
    X = np.array([[0.7, -0.3]], dtype=np.float64)
    y, g, ok = oracle.eval_grad_tree_array(to_wire(tree), ops.opcodes, X, oracle.GRAD_CONSTANTS)
    assert ok and g.shape == (4, 2)
    # d/dc_k by central differences, constants perturbed in get_scalar_constants order
    for k in range(4):
        for sgn, store in ((1, "hi"), (-1, "lo")):
            c2 = consts.copy()
            c2[k] += sgn * 1e-6
            t2 = dexb200.set_scalar_constants(tree.copy(), c2)
            v, _ = oracle.eval_tree_array(to_wire(t2), ops.opcodes, X)
            if sgn == 1:
                hi = v
            else:
                lo = v
        np.testing.assert_allclose(g[k], (hi - lo) / 2e-6, rtol=1e-5, atol=1e-6)


def test_oracle_partials_match_finite_differences(oracle):
    """The reference takes operator partials from Zygote (ext/DynamicExpressionsZygoteExt.jl:7-15),
    i.e. the analytic derivative; the oracle's closed forms (dex_oracle_ops.inc) are pinned here
    for EVERY opcode of include/dex_ops.def against central differences of the oracle's own
    Float64 operator values at points where the operator is smooth."""
    import dexb200
    rng = np.random.default_rng(0)
    # piecewise-constant operators: derivative 0 away from the jumps
    flat = {"ROUND", "FLOOR", "CEIL", "TRUNC", "SIGN", "GREATER", "LESS", "LOGICAL_OR", "LOGICAL_AND",
            "GREATER_EQ", "LESS_EQ"}
    domains = {"ASIN": (-0.9, 0.9), "ACOS": (-0.9, 0.9), "ATANH": (-0.9, 0.9), "ACOSH": (1.2, 3.0),
               "SAFE_ACOSH": (1.2, 3.0), "LOG": (0.2, 3.0), "LOG2": (0.2, 3.0), "LOG10": (0.2, 3.0),
               "LOG1P": (-0.8, 3.0), "SAFE_LOG": (0.2, 3.0), "SAFE_LOG2": (0.2, 3.0), "SAFE_LOG10": (0.2, 3.0),
               "SAFE_LOG1P": (-0.8, 3.0), "SQRT": (0.2, 3.0), "SAFE_SQRT": (0.2, 3.0), "POW": (0.3, 2.5),
               "INV": (0.3, 2.5)}
    n_checked = 0
    for code, (sym, deg, name) in sorted(dexb200.OPCODE_INFO.items()):
        lo, hi = domains.get(sym, (-2.0, 2.0))
        for _ in range(40):
            a = [float(v) for v in rng.uniform(lo, hi, 3)]
            # keep away from kinks / jumps / poles of the piecewise operators
            if min(abs(a[0] - a[1]), abs(a[1] - a[2]), abs(a[0] - a[2]), abs(a[0]), abs(a[1]), abs(a[2])) < 0.05:
                continue
            if sym in flat and any(abs(v - round(v)) < 0.05 for v in a):
                continue
            if sym in ("MOD",) and abs(a[0] / a[1] - round(a[0] / a[1])) < 0.05:
                continue
            if sym == "CLAMP" and not (a[1] < a[2]):
                continue
            g = oracle.partials(code, *a)
            for k in range(deg):
                h = 1e-6
                p, m = list(a), list(a)
                p[k] += h
                m[k] -= h
                fd = (oracle.apply(code, *p) - oracle.apply(code, *m)) / (2 * h)
                assert abs(g[k] - fd) <= 1e-6 * max(1.0, abs(fd)) + 1e-7, (sym, k, a, g[k], fd)
            n_checked += 1
    assert n_checked > 30 * len(dexb200.OPCODE_INFO) * 0.5
