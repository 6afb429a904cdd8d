"""Parity of the CUDA path (through the C ABI) with the CPU oracle and with the
reference's known answers.  Needs a GPU: run with `-m gpu` on the B200 box.

Tolerances (BASELINE.json north_star): <= 1e-4 relative for Float32, <= 1e-6 relative for
Float64, norm-wise per tree as the reference's own `≈` does; `complete` flags, feature /
parameter / constant indexing: exact.
"""
import numpy as np
import pytest

import dexb200
from dexb200 import device as D
from dexb200 import treegen
from tests.golden_util import (CONTEXTS, expected_grad, expected_y, load_cases, make_matrix,
                               make_operators, make_tree)
from tests.parity_util import check_trees, tree_verdict

pytestmark = pytest.mark.gpu

RTOL = {np.float32: 1e-4, np.float64: 1e-6}


def _oflags(o, ctx):
    return ((o.EARLY_EXIT if ctx.get("early_exit", True) else 0) |
            (o.USE_FUSED if ctx.get("use_fused", True) else 0) | (o.BUMPER if ctx.get("bumper") else 0))


def _relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(float(np.max(np.abs(b), initial=0.0)), 1e-300)   # avoid overflow inside norm()
    with np.errstate(all="ignore"):
        den = max(np.linalg.norm(b / scale), 1e-300)
        return float(np.linalg.norm(a / scale - b / scale) / den)


def _flags_agree(ok, rok_ref, rok_elem, label=""):
    """Device `complete` flags vs the oracle.

    rok_ref  = the reference's rule: an array is valid iff isfinite(sum(array))
    rok_elem = the same walk with "every element finite" (oracle flag ELEMENTWISE)
    The two differ only when a SUM of finite values overflows (e.g. 3000 samples of 1e36 in
    Float32); the device implements the element-wise rule (DESIGN.md "completion flag").
    Everything else must agree exactly."""
    ok = np.asarray(ok, dtype=bool)
    rok_ref = np.asarray(rok_ref, dtype=bool)
    rok_elem = np.asarray(rok_elem, dtype=bool)
    assert (rok_ref <= rok_elem).all(), f"{label}: sum rule passed where the element rule failed?"
    assert (ok == rok_elem).all(), \
        f"{label}: complete flags differ at trees {np.nonzero(ok != rok_elem)[0][:10]}"
    return int((rok_ref != rok_elem).sum())   # number of sum-overflow-only cases


def _same_nonfinite(a, b):
    return (np.isnan(a) == np.isnan(b)).all() and (np.isposinf(a) == np.isposinf(b)).all() and \
        (np.isneginf(a) == np.isneginf(b)).all()


# ---------------------------------------------------------------------------------------
# 1. the reference's known-answer tests, through the reference-shaped public API
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", load_cases(), ids=lambda c: c["id"])
def test_known_answers_through_public_api(case, oracle):
    for dt in case["dtypes"]:
        dtype = np.dtype(dt).type
        ops = make_operators(case)
        tree = make_tree(case["tree"], ops, dtype)
        wire = dexb200.to_wire(tree)
        X = make_matrix(case["X"], dtype)
        e = case["expect"]
        for cname in case.get("contexts", ["default"]):
            c = CONTEXTS[cname]
            ectx = dexb200.EvalContext(**c)
            if "parameters" in case:
                P = make_matrix(case["parameters"], dtype)
                cls = np.array(case["classes"])
                ex = dexb200.ParametricExpression(tree, operators=ops, parameters=P)
                y, ok = ex.eval_tree_array(X, cls, eval_context=ectx)
                want = expected_y(case, X, P, cls - 1)
                ry, rok = oracle.eval_parametric(wire, ops.opcodes, X, P, cls - 1, _oflags(oracle, c))
                yc = ex(X, cls, eval_context=ectx)
            else:
                y, ok = dexb200.eval_tree_array(tree, X, ops, eval_context=ectx)
                want = expected_y(case, X)
                ry, rok = oracle.eval_tree_array(wire, ops.opcodes, X, _oflags(oracle, c))
                yc = tree(X, ops, eval_context=ectx)
            assert y.dtype == dtype and y.shape == (X.shape[1],)
            assert ok == e["ok"] == rok, (case["id"], dt, cname)
            if e.get("call_all_nan"):
                assert np.isnan(yc).all()           # tree(X) NaN-fills, EvaluationHelpers.jl:31
            if ok:
                assert (yc == y).all()              # test_evaluation.jl:88-89
                rtol = max(RTOL[dtype], e.get("rtol32", 0) if dtype == np.float32 else 0)
                atol = e.get("atol", 1e-5 if dtype == np.float32 else 1e-9)
                if want is not None:
                    for j, w in enumerate(want):
                        if w is None:
                            assert not np.isfinite(y[j])
                        elif np.isfinite(w):
                            assert abs(y[j] - w) <= atol + rtol * abs(w), (case["id"], dt, cname, j)
                        else:
                            assert not np.isfinite(y[j])
                fin = np.isfinite(ry)
                assert _same_nonfinite(y, ry)
                np.testing.assert_allclose(y[fin], ry[fin], rtol=rtol, atol=atol)


GRAD_CASES = [c for c in load_cases() if "grad_mode" in c]


@pytest.mark.parametrize("case", GRAD_CASES, ids=lambda c: c["id"])
def test_known_gradients_through_public_api(case):
    variable = {"features": True, "constants": False, "both": "both"}[case["grad_mode"]]
    for dt in case["dtypes"]:
        dtype = np.dtype(dt).type
        ops = make_operators(case)
        tree = make_tree(case["tree"], ops, dtype)
        X = make_matrix(case["X"], dtype)
        y, g, ok = dexb200.eval_grad_tree_array(tree, X, ops, variable=variable)
        e = case["expect"]
        assert ok == e["ok"]
        if e.get("grad_all_nan"):
            assert np.isnan(dexb200.grad_tree(tree, X, ops, variable=variable)).all()
            continue
        want = expected_grad(case, X)
        rtol = e.get("grad_rtol", RTOL[dtype])
        atol = e.get("grad_atol", 1e-5 if dtype == np.float32 else 1e-9)
        assert g.shape == want.shape and g.dtype == dtype
        np.testing.assert_allclose(g, want, rtol=rtol, atol=atol)
        if variable is True:   # eval_diff == rows of eval_grad == tree' (test_derivatives.jl:96-117)
            for k in range(X.shape[0]):
                y1, d1, ok1 = dexb200.eval_diff_tree_array(tree, X, ops, k + 1)
                assert ok1
                np.testing.assert_allclose(d1, g[k], rtol=1e-6, atol=1e-7)
            ex = dexb200.Expression(tree, operators=ops)
            np.testing.assert_array_equal(ex.gradient(X, variable=True), g)


# ---------------------------------------------------------------------------------------
# 2. random populations against the oracle
# ---------------------------------------------------------------------------------------
def _grad_verdict(dtype, out, g, ref, rg, yard_pairs):
    """Verdict of one tree of a gradient call: the value row and the (G x N) gradient block, each
    against the oracle with the oracle's own yardstick evaluations [(value, gradient), ...]."""
    parts = [(out, ref, [y for y, _ in yard_pairs])]
    if rg.shape[0]:
        parts.append((g, rg, [yg for _, yg in yard_pairs]))
    return tree_verdict(dtype, parts)


def _check_population(oracle, nodes, offsets, ops, X, dtype, *, ctx=None, label="", min_strict=0.85, pop=None):
    """Device vs oracle for a population: flags exactly, values per tree in the classes of
    tests/parity_util.py (strict / loose / elementwise; no tree is skipped).  Returns (errors of the compared trees, ok)."""
    c = ctx or {}
    early = c.get("early_exit", True)
    if pop is None:
        pop = D.Population(None, ops, dtype, wire=(nodes, offsets), bumper=c.get("bumper", False),
                           use_fused=c.get("use_fused", True))
    out, ok = pop.eval(X, early_exit=early)
    out = out.cpu().numpy()
    ok = ok.cpu().numpy().astype(bool)
    ref, rok = oracle.eval_population(nodes, offsets, ops.opcodes, X, _oflags(oracle, c))
    _, rok_elem = oracle.eval_population(nodes, offsets, ops.opcodes, X, _oflags(oracle, c) | oracle.ELEMENTWISE)
    _flags_agree(ok, rok, rok_elem, label)
    # conditioning yardsticks: how far the oracle's own result moves (a) when X is perturbed
    # by one ulp, (b) for float32, when it is evaluated in float64 (float64: with 80-bit
    # intermediates), (c) with every transcendental result moved by one ulp
    # (oracle.set_ulp_nudge): two faithful implementations of sin / exp / ... may differ by exactly
    # that, and a pole such as x3 / (0.997 - sin(x3 + x2)) amplifies it without any sensitivity
    # to X showing in (a)
    ref_p, _ = oracle.eval_population(nodes, offsets, ops.opcodes, np.nextafter(X, dtype(np.inf)),
                                      _oflags(oracle, c))
    if dtype == np.float32:
        ref2, _ = oracle.eval_population(nodes, offsets, ops.opcodes, X.astype(np.float64), _oflags(oracle, c))
    else:
        ref2, _ = oracle.eval_population_f80(nodes, offsets, ops.opcodes, X, _oflags(oracle, c))
    try:
        oracle.set_ulp_nudge(1)
        ref_n, _ = oracle.eval_population(nodes, offsets, ops.opcodes, X, _oflags(oracle, c))
    finally:
        oracle.set_ulp_nudge(0)
    verdicts, ids = [], []
    n_pattern = 0
    for t in np.nonzero(rok)[0]:
        r = ref[t]
        yards = (ref_p[t], ref_n[t], ref2[t])
        fin = np.isfinite(r)
        mask = None
        if not fin.all():
            # only possible with early_exit = false: the row holds the Inf / NaN the reference's
            # arithmetic produces.  The PATTERN is part of the answer wherever the oracle's own
            # pattern is stable under the yardstick perturbations.
            assert not early, f"{label}: tree {t} is complete but the oracle row is not finite"
            if all(_same_nonfinite(np.asarray(y, dtype=r.dtype), r) for y in yards):
                assert _same_nonfinite(out[t], r), f"{label}: tree {t}: non-finite pattern differs from the oracle"
                n_pattern += 1
            mask = fin & np.isfinite(out[t])
        verdicts.append(tree_verdict(dtype, [(out[t], r, yards, mask)]))
        ids.append(int(t))
    stats = check_trees(label, dtype, verdicts, min_strict=min_strict, ids=ids)
    stats["n_pattern_checked"] = n_pattern
    errs = np.array([v[1] for v in verdicts])
    return errs, ok


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("opset", ["A", "B"])
def test_random_population_matches_oracle(dtype, opset, oracle):
    spec = treegen.OPSET_A if opset == "A" else treegen.OPSET_B
    ops = dexb200.OperatorEnum(spec)
    nodes, offsets = treegen.gen_population(300, 8, len(spec[1]), len(spec[2]), 5, seed=11, dtype=dtype)
    X = np.random.default_rng(0).standard_normal((5, 3000)).astype(dtype)
    errs, ok = _check_population(oracle, nodes, offsets, ops, X, dtype, label=f"{opset}/{dtype.__name__}")
    assert ok.sum() > 30
    assert np.median(errs) < (1e-6 if dtype == np.float32 else 1e-14)


@pytest.mark.parametrize("cname", ["bumper", "unfused", "no_early_exit", "bumper_no_early_exit"])
def test_population_policies_match_oracle(cname, oracle):
    ops = dexb200.OperatorEnum(treegen.OPSET_A)
    nodes, offsets = treegen.gen_population(200, 7, 2, 4, 5, seed=5)
    X = np.random.default_rng(1).standard_normal((5, 777)).astype(np.float32)
    X[2, 5] = np.inf
    X[0, 700] = np.nan
    _check_population(oracle, nodes, offsets, ops, X, np.float32, ctx=CONTEXTS[cname], label=cname)


def test_depth12_ten_features_population(oracle):
    """C4's tree shape (depth 12, 10 features) at a size the oracle finishes quickly."""
    ops = dexb200.OperatorEnum(treegen.OPSET_A)
    nodes, offsets = treegen.gen_population(120, 12, 2, 4, 10, seed=3)
    X = np.random.default_rng(2).standard_normal((10, 2048)).astype(np.float32)
    _check_population(oracle, nodes, offsets, ops, X, np.float32, label="depth12", min_strict=0.65)


@pytest.mark.parametrize("N", [1, 2, 3, 5, 127, 128, 129, 1023, 1024, 1025, 4097])
def test_ragged_sample_counts(N, oracle):
    ops = dexb200.OperatorEnum(treegen.OPSET_B)
    nodes, offsets = treegen.gen_population(40, 6, 4, 4, 3, seed=9)
    X = np.random.default_rng(N).standard_normal((3, N)).astype(np.float32)
    _check_population(oracle, nodes, offsets, ops, X, np.float32, label=f"N={N}")


def test_edge_shapes(oracle):
    ops = dexb200.OperatorEnum({1: ("cos", "exp"), 2: ("+", "*"), 3: ("fma", "clamp")})
    N_ = dexb200.Node
    trees = [
        N_(feature=1),                                      # bare feature leaf
        N_(val=2.5),                                        # bare constant
        N_(val=float("inf")),                               # non-finite constant => incomplete
        N_(1, N_(val=0.5)),                                 # constant subtree
        N_(2, N_(val=float("inf")), N_(val=0.0)),           # constant subtree -> NaN
        N_(1, N_(feature=7), N_(feature=7)),                # only the last feature
        N_(op=1, children=(N_(feature=1), N_(val=2.0), N_(feature=3))),          # ternary, leaves
        N_(op=2, children=(N_(1, N_(feature=2)), N_(val=-0.5), N_(val=0.5))),    # clamp
        N_(op=1, children=(N_(val=1.0), N_(val=2.0), N_(val=3.0))),              # all-constant ternary
    ]
    # a chain of 200 unary ops and a maximally unbalanced binary tree
    chain = N_(feature=1)
    for _ in range(200):
        chain = N_(1, chain)
    trees.append(chain)
    comb = N_(1, N_(feature=1))
    for k in range(60):
        comb = N_(2, N_(1, N_(feature=1 + k % 7)), comb)
    trees.append(comb)
    nodes, offsets = dexb200.to_wire_population(trees)
    for dtype in (np.float32, np.float64):
        X = np.random.default_rng(4).standard_normal((7, 300)).astype(dtype)
        _check_population(oracle, nodes, offsets, ops, X, dtype, label="edge")
        for early in (True, False):
            _check_population(oracle, nodes, offsets, ops, X, dtype, ctx={"early_exit": early}, label="edge")


def test_shared_subexpressions_on_device(oracle):
    """Repeated subtrees (GraphNode sharing) computed once: values, flags, d/dX gradients and the
    fused loss against the oracle on the expanded trees."""
    from tests.test_abi_and_flatten import _shared_trees
    ops = dexb200.OperatorEnum({1: ("cos", "exp"), 2: ("+", "-", "*", "/")})
    for dtype in (np.float32, np.float64):
        trees = _shared_trees(dtype) * 3
        nodes, offsets = dexb200.to_wire_population(trees)
        X = np.random.default_rng(5).standard_normal((3, 2500)).astype(dtype)
        errs, ok = _check_population(oracle, nodes, offsets, ops, X, dtype, label=f"shared/{np.dtype(dtype).name}")
        _check_population(oracle, nodes, offsets, ops, X, dtype, ctx={"early_exit": False}, label="shared/no_early_exit")
        assert ok.sum() >= 6
        pop = D.Population(None, ops, dtype, wire=(nodes, offsets))
        assert pop.info["n_folded_instructions"] < pop.info["n_instructions"]
        out, grad, off, gok = pop.eval_grad(X, D.GRAD_FEATURES)
        out, grad, gok = out.cpu().numpy(), grad.cpu().numpy(), gok.cpu().numpy().astype(bool)
        ref, rgrads, rok = oracle.eval_grad_population(nodes, offsets, ops.opcodes, X, oracle.GRAD_FEATURES)
        _, _, rok_e = oracle.eval_grad_population(nodes, offsets, ops.opcodes, X, oracle.GRAD_FEATURES | oracle.GRAD_ELEMENTWISE)
        _flags_agree(gok, rok, rok_e, "shared grad")
        ref64, rg64, _ = oracle.eval_grad_population(nodes, offsets, ops.opcodes, X.astype(np.float64), oracle.GRAD_FEATURES)
        ref_p, rg_p, _ = oracle.eval_grad_population(nodes, offsets, ops.opcodes, np.nextafter(X, dtype(np.inf)), oracle.GRAD_FEATURES)
        N = X.shape[1]
        verdicts = [_grad_verdict(dtype, out[t], grad[off[t]:off[t + 1]].reshape(N, 3).T, ref[t], rgrads[t],
                                  [(ref_p[t], rg_p[t]), (ref64[t], rg64[t])]) for t in np.nonzero(rok)[0]]
        check_trees(f"shared grad/{np.dtype(dtype).name}", dtype, verdicts, min_strict=0.6)
        y = np.random.default_rng(6).standard_normal(N).astype(dtype)
        loss, lok = pop.eval_loss(X, y)
        o, okk = pop.eval(X)
        good = okk.cpu().numpy().astype(bool)
        want = ((o.cpu().numpy().astype(np.float64) - y.astype(np.float64)[None]) ** 2).mean(axis=1)
        np.testing.assert_allclose(loss.cpu().numpy()[good], want[good], rtol=1e-10)


def test_empty_inputs():
    ops = dexb200.OperatorEnum(treegen.OPSET_A)
    nodes, offsets = treegen.gen_population(3, 4, 2, 4, 2, seed=1)
    pop = D.Population(None, ops, np.float32, wire=(nodes, offsets))
    out, ok = pop.eval(np.zeros((2, 0), np.float32))
    assert out.shape == (3, 0)
    assert ok.cpu().numpy().tolist() == [1, 1, 1]          # is_valid_array(empty) is true in the reference
    y, g, off, gok = pop.eval_grad(np.zeros((2, 0), np.float32), D.GRAD_FEATURES)
    assert y.shape == (3, 0) and g.numel() == 0 and gok.cpu().numpy().tolist() == [1, 1, 1]
    # ... unless a folded constant subtree is invalid (Evaluate.jl:347-354, whatever N is)
    N_ = dexb200.Node
    bad = D.Population([N_(1, N_(val=1.0), N_(3, N_(val=1.0), N_(val=0.0))), N_(feature=1)], ops, np.float32)
    out, ok = bad.eval(np.zeros((2, 0), np.float32))
    assert ok.cpu().numpy().tolist() == [0, 1]
    oh, kh = np.zeros((2, 0), np.float32), np.full(2, 7, np.uint8)
    bad.eval_host(np.zeros((0, 2), np.float32), oh, kh)
    assert kh.tolist() == [0, 1]
    empty = D.Population(None, ops, np.float32, wire=(nodes[:0], np.zeros(1, np.int64)))
    out, ok = empty.eval(np.zeros((2, 10), np.float32))
    assert out.shape == (0, 10)


def test_calls_on_different_streams_do_not_race_on_the_context_scratch(oracle):
    """One context, two torch streams: the transposed copy of X and the other scratch buffers
    belong to the context, so a call issued on stream B must run after the previous call on
    stream A (dex_ctx_set_stream orders them)."""
    import torch
    ops = dexb200.OperatorEnum(treegen.OPSET_A)
    nodes, offsets = treegen.gen_population(400, 8, 2, 4, 5, seed=3)
    pop = D.Population(None, ops, np.float32, wire=(nodes, offsets))
    rng = np.random.default_rng(5)
    N = 1 << 15
    Xa = torch.from_numpy(rng.standard_normal((N, 5)).astype(np.float32)).cuda()
    Xb = torch.from_numpy(rng.standard_normal((N, 5)).astype(np.float32)).cuda()
    ref_a, _ = pop.eval(Xa.T)
    ref_b, _ = pop.eval(Xb.T)
    ya, ga, _, _ = pop.eval_grad(Xa.T, D.GRAD_FEATURES)
    torch.cuda.synchronize()
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(20):
        with torch.cuda.stream(sa):
            oa, _ = pop.eval(Xa.T)
        with torch.cuda.stream(sb):
            ob, _ = pop.eval(Xb.T)
        with torch.cuda.stream(sa):
            y2, g2, _, _ = pop.eval_grad(Xa.T, D.GRAD_FEATURES)
        with torch.cuda.stream(sb):
            ob2, _ = pop.eval(Xb.T)
        torch.cuda.synchronize()
        same = lambda u, v: bool(((u == v) | (torch.isnan(u) & torch.isnan(v))).all())
        assert same(oa, ref_a) and same(ob, ref_b) and same(ob2, ref_b) and same(g2, ga) and same(y2, ya)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_max_min_gradient_ties_follow_the_operand_order(dtype, oracle):
    """max(0.0, x1) at x1 == 0: the reference gives the tie to the second operand
    (partials (x > y, !(x > y))), whatever order the flattener evaluates in."""
    ops = dexb200.OperatorEnum({1: ("cos", "relu"), 2: ("max", "min", "*", "+")})
    N_ = dexb200.Node
    x1, x2 = N_(feature=1, T=dtype), N_(feature=2, T=dtype)
    c0 = lambda: N_(val=0.0, T=dtype)
    prod = lambda: N_(3, x1, x2)
    trees = []
    for op in (1, 2):
        trees += [N_(op, c0(), x1), N_(op, x1, c0()), N_(op, x1, x2), N_(op, x2, x1), N_(op, x2, prod()),
                  N_(op, prod(), x2), N_(op, c0(), prod()), N_(op, prod(), c0()),
                  N_(op, N_(4, prod(), x1), N_(4, x1, prod())), N_(op, N_(4, x1, prod()), N_(4, prod(), x1))]
    X = np.random.default_rng(0).standard_normal((2, 512)).astype(dtype)
    X[0, ::3] = 0.0                 # ties with the constant and (through x1 * x2 = 0) with products
    X[1, ::5] = X[0, ::5]           # ties between features
    X[1, 1::7] = 1.0                # x1 * x2 == x1
    nodes, offsets = dexb200.to_wire_population(trees)
    pop = D.Population(None, ops, dtype, wire=(nodes, offsets))
    for dmode, omode in ((D.GRAD_FEATURES, oracle.GRAD_FEATURES), (D.GRAD_BOTH, oracle.GRAD_BOTH),
                         (D.GRAD_CONSTANTS, oracle.GRAD_CONSTANTS)):
        out, grad, off, ok = pop.eval_grad(X, dmode)
        out, grad = out.cpu().numpy(), grad.cpu().numpy()
        ref, rgrads, rok = oracle.eval_grad_population(nodes, offsets, ops.opcodes, X, omode)
        assert rok.all() and ok.cpu().numpy().all()
        for t in range(len(trees)):
            G = rgrads[t].shape[0]
            g = grad[off[t]:off[t + 1]].reshape(X.shape[1], G).T
            np.testing.assert_array_equal(out[t], ref[t])
            np.testing.assert_allclose(g, rgrads[t], rtol=1e-6, atol=0, err_msg=f"tree {t} mode {dmode}")


def test_feature_out_of_range_is_an_error():
    ops = dexb200.OperatorEnum(treegen.OPSET_A)
    tree = dexb200.Node(1, dexb200.Node(feature=4))
    with pytest.raises(D.DexError, match="feature 4"):
        dexb200.eval_tree_array(tree, np.ones((3, 8), np.float32), ops)
    with pytest.raises(AssertionError):
        dexb200.Expression(tree, operators=ops)(np.ones((3, 8), np.float32))   # _validate_input


def test_strided_inputs_and_outputs(oracle):
    import torch
    ops = dexb200.OperatorEnum(treegen.OPSET_A)
    nodes, offsets = treegen.gen_population(50, 6, 2, 4, 4, seed=2)
    rng = np.random.default_rng(3)
    N = 1500
    Xpad = rng.standard_normal((N, 6)).astype(np.float32)       # ldx = 6 > F = 4
    Xd = torch.from_numpy(Xpad).cuda()
    Xview = Xd[:, :4].T                                            # (F, N) with strides (1, 6): zero copy
    pop = D.Population(None, ops, np.float32, wire=(nodes, offsets))
    big = torch.full((50, N + 8), -7.0, device="cuda")
    out, ok = pop.eval(Xview, out=big[:, :N])                      # ldo = N + 8
    ref, rok = oracle.eval_population(nodes, offsets, ops.opcodes, Xpad[:, :4].T.copy())
    _, rok_elem = oracle.eval_population(nodes, offsets, ops.opcodes, Xpad[:, :4].T.copy(),
                                         oracle.DEFAULT_FLAGS | oracle.ELEMENTWISE)
    _flags_agree(ok.cpu().numpy(), rok, rok_elem, "strided")
    assert (big[:, N:] == -7.0).all()
    o = out.cpu().numpy()
    ref64, _ = oracle.eval_population(nodes, offsets, ops.opcodes, Xpad[:, :4].T.astype(np.float64))
    check_trees("strided", np.float32, [tree_verdict(np.float32, [(o[t], ref[t], (ref64[t],))]) for t in np.nonzero(rok)[0]])
    # torch in => torch out, same values
    y, okk = dexb200.eval_trees_array([dexb200.from_wire(nodes[offsets[0]:offsets[1]])], Xview, ops)
    assert y.is_cuda and torch.equal(y[0], out[0])


# ---------------------------------------------------------------------------------------
# 3. every operator of the table, scalar by scalar
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_every_builtin_operator(dtype, oracle):
    grid = np.array([-1e30, -300.0, -7.5, -2.0, -1.0, -0.5, -1e-3, -0.0, 0.0, 1e-3, 0.5, 1.0, 1.5, 2.0,
                     7.5, 88.0, 300.0, 1e30, np.inf, -np.inf, np.nan, 3.0, -3.0, 0.25], dtype=dtype)
    N_ = dexb200.Node
    for code, (sym, deg, name) in sorted(dexb200.OPCODE_INFO.items()):
        ops = dexb200.OperatorEnum({deg: (name,)})
        if deg == 1:
            X = grid[None, :]
            tree = N_(1, N_(feature=1))
        elif deg == 2:
            a, b = np.meshgrid(grid, grid)
            X = np.stack([a.ravel(), b.ravel()]).astype(dtype)
            tree = N_(1, N_(feature=1), N_(feature=2))
        else:
            g3 = grid[[3, 5, 8, 10, 12, 14, 18, 20]]
            a, b, c = np.meshgrid(g3, g3, g3)
            X = np.stack([a.ravel(), b.ravel(), c.ravel()]).astype(dtype)
            tree = N_(op=1, children=(N_(feature=1), N_(feature=2), N_(feature=3)))
        y, _ = dexb200.eval_tree_array(tree, X, ops, eval_context=dexb200.EvalContext(early_exit=False))
        ry, _ = oracle.eval_tree_array(dexb200.to_wire(tree), ops.opcodes, X, oracle.USE_FUSED)
        assert _same_nonfinite(y, ry), f"{sym}: non-finite pattern differs"
        fin = np.isfinite(ry)
        big = fin & (np.abs(ry) > 1e-30)
        tol = 2e-6 if dtype == np.float32 else 1e-13
        if sym in ("POW", "POW_ABS", "TAN", "SINH", "COSH", "EXP10", "EXP2", "EXP", "EXPM1", "ERFC"):
            tol *= 40    # a few ulp of the argument, amplified by the function's growth
        np.testing.assert_allclose(y[big], ry[big], rtol=tol, atol=0, err_msg=sym)
        np.testing.assert_allclose(y[fin & ~big], ry[fin & ~big], rtol=0, atol=1e-30, err_msg=sym)


# ---------------------------------------------------------------------------------------
# 4. parametric populations, gradients, constants, fused loss, host entry point
# ---------------------------------------------------------------------------------------
def test_parametric_population_matches_oracle(oracle):
    ops = dexb200.OperatorEnum(treegen.OPSET_A)
    P_, n_params, n_classes, F, N = 150, 3, 10, 5, 2500
    nodes, offsets = treegen.gen_population(P_, 8, 2, 4, F, seed=21, n_params=n_params)
    rng = np.random.default_rng(5)
    X = rng.standard_normal((F, N)).astype(np.float32)
    params = rng.standard_normal((P_, n_params, n_classes)).astype(np.float32)
    cls0 = rng.integers(0, n_classes, N)
    pop = D.Population(None, ops, np.float32, wire=(nodes, offsets))
    assert pop.info["max_parameter"] == n_params - 1
    out, ok = pop.eval_parametric(X, params, cls0)
    out, ok = out.cpu().numpy(), ok.cpu().numpy().astype(bool)
    ref, rok = oracle.eval_parametric_population(nodes, offsets, ops.opcodes, X, params, cls0)
    _, rok_elem = oracle.eval_parametric_population(nodes, offsets, ops.opcodes, X, params, cls0,
                                                    oracle.DEFAULT_FLAGS | oracle.ELEMENTWISE)
    _flags_agree(ok, rok, rok_elem, "parametric")
    ref64, _ = oracle.eval_parametric_population(nodes, offsets, ops.opcodes, X.astype(np.float64),
                                                 params.astype(np.float64), cls0)
    st = check_trees("parametric", np.float32,
                     [tree_verdict(np.float32, [(out[t], ref[t], (ref64[t],))]) for t in np.nonzero(rok)[0]])
    assert st["n_complete"] > 30
    with pytest.raises(D.DexError):
        pop.eval(X)                                   # parameter leaves need the parametric entry point
    with pytest.raises(D.DexError):
        pop.eval_parametric(X, params, cls0 + n_classes)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mode", ["features", "constants", "both"])
def test_gradient_population_matches_oracle(dtype, mode, oracle):
    spec = {1: ("cos", "exp", "sin"), 2: ("+", "-", "*", "/")}
    ops = dexb200.OperatorEnum(spec)
    nodes, offsets = treegen.gen_population(120, 7, 3, 4, 5, seed=31, dtype=dtype)
    X = np.random.default_rng(6).standard_normal((5, 700)).astype(dtype)
    omode = {"features": oracle.GRAD_FEATURES, "constants": oracle.GRAD_CONSTANTS, "both": oracle.GRAD_BOTH}[mode]
    dmode = {"features": D.GRAD_FEATURES, "constants": D.GRAD_CONSTANTS, "both": D.GRAD_BOTH}[mode]
    pop = D.Population(None, ops, dtype, wire=(nodes, offsets))
    out, grad, off, ok = pop.eval_grad(X, dmode)
    out, grad, ok = out.cpu().numpy(), grad.cpu().numpy(), ok.cpu().numpy().astype(bool)
    ref, rgrads, rok = oracle.eval_grad_population(nodes, offsets, ops.opcodes, X, omode)
    _, _, rok_elem = oracle.eval_grad_population(nodes, offsets, ops.opcodes, X, omode | oracle.GRAD_ELEMENTWISE)
    ref64, rgrads64, _ = oracle.eval_grad_population(nodes, offsets, ops.opcodes, X.astype(np.float64), omode)
    _flags_agree(ok, rok, rok_elem, f"grad/{mode}")
    ref_p, rgrads_p, _ = oracle.eval_grad_population(nodes, offsets, ops.opcodes, np.nextafter(X, dtype(np.inf)), omode)
    N = X.shape[1]
    verdicts, n_with_grad = [], 0
    for t in np.nonzero(rok)[0]:
        G = rgrads[t].shape[0]
        assert off[t + 1] - off[t] == G * N                      # layout: (G x N), gradient index fastest
        g = grad[off[t]:off[t + 1]].reshape(N, G).T
        verdicts.append(_grad_verdict(dtype, out[t], g, ref[t], rgrads[t], [(ref_p[t], rgrads_p[t]), (ref64[t], rgrads64[t])]))
        n_with_grad += bool(G)
    check_trees(f"grad/{mode}/{np.dtype(dtype).name}", dtype, verdicts)
    assert n_with_grad > 10


def test_many_constants_gradient_passes(oracle):
    """More constants than one pass of the kernel carries (GC <= 8)."""
    ops = dexb200.OperatorEnum({1: ("cos",), 2: ("+", "*")})
    N_ = dexb200.Node
    tree = N_(2, N_(val=0.1), N_(feature=1))
    for k in range(30):
        tree = N_(1, N_(2, N_(val=0.5 + 0.01 * k), tree), N_(1, N_(2, N_(val=1.0 + k), N_(feature=1 + k % 2))))
    X = np.random.default_rng(7).standard_normal((2, 333))
    y, g, ok = dexb200.eval_grad_tree_array(tree, X, ops, variable="both")
    ry, rg, rok = oracle.eval_grad_tree_array(dexb200.to_wire(tree), ops.opcodes, X, oracle.GRAD_BOTH)
    assert ok and rok and g.shape == rg.shape == (2 + 61, 333)
    np.testing.assert_allclose(g, rg, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(y, ry, rtol=1e-12)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_gradient_constant_subtrees_follow_the_gradient_rule(dtype, oracle):
    """d/dX runs the constant-folded tape; the folded subtrees must still obey the gradient
    path's rule (value AND gradient of every node finite, EvaluateDerivative.jl:238-243):
    sqrt(0.0) is a finite value with an infinite partial, so its zero-seeded gradient is NaN."""
    ops = dexb200.OperatorEnum({1: ("sqrt", "exp", "cos"), 2: ("+", "*", "/")})
    N_ = dexb200.Node
    x1, x2 = N_(feature=1), N_(feature=2)
    trees = [
        N_(1, x1, N_(1, N_(val=0.0))),                                   # x1 + sqrt(0): d = Inf * 0
        N_(1, x1, N_(1, N_(val=4.0))),                                   # x1 + sqrt(4): fine
        N_(2, x2, N_(2, N_(val=1000.0))),                                # x2 * exp(1000): Inf value
        N_(1, N_(2, x1, x2), N_(3, N_(2, N_(val=2.0), N_(val=3.0)))),    # x1*x2 + cos(2*3)
        N_(1, x1, N_(3, N_(val=1.0), N_(1, N_(val=0.0)))),               # x1 + 1/sqrt(0): Inf value
        N_(1, N_(val=0.0)),                                              # sqrt(0) alone
        N_(1, x1, N_(1, N_(1, N_(val=0.0), N_(val=0.0)))),               # x1 + sqrt(0 + 0)
    ]
    X = np.random.default_rng(3).standard_normal((2, 257)).astype(dtype)
    for i, tree in enumerate(trees):
        y, g, ok = dexb200.eval_grad_tree_array(tree, X, ops, variable=True)
        ry, rg, rok = oracle.eval_grad_tree_array(dexb200.to_wire(tree), ops.opcodes, X,
                                                  oracle.GRAD_FEATURES | oracle.GRAD_ELEMENTWISE)
        assert bool(ok) == bool(rok), (i, ok, rok)
        if rok:
            np.testing.assert_allclose(y, ry, rtol=1e-5)
            np.testing.assert_allclose(g, rg, rtol=1e-5, atol=1e-6)


FAST_UNARY = ("neg", "abs", "square", "cube", "inv", "sqrt", "exp", "log", "sin", "cos", "tanh", "relu",
              "safe_log", "safe_sqrt")
FAST_BINARY = ("+", "-", "*", "/", "max", "min")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mode", ["features", "constants", "both"])
def test_gradient_of_every_fast_handler_and_operand_form(dtype, mode, oracle):
    """Every operator with a specialised gradient code path (the generated PTX loop for Float32,
    csrc/gen_grad_ptx.py) in every operand form — accumulator, stack slot, feature leaf, inline
    constant on either side — against the oracle, for all three gradient modes."""
    ops = dexb200.OperatorEnum({1: FAST_UNARY, 2: FAST_BINARY})
    N_ = dexb200.Node
    x1, x2, x3 = (N_(feature=k, T=dtype) for k in (1, 2, 3))
    c = lambda v: N_(val=v, T=dtype)
    mul = FAST_BINARY.index("*") + 1
    add = FAST_BINARY.index("+") + 1
    inner = lambda: N_(mul, x1, x2)            # an operator child: arrives in the accumulator
    other = lambda: N_(add, x2, x3)            # a second operator child: one of the two is pushed
    trees, labels = [x2, c(3.0)], ["load/feature", "load/constant"]
    for i, name in enumerate(FAST_UNARY, start=1):
        for form, t in (("leaf", N_(i, x1)), ("acc", N_(i, inner())), ("const", N_(add, x1, N_(i, c(0.7)))),
                        ("acc2", N_(i, N_(add, inner(), c(1.25))))):
            trees.append(t)
            labels.append(f"{name}/{form}")
    for i, name in enumerate(FAST_BINARY, start=1):
        forms = {"AR": N_(i, inner(), x3), "RA": N_(i, x3, inner()), "AC": N_(i, inner(), c(1.5)),
                 "CA": N_(i, c(-0.75), inner()), "RR": N_(i, x1, x2), "RR_same": N_(i, x1, x1),
                 "RC": N_(i, x1, c(2.5)), "CR": N_(i, c(2.5), x3), "slot_acc": N_(i, inner(), other()),
                 "deep": N_(i, N_(i, inner(), other()), N_(i, other(), inner()))}
        for form, t in forms.items():
            trees.append(t)
            labels.append(f"{name}/{form}")
    nodes, offsets = dexb200.to_wire_population(trees)
    rng = np.random.default_rng(17)
    omode = {"features": oracle.GRAD_FEATURES, "constants": oracle.GRAD_CONSTANTS, "both": oracle.GRAD_BOTH}[mode]
    dmode = {"features": D.GRAD_FEATURES, "constants": D.GRAD_CONSTANTS, "both": D.GRAD_BOTH}[mode]
    pop = D.Population(None, ops, dtype, wire=(nodes, offsets))
    n_ok = 0
    # positive inputs keep sqrt/log/inv finite; signed inputs exercise the failure paths
    for X in (rng.uniform(0.5, 2.0, (3, 600)).astype(dtype), rng.standard_normal((3, 600)).astype(dtype)):
        out, grad, off, ok = pop.eval_grad(X, dmode)
        out, grad, ok = out.cpu().numpy(), grad.cpu().numpy(), ok.cpu().numpy().astype(bool)
        ref, rgrads, rok = oracle.eval_grad_population(nodes, offsets, ops.opcodes, X, omode)
        _, _, rok_elem = oracle.eval_grad_population(nodes, offsets, ops.opcodes, X, omode | oracle.GRAD_ELEMENTWISE)
        _flags_agree(ok, rok, rok_elem, f"grad handlers/{mode}")
        N = X.shape[1]
        tol = 2e-5 if dtype == np.float32 else 1e-11
        for t in np.nonzero(rok)[0]:
            G = rgrads[t].shape[0]
            g = grad[off[t]:off[t + 1]].reshape(N, G).T
            assert _relerr(out[t], ref[t]) <= tol, labels[t]
            if G:
                assert _relerr(g, rgrads[t]) <= tol, (labels[t], mode)
            n_ok += 1
    assert n_ok > len(trees)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_random_trees_with_ternary_operators_value_and_gradient(dtype, oracle):
    """The reference's property test (test/test_supposition_consistency.jl:16-104: random trees
    over ((abs, cos, exp), (+, -, *, /), (fma, clamp, +, max)), X 5 x (1..16)) as a population:
    values, flags and gradients (ternary operators take the generic gradient path)."""
    from tests.test_abi_and_flatten import _random_wire
    spec = {1: ("abs", "cos", "exp"), 2: ("+", "-", "*", "/"), 3: ("fma", "clamp", "+", "max")}
    ops = dexb200.OperatorEnum(spec)
    rng = np.random.default_rng(2024)
    wires = []
    while len(wires) < 250:
        w = _random_wire(rng, ops, 5, max_nodes=int(rng.integers(1, 60)))
        if np.isfinite(w["val"][(w["degree"] == 0) & (w["kind"] == 0)]).all():
            wires.append(w)
    nodes = np.concatenate(wires)
    offsets = np.concatenate([[0], np.cumsum([len(w) for w in wires])]).astype(np.int64)
    for N in (1, 16, 333):
        X = rng.standard_normal((5, N)).astype(dtype)
        errs, ok = _check_population(oracle, nodes, offsets, ops, X, dtype, label=f"ternary N={N}")
        assert len(errs) > 20
        pop = D.Population(None, ops, dtype, wire=(nodes, offsets))
        for dmode, omode in ((D.GRAD_FEATURES, oracle.GRAD_FEATURES), (D.GRAD_BOTH, oracle.GRAD_BOTH)):
            out, grad, off, gok = pop.eval_grad(X, dmode)
            out, grad, gok = out.cpu().numpy(), grad.cpu().numpy(), gok.cpu().numpy().astype(bool)
            ref, rgrads, rok = oracle.eval_grad_population(nodes, offsets, ops.opcodes, X, omode)
            _, _, rok_elem = oracle.eval_grad_population(nodes, offsets, ops.opcodes, X, omode | oracle.GRAD_ELEMENTWISE)
            ref_p, rgrads_p, _ = oracle.eval_grad_population(nodes, offsets, ops.opcodes,
                                                             np.nextafter(X, dtype(np.inf)), omode)
            ref64, rgrads64, _ = oracle.eval_grad_population(nodes, offsets, ops.opcodes, X.astype(np.float64), omode)
            _flags_agree(gok, rok, rok_elem, f"ternary grad N={N}")
            verdicts = []
            for t in np.nonzero(rok)[0]:
                G = rgrads[t].shape[0]
                g = grad[off[t]:off[t + 1]].reshape(N, G).T
                verdicts.append(_grad_verdict(dtype, out[t], g, ref[t], rgrads[t],
                                              [(ref_p[t], rgrads_p[t]), (ref64[t], rgrads64[t])]))
            st = check_trees(f"ternary grad N={N} mode={dmode} {np.dtype(dtype).name}", dtype, verdicts,
                             min_strict=0.8)          # N = 1: a single sample per tree
            assert st["n_strict"] + st["n_loose"] > 20


def test_native_float32_handlers_are_accurate_to_a_few_ulp():
    """The specialised Float32 handlers of the generated PTX loops (default early_exit=True path)
    against float64 numpy on inputs where every result is finite: a few ulp, not just the 1e-4 of
    the population tests.  (test_every_builtin_operator runs early_exit=False: the _noexit form of the
    loop and, for the operators it does not implement, the C++ handlers.)"""
    rng = np.random.default_rng(99)
    n = 4096
    no_exit = dexb200.EvalContext(early_exit=False)
    pos = np.exp(rng.uniform(-80.0, 80.0, n)).astype(np.float32)            # wide-range positive
    sgn = (rng.standard_normal(n) * 10.0).astype(np.float32)
    small = rng.uniform(-80.0, 80.0, n).astype(np.float32)
    unary = {
        "log": (pos, np.log), "safe_log": (pos, np.log), "sqrt": (pos, np.sqrt), "safe_sqrt": (pos, np.sqrt),
        "inv": (pos, lambda x: 1.0 / x), "exp": (small, np.exp), "sin": (sgn * 50, np.sin), "cos": (sgn * 50, np.cos),
        "abs": (sgn, np.abs), "neg": (sgn, np.negative), "square": (sgn, np.square), "cube": (sgn, lambda x: x ** 3),
        "relu": (sgn, lambda x: np.maximum(x, 0.0)),
    }
    N_ = dexb200.Node
    for name, (x, f) in unary.items():
        ops = dexb200.OperatorEnum({1: (name,), 2: ("*",)})
        X = np.stack([x, np.ones_like(x)])
        for form, tree in (("R", N_(1, N_(feature=1))), ("A", N_(1, N_(1, N_(feature=1), N_(feature=2))))):
            y, ok = dexb200.eval_tree_array(tree, X, ops)
            want = f(x.astype(np.float64))
            assert ok, (name, form)
            np.testing.assert_allclose(y, want, rtol=4e-7, atol=1e-37, err_msg=f"{name}/{form}")
            # the early_exit = false form of the loop (or the C++ handler it hands over to) and the
            # Float32 value of the eval_diff kernel's scalar C++ functions: within 2 ulp of the same
            y2, ok2 = dexb200.eval_tree_array(tree, X, ops, eval_context=no_exit)
            assert ok2 and np.allclose(y2, y, rtol=2.5e-7, atol=1e-37), (name, form)
            if name not in ("log", "safe_log"):      # log: polynomial (PTX) vs library (C++), both a few ulp
                assert np.array_equal(y2, y), (name, form)
    a = (rng.standard_normal(n) * 100).astype(np.float32)
    b = np.where(rng.random(n) < 0.5, -1, 1).astype(np.float32) * np.exp(rng.uniform(-20, 20, n)).astype(np.float32)
    binary = {"+": np.add, "-": np.subtract, "*": np.multiply, "/": np.divide, "max": np.maximum, "min": np.minimum}
    for name, f in binary.items():
        ops = dexb200.OperatorEnum({2: (name, "*")})
        X = np.stack([a, b, np.ones_like(a)])
        x1, x2, x3 = (N_(feature=k) for k in (1, 2, 3))
        forms = {"RR": N_(1, x1, x2), "AR": N_(1, N_(2, x1, x3), x2), "RA": N_(1, x1, N_(2, x2, x3)),
                 "RC": N_(1, x1, N_(val=2.5)), "CR": N_(1, N_(val=2.5), x2)}
        for form, tree in forms.items():
            y, ok = dexb200.eval_tree_array(tree, X, ops)
            lhs = np.full(n, 2.5) if form == "CR" else a.astype(np.float64)
            rhs = np.full(n, 2.5) if form == "RC" else b.astype(np.float64)
            assert ok, (name, form)
            np.testing.assert_allclose(y, f(lhs, rhs), rtol=2e-7, atol=1e-37, err_msg=f"{name}/{form}")
            y2, ok2 = dexb200.eval_tree_array(tree, X, ops, eval_context=no_exit)
            assert ok2 and np.array_equal(y2, y), (name, form)


@pytest.mark.parametrize("name", ["sin", "cos"])
def test_sin_cos_beyond_the_cody_waite_range(name):
    """|x| > 105615 up to the largest float: integer Payne-Hanek reduction (dex::large_sincosf),
    served by the PTX loop without leaving the warp and without a library call.  Against float64
    numpy (exact argument reduction); a sample's result must not depend on which code path its
    warp neighbours force."""
    rng = np.random.default_rng(7)
    n = 8192
    big = (10.0 ** rng.uniform(5.03, 14.4, n) * rng.choice([-1.0, 1.0], n)).astype(np.float32)
    small = (rng.standard_normal(n) * 1000).astype(np.float32)
    huge = (10.0 ** rng.uniform(14.5, 38.5, n) * rng.choice([-1.0, 1.0], n)).astype(np.float32)
    huge[:4] = [np.finfo(np.float32).max, -np.finfo(np.float32).max, 2.0 ** 120, -(2.0 ** 100)]
    f = np.sin if name == "sin" else np.cos
    ops = dexb200.OperatorEnum({1: (name,), 2: ("*",)})
    N_ = dexb200.Node
    for form, tree in (("R", N_(1, N_(feature=1))), ("A", N_(1, N_(1, N_(feature=1), N_(feature=2))))):
        alone = {}
        for label, x in (("big", big), ("small", small), ("huge", huge)):
            X = np.stack([x, np.ones_like(x)])
            y, ok = dexb200.eval_tree_array(tree, X, ops)
            assert ok, (name, form, label)
            np.testing.assert_allclose(y, f(x.astype(np.float64)), rtol=6e-7, atol=2e-9, err_msg=f"{name}/{form}/{label}")
            alone[label] = y
        # interleaved: every warp sees all kinds; each sample keeps its own result bit for bit
        mix = np.empty(3 * n, np.float32)
        mix[0::3], mix[1::3], mix[2::3] = big, small, huge
        X = np.stack([mix, np.ones_like(mix)])
        y, ok = dexb200.eval_tree_array(tree, X, ops)
        assert ok
        assert np.array_equal(y[0::3], alone["big"]) and np.array_equal(y[1::3], alone["small"]) \
            and np.array_equal(y[2::3], alone["huge"])
        # the early_exit = false form of the loop gives the same bits, and so do the scalar C++ functions
        # (dex::m_sin / m_cos, which the eval_diff kernel runs for the value)
        y3, _ = dexb200.eval_tree_array(tree, X, ops, eval_context=dexb200.EvalContext(early_exit=False))
        assert np.array_equal(y3, y)
        y4, _, _ = dexb200.eval_diff_tree_array(tree, X, ops, 1)
        assert np.array_equal(y4, y)
        # Inf -> NaN, flagged incomplete; the other samples of the warp are unaffected
        mix2 = mix.copy()
        mix2[5::97] = np.inf
        mix2[6::97] = -np.inf
        y2, ok2 = dexb200.eval_tree_array(tree, np.stack([mix2, np.ones_like(mix2)]), ops)
        assert not ok2 and np.isnan(y2[5::97]).all() and np.isnan(y2[6::97]).all()
        keep = np.ones(3 * n, bool)
        keep[5::97] = False
        keep[6::97] = False
        assert np.array_equal(y2[keep], y[keep])


@pytest.mark.parametrize("name", ["sin", "cos"])
def test_sin_cos_gradients_beyond_the_cody_waite_range(name):
    """|x| > 105615 in the gradient kernel (its PTX loop hands such a warp to the C++ handler, which
    uses dex::large_sincosf): value and derivative against float64 numpy, and bit-equal to
    eval_diff for the samples that take the large-argument form."""
    rng = np.random.default_rng(17)
    n = 4096
    big = (10.0 ** rng.uniform(5.03, 14.4, n) * rng.choice([-1.0, 1.0], n)).astype(np.float32)
    small = (rng.standard_normal(n) * 1000).astype(np.float32)
    huge = (10.0 ** rng.uniform(14.5, 38.5, n) * rng.choice([-1.0, 1.0], n)).astype(np.float32)
    huge[:2] = [np.finfo(np.float32).max, -(2.0 ** 100)]
    mix = np.empty(3 * n, np.float32)
    mix[0::3], mix[1::3], mix[2::3] = big, small, huge
    f, df = (np.sin, np.cos) if name == "sin" else (np.cos, lambda v: -np.sin(v))
    ops = dexb200.OperatorEnum({1: (name,), 2: ("*",)})
    N_ = dexb200.Node
    X = np.stack([mix, np.ones_like(mix)])
    x64 = mix.astype(np.float64)
    for form, tree in (("R", N_(1, N_(feature=1))), ("A", N_(1, N_(1, N_(feature=1), N_(feature=2))))):
        y, g, ok = dexb200.eval_grad_tree_array(tree, X, ops, variable=True)
        assert ok, (name, form)
        np.testing.assert_allclose(y, f(x64), rtol=6e-7, atol=2e-9, err_msg=f"{name}/{form} value")
        np.testing.assert_allclose(g[0], df(x64), rtol=6e-7, atol=2e-9, err_msg=f"{name}/{form} d/dx1")
        if form == "A":     # d/dx2 of op(x1 * x2) at x2 = 1 is x1 * op'(x1)
            ref = x64 * df(x64)
            fin = np.abs(ref) < 3e38
            np.testing.assert_allclose(g[1][fin], ref[fin], rtol=1e-6, atol=2e-9, err_msg=f"{name}/{form} d/dx2")
        yd, gd, _ = dexb200.eval_diff_tree_array(tree, X, ops, 1)
        far = np.abs(mix) > 105615.0
        assert np.array_equal(y[far], yd[far]) and np.array_equal(g[0][far], gd[far])
        # Inf: NaN value and derivative, incomplete
        X2 = X.copy()
        X2[0, 7::101] = np.inf
        y2, g2, ok2 = dexb200.eval_grad_tree_array(tree, X2, ops, variable=True)
        assert not ok2


@pytest.mark.parametrize("P_,N", [(64, 4096), (400, 8192)])   # the second takes the sliced D2H pipeline
def test_host_entry_point_can_leave_incomplete_rows_on_the_device(P_, N):
    """DEX_EVAL_SKIP_INCOMPLETE: flags and the rows of complete trees as usual; the rows of incomplete
    trees are not transferred (out_host keeps what it held)."""
    import torch
    ops = dexb200.OperatorEnum(treegen.OPSET_A)
    nodes, offsets = treegen.gen_population(P_, 7, 2, 4, 5, seed=62)
    Xh = torch.from_numpy(np.random.default_rng(11).standard_normal((N, 5)).astype(np.float32)).pin_memory()
    pop = D.Population(None, ops, np.float32, wire=(nodes, offsets))
    out_d, ok_d = pop.eval(Xh.cuda().T)
    okd = ok_d.cpu().numpy().astype(bool)
    assert 0 < okd.sum() < P_
    for pinned in (True, False):
        out_h = torch.full((P_, N), -7.0, dtype=torch.float32)
        ok_h = torch.full((P_,), 9, dtype=torch.uint8)
        if pinned:
            out_h, ok_h = out_h.pin_memory(), ok_h.pin_memory()
        pop.eval_host(Xh, out_h, ok_h, skip_incomplete=True)
        assert (ok_h.numpy().astype(bool) == okd).all() and set(np.unique(ok_h.numpy())) <= {0, 1}
        a = out_h.numpy()
        assert np.array_equal(a[okd], out_d.cpu().numpy()[okd])
        assert (a[~okd] == -7.0).all()
    # without early exit the flag has no effect: every row is an output
    out_h = torch.full((P_, N), -7.0, dtype=torch.float32).pin_memory()
    ok_h = torch.zeros(P_, dtype=torch.uint8).pin_memory()
    C_ = D.lib().dex_eval_host
    pop.ctx.check(C_(pop.ctx.h, pop.h, D._ptr(Xh), 5, N, 5, D._ptr(out_h), N, D._ptr(ok_h), D.EVAL_SKIP_INCOMPLETE))
    full, _ = pop.eval(Xh.cuda().T, early_exit=False)
    f, a = full.cpu().numpy(), out_h.numpy()
    assert ((a == f) | (np.isnan(a) & np.isnan(f))).all()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_early_exit_leaves_the_other_trees_alone(dtype):
    """A warp that meets a non-finite checked value skips the rest of that tree's tape
    (`@return_on_nonfinite_array`, /root/reference/src/Evaluate.jl:26-32): the tree is flagged, and
    the trees that follow it in the same CTA are evaluated exactly as if it were not there —
    wherever in its tape the failure happens and however many samples trip it."""
    ops = dexb200.OperatorEnum({1: ("log", "exp", "cos"), 2: ("+", "*", "/")})
    N_ = dexb200.Node
    x1, x2 = N_(feature=1, T=dtype), N_(feature=2, T=dtype)
    c = lambda v: N_(val=v, T=dtype)
    log, exp, cos = (lambda a, i=i: N_(i, a) for i in (1, 2, 3))
    add, mul, div = (lambda a, b, i=i: N_(i, a, b) for i in (1, 2, 3))
    good = [add(cos(mul(x1, x2)), mul(exp(cos(x2)), x1)), mul(add(x1, c(2.5)), cos(add(exp(cos(x1)), x2))),
            div(exp(cos(x1)), add(exp(x2), c(1.0)))]
    bad = [add(log(x1), mul(cos(x2), x1)),                                       # fails at once, in about half of the samples
           mul(cos(add(exp(cos(mul(x1, x2))), x1)), log(add(mul(x1, c(0.0)), c(-1.0)))),   # fails late, in every sample
           add(div(add(cos(x1), x2), mul(x1, c(0.0))), exp(cos(mul(x2, x1)))),   # division by zero in the middle
           add(exp(exp(exp(x1))), cos(x2))]                                      # overflows in a few percent of the samples
    rng = np.random.default_rng(23)
    X = rng.standard_normal((2, 5000)).astype(dtype)
    trees = []
    for r in range(40):
        trees += [bad[r % 4], good[r % 3], bad[(r + 1) % 4], bad[(r + 2) % 4], good[(r + 1) % 3]]
    pop = D.Population(trees, ops, dtype)
    out, ok = pop.eval(X)
    ok = ok.cpu().numpy().astype(bool)
    out = out.cpu().numpy()
    alone, ok_alone = D.Population(good, ops, dtype).eval(X)
    assert ok_alone.cpu().numpy().all()
    alone = alone.cpu().numpy()
    for i, t in enumerate(trees):
        is_good = any(t is g for g in good)
        assert ok[i] == is_good, i
        if is_good:
            assert np.array_equal(out[i], alone[[t is g for g in good].index(True)]), i
    # with early_exit off nothing is skipped: every row is computed to the end
    full, ok_full = pop.eval(X, early_exit=False)
    full = full.cpu().numpy()
    assert not np.isfinite(full[0]).all() and np.isfinite(full[0]).any()       # log(x1): NaN only where x1 < 0
    assert np.array_equal(full[ok], out[ok])


def test_contexts_on_several_host_threads_are_independent():
    """One context per host thread, like the reference's task-local state (SURVEY §8b "Threading"): packing,
    evaluation and gradients from four threads at once (ctypes releases the GIL inside the library) give
    the single-threaded answers."""
    import threading
    import torch
    ops = dexb200.OperatorEnum(treegen.OPSET_A)
    # a different number of features per thread: the launches need different amounts of dynamic shared
    # memory, and that limit is an attribute of the (device, kernel), shared by all threads
    Fs = (2, 5, 8, 11)
    Xs = [np.random.default_rng(12 + i).standard_normal((F, 3000)).astype(np.float32) for i, F in enumerate(Fs)]
    pops = [treegen.gen_population(120, 6, 2, 4, F, seed=70 + i) for i, F in enumerate(Fs)]
    want = []
    for (nodes, offsets), X in zip(pops, Xs):
        pop = D.Population(None, ops, np.float32, wire=(nodes, offsets))
        o, k = pop.eval(X)
        _, g, off, gk = pop.eval_grad(X, D.GRAD_FEATURES)
        want.append((o.cpu().numpy(), k.cpu().numpy(), g.cpu().numpy(), gk.cpu().numpy()))
    errors = []

    def work(i):
        try:
            torch.cuda.set_device(0)
            ctx = D.Context(0)
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                X = Xs[i]
                for rep in range(6):
                    nodes, offsets = pops[i]
                    pop = D.Population(None, ops, np.float32, wire=(nodes, offsets), ctx=ctx)      # re-packed every time
                    o, k = pop.eval(X)
                    _, g, off, gk = pop.eval_grad(X, D.GRAD_FEATURES)
                    stream.synchronize()
                    o, k, g, gk = o.cpu().numpy(), k.cpu().numpy(), g.cpu().numpy(), gk.cpu().numpy()
                    good = want[i][1].astype(bool)
                    ggood = want[i][3].astype(bool)
                    assert (k == want[i][1]).all() and (gk == want[i][3]).all()
                    assert np.array_equal(o[good], want[i][0][good])
                    G = g.reshape(120, -1)
                    assert np.array_equal(G[ggood], want[i][2].reshape(120, -1)[ggood])
        except BaseException as e:     # noqa: BLE001 - reported by the main thread
            errors.append((i, repr(e)))

    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors


def test_pack_evaluate_destroy_cycles_do_not_leak():
    """Callers whose trees change every generation pack and destroy a population per step; contexts come
    and go with their threads.  Device memory must come back (the population pool recycles blocks)."""
    import gc
    import torch
    ops = dexb200.OperatorEnum(treegen.OPSET_A)
    X = torch.randn((1000, 5), device="cuda")

    def cycle(ctx, n, seed0):
        for i in range(n):
            nodes, offsets = treegen.gen_population(50 + 37 * (i % 5), 5 + i % 3, 2, 4, 5, seed=seed0 + i)
            pop = D.Population(None, ops, np.float32, wire=(nodes, offsets), ctx=ctx)
            pop.eval(X.T)
            if i % 7 == 0:
                pop.eval_grad(X.T, D.GRAD_BOTH)
            del pop
    ctx = D.Context(0)
    cycle(ctx, 40, 0)                     # fills the pools / scratch buffers once
    torch.cuda.synchronize()
    gc.collect()
    free0, _ = torch.cuda.mem_get_info()
    cycle(ctx, 300, 1000)
    for k in range(10):                   # short-lived contexts
        c = D.Context(0)
        cycle(c, 3, 5000 + 10 * k)
        c.synchronize()
        del c
    torch.cuda.synchronize()
    gc.collect()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < 64 << 20, f"device memory shrank by {(free0 - free1) >> 20} MiB over 330 pack/destroy cycles"


def test_set_constants_equals_repack(oracle):
    ops = dexb200.OperatorEnum(treegen.OPSET_A)
    nodes, offsets = treegen.gen_population(80, 6, 2, 4, 3, seed=41)
    X = np.random.default_rng(8).standard_normal((3, 500)).astype(np.float32)
    pop = D.Population(None, ops, np.float32, wire=(nodes, offsets))
    c = pop.get_constants()
    isc = (nodes["degree"] == 0) & (nodes["kind"] == 0)
    np.testing.assert_array_equal(c, nodes["val"][isc].astype(np.float32))      # tree order, leaf order
    newc = (c * 0.5 + 0.25).astype(np.float32)
    pop.set_constants(newc)
    out1, ok1 = pop.eval(X)
    nodes2 = nodes.copy()
    nodes2["val"][isc] = newc
    pop2 = D.Population(None, ops, np.float32, wire=(nodes2, offsets))
    out2, ok2 = pop2.eval(X)
    assert (ok1 == ok2).all()
    a, b = out1.cpu().numpy(), out2.cpu().numpy()
    assert ((a == b) | (np.isnan(a) & np.isnan(b))).all()
    # gradient tape sees the new constants too
    y1, g1, _, _ = pop.eval_grad(X, D.GRAD_CONSTANTS)
    y2, g2, _, _ = pop2.eval_grad(X, D.GRAD_CONSTANTS)
    ga, gb = g1.cpu().numpy(), g2.cpu().numpy()
    assert ((ga == gb) | (np.isnan(ga) & np.isnan(gb))).all()


def test_fused_loss_matches_materialised_results():
    ops = dexb200.OperatorEnum(treegen.OPSET_A)
    nodes, offsets = treegen.gen_population(100, 7, 2, 4, 5, seed=51)
    rng = np.random.default_rng(9)
    X = rng.standard_normal((5, 5000)).astype(np.float32)
    y = rng.standard_normal(5000).astype(np.float32)
    pop = D.Population(None, ops, np.float32, wire=(nodes, offsets))
    out, ok = pop.eval(X)
    loss, ok2 = pop.eval_loss(X, y)
    assert (ok == ok2).all()
    o = out.cpu().numpy().astype(np.float64)
    want = ((o - y[None, :].astype(np.float64)) ** 2).mean(axis=1)
    got = loss.cpu().numpy()
    good = ok.cpu().numpy().astype(bool)
    np.testing.assert_allclose(got[good], want[good], rtol=1e-10)


def test_weighted_fused_loss_matches_materialised_results():
    ops = dexb200.OperatorEnum(treegen.OPSET_A)
    nodes, offsets = treegen.gen_population(60, 6, 2, 4, 4, seed=52)
    rng = np.random.default_rng(12)
    X = rng.standard_normal((4, 3001)).astype(np.float32)
    y = rng.standard_normal(3001).astype(np.float32)
    w = rng.uniform(0.1, 3.0, 3001).astype(np.float32)
    pop = D.Population(None, ops, np.float32, wire=(nodes, offsets))
    out, ok = pop.eval(X)
    loss, ok2 = pop.eval_loss(X, y, weights=w)
    assert (ok == ok2).all()
    o = out.cpu().numpy().astype(np.float64)
    w64 = w.astype(np.float64)
    want = (w64[None, :] * (o - y[None, :].astype(np.float64)) ** 2).sum(axis=1) / w64.sum()
    good = ok.cpu().numpy().astype(bool)
    np.testing.assert_allclose(loss.cpu().numpy()[good], want[good], rtol=1e-10)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mode", ["constants", "features", "both"])
@pytest.mark.parametrize("weighted", [False, True])
def test_fused_loss_gradient_matches_oracle(dtype, mode, weighted, oracle):
    """dex_eval_loss_grad: loss and d loss / d theta per tree, reduced inside the interpreter,
    against the same contraction of the oracle's (value, gradient) arrays
    (the reference's pullback, src/ChainRules.jl:56-77, with dY = 2 w (y_pred - y) / sum w)."""
    spec = {1: ("cos", "exp", "sin"), 2: ("+", "-", "*", "/")}
    ops = dexb200.OperatorEnum(spec)
    nodes, offsets = treegen.gen_population(90, 6, 3, 4, 3, seed=77, dtype=dtype)
    rng = np.random.default_rng(21)
    N = 1500
    X = rng.standard_normal((3, N)).astype(dtype)
    y = rng.standard_normal(N).astype(dtype)
    w = rng.uniform(0.2, 2.0, N).astype(dtype) if weighted else None
    omode = {"features": oracle.GRAD_FEATURES, "constants": oracle.GRAD_CONSTANTS, "both": oracle.GRAD_BOTH}[mode]
    dmode = {"features": D.GRAD_FEATURES, "constants": D.GRAD_CONSTANTS, "both": D.GRAD_BOTH}[mode]
    pop = D.Population(None, ops, dtype, wire=(nodes, offsets))
    loss, grad, off, ok = pop.eval_loss_grad(X, y, dmode, weights=w)
    loss, grad, ok = loss.cpu().numpy(), grad.cpu().numpy(), ok.cpu().numpy().astype(bool)
    ref, rgrads, rok = oracle.eval_grad_population(nodes, offsets, ops.opcodes, X, omode)
    _, _, rok_elem = oracle.eval_grad_population(nodes, offsets, ops.opcodes, X, omode | oracle.GRAD_ELEMENTWISE)
    _flags_agree(ok, rok, rok_elem, f"loss grad/{mode}")
    # conditioning yardsticks (as in _check_population): how far the oracle's own value / gradient
    # move when X is perturbed by one ulp and, for Float32, when evaluated in Float64; chaotic
    # compositions such as cos(exp(exp(x))) are compared at 30x that movement
    ref_p, rgrads_p, _ = oracle.eval_grad_population(nodes, offsets, ops.opcodes, np.nextafter(X, dtype(np.inf)), omode)
    ref64, rgrads64, _ = oracle.eval_grad_population(nodes, offsets, ops.opcodes, X.astype(np.float64), omode)
    w64 = np.ones(N) if w is None else w.astype(np.float64)
    n_checked = n_chaotic = 0
    for t in np.nonzero(rok)[0]:
        G = rgrads[t].shape[0]
        cond = max(_relerr(ref_p[t], ref[t]), _relerr(rgrads_p[t], rgrads[t]) if G else 0.0)
        if dtype == np.float32:
            cond = max(cond, _relerr(ref[t], ref64[t]), _relerr(rgrads[t], rgrads64[t]) if G else 0.0)
        if not np.isfinite(cond) or 30 * cond > 1e-2:
            n_chaotic += 1
            continue
        tol = max(2e-4 if dtype == np.float32 else 1e-6, 30 * cond)
        r = ref[t].astype(np.float64) - y.astype(np.float64)
        want_loss = (w64 * r * r).sum() / w64.sum()
        assert abs(loss[t] - want_loss) <= tol * max(want_loss, 1e-30), (t, loss[t], want_loss)
        assert off[t + 1] - off[t] == G
        if G:
            terms = 2.0 * w64[None, :] * r[None, :] * rgrads[t].astype(np.float64)     # (G, N)
            want = terms.sum(axis=1) / w64.sum()
            scale = np.abs(terms).sum(axis=1) / w64.sum()                                # conditioning of the sum
            got = grad[off[t]:off[t + 1]]
            assert np.all(np.abs(got - want) <= tol * np.maximum(scale, 1e-30)), (t, mode, got, want)
            n_checked += 1
    assert n_checked > 10 and n_chaotic <= max(2, 0.12 * int(rok.sum()))


@pytest.mark.parametrize("P_,N", [(64, 4096), (400, 8192)])   # the second takes the sliced D2H pipeline
def test_host_entry_point_matches_device_entry_point(P_, N):
    import torch
    ops = dexb200.OperatorEnum(treegen.OPSET_A)
    nodes, offsets = treegen.gen_population(P_, 6, 2, 4, 5, seed=61)
    Xh = torch.from_numpy(np.random.default_rng(10).standard_normal((N, 5)).astype(np.float32)).pin_memory()
    pop = D.Population(None, ops, np.float32, wire=(nodes, offsets))
    out_h = torch.empty((P_, N), dtype=torch.float32).pin_memory()
    ok_h = torch.empty(P_, dtype=torch.uint8).pin_memory()
    pop.eval_host(Xh, out_h, ok_h)
    out_d, ok_d = pop.eval(Xh.cuda().T)
    a, b = out_h.numpy(), out_d.cpu().numpy()
    assert ((a == b) | (np.isnan(a) & np.isnan(b))).all()
    assert (ok_h.numpy() == ok_d.cpu().numpy()).all()
    # pageable numpy buffers work too
    out_n = np.empty((P_, N), np.float32)
    ok_n = np.empty(P_, np.uint8)
    pop.eval_host(Xh.numpy().copy(), out_n, ok_n)
    assert ((out_n == b) | (np.isnan(out_n) & np.isnan(b))).all()


# ---------------------------------------------------------------------------------------
# 5. BASELINE.json sizes: size-independent properties + a sampled oracle comparison
# ---------------------------------------------------------------------------------------
def test_full_size_config2_properties(oracle):
    """configs[1]: 1k depth-8 trees, 5 features, Float32, 2^16 samples."""
    import torch
    ops = dexb200.OperatorEnum(treegen.OPSET_A)
    nodes, offsets = treegen.gen_population(1000, 8, 2, 4, 5, seed=0)
    N = 1 << 16
    X = np.random.default_rng(0).standard_normal((5, N)).astype(np.float32)
    pop = D.Population(None, ops, np.float32, wire=(nodes, offsets))
    Xd = torch.from_numpy(np.ascontiguousarray(X.T)).cuda()
    out, ok = pop.eval(Xd.T)
    good = ok.bool()      # rows of incomplete trees are unspecified (early exit), as in the reference
    out = out[good]
    # (a) column permutation equivariance: tiles / lanes see different samples, same answers
    perm = torch.randperm(N, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    out_p, ok_p = pop.eval(Xd[perm].T)
    assert torch.equal(ok, ok_p)
    assert torch.equal(out[:, perm], out_p[good])
    # (b) evaluating two halves == evaluating the whole (tile boundaries do not matter)
    h1, k1 = pop.eval(Xd[: N // 2 + 17].T)
    h2, k2 = pop.eval(Xd[N // 2 + 17:].T)
    assert torch.equal(k1 & k2, ok)
    assert torch.equal(torch.cat([h1, h2], dim=1)[good], out)
    # (c) idempotence
    out2, ok2 = pop.eval(Xd.T)
    assert torch.equal(ok, ok2) and torch.equal(out2[good], out)
    okh = ok.cpu().numpy().astype(bool)
    # (d) all 1 000 trees against the oracle (flags exactly, values in the classes of parity_util)
    errs, ok_again = _check_population(oracle, nodes, offsets, ops, X, np.float32, label="C2 full size", pop=pop)
    assert (ok_again == okh).all() and len(errs) > 600
    assert bool(torch.isfinite(out).all())       # complete => every output finite


def _subset(nodes, offsets, sel):
    parts = [nodes[offsets[t]:offsets[t + 1]] for t in sel]
    off = np.zeros(len(sel) + 1, dtype=np.int64)
    np.cumsum([len(p) for p in parts], out=off[1:])
    return np.concatenate(parts), off
