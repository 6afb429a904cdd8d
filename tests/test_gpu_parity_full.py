"""GPU parity at BASELINE.json sizes, gradients of EVERY operator of include/dex_ops.def, and
eval_diff populations — all against the CPU oracle, through the C ABI.  Needs a GPU (`-m gpu`).

Sizes: C2 in full is in test_gpu_parity.py::test_full_size_config2_properties; here C3 (gradients
of 250 of the C2 trees at the full 2^16 samples), one C4 shard shape (600 depth-12 trees, 10
features, 2^14 samples), C5 (parametric, the full 2^18 samples on 250 of the 1 000 trees).  The
oracle evaluates each of them in seconds on the box's host cores.
"""
import numpy as np
import pytest

import dexb200
from dexb200 import device as D
from dexb200 import treegen
from tests.parity_util import FAILED, RTOL, check_trees, part_verdict, relerr, same_nonfinite, tree_verdict
from tests.test_gpu_parity import _check_population, _flags_agree, _grad_verdict, _subset

pytestmark = pytest.mark.gpu


def _grad_yardsticks(oracle, nodes, offsets, ops, X, mode):
    """The oracle's (value, gradient) under the conditioning perturbations of tests/parity_util.py:
    X one ulp up / down, Float64 arithmetic (Float32 inputs only), every transcendental result one
    ulp up / down."""
    dtype = X.dtype.type
    ys = []
    for direction in (np.inf, -np.inf):
        r, g, _ = oracle.eval_grad_population(nodes, offsets, ops.opcodes, np.nextafter(X, dtype(direction)), mode)
        ys.append((r, g))
    if dtype == np.float32:
        r, g, _ = oracle.eval_grad_population(nodes, offsets, ops.opcodes, X.astype(np.float64), mode)
        ys.append((r, g))
    for n in (1, -1):
        try:
            oracle.set_ulp_nudge(n)
            r, g, _ = oracle.eval_grad_population(nodes, offsets, ops.opcodes, X, mode)
        finally:
            oracle.set_ulp_nudge(0)
        ys.append((r, g))
    return ys


# ---------------------------------------------------------------------------------------
# BASELINE.json sizes
# ---------------------------------------------------------------------------------------
def test_config3_gradients_at_full_sample_count(oracle):
    """configs[2]: eval_grad_tree_array d/dX on the C2 population, 2^16 samples; every 4th tree."""
    ops = dexb200.OperatorEnum(treegen.OPSET_A)
    nodes, offsets = treegen.gen_population(1000, 8, 2, 4, 5, seed=0)
    sn, so = _subset(nodes, offsets, np.arange(0, 1000, 4))
    N = 1 << 16
    X = np.random.default_rng(0).standard_normal((5, N)).astype(np.float32)
    pop = D.Population(None, ops, np.float32, wire=(sn, so))
    out, grad, off, ok = pop.eval_grad(X, D.GRAD_FEATURES)
    out, grad, ok = out.cpu().numpy(), grad.cpu().numpy(), ok.cpu().numpy().astype(bool)
    mode = oracle.GRAD_FEATURES
    ref, rgrads, rok = oracle.eval_grad_population(sn, so, ops.opcodes, X, mode)
    _, _, rok_elem = oracle.eval_grad_population(sn, so, ops.opcodes, X, mode | oracle.GRAD_ELEMENTWISE)
    _flags_agree(ok, rok, rok_elem, "C3")
    yards = _grad_yardsticks(oracle, sn, so, ops, X, mode)
    verdicts = []
    for t in np.nonzero(rok)[0]:
        g = grad[off[t]:off[t + 1]].reshape(N, 5).T
        verdicts.append(_grad_verdict(np.float32, out[t], g, ref[t], rgrads[t], [(y[t], yg[t]) for y, yg in yards]))
    st = check_trees("C3 full sample count (250 trees x 2^16)", np.float32, verdicts, min_strict=0.75,
                     ids=[int(t) for t in np.nonzero(rok)[0]])
    assert st["n_complete"] > 100


def test_config4_shard_shape(oracle):
    """configs[3]: depth-12 trees, 10 features; 600 trees x 2^14 samples of one shard."""
    ops = dexb200.OperatorEnum(treegen.OPSET_A)
    nodes, offsets = treegen.gen_population(600, 12, 2, 4, 10, seed=0)
    X = np.random.default_rng(3).standard_normal((10, 1 << 14)).astype(np.float32)
    errs, ok = _check_population(oracle, nodes, offsets, ops, X, np.float32, label="C4 shard shape (600 x 2^14)",
                                 min_strict=0.65)      # depth-12 trees: more poles per tree
    assert len(errs) > 100


_WIDE_CHILD = """
import sys, numpy as np
import dexb200
from dexb200 import device as D, treegen
F, N, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
ops = dexb200.OperatorEnum(treegen.OPSET_A)
nodes, offsets = treegen.gen_population(200, 9, 2, 4, F, seed=8)
X = np.random.default_rng(11).standard_normal((F, N)).astype(np.float32)
pop = D.Population(None, ops, np.float32, wire=(nodes, offsets))
o, ok = pop.eval(X)
np.savez(out, o=o.cpu().numpy(), ok=ok.cpu().numpy())
"""


@pytest.mark.parametrize("F,N", [(12, 5000), (24, 3 * 2048), (60, 4099)])
def test_wide_inputs_read_feature_rows_through_l1(F, N, oracle, tmp_path):
    """More than 9 rows per 2 048-sample tile: the wide-input kernel keeps 8 rows in shared memory and
    reads the other feature rows from the feature-major global copy (dex_eval.cu eval_num_tiles, GX).
    Values vs the oracle, bit-equality with the all-shared-memory kernel (a child process with
    DEXB200_NO_GX=1), fused loss vs the materialised rows."""
    import os
    import subprocess
    import sys
    ops = dexb200.OperatorEnum(treegen.OPSET_A)
    nodes, offsets = treegen.gen_population(200, 9, 2, 4, F, seed=8)
    X = np.random.default_rng(11).standard_normal((F, N)).astype(np.float32)
    errs, ok = _check_population(oracle, nodes, offsets, ops, X, np.float32, label=f"wide input F={F} N={N}",
                                 min_strict=0.7)
    assert ok.sum() > 40
    _check_population(oracle, nodes, offsets, ops, X, np.float32, ctx={"early_exit": False},
                      label=f"wide input F={F} N={N} early_exit=false", min_strict=0.7)
    pop = D.Population(None, ops, np.float32, wire=(nodes, offsets))
    out, dok = pop.eval(X)
    out, dok = out.cpu().numpy(), dok.cpu().numpy().astype(bool)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, DEXB200_NO_GX="1", PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
    dst = str(tmp_path / "nogx.npz")
    subprocess.run([sys.executable, "-c", _WIDE_CHILD, str(F), str(N), dst], check=True, env=env, cwd=root)
    ref = np.load(dst)
    assert (ref["ok"].astype(bool) == dok).all()
    assert np.array_equal(ref["o"][dok].view(np.uint32), out[dok].view(np.uint32))
    y = np.random.default_rng(13).standard_normal(N).astype(np.float32)
    loss, lok = pop.eval_loss(X, y)
    assert (lok.cpu().numpy().astype(bool) == dok).all()
    want = ((out.astype(np.float64) - y.astype(np.float64)[None]) ** 2).mean(axis=1)
    np.testing.assert_allclose(loss.cpu().numpy()[dok], want[dok], rtol=1e-10)


def test_config5_parametric_at_full_sample_count(oracle):
    """configs[4]: ParametricExpression, 3 parameters x 10 classes, 2^18 samples; every 4th tree."""
    ops = dexb200.OperatorEnum(treegen.OPSET_A)
    n_params, n_classes, F, N = 3, 10, 5, 1 << 18
    nodes, offsets = treegen.gen_population(1000, 8, 2, 4, F, seed=0, n_params=n_params)
    sn, so = _subset(nodes, offsets, np.arange(0, 1000, 4))
    P_ = len(so) - 1
    rng = np.random.default_rng(5)
    X = rng.standard_normal((F, N)).astype(np.float32)
    params = rng.standard_normal((P_, n_params, n_classes)).astype(np.float32)
    cls0 = rng.integers(0, n_classes, N)
    pop = D.Population(None, ops, np.float32, wire=(sn, so))
    out, ok = pop.eval_parametric(X, params, cls0)
    out, ok = out.cpu().numpy(), ok.cpu().numpy().astype(bool)
    ref, rok = oracle.eval_parametric_population(sn, so, ops.opcodes, X, params, cls0)
    _, rok_elem = oracle.eval_parametric_population(sn, so, ops.opcodes, X, params, cls0,
                                                    oracle.DEFAULT_FLAGS | oracle.ELEMENTWISE)
    _flags_agree(ok, rok, rok_elem, "C5")
    yards = [oracle.eval_parametric_population(sn, so, ops.opcodes, X.astype(np.float64), params.astype(np.float64), cls0)[0]]
    for direction in (np.inf, -np.inf):
        yards.append(oracle.eval_parametric_population(sn, so, ops.opcodes, np.nextafter(X, np.float32(direction)), params, cls0)[0])
    for n in (1, -1):
        try:
            oracle.set_ulp_nudge(n)
            yards.append(oracle.eval_parametric_population(sn, so, ops.opcodes, X, params, cls0)[0])
        finally:
            oracle.set_ulp_nudge(0)
    verdicts = [tree_verdict(np.float32, [(out[t], ref[t], [y[t] for y in yards])]) for t in np.nonzero(rok)[0]]
    st = check_trees("C5 full sample count (250 trees x 2^18)", np.float32, verdicts, min_strict=0.8,
                     ids=[int(t) for t in np.nonzero(rok)[0]])
    assert st["n_complete"] > 80


# ---------------------------------------------------------------------------------------
# gradients of every operator of the table (the reference gets them from Zygote for ANY
# operator, ext/DynamicExpressionsZygoteExt.jl:7-15, src/EvaluateDerivative.jl:340-365)
# ---------------------------------------------------------------------------------------
def _operator_trees(code, deg, dtype):
    """Trees exercising operator index 1 of its degree in leaf / accumulator / constant / stack
    forms.  Operator sets: {1: (op, 'cos'), 2: ('*', '+', op)} etc. built by the caller."""
    N_ = dexb200.Node
    x1, x2, x3 = (N_(feature=k, T=dtype) for k in (1, 2, 3))
    c = lambda v: N_(val=v, T=dtype)
    return N_, x1, x2, x3, c


def _all_operator_cases(dtype):
    """[(label, OperatorEnum, trees)] for every opcode of include/dex_ops.def."""
    cases = []
    for code, (sym, deg, name) in sorted(dexb200.OPCODE_INFO.items()):
        N_, x1, x2, x3, c = _operator_trees(code, deg, dtype)
        if deg == 1:
            ops = dexb200.OperatorEnum({1: (name,), 2: ("*", "+")})
            mul, add = 1, 2
            inner = lambda: N_(mul, x1, x2)
            trees = {"leaf": N_(1, x1), "acc": N_(1, inner()), "const": N_(add, x1, N_(1, c(0.7))),
                     "acc+const": N_(1, N_(add, inner(), c(0.25))), "nested": N_(1, N_(1, x2)),
                     "slot": N_(add, N_(1, inner()), N_(1, N_(add, x2, x3)))}
        elif deg == 2:
            base = ("*", "+")
            ops = dexb200.OperatorEnum({2: base + (name,)}) if name not in base else dexb200.OperatorEnum({2: base})
            op = (base + (name,)).index(name) + 1
            mul, add = 1, 2
            inner = lambda: N_(mul, x1, x2)
            other = lambda: N_(add, x2, x3)
            trees = {"RR": N_(op, x1, x2), "RR_same": N_(op, x1, x1), "AR": N_(op, inner(), x3), "RA": N_(op, x3, inner()),
                     "AC": N_(op, inner(), c(1.5)), "CA": N_(op, c(0.75), inner()), "RC": N_(op, x1, c(2.5)),
                     "CR": N_(op, c(2.5), x3), "CC": N_(add, x1, N_(op, c(0.5), c(1.5))),
                     "slot_acc": N_(op, inner(), other()), "acc_slot": N_(op, other(), inner())}
        else:
            ops = dexb200.OperatorEnum({2: ("*", "+"), 3: (name,)})
            mul, add = 1, 2
            inner = lambda: N_(mul, x1, x2)
            other = lambda: N_(add, x2, x3)
            t = lambda a, b, cc: N_(op=1, children=(a, b, cc))
            trees = {"leaves": t(x1, x2, x3), "ops": t(inner(), other(), N_(mul, x3, x1)),
                     "mixed1": t(c(0.5), x2, inner()), "mixed2": t(inner(), c(1.25), x3),
                     "mixed3": t(x3, other(), c(2.0)), "consts": N_(add, x1, t(c(0.5), c(0.25), c(1.5)))}
        cases.append((sym, ops, trees))
    return cases


# input ranges chosen so that every operator has a range on which it is finite and smooth:
# (0.1, 0.9): |x| < 1 for asin / acos / atanh, positive for log / sqrt;  (1.1, 2): acosh;
# signed normal: the failure paths and sign-dependent branches
X_RANGES = [("unit", lambda r, n: r.uniform(0.1, 0.9, (3, n))), ("above1", lambda r, n: r.uniform(1.1, 2.0, (3, n))),
            ("normal", lambda r, n: r.standard_normal((3, n)))]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mode", ["features", "constants", "both"])
def test_gradient_of_every_operator_in_the_table(dtype, mode, oracle):
    """Every opcode of include/dex_ops.def x operand forms x gradient modes vs the oracle."""
    omode = {"features": oracle.GRAD_FEATURES, "constants": oracle.GRAD_CONSTANTS, "both": oracle.GRAD_BOTH}[mode]
    dmode = {"features": D.GRAD_FEATURES, "constants": D.GRAD_CONSTANTS, "both": D.GRAD_BOTH}[mode]
    rng = np.random.default_rng(123)
    N = 520
    tol = 3e-5 if dtype == np.float32 else 1e-10
    covered, never_ok = 0, []
    n_strict = n_total = 0
    for sym, ops, trees in _all_operator_cases(dtype):
        labels = list(trees)
        nodes, offsets = dexb200.to_wire_population([trees[k] for k in labels])
        pop = D.Population(None, ops, dtype, wire=(nodes, offsets))
        compared = set()
        for rname, gen in X_RANGES:
            X = gen(rng, N).astype(dtype)
            out, grad, off, ok = pop.eval_grad(X, dmode)
            out, grad, ok = out.cpu().numpy(), grad.cpu().numpy(), ok.cpu().numpy().astype(bool)
            ref, rgrads, rok = oracle.eval_grad_population(nodes, offsets, ops.opcodes, X, omode)
            _, _, rok_elem = oracle.eval_grad_population(nodes, offsets, ops.opcodes, X, omode | oracle.GRAD_ELEMENTWISE)
            _flags_agree(ok, rok, rok_elem, f"{sym}/{rname}/{mode}")
            yards = _grad_yardsticks(oracle, nodes, offsets, ops, X, omode)
            for t in np.nonzero(rok)[0]:
                G = rgrads[t].shape[0]
                g = grad[off[t]:off[t + 1]].reshape(N, G).T
                v = _grad_verdict(dtype, out[t], g, ref[t], rgrads[t], [(y[t], yg[t]) for y, yg in yards])
                assert v[0] != FAILED, (sym, labels[t], rname, mode, v)
                n_strict += v[0] == 0 and v[1] <= tol
                n_total += 1
                compared.add(labels[t])
        if compared:
            covered += 1
        else:
            never_ok.append(sym)
    assert not never_ok, f"operators whose trees were never complete on any input range: {never_ok}"
    assert covered == len(dexb200.OPCODE_INFO)
    # the bulk agrees to a few ulp (not just the north_star tolerance)
    assert n_strict >= 0.9 * n_total, (n_strict, n_total)


# ---------------------------------------------------------------------------------------
# eval_diff_tree_array for populations (src/EvaluateDerivative.jl:40-168): never checks,
# so rows keep the Inf / NaN of the reference's arithmetic
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_eval_diff_population_matches_oracle(dtype, oracle):
    spec = {1: ("cos", "exp", "sin", "sqrt", "log"), 2: ("+", "-", "*", "/")}
    ops = dexb200.OperatorEnum(spec)
    F, N, P_ = 4, 600, 160
    nodes, offsets = treegen.gen_population(P_, 7, 5, 4, F, seed=77, dtype=dtype)
    rng = np.random.default_rng(8)
    X = rng.standard_normal((F, N)).astype(dtype)
    X[1, 17] = np.inf
    X[2, 99] = -np.inf
    X[0, 300] = np.nan
    X[3, 5] = 0.0
    pop = D.Population(None, ops, dtype, wire=(nodes, offsets))
    n_pattern = n_values = 0
    for direction in range(F):
        out, dout, ok = pop.eval_diff(X, direction)
        out, dout, ok = out.cpu().numpy(), dout.cpu().numpy(), ok.cpu().numpy()
        assert ok.all()                                    # eval_diff never reports failure (:68-85)
        pairs = []
        for t in range(P_):
            w = nodes[offsets[t]:offsets[t + 1]]
            ry, rd, rok = oracle.eval_diff_tree_array(w, ops.opcodes, X, direction)
            assert rok
            ry64, rd64, _ = oracle.eval_diff_tree_array(w, ops.opcodes, X.astype(np.float64), direction)
            ryp, rdp, _ = oracle.eval_diff_tree_array(w, ops.opcodes, np.nextafter(X, dtype(np.inf)), direction)
            # the non-finite pattern is part of the answer wherever the oracle's own pattern does not
            # depend on the precision or on a 1-ulp nudge
            stable = same_nonfinite(np.asarray(ry64, dtype=dtype), ry) and same_nonfinite(np.asarray(rd64, dtype=dtype), rd) \
                and same_nonfinite(ryp, ry) and same_nonfinite(rdp, rd)
            if stable:
                assert same_nonfinite(out[t], ry), (t, direction, "value pattern")
                assert same_nonfinite(dout[t], rd), (t, direction, "derivative pattern")
                n_pattern += 1
            fin = np.isfinite(ry) & np.isfinite(rd) & np.isfinite(out[t]) & np.isfinite(dout[t])
            if not fin.any():
                continue
            pairs.append(tree_verdict(dtype, [(out[t], ry, (ry64, ryp), fin), (dout[t], rd, (rd64, rdp), fin)]))
            n_values += 1
        check_trees(f"eval_diff direction {direction} {np.dtype(dtype).name}", dtype, pairs, min_strict=0.8)
    assert n_pattern > 2 * P_ and n_values > 2 * P_


# ---------------------------------------------------------------------------------------
# derivatives of ParametricExpressions: the reference differentiates
# eval_tree_array(convert(Node, ex), vcat(parameters[:, classes], X)) with variable = :both
# (src/ParametricExpression.jl:305-350, 380-389; src/ChainRules.jl:56-77)
# ---------------------------------------------------------------------------------------
def _converted(wire, n_params):
    """convert(Node, ex): parameter p -> feature p, feature f -> feature f + n_params (LeafConverter)."""
    w = wire.copy()
    leaf = w["degree"] == 0
    isf, isp = leaf & (w["kind"] == 1), leaf & (w["kind"] == 2)
    w["feature"][isf] += n_params
    w["kind"][isp] = 1
    return w


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mode", ["features", "constants", "both"])
def test_parametric_gradients_match_the_converted_tree(dtype, mode, oracle):
    spec = {1: ("cos", "exp", "sin"), 2: ("+", "-", "*", "/")}
    ops = dexb200.OperatorEnum(spec)
    P_, n_params, n_classes, F, N = 90, 4, 7, 3, 900          # the trees use parameters 0..2 of 4
    nodes, offsets = treegen.gen_population(P_, 7, 3, 4, F, seed=91, n_params=3, dtype=dtype)
    rng = np.random.default_rng(13)
    X = rng.standard_normal((F, N)).astype(dtype)
    params = rng.standard_normal((P_, n_params, n_classes)).astype(dtype)
    cls0 = rng.integers(0, n_classes, N)
    omode = {"features": oracle.GRAD_FEATURES, "constants": oracle.GRAD_CONSTANTS, "both": oracle.GRAD_BOTH}[mode]
    dmode = {"features": D.GRAD_FEATURES, "constants": D.GRAD_CONSTANTS, "both": D.GRAD_BOTH}[mode]
    pop = D.Population(None, ops, dtype, wire=(nodes, offsets), n_params=n_params)
    out, grad, off, ok = pop.eval_grad_parametric(X, params, cls0, dmode)
    out, grad, ok = out.cpu().numpy(), grad.cpu().numpy(), ok.cpu().numpy().astype(bool)
    verdicts, n_with = [], 0
    for t in range(P_):
        w = _converted(nodes[offsets[t]:offsets[t + 1]], n_params)
        Xv = np.vstack([params[t][:, cls0], X])                 # vcat(indexed_parameters, X)
        so = np.array([0, len(w)])
        ref, rg, rok = oracle.eval_grad_population(w, so, ops.opcodes, Xv, omode)
        _, _, rok_e = oracle.eval_grad_population(w, so, ops.opcodes, Xv, omode | oracle.GRAD_ELEMENTWISE)
        assert rok[0] <= rok_e[0] and ok[t] == rok_e[0], (t, ok[t], rok[0], rok_e[0])
        if not rok[0]:
            continue
        G = rg[0].shape[0]
        assert off[t + 1] - off[t] == G * N
        g = grad[off[t]:off[t + 1]].reshape(N, G).T
        yards = _grad_yardsticks(oracle, w, so, ops, Xv, omode)
        verdicts.append(_grad_verdict(dtype, out[t], g, ref[0], rg[0], [(y[0], yg[0]) for y, yg in yards]))
        n_with += 1
        if mode != "constants":
            assert not g[3].any()                               # the unused 4th parameter: a zero row
    check_trees(f"parametric grad/{mode}/{np.dtype(dtype).name}", dtype, verdicts, min_strict=0.8)
    assert n_with > 25
    # without the announced parameter count the library refuses (one gradient row per parameter)
    short = D.Population(None, ops, dtype, wire=(nodes, offsets))
    with pytest.raises(D.DexError, match="PARAM_ROWS"):
        short.eval_grad_parametric(X, params, cls0, dmode)
    with pytest.raises(D.DexError, match="parametric"):
        pop.eval_grad(X, dmode)


def test_parametric_expression_gradient_api(oracle):
    """ParametricExpression.eval_grad_tree_array / parameter_gradient through the host mirror, on the
    reference's own example tree sin(x) + y + p1 * p2 (test/test_parametric_expression.jl:103-128)."""
    ops = dexb200.OperatorEnum({1: ("sin",), 2: ("+", "*")})
    N_ = dexb200.Node
    tree = N_(1, N_(1, N_(1, N_(feature=1)), N_(feature=2)), N_(2, N_(parameter=1), N_(parameter=2)))
    P = np.array([[1.0, 1.0, 0.8], [2.0, 3.0, 5.0]])
    X = np.array([[0.0, np.pi / 2, np.pi, 1.2], [0.0, 0.0, 1.5, 0.1]])
    cls = np.array([1, 1, 2, 3])
    ex = dexb200.ParametricExpression(tree, operators=ops, parameters=P)
    y, g, ok = ex.eval_grad_tree_array(X, cls, variable=True)
    assert ok and g.shape == (4, 4)
    np.testing.assert_allclose(y, [2, 3, 4.5, 5.032039085967226], rtol=1e-12)
    p1, p2 = P[0, cls - 1], P[1, cls - 1]
    np.testing.assert_allclose(g, np.stack([p2, p1, np.cos(X[0]), np.ones(4)]), rtol=1e-12, atol=1e-15)
    dY = np.array([1.0, 2.0, 3.0, 4.0])
    dP = ex.parameter_gradient(X, cls, dY)
    want = np.zeros((2, 3))
    for j in range(4):
        want[0, cls[j] - 1] += dY[j] * p2[j]
        want[1, cls[j] - 1] += dY[j] * p1[j]
    np.testing.assert_allclose(dP, want, rtol=1e-12)
