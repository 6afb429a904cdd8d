"""CPU-only checks of the native library: the C ABI exports every declared symbol, the
host-only context packs/validates, and the flattened tapes (executed by a numpy tape
interpreter, tests/tape_sim.py) agree with the CPU oracle — values AND the `complete`
flag, for every EvalContext policy.  No compute call touches a GPU here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import dexb200
from dexb200 import device as D
from dexb200 import treegen
from tests.golden_util import (CONTEXTS, load_cases, make_matrix, make_operators, make_tree)
from tests.tape_sim import run_folded, run_tape

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "dexb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(dex_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    raw = C.CDLL(D.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(raw, name), f"{name} is declared in include/dexb200.h but not exported"
    assert declared <= set(D.ABI), declared - set(D.ABI)
    assert D.lib().dex_abi_version() == 1


def test_enum_values_agree_between_the_header_and_its_bindings():
    """include/dexb200.h is the contract: the ctypes binding and the Julia extension restate its enum
    values by hand."""
    hdr = open(os.path.join(ROOT, "include", "dexb200.h")).read() + open(os.path.join(ROOT, "include", "dex_wire.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    enums = {k: int(v) for k, v in re.findall(r"\b(DEX_[A-Z0-9_]+)\s*=\s*(-?\d+)", hdr)}
    py = {"DEX_EVAL_EARLY_EXIT": D.EVAL_EARLY_EXIT, "DEX_EVAL_SKIP_INCOMPLETE": D.EVAL_SKIP_INCOMPLETE,
          "DEX_PACK_FUSED": D.PACK_FUSED, "DEX_PACK_BUMPER": D.PACK_BUMPER, "DEX_GRAD_CONSTANTS": D.GRAD_CONSTANTS,
          "DEX_GRAD_FEATURES": D.GRAD_FEATURES, "DEX_GRAD_BOTH": D.GRAD_BOTH, "DEX_F32": D.F32, "DEX_F64": D.F64,
          "DEX_OK": D.OK}
    for k, v in py.items():
        assert enums[k] == v, k
    jl = open(os.path.join(ROOT, "ext", "DynamicExpressionsB200Ext.jl")).read()
    consts = {}
    for names, values in re.findall(r"^const (DEX_[A-Z0-9_, ]+?) = ((?:Cint\(-?\d+\)(?:, )?)+)", jl, flags=re.M):
        for k, v in zip(names.split(", "), re.findall(r"-?\d+", values)):
            consts[k] = int(v)
    assert len(consts) >= 9, consts
    for k, v in consts.items():
        assert enums.get(k) == v, f"{k} = {v} in the Julia extension, {enums.get(k)} in the header"


def test_safe_pow_is_the_power_operator(oracle):
    """`safe_pow(x, y) = x < 0 && y != round(y) ? NaN : x^y` of the reference's tests (test/test_parse.jl:66,
    test/test_symbolic_utils.jl:9-10) is what the table's `^` computes: C pow returns NaN exactly where
    Julia's `^` throws."""
    assert dexb200.opcode_of("safe_pow", 2) == dexb200.opcode_of("^", 2)
    ops = dexb200.OperatorEnum({2: ("safe_pow", "+")})
    N_ = dexb200.Node
    tree = N_(1, N_(feature=1, T=np.float64), N_(feature=2, T=np.float64))
    X = np.array([[-2.0, -2.0, 2.0, 0.0], [2.0, 0.5, 0.5, 3.0]])
    y, _ = oracle.eval_tree_array(dexb200.to_wire(tree), ops.opcodes, X, 0)      # early_exit off: every value
    assert np.isnan(y[1]) and y[0] == 4.0 and y[2] == 2.0 ** 0.5 and y[3] == 0.0
    _, ok = oracle.eval_tree_array(dexb200.to_wire(tree), ops.opcodes, X)
    assert not ok


def test_opcode_lookup_matches_def_file():
    l = D.lib()
    for (name, deg), code in dexb200.OPCODE_TABLE.items():
        got = l.dex_opcode_from_name(name.encode(), deg)
        if got != code:
            # python-side table also knows lower-cased symbol names; C side knows name + aliases
            assert name == dexb200.OPCODE_INFO[code][0].lower()
            continue
        assert l.dex_opcode_degree(code) == deg
    assert l.dex_opcode_from_name(b"no_such_operator", 1) == -1
    assert l.dex_opcode_from_name(b"cos", 2) == -1


def test_compute_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ops = dexb200.OperatorEnum({1: ("cos",), 2: ("+", "*")})
    tree = dexb200.Node(1, dexb200.Node(feature=1))
    with pytest.raises(D.DexError):
        dexb200.eval_tree_array(tree, np.ones((1, 4), np.float32), ops)
    # and through the raw ABI on a host-only context
    ctx = D.host_context()
    pop = D.Population([tree], ops, np.float32, ctx=ctx)
    out = np.zeros((1, 4), np.float32)
    ok = np.zeros(1, np.uint8)
    X = np.ones((4, 1), np.float32)
    rc = D.lib().dex_eval(ctx.h, pop.h, X.ctypes.data, 1, 4, 1, out.ctypes.data, 4, ok.ctypes.data, 1)
    assert rc == -3 and b"no CPU fallback" in D.lib().dex_last_error(ctx.h)


def test_launch_geometry_of_narrow_wide_and_short_inputs():
    """Host logic of dex_eval.cu eval_num_tiles through dex_eval_launch_info (no device): Float32 early-exit
    launches of at least one 2 048-sample tile keep at most six rows of a tile in shared memory (four CTAs of
    48 KB per SM; the other feature rows are read through L1), the fused loss takes that layout for wide inputs
    only, Float64 and inputs shorter than a tile keep every row in shared memory."""
    N = dexb200.Node
    ctx = D.host_context()
    ops = dexb200.OperatorEnum({1: ("cos",), 2: ("+", "*")})
    x = [N(feature=i + 1) for i in range(3)]
    tree = N(2, N(1, N(1, x[0], x[1])), N(1, x[2], N(val=2.0)))        # needs a couple of stack rows
    pop = D.Population([tree], ops, np.float32, ctx=ctx)
    for F, N_ in ((5, 1 << 16), (10, 1 << 17), (60, 4099), (100, 1 << 20)):
        li = pop.launch_info(F, N_)
        assert li["threads"] == 256 and li["n_tiles"] == -(-N_ // 2048)
        assert 0 < li["smem_rows"] <= 6 and li["smem_bytes"] == li["smem_rows"] * 2048 * 4
        assert 4 * (li["smem_bytes"] + 1024) <= 228 * 1024             # four CTAs resident
        assert pop.launch_info(F, N_, early_exit=False)["smem_rows"] == li["smem_rows"]
    # the fused loss: all rows in shared memory for narrow inputs, eight for wide ones
    narrow, wide = pop.launch_info(5, 1 << 16, loss=True), pop.launch_info(24, 1 << 16, loss=True)
    assert narrow["smem_rows"] == 0 and wide["smem_rows"] == 8 and wide["threads"] == 256
    assert 3 * (wide["smem_bytes"] + 1024) <= 228 * 1024
    # shorter than one tile: the block shrinks, every row in shared memory
    short = pop.launch_info(5, 100)
    assert short["smem_rows"] == 0 and short["threads"] < 256 and short["n_tiles"] == 1
    # Float64 keeps the all-shared-memory layout
    pop64 = D.Population([tree], ops, np.float64, ctx=ctx)
    assert pop64.launch_info(5, 1 << 16)["smem_rows"] == 0
    assert D.lib().dex_eval_launch_info(None, 1, 1, 1, 0, 0, None, None, None, None) != D.OK


def test_pack_validation_errors():
    ctx = D.host_context()
    ops = dexb200.OperatorEnum({1: ("cos",), 2: ("+",)})
    N = dexb200.Node
    bad_op = N(3, N(feature=1), N(val=1.0))          # only one binary operator
    with pytest.raises(D.DexError, match="op index"):
        D.Population([N(1, N(feature=1)), bad_op], ops, np.float32, ctx=ctx)
    # truncated wire array
    w = dexb200.to_wire(N(1, N(feature=1), N(val=2.0)))[:2]
    with pytest.raises(D.DexError, match="tree 0"):
        D.Population(None, ops, np.float32, ctx=ctx, wire=(w, np.array([0, 2])))
    with pytest.raises(ValueError, match="no device implementation"):
        dexb200.OperatorEnum({1: ("my_custom_function",)})
    # offsets must start at >= 0 and increase: a negative or non-monotonic table would make the
    # flattener read outside the node array
    w2 = dexb200.to_wire(N(1, N(feature=1)))
    nodes = np.concatenate([w2, w2])
    for bad in ([-5, 2, 4], [0, 4, 2], [0, 0, 2], [2, 0, 4]):
        with pytest.raises(D.DexError, match="offsets"):
            D.Population(None, ops, np.float32, ctx=ctx, wire=(nodes, np.array(bad, dtype=np.int64)))


def test_max_min_operand_exchange_is_recorded_for_the_gradient():
    """The flattener brings commutative operators into (ACC|ROW, ROW|CONST) order.  For max / min
    the reference's partials (x > y, !(x > y)) break ties by operand order, so an exchange must be
    visible to the gradient interpreters: bit 26 of w0 (csrc/dex_tape.h)."""
    ctx = D.host_context()
    ops = dexb200.OperatorEnum({1: ("cos",), 2: ("max", "min", "+")})
    N = dexb200.Node
    inner = lambda: N(1, N(feature=1))
    trees = [N(1, N(val=0.0), N(feature=1)),     # max(c, x1)    -> (ROW, CONST): exchanged
             N(1, N(feature=1), N(val=0.0)),     # max(x1, c)    -> as written
             N(2, N(feature=2), inner()),        # min(x2, cos)  -> (ACC, ROW): exchanged
             N(2, inner(), N(feature=2)),        # min(cos, x2)  -> as written
             N(3, N(val=1.0), N(feature=1))]     # c + x1: exchanged, but + has no tie rule -> no flag
    pop = D.Population(trees, ops, np.float32, ctx=ctx)
    ins, off = pop.tape()
    last = ins[off[1:] - 1, 0]
    assert [int(w >> 26) & 1 for w in last] == [1, 0, 1, 0, 0]


def test_population_info_and_constants_roundtrip():
    ctx = D.host_context()
    ops = dexb200.OperatorEnum({1: ("cos",), 2: ("+", "*", "-")})
    N = dexb200.Node
    t1 = N(2, N(1, N(val=1.5), N(3, N(feature=1), N(val=2.5))), N(1, N(1, N(val=3.5), N(val=4.5))))
    t2 = N(1, N(feature=3))
    pop = D.Population([t1, t2], ops, np.float64, ctx=ctx)
    assert pop.info["n_trees"] == 2 and pop.info["n_nodes"] == 10 + 2
    assert pop.info["max_feature"] == 2 and pop.info["n_constants"] == 4
    assert list(pop.constant_counts()) == [4, 0]
    assert list(pop.get_constants()) == [1.5, 2.5, 3.5, 4.5]      # leaf order, NodeUtils.jl:99-116
    pop.set_constants([10.0, 20.0, 30.0, 40.0])
    assert list(pop.get_constants()) == [10.0, 20.0, 30.0, 40.0]


def _flags(o, ctx):
    return ((o.EARLY_EXIT if ctx.get("early_exit", True) else 0) |
            (o.USE_FUSED if ctx.get("use_fused", True) else 0) | (o.BUMPER if ctx.get("bumper") else 0))


ALL_CTX = ["default", "bumper", "unfused", "no_early_exit", "bumper_no_early_exit"]


@pytest.mark.parametrize("case", load_cases(), ids=lambda c: c["id"])
def test_flattened_tape_matches_oracle_on_golden_cases(case, oracle):
    hctx = D.host_context()
    for dt in case["dtypes"]:
        dtype = np.dtype(dt).type
        ops = make_operators(case)
        tree = make_tree(case["tree"], ops, dtype)
        wire = dexb200.to_wire(tree)
        X = make_matrix(case["X"], dtype)
        P = cls0 = None
        if "parameters" in case:
            P = make_matrix(case["parameters"], dtype)
            cls0 = np.array(case["classes"], dtype=np.int64) - 1
        for cname in ALL_CTX:
            c = CONTEXTS[cname]
            pop = D.Population([tree], ops, dtype, ctx=hctx, bumper=c.get("bumper", False),
                               use_fused=c.get("use_fused", True))
            if P is not None:
                ry, rok = oracle.eval_parametric(wire, ops.opcodes, X, P, cls0, _flags(oracle, c))
            else:
                ry, rok = oracle.eval_tree_array(wire, ops.opcodes, X, _flags(oracle, c))
            # the full tape (gradients) and the folded image (evaluation) must both agree
            ins, off = pop.tape()
            full = run_tape(ins, X, pop.info["max_stack"], dexb200.OPCODE_INFO, dtype,
                            early_exit=c.get("early_exit", True), params=P, classes0=cls0,
                            n_param_rows=pop.info["max_parameter"] + 1)
            fold = run_folded(pop.folded(), 0, X, pop.info["folded_max_stack"], dexb200.OPCODE_INFO, dtype,
                              early_exit=c.get("early_exit", True), params=P, classes0=cls0,
                              n_param_rows=pop.info["max_parameter"] + 1)
            for which, (y, ok) in (("full", full), ("folded", fold)):
                assert ok == rok, (case["id"], dt, cname, which)
                if rok:
                    tol = 1e-4 if dtype == np.float32 else 1e-9
                    fin = np.isfinite(ry)
                    np.testing.assert_allclose(y[fin], ry[fin], rtol=tol, atol=tol)
                    assert (np.isfinite(y) == fin).all()


def _nonfinite_X(rng, F, N, dtype):
    X = rng.standard_normal((F, N)).astype(dtype)
    # sprinkle Inf / NaN into some columns so that the path-dependent validity rules matter
    for _ in range(3):
        X[rng.integers(F), rng.integers(N)] = rng.choice([np.inf, -np.inf, np.nan])
    return X


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("cname", ALL_CTX)
def test_flattened_tape_matches_oracle_on_random_trees(seed, cname, oracle):
    """Random trees incl. ternary operators, >15-operator enums, non-finite constants and
    non-finite X: the `complete` flag of the tape equals the oracle's for every policy."""
    rng = np.random.default_rng(1000 + seed)
    hctx = D.host_context()
    big = seed % 3 == 2
    spec = {1: ("cos", "exp", "abs", "sin") + (("square",) * 14 if big else ()),
            2: ("+", "-", "*", "/", "max") + (("+",) * 12 if big else ()),
            3: ("fma", "clamp")}
    ops = dexb200.OperatorEnum(spec)
    c = CONTEXTS[cname]
    F, N = 4, 24
    n_checked = n_folded = 0
    for k in range(60):
        w = _random_wire(rng, ops, F, max_nodes=int(rng.integers(1, 40)))
        dtype = np.float32 if k % 2 else np.float64
        X = _nonfinite_X(rng, F, N, dtype) if k % 3 == 0 else rng.standard_normal((F, N)).astype(dtype)
        pop = D.Population(None, ops, dtype, ctx=hctx, wire=(w, np.array([0, len(w)])),
                           bumper=c.get("bumper", False), use_fused=c.get("use_fused", True))
        ins, _ = pop.tape()
        ry, rok = oracle.eval_tree_array(w, ops.opcodes, X, _flags(oracle, c))
        full = run_tape(ins, X, pop.info["max_stack"], dexb200.OPCODE_INFO, dtype,
                        early_exit=c.get("early_exit", True))
        fold = run_folded(pop.folded(), 0, X, pop.info["folded_max_stack"], dexb200.OPCODE_INFO, dtype,
                          early_exit=c.get("early_exit", True))
        n_folded += pop.info["n_folded_subtrees"]
        for which, (y, ok) in (("full", full), ("folded", fold)):
            assert ok == rok, (seed, k, cname, which, dexb200.string_tree(dexb200.from_wire(w), ops))
            if rok:
                tol = 2e-4 if dtype == np.float32 else 1e-8
                fin = np.isfinite(ry) & (np.abs(ry) < 1e30)
                np.testing.assert_allclose(y[fin], ry[fin], rtol=tol, atol=tol)
        n_checked += bool(rok)
    assert n_checked > 5
    assert n_folded > 0 or c.get("bumper")       # the Bumper evaluator does not fold


def _random_wire(rng, ops, F, max_nodes):
    """Random tree with unary/binary/ternary nodes and occasional Inf/NaN constants."""
    N = dexb200.Node

    def leaf():
        r = rng.random()
        if r < 0.45:
            return N(feature=int(rng.integers(F)) + 1)
        if r < 0.48:
            return N(val=float(rng.choice([np.inf, -np.inf, np.nan])))
        return N(val=float(np.float32(rng.standard_normal())))

    def grow(budget):
        if budget <= 1 or rng.random() < 0.15:
            return leaf(), 1
        deg = int(rng.choice([1, 2, 2, 2, 3]))
        deg = min(deg, budget - 1)
        used = 1
        ch = []
        for k in range(deg):
            sub, n = grow(max(1, (budget - used) // (deg - k)))
            ch.append(sub)
            used += n
        return N(int(rng.integers(ops.nops(deg))) + 1, *ch), used

    t, _ = grow(max_nodes)
    return dexb200.to_wire(t)


def _shared_trees(dtype):
    """Trees in which one operator subtree (>= 4 nodes) occurs several times — GraphNode sharing of the
    reference (src/Node.jl:137-166), which its evaluators expand; here the SAME Python object is
    referenced from several parents and `to_wire` expands it likewise."""
    N = dexb200.Node
    x1, x2, x3 = (N(feature=k, T=dtype) for k in (1, 2, 3))
    c = lambda v: N(val=v, T=dtype)
    # operators: 1: cos exp ; 2: + - * /
    S = N(3, N(1, x1, c(0.5)), N(1, x2))                      # (x1 + 0.5) * cos(x2)         6 nodes
    D2 = N(4, x1, N(1, x3, c(2.0)))                           # x1 / (x3 + 2)               5 nodes
    E = N(2, N(1, x2, x3))                                    # exp(x2 + x3)                4 nodes
    return [
        N(1, S, N(1, S)),                                     # S + cos(S)
        N(4, N(1, S), N(1, S, c(3.0))),                       # cos(S) / (S + 3): reuse as divisor operand
        N(3, N(2, D2), D2),                                   # exp(D2) * D2: the shared value can be Inf / NaN
        N(2, N(1, E, N(3, E, E))),                            # exp(E + E * E): three occurrences
        N(1, N(1, S, D2), N(3, S, D2)),                       # two repeated subtrees: one of them is shared
        N(1, N(1, N(1, N(1, S, x1), x2), x3), S),             # deep left spine
    ]


@pytest.mark.parametrize("cname", ALL_CTX)
def test_shared_subexpressions_are_computed_once(cname, oracle):
    """The evaluation image KEEPs the value of a repeated subtree in a row and loads it at the later
    occurrences; values and `complete` flags stay those of the expanded tree for every policy."""
    hctx = D.host_context()
    ops = dexb200.OperatorEnum({1: ("cos", "exp"), 2: ("+", "-", "*", "/")})
    c = CONTEXTS[cname]
    rng = np.random.default_rng(3)
    n_keep = 0
    for dtype in (np.float32, np.float64):
        for k, tree in enumerate(_shared_trees(dtype)):
            w = dexb200.to_wire(tree)
            pop = D.Population([tree], ops, dtype, ctx=hctx, bumper=c.get("bumper", False), use_fused=c.get("use_fused", True))
            f = pop.folded()
            keeps = [ins for ins in f["tape"] if D.lib().dex_handler_name(int(ins[0]) & 63) == b"KEEP"]
            n_keep += len(keeps)
            if cname in ("default", "no_early_exit"):
                # tree 2's shared subtree is two instructions: keeping + loading it saves nothing
                assert len(keeps) == (0 if k == 2 else 1), (k, cname)
                if keeps:
                    assert pop.info["n_folded_instructions"] < pop.info["n_instructions"]
                assert pop.info["folded_max_stack"] <= 3
            else:
                assert not keeps                       # Bumper / unfused policies flatten the expanded tree
            for trial in range(3):
                X = rng.standard_normal((3, 40)).astype(dtype)
                if trial == 1:
                    X[rng.integers(3), rng.integers(40)] = np.inf
                    X[2, 7] = -2.0                     # x3 + 2 == 0: the shared divisor is 0
                if trial == 2:
                    X[rng.integers(3), rng.integers(40)] = np.nan
                ry, rok = oracle.eval_tree_array(w, ops.opcodes, X, _flags(oracle, c))
                y, ok = run_folded(f, 0, X, pop.info["folded_max_stack"], dexb200.OPCODE_INFO, dtype,
                                   early_exit=c.get("early_exit", True))
                assert ok == rok, (k, cname, trial)
                if rok:
                    fin = np.isfinite(ry)
                    np.testing.assert_allclose(y[fin], ry[fin], rtol=2e-5 if dtype == np.float32 else 1e-12)
                    assert (np.isfinite(y) == fin).all()
    assert (n_keep > 0) == (cname in ("default", "no_early_exit"))


def test_parallel_flattening_equals_the_one_thread_walk(monkeypatch):
    """Large populations are flattened by several threads over tree ranges and the partial tapes
    merged (csrc/dex_flatten.cpp flatten_image): every array of both images must be identical to
    the one-thread result, including constant positions (checked through a constants round trip)
    and the error message of a malformed tree."""
    spec = {1: ("cos", "exp", "abs"), 2: ("+", "-", "*", "/"), 3: ("fma", "clamp")}
    ops = dexb200.OperatorEnum(spec)
    rng = np.random.default_rng(5)
    wires = [_random_wire(rng, ops, 6, max_nodes=int(rng.integers(1, 50))) for _ in range(1500)]
    nodes = np.concatenate(wires)
    offsets = np.concatenate([[0], np.cumsum([len(w) for w in wires])]).astype(np.int64)
    hctx = D.host_context()
    images = {}
    for thr in ("1", "3", "8"):
        monkeypatch.setenv("DEXB200_PACK_THREADS", thr)
        pop = D.Population(None, ops, np.float64, ctx=hctx, wire=(nodes, offsets))
        ins, off = pop.tape()
        c = pop.get_constants()
        pop.set_constants(c * 2.0 + 1.0)                      # exercises const_pos of both images
        ins2, _ = pop.tape()
        images[thr] = (ins, off, pop.folded(), c, dict(pop.info), ins2, pop.constant_counts())
    a = images["1"]
    for thr, b in images.items():
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[3].tobytes() == b[3].tobytes(), thr
        assert a[4] == b[4] and np.array_equal(a[5], b[5]) and np.array_equal(a[6], b[6]), thr
        for k in a[2]:
            assert np.array_equal(a[2][k], b[2][k]), (thr, k)
    # a malformed tree in the middle: same error (with the global tree index) for any thread count
    bad = nodes.copy()
    t_bad = 777
    bad["op"][offsets[t_bad]] = 200 if bad["degree"][offsets[t_bad]] > 0 else bad["op"][offsets[t_bad]]
    bad["degree"][offsets[t_bad]] = max(bad["degree"][offsets[t_bad]], 1)
    msgs = []
    for thr in ("1", "8"):
        monkeypatch.setenv("DEXB200_PACK_THREADS", thr)
        with pytest.raises(D.DexError) as ei:
            D.Population(None, ops, np.float64, ctx=hctx, wire=(bad, offsets))
        msgs.append(str(ei.value))
    assert msgs[0] == msgs[1] and str(t_bad) in msgs[0], msgs


def test_stack_need_uses_sethi_ullman_order():
    """A right-deep tree needs no more stack rows than its left-deep mirror."""
    hctx = D.host_context()
    ops = dexb200.OperatorEnum({1: ("cos",), 2: ("+", "*")})
    N = dexb200.Node

    def deep(side, d):
        if d == 0:
            return N(1, N(feature=1))
        sub = deep(side, d - 1)
        other = N(1, N(feature=2))
        return N(2, sub, other) if side == "l" else N(2, other, sub)

    a = D.Population([deep("l", 12)], ops, np.float32, ctx=hctx).info["max_stack"]
    b = D.Population([deep("r", 12)], ops, np.float32, ctx=hctx).info["max_stack"]
    assert a == b == 1


def test_treegen_depth_and_determinism():
    nodes, off = treegen.gen_population(50, 8, 2, 4, 5, seed=7)
    nodes2, off2 = treegen.gen_population(50, 8, 2, 4, 5, seed=7)
    assert (off == off2).all() and (nodes == nodes2).all()
    for i in range(50):
        t = dexb200.from_wire(nodes[off[i]:off[i + 1]])
        assert dexb200.count_depth(t) == 8
        assert dexb200.count_nodes(t) == off[i + 1] - off[i] <= 255
