"""ctypes wrapper around the CPU oracle (oracle/dex_oracle.c).

TEST INFRASTRUCTURE ONLY — imported by tests/, by __graft_entry__.smoke() and by
bench.py's cpu_baseline / ``--impl reference`` legs, never by the product package.

Also holds :func:`numpy_eval`, a deliberately dumb second opinion (plain
recursion over the wire array with numpy ufuncs, no fusion, no early exit) used
to cross-check the C restatement itself.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libdexoracle.so")

EARLY_EXIT, USE_FUSED, BUMPER, ELEMENTWISE = 1, 2, 4, 8
GRAD_ELEMENTWISE = 8
DEFAULT_FLAGS = EARLY_EXIT | USE_FUSED
GRAD_CONSTANTS, GRAD_FEATURES, GRAD_BOTH = 0, 1, 2
F32, F64 = 0, 1
MAX_DEGREE = 3


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in
            ("dex_oracle.c", "dex_oracle_impl.inc", "dex_oracle_ops.inc", "dex_oracle.h")]
    srcs += [os.path.join(_HERE, "..", "include", f) for f in ("dex_wire.h", "dex_ops.def")]
    if not force and os.path.exists(_SO):
        if all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs if os.path.exists(s)):
            return _SO
    subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


class _OpTable(C.Structure):
    _fields_ = [("nops", C.c_int32 * MAX_DEGREE), ("ops", C.POINTER(C.c_int32) * MAX_DEGREE)]


_lib = None
_flags = "gcc -O3 -march=x86-64-v3 (AVX2) -ffp-contract=off -fopenmp"


def use_native_build():
    """Switch to a `-march=native` build made on THIS machine (bench.py's CPU-baseline legs: the
    timed stand-in should use the host's full vector width).  Falls back to the portable
    x86-64-v3 build when gcc is missing or the build fails.  Returns the flags in use."""
    global _lib, _flags
    out = os.path.join(_HERE, "_build", "libdexoracle_native.so")
    try:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "MARCH=native", "OUT=_build/libdexoracle_native.so"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        native = C.CDLL(out)
        _lib = None
        _bind(native)
        _lib = native
        _flags = "gcc -O3 -march=native -ffp-contract=off -fopenmp, built on this host"
    except Exception:
        pass
    return _flags


def _bind(l):
    l.dexo_apply_f64.restype = C.c_double
    l.dexo_apply_f64.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double]
    l.dexo_apply_f32.restype = C.c_float
    l.dexo_apply_f32.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float]
    l.dexo_partials_f64.restype = None
    l.dexo_partials_f64.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double)]
    l.dexo_count_constants.restype = C.c_int32
    l.dexo_count_constants.argtypes = [C.c_void_p, C.c_int64]


def lib():
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(_SO)
        _bind(l)
        _lib = l
    return _lib


def _optable(opcodes_per_degree):
    """opcodes_per_degree: sequence (len 3) of int32 arrays of builtin opcodes."""
    t = _OpTable()
    keep = []
    for d in range(MAX_DEGREE):
        a = np.ascontiguousarray(opcodes_per_degree[d] if d < len(opcodes_per_degree) else [],
                                 dtype=np.int32)
        keep.append(a)
        t.nops[d] = len(a)
        t.ops[d] = a.ctypes.data_as(C.POINTER(C.c_int32))
    return t, keep


def _dt(X):
    if X.dtype == np.float32:
        return F32
    if X.dtype == np.float64:
        return F64
    raise TypeError("oracle handles float32/float64")


def _prep_X(X):
    X = np.asarray(X)
    if X.ndim == 1:
        X = X.reshape(-1, 1)
    # Julia column-major F x N  ==  C-contiguous N x F
    Xc = np.ascontiguousarray(X.T)
    return X.shape[0], X.shape[1], Xc


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def eval_tree_array(wire, opcodes, X, flags=DEFAULT_FLAGS):
    """(out[N], ok) for one wire tree; X is (F, N) like the reference's cX."""
    F, N, Xc = _prep_X(X)
    t, keep = _optable(opcodes)
    out = np.full(N, np.nan, dtype=Xc.dtype)
    ok = C.c_uint8(0)
    wire = np.ascontiguousarray(wire)
    rc = lib().dexo_eval_tree_array(_p(wire), C.c_int64(len(wire)), C.byref(t), _dt(Xc), _p(Xc),
                                    C.c_int32(F), C.c_int64(N), C.c_int64(F), C.c_int(flags),
                                    _p(out), C.byref(ok))
    if rc:
        raise ValueError(f"oracle rejected the tree (rc={rc})")
    return out, bool(ok.value)


def eval_diff_tree_array(wire, opcodes, X, direction0):
    F, N, Xc = _prep_X(X)
    t, keep = _optable(opcodes)
    out = np.full(N, np.nan, dtype=Xc.dtype)
    dout = np.full(N, np.nan, dtype=Xc.dtype)
    ok = C.c_uint8(0)
    wire = np.ascontiguousarray(wire)
    rc = lib().dexo_eval_diff_tree_array(_p(wire), C.c_int64(len(wire)), C.byref(t), _dt(Xc),
                                         _p(Xc), C.c_int32(F), C.c_int64(N), C.c_int64(F),
                                         C.c_int32(direction0), _p(out), _p(dout), C.byref(ok))
    if rc:
        raise ValueError(f"oracle rejected the tree (rc={rc})")
    return out, dout, bool(ok.value)


def count_constants(wire):
    wire = np.ascontiguousarray(wire)
    return int(lib().dexo_count_constants(_p(wire), C.c_int64(len(wire))))


def eval_grad_tree_array(wire, opcodes, X, mode):
    """(out[N], grad[G, N], ok); grad[k, j] = d out[j] / d (k-th variable)."""
    F, N, Xc = _prep_X(X)
    t, keep = _optable(opcodes)
    nc = count_constants(wire)
    m = mode & 3
    G = F if m == GRAD_FEATURES else nc if m == GRAD_CONSTANTS else F + nc
    out = np.full(N, np.nan, dtype=Xc.dtype)
    grad = np.full((N, G), np.nan, dtype=Xc.dtype)  # column-major (G x N)
    ok = C.c_uint8(0)
    ng = C.c_int32(0)
    wire = np.ascontiguousarray(wire)
    rc = lib().dexo_eval_grad_tree_array(_p(wire), C.c_int64(len(wire)), C.byref(t), _dt(Xc),
                                         _p(Xc), C.c_int32(F), C.c_int64(N), C.c_int64(F),
                                         C.c_int(mode), _p(out), _p(grad), C.c_int64(N * G),
                                         C.byref(ng), C.byref(ok))
    if rc:
        raise ValueError(f"oracle rejected the tree (rc={rc})")
    assert ng.value == G
    return out, grad.T, bool(ok.value)


def eval_parametric(wire, opcodes, X, parameters, classes0, flags=DEFAULT_FLAGS):
    """parameters: (n_params, n_classes); classes0: 0-based int array of length N."""
    F, N, Xc = _prep_X(X)
    t, keep = _optable(opcodes)
    P = np.ascontiguousarray(np.asarray(parameters, dtype=Xc.dtype).T)  # column-major
    n_params, n_classes = np.asarray(parameters).shape
    cl = np.ascontiguousarray(classes0, dtype=np.int32)
    out = np.full(N, np.nan, dtype=Xc.dtype)
    ok = C.c_uint8(0)
    wire = np.ascontiguousarray(wire)
    rc = lib().dexo_eval_parametric(_p(wire), C.c_int64(len(wire)), C.byref(t), _dt(Xc), _p(Xc),
                                    C.c_int32(F), C.c_int64(N), C.c_int64(F), _p(P),
                                    C.c_int32(n_params), C.c_int32(n_classes), _p(cl),
                                    C.c_int(flags), _p(out), C.byref(ok))
    if rc:
        raise ValueError(f"oracle rejected the tree (rc={rc})")
    return out, bool(ok.value)


def set_ulp_nudge(n):
    """TEST-ONLY conditioning yardstick: move the result of every transcendental unary operator
    by ``n`` ulps in later evaluations (0 switches it off)."""
    lib().dexo_set_ulp_nudge(C.c_int(int(n)))


def eval_population(nodes, offsets, opcodes, X, flags=DEFAULT_FLAGS, nthreads=0, out=None):
    """(out[P, N], ok[P]); OpenMP over trees when nthreads != 1."""
    F, N, Xc = _prep_X(X)
    t, keep = _optable(opcodes)
    P = len(offsets) - 1
    if out is None:
        out = np.empty((P, N), dtype=Xc.dtype)
    ok = np.zeros(P, dtype=np.uint8)
    nodes = np.ascontiguousarray(nodes)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    rc = lib().dexo_eval_population(_p(nodes), _p(offsets), C.c_int64(P), C.byref(t), _dt(Xc),
                                    _p(Xc), C.c_int32(F), C.c_int64(N), C.c_int64(F),
                                    C.c_int(flags), C.c_int(nthreads), _p(out), _p(ok))
    if rc:
        raise ValueError(f"oracle rejected a tree (rc={rc})")
    return out, ok.astype(bool)


def eval_population_f80(nodes, offsets, opcodes, X, flags=DEFAULT_FLAGS, nthreads=0):
    """Float64 in/out with 80-bit intermediates: the rounding-sensitivity yardstick."""
    F, N, Xc = _prep_X(np.asarray(X, dtype=np.float64))
    t, keep = _optable(opcodes)
    P = len(offsets) - 1
    out = np.empty((P, N), dtype=np.float64)
    ok = np.zeros(P, dtype=np.uint8)
    nodes = np.ascontiguousarray(nodes)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    rc = lib().dexo_eval_population_f80(_p(nodes), _p(offsets), C.c_int64(P), C.byref(t), _p(Xc),
                                        C.c_int32(F), C.c_int64(N), C.c_int64(F), C.c_int(flags),
                                        C.c_int(nthreads), _p(out), _p(ok))
    if rc:
        raise ValueError(f"oracle rejected a tree (rc={rc})")
    return out, ok.astype(bool)


def eval_grad_population(nodes, offsets, opcodes, X, mode, nthreads=0):
    """(out[P, N], [grad_t (G_t, N)], ok[P])."""
    F, N, Xc = _prep_X(X)
    t, keep = _optable(opcodes)
    P = len(offsets) - 1
    nodes = np.ascontiguousarray(nodes)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    G = np.zeros(P, dtype=np.int64)
    for i in range(P):
        nc = count_constants(nodes[offsets[i]:offsets[i + 1]])
        m = mode & 3
        G[i] = F if m == GRAD_FEATURES else nc if m == GRAD_CONSTANTS else F + nc
    goff = np.zeros(P + 1, dtype=np.int64)
    np.cumsum(G * N, out=goff[1:])
    out = np.empty((P, N), dtype=Xc.dtype)
    grad = np.empty(int(goff[-1]), dtype=Xc.dtype)
    ok = np.zeros(P, dtype=np.uint8)
    rc = lib().dexo_eval_grad_population(_p(nodes), _p(offsets), C.c_int64(P), C.byref(t),
                                         _dt(Xc), _p(Xc), C.c_int32(F), C.c_int64(N),
                                         C.c_int64(F), C.c_int(mode), C.c_int(nthreads), _p(out),
                                         _p(grad), _p(goff), _p(ok))
    if rc:
        raise ValueError(f"oracle rejected a tree (rc={rc})")
    grads = [grad[goff[i]:goff[i + 1]].reshape(N, int(G[i])).T for i in range(P)]
    return out, grads, ok.astype(bool)


def eval_parametric_population(nodes, offsets, opcodes, X, parameters, classes0,
                               flags=DEFAULT_FLAGS, nthreads=0):
    """parameters: (P, n_params, n_classes)."""
    F, N, Xc = _prep_X(X)
    t, keep = _optable(opcodes)
    P = len(offsets) - 1
    params = np.asarray(parameters, dtype=Xc.dtype)
    _, n_params, n_classes = params.shape
    pc = np.ascontiguousarray(params.transpose(0, 2, 1))  # per tree column-major
    cl = np.ascontiguousarray(classes0, dtype=np.int32)
    out = np.empty((P, N), dtype=Xc.dtype)
    ok = np.zeros(P, dtype=np.uint8)
    nodes = np.ascontiguousarray(nodes)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    rc = lib().dexo_eval_parametric_population(
        _p(nodes), _p(offsets), C.c_int64(P), C.byref(t), _dt(Xc), _p(Xc), C.c_int32(F),
        C.c_int64(N), C.c_int64(F), _p(pc), C.c_int32(n_params), C.c_int32(n_classes), _p(cl),
        C.c_int(flags), C.c_int(nthreads), _p(out), _p(ok))
    if rc:
        raise ValueError(f"oracle rejected a tree (rc={rc})")
    return out, ok.astype(bool)


def max_threads():
    return int(lib().dexo_max_threads())


def apply(opcode, a, b=0.0, c=0.0, dtype=np.float64):
    if dtype == np.float32:
        return np.float32(lib().dexo_apply_f32(opcode, a, b, c))
    return float(lib().dexo_apply_f64(opcode, a, b, c))


def partials(opcode, a, b=0.0, c=0.0):
    g = (C.c_double * 3)()
    lib().dexo_partials_f64(opcode, a, b, c, g)
    return [g[0], g[1], g[2]]


# ---------------------------------------------------------------------------------
# Second opinion: plain numpy recursion over the wire array (no fusion, no early
# exit, elementwise like the closed-form lambdas of test/test_evaluation.jl:71-86).
# ---------------------------------------------------------------------------------
def _np_ops():
    import scipy.special as sp

    def jmax(x, y):
        return np.where(np.isnan(x) | np.isnan(y), np.nan, np.maximum(x, y))

    def jmin(x, y):
        return np.where(np.isnan(x) | np.isnan(y), np.nan, np.minimum(x, y))

    def jmod(x, y):
        return x - np.floor(x / y) * y

    u = {
        "NEG": lambda x: -x, "ABS": np.abs, "ABS2": lambda x: x * x, "SQUARE": lambda x: x * x,
        "CUBE": lambda x: x * x * x, "INV": lambda x: 1 / x, "SQRT": np.sqrt, "CBRT": np.cbrt,
        "EXP": np.exp, "EXP2": np.exp2, "EXP10": lambda x: np.power(x.dtype.type(10), x),
        "EXPM1": np.expm1, "LOG": np.log, "LOG2": np.log2, "LOG10": np.log10, "LOG1P": np.log1p,
        "SIN": np.sin, "COS": np.cos, "TAN": np.tan, "ASIN": np.arcsin, "ACOS": np.arccos,
        "ATAN": np.arctan, "SINH": np.sinh, "COSH": np.cosh, "TANH": np.tanh,
        "ASINH": np.arcsinh, "ACOSH": np.arccosh, "ATANH": np.arctanh, "ROUND": np.rint,
        "FLOOR": np.floor, "CEIL": np.ceil, "TRUNC": np.trunc, "SIGN": np.sign,
        "RELU": lambda x: np.where(x < 0, 0, x), "IDENTITY": lambda x: x,
        "SAFE_LOG": lambda x: np.where(x <= 0, np.nan, np.log(np.where(x <= 0, 1, x))),
        "SAFE_LOG2": lambda x: np.where(x <= 0, np.nan, np.log2(np.where(x <= 0, 1, x))),
        "SAFE_LOG10": lambda x: np.where(x <= 0, np.nan, np.log10(np.where(x <= 0, 1, x))),
        "SAFE_LOG1P": lambda x: np.where(x <= -1, np.nan, np.log1p(np.where(x <= -1, 0, x))),
        "SAFE_SQRT": lambda x: np.where(x < 0, np.nan, np.sqrt(np.where(x < 0, 0, x))),
        "SAFE_ACOSH": lambda x: np.where(x < 1, np.nan, np.arccosh(np.where(x < 1, 1, x))),
        "COS2": lambda x: np.cos(x) ** 2, "ERF": sp.erf, "ERFC": sp.erfc,
    }
    b = {
        "ADD": np.add, "SUB": np.subtract, "MUL": np.multiply, "DIV": np.divide, "POW": np.power,
        "MAX": jmax, "MIN": jmin, "MOD": jmod, "ATAN2": np.arctan2, "COPYSIGN": np.copysign,
        "GREATER": lambda x, y: (x > y).astype(x.dtype), "LESS": lambda x, y: (x < y).astype(x.dtype),
        "POW_ABS": lambda x, y: np.exp(y * np.log(np.abs(x))),
        "COND": lambda x, y: np.where(x > 0, y, 0),
        "LOGICAL_OR": lambda x, y: ((x > 0) | (y > 0)).astype(x.dtype),
        "LOGICAL_AND": lambda x, y: ((x > 0) & (y > 0)).astype(x.dtype),
        "GREATER_EQ": lambda x, y: (x >= y).astype(x.dtype),
        "LESS_EQ": lambda x, y: (x <= y).astype(x.dtype),
    }
    t = {
        "FMA": lambda a, b_, c: a * b_ + c, "MULADD": lambda a, b_, c: a * b_ + c,
        "CLAMP": lambda a, lo, hi: np.where(a > hi, hi, np.where(a < lo, lo, a)),
        "MAX3": lambda a, b_, c: jmax(jmax(a, b_), c), "MIN3": lambda a, b_, c: jmin(jmin(a, b_), c),
        "ADD3": lambda a, b_, c: a + b_ + c, "MUL3": lambda a, b_, c: a * b_ * c,
    }
    return u, b, t


def numpy_eval(wire, opcodes, X, opcode_info, parameters=None, classes0=None):
    """Elementwise value of the tree on X (F, N); returns array[N] (may hold NaN/Inf)."""
    X = np.asarray(X)
    if X.ndim == 1:
        X = X.reshape(-1, 1)
    N = X.shape[1]
    u, b, t = _np_ops()
    tables = (u, b, t)
    pos = [0]

    def rec():
        r = wire[pos[0]]
        pos[0] += 1
        d = int(r["degree"])
        if d == 0:
            k = int(r["kind"])
            if k == 0:
                return np.full(N, X.dtype.type(r["val"]), dtype=X.dtype)
            if k == 2:
                return np.asarray(parameters, dtype=X.dtype)[int(r["feature"]), classes0]
            return X[int(r["feature"]), :].copy()
        code = int(opcodes[d - 1][int(r["op"])])
        sym = opcode_info[code][0]
        args = [rec() for _ in range(d)]
        with np.errstate(all="ignore"):
            return np.asarray(tables[d - 1][sym](*args), dtype=X.dtype)

    return rec()
