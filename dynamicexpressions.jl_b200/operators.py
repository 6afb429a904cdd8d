"""Host-side mirror of the reference's operator API.

``OperatorEnum`` (/root/reference/src/OperatorEnum.jl:14-49) holds, per degree, a
tuple of functions; nodes store only ``op`` = 1-based index into
``operators[degree]``.  On the device an operator is a *builtin opcode* from
``include/dex_ops.def``; this module maps each function of an ``OperatorEnum`` to
its opcode once, at construction.  A function that has no builtin opcode cannot
run on the device and construction fails (no CPU fallback).

Operators may be given as names (``"cos"``, ``"+"``, ``"safe_log"`` ...), as
Python/numpy callables whose ``__name__`` is a known name (``math.cos``,
``np.cos``, ``operator.add`` ...), or as ``(name, degree)`` for names that exist at
several arities (``"+"``, ``"-"``, ``"max"``).
"""
from __future__ import annotations

import os
import re

import numpy as np

from .node import MAX_DEGREE, Node

_DEF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "dex_ops.def")
_ROW = re.compile(r'DEX_OP\(\s*(\w+)\s*,\s*(\d+)\s*,\s*(\d+)\s*,\s*"([^"]*)"\s*,\s*"([^"]*)"\s*\)')


def _parse_def(path=_DEF):
    table = {}   # (name, degree) -> opcode
    info = {}    # opcode -> (SYMBOL, degree, julia name)
    with open(path) as f:
        for line in f:
            m = _ROW.search(line)
            if not m:
                continue
            sym, code, deg, name, aliases = m.group(1), int(m.group(2)), int(m.group(3)), m.group(4), m.group(5)
            info[code] = (sym, deg, name)
            for nm in [name, sym.lower()] + [a for a in aliases.split("|") if a]:
                table.setdefault((nm, deg), code)
    return table, info


OPCODE_TABLE, OPCODE_INFO = _parse_def()

# python callables -> julia-style names
_PY_NAMES = {
    "add": "+", "sub": "-", "mul": "*", "truediv": "/", "div": "/", "pow": "^", "power": "^",
    "neg": "-", "negative": "-", "multiply": "*", "subtract": "-", "divide": "/",
    "true_divide": "/", "maximum": "max", "minimum": "min", "absolute": "abs", "fabs": "abs",
    "arcsin": "asin", "arccos": "acos", "arctan": "atan", "arcsinh": "asinh",
    "arccosh": "acosh", "arctanh": "atanh", "arctan2": "atan", "rint": "round",
}


def opcode_of(fn, degree: int) -> int:
    """Builtin opcode for operator ``fn`` at arity ``degree`` (raises if unknown)."""
    name = fn if isinstance(fn, str) else getattr(fn, "__name__", None)
    if name is None:
        raise TypeError(f"cannot derive an operator name from {fn!r}")
    name = _PY_NAMES.get(name, name)
    code = OPCODE_TABLE.get((name, degree))
    if code is None:
        raise ValueError(
            f"operator {name!r} of degree {degree} has no device implementation in "
            f"include/dex_ops.def; the B200 path has no CPU fallback"
        )
    return code


class OperatorEnum:
    """``OperatorEnum(1 => (cos, exp), 2 => (+, -, *, /))``
    (/root/reference/src/OperatorEnumConstruction.jl:476-516); also accepts the
    deprecated keyword form ``binary_operators=..., unary_operators=...``
    (/root/reference/src/deprecated.jl:95-106)."""

    def __init__(self, ops=None, *, unary_operators=None, binary_operators=None,
                 ternary_operators=None):
        per_degree = {1: (), 2: (), 3: ()}
        if ops is not None:
            if isinstance(ops, dict):
                for d, fs in ops.items():
                    per_degree[int(d)] = tuple(fs)
            else:  # tuple of tuples, degree = position
                for d, fs in enumerate(ops, start=1):
                    per_degree[d] = tuple(fs)
        if unary_operators is not None:
            per_degree[1] = tuple(unary_operators)
        if binary_operators is not None:
            per_degree[2] = tuple(binary_operators)
        if ternary_operators is not None:
            per_degree[3] = tuple(ternary_operators)
        if any(d < 1 or d > MAX_DEGREE for d in per_degree):
            raise ValueError(f"operator degree must be 1..{MAX_DEGREE}")
        for d, fs in per_degree.items():
            if len(fs) > 255:
                raise ValueError("op index must fit UInt8")
        self.ops = tuple(per_degree[d] for d in range(1, MAX_DEGREE + 1))
        self.opcodes = tuple(
            np.array([opcode_of(f, d) for f in per_degree[d]], dtype=np.int32)
            for d in range(1, MAX_DEGREE + 1)
        )
        self.names = tuple(
            tuple(OPCODE_INFO[int(c)][2] for c in codes) for codes in self.opcodes
        )

    unaops = property(lambda self: self.ops[0])
    binops = property(lambda self: self.ops[1])

    def __getitem__(self, degree):
        return self.ops[degree - 1]

    def nops(self, degree):
        return len(self.ops[degree - 1])

    def flat_opcodes(self):
        """(opcodes, degree_offsets) as handed to ``dex_optable_create``."""
        offs = np.zeros(MAX_DEGREE + 1, dtype=np.int32)
        for d in range(MAX_DEGREE):
            offs[d + 1] = offs[d] + len(self.opcodes[d])
        flat = np.concatenate(self.opcodes).astype(np.int32) if offs[-1] else np.zeros(0, np.int32)
        return flat, offs

    def key(self):
        return tuple(tuple(int(c) for c in codes) for codes in self.opcodes)

    def index_of(self, name, degree):
        """1-based op index of the operator called ``name`` at ``degree``."""
        code = opcode_of(name, degree)
        hits = np.nonzero(self.opcodes[degree - 1] == code)[0]
        if len(hits) == 0:
            raise KeyError(f"operator {name!r} (degree {degree}) is not in this OperatorEnum")
        return int(hits[0]) + 1

    # -- tree-building helpers (the role of @extend_operators) -------------------
    def build(self, name, *children):
        ch = tuple(c if isinstance(c, Node) else Node(val=float(c)) for c in children)
        return Node(self.index_of(name, len(ch)), *ch)

    def __repr__(self):
        return "OperatorEnum(" + ", ".join(f"{d + 1}=>{self.names[d]}" for d in range(MAX_DEGREE) if self.names[d]) + ")"


_LATEST = [None]


def extend_operators(operators: OperatorEnum):
    """Install ``operators`` as the enum used by ``Node`` operator overloading and
    by :func:`call` — the role of ``@extend_operators``
    (/root/reference/src/OperatorEnumConstruction.jl:343-420)."""
    _LATEST[0] = operators
    return operators


def _latest():
    if _LATEST[0] is None:
        raise RuntimeError("no OperatorEnum installed; call extend_operators(operators) first")
    return _LATEST[0]


def call(name, *children):
    """``call("cos", x1)`` builds ``cos(x1)`` with the installed operators."""
    return _latest().build(name, *children)


def _bin(name, swap=False):
    def f(a, b):
        return _latest().build(name, b, a) if swap else _latest().build(name, a, b)
    return f


Node.__add__ = _bin("+")
Node.__radd__ = _bin("+", True)
Node.__sub__ = _bin("-")
Node.__rsub__ = _bin("-", True)
Node.__mul__ = _bin("*")
Node.__rmul__ = _bin("*", True)
Node.__truediv__ = _bin("/")
Node.__rtruediv__ = _bin("/", True)
Node.__pow__ = _bin("^")
Node.__rpow__ = _bin("^", True)
Node.__neg__ = lambda a: _latest().build("-", a)
