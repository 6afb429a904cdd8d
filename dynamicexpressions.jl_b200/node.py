"""Host-side mirror of the reference's tree node types.

Mirrors ``Node{T,D}`` (/root/reference/src/Node.jl:74-90) and ``ParametricNode``
(/root/reference/src/ParametricExpression.jl:52-74): fields ``degree``,
``constant``, ``val``, ``feature`` (1-based, as in Julia), ``op`` (1-based index
into ``operators[degree]``), ``children``; parametric nodes add ``is_parameter``
and ``parameter`` (1-based).

The only thing the device path needs from a tree is its *wire form*
(include/dex_wire.h): a preorder array of 16-byte records, produced by
:func:`to_wire` with the traversal order of the reference's ``tree_mapreduce``
(parent first, children left to right; /root/reference/src/base.jl:123-158).
The flattening into device tapes happens inside libdexb200 (csrc/dex_flatten.cpp).
"""
from __future__ import annotations

import numpy as np

# numpy image of `struct dex_node` (include/dex_wire.h) — 16 bytes
WIRE_DTYPE = np.dtype(
    [
        ("degree", np.uint8),
        ("kind", np.uint8),
        ("op", np.uint8),
        ("reserved0", np.uint8),
        ("feature", np.uint16),
        ("reserved1", np.uint16),
        ("val", np.float64),
    ],
    align=True,
)
assert WIRE_DTYPE.itemsize == 16

LEAF_CONST, LEAF_FEATURE, LEAF_PARAMETER = 0, 1, 2
MAX_DEGREE = 3

_DTYPES = {"float32": np.float32, "float64": np.float64, np.float32: np.float32, np.float64: np.float64}


def _as_dtype(T):
    if T is None:
        return None
    if T is float:
        return np.float64
    T = np.dtype(T).type
    if T not in (np.float32, np.float64):
        raise TypeError(f"only Float32/Float64 trees run on the device path, got {T}")
    return T


class Node:
    """Expression-tree node; see module docstring.

    Constructors (same forms as the reference, /root/reference/src/Node.jl:118-131)::

        Node(val=3.0)             constant leaf            Node(T; val=3.0)
        Node(feature=2)           feature leaf x2          Node(T; feature=2)
        Node(1, child)            unary operator #1        Node(1, child)
        Node(2, left, right)      binary operator #2       Node(2, l, r)
        Node(op=1, children=(a, b, c))                     Node(; op=1, children=(a,b,c))
        Node("x2")                feature leaf by name     Node("x2")
    """

    __slots__ = ("degree", "constant", "val", "feature", "op", "children", "dtype",
                 "is_parameter", "parameter")

    def __init__(self, *args, val=None, feature=None, op=None, children=None, T=None,
                 parameter=None):
        self.dtype = _as_dtype(T)
        self.is_parameter = False
        self.parameter = 0
        self.constant = False
        self.val = 0.0
        self.feature = 0
        self.op = 0
        self.children = ()
        if args:
            a0 = args[0]
            if isinstance(a0, str):  # Node("x3")
                if not a0.startswith("x"):
                    raise ValueError("feature name must look like 'x3'")
                feature = int(a0[1:])
            elif isinstance(a0, (type, np.dtype)) and len(args) == 1:  # Node(T; ...)
                self.dtype = _as_dtype(a0)
            else:  # Node(op, children...)
                op = int(a0)
                children = tuple(args[1:])
        if children is not None:
            children = tuple(children)
            if not (1 <= len(children) <= MAX_DEGREE):
                raise ValueError(f"operator arity must be 1..{MAX_DEGREE}")
            if op is None or op < 1:
                raise ValueError("operator nodes need a 1-based op index")
            self.degree = len(children)
            self.op = int(op)
            self.children = children
            if self.dtype is None:
                for c in children:
                    if c.dtype is not None:
                        self.dtype = c.dtype
                        break
        elif parameter is not None:
            self.degree = 0
            self.is_parameter = True
            self.parameter = int(parameter)
            if self.parameter < 1:
                raise ValueError("parameter index is 1-based")
        elif feature is not None:
            self.degree = 0
            self.feature = int(feature)
            if not (1 <= self.feature <= 65535):
                raise ValueError("feature index is 1-based and must fit UInt16")
        elif val is not None:
            self.degree = 0
            self.constant = True
            if self.dtype is None and isinstance(val, (np.float32, np.float64)):
                self.dtype = type(val)
            self.val = float(val)
        else:
            raise ValueError("need one of val=, feature=, parameter= or (op, children...)")

    # -- traversal helpers (src/base.jl, src/NodeUtils.jl) -----------------------
    def __iter__(self):
        """Preorder iteration: parent first, children left to right."""
        stack = [self]
        while stack:
            n = stack.pop()
            yield n
            stack.extend(reversed(n.children))

    def copy(self):
        if self.degree == 0:
            n = Node.__new__(Node)
            for s in Node.__slots__:
                setattr(n, s, getattr(self, s))
            return n
        n = Node(self.op, *[c.copy() for c in self.children])
        n.dtype = self.dtype
        return n

    def __call__(self, X, operators, **kws):
        from .evaluate import call_tree
        return call_tree(self, X, operators, **kws)

    def __repr__(self):
        return string_tree(self)

    # operator overloading is installed by operators.extend_operators()


def count_nodes(tree: Node) -> int:
    """/root/reference/src/base.jl:271-280"""
    return sum(1 for _ in tree)


def count_depth(tree: Node) -> int:
    """Leaf = 1 (/root/reference/src/NodeUtils.jl:25-29)."""
    best = 0
    stack = [(tree, 1)]
    while stack:
        n, d = stack.pop()
        best = max(best, d)
        for c in n.children:
            stack.append((c, d + 1))
    return best


def is_node_constant(n: Node) -> bool:
    return n.degree == 0 and n.constant


def count_constant_nodes(tree: Node) -> int:
    """/root/reference/src/NodeUtils.jl:43-51"""
    return sum(1 for n in tree if is_node_constant(n))


def is_constant(tree: Node) -> bool:
    """True when the subtree holds no feature/parameter leaf (src/NodeUtils.jl:73)."""
    return all(n.constant for n in tree if n.degree == 0)


def get_scalar_constants(tree: Node):
    """Constants in depth-first left-to-right order (src/NodeUtils.jl:99-116) plus
    the node references that :func:`set_scalar_constants` writes back into."""
    refs = [n for n in tree if is_node_constant(n)]
    return np.array([n.val for n in refs], dtype=np.float64), refs


def set_scalar_constants(tree: Node, constants, refs=None):
    """/root/reference/src/NodeUtils.jl:118-143"""
    if refs is None:
        refs = [n for n in tree if is_node_constant(n)]
    if len(refs) != len(constants):
        raise ValueError("constant count mismatch")
    for n, v in zip(refs, constants):
        n.val = float(v)
    return tree


def max_feature(tree: Node) -> int:
    return max((n.feature for n in tree if n.degree == 0 and not n.constant and not n.is_parameter),
               default=0)


def tree_dtype(tree: Node, default=None):
    for n in tree:
        if n.dtype is not None:
            return n.dtype
    return default


def to_wire(tree: Node, out=None, offset=0):
    """Serialise ``tree`` to the preorder wire array (include/dex_wire.h), converting
    the 1-based ``feature`` / ``op`` / ``parameter`` of the host mirror to 0-based."""
    n = count_nodes(tree)
    if out is None:
        out = np.zeros(n, dtype=WIRE_DTYPE)
        offset = 0
    i = offset
    for nd in tree:
        rec = out[i]
        rec["degree"] = nd.degree
        if nd.degree == 0:
            if nd.constant:
                rec["kind"] = LEAF_CONST
                rec["val"] = nd.val
            elif nd.is_parameter:
                rec["kind"] = LEAF_PARAMETER
                rec["feature"] = nd.parameter - 1
            else:
                rec["kind"] = LEAF_FEATURE
                rec["feature"] = nd.feature - 1
        else:
            rec["op"] = nd.op - 1
        i += 1
    return out


def to_wire_population(trees):
    """Concatenate the wire arrays of many trees: returns (nodes, offsets[P+1])."""
    counts = np.fromiter((count_nodes(t) for t in trees), dtype=np.int64, count=len(trees))
    offsets = np.zeros(len(trees) + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    nodes = np.zeros(int(offsets[-1]), dtype=WIRE_DTYPE)
    for t, off in zip(trees, offsets[:-1]):
        to_wire(t, nodes, int(off))
    return nodes, offsets


def from_wire(nodes, T=None) -> Node:
    """Inverse of :func:`to_wire` (used by tests and the tree generator)."""
    pos = [0]

    def rec():
        r = nodes[pos[0]]
        pos[0] += 1
        d = int(r["degree"])
        if d == 0:
            k = int(r["kind"])
            if k == LEAF_CONST:
                return Node(val=float(r["val"]), T=T)
            if k == LEAF_PARAMETER:
                return Node(parameter=int(r["feature"]) + 1, T=T)
            return Node(feature=int(r["feature"]) + 1, T=T)
        op = int(r["op"]) + 1
        ch = [rec() for _ in range(d)]
        n = Node(op, *ch)
        n.dtype = _as_dtype(T) if T is not None else n.dtype
        return n

    return rec()


def string_tree(tree: Node, operators=None) -> str:
    """Debug printer in the spirit of /root/reference/src/Strings.jl (not byte-compatible)."""
    def name(deg, op):
        if operators is not None:
            try:
                return operators.names[deg - 1][op - 1]
            except Exception:
                pass
        return f"op{deg}_{op}"

    def rec(n):
        if n.degree == 0:
            if n.constant:
                return repr(n.val)
            if n.is_parameter:
                return f"p{n.parameter}"
            return f"x{n.feature}"
        nm = name(n.degree, n.op)
        args = [rec(c) for c in n.children]
        if n.degree == 2 and nm in ("+", "-", "*", "/", "^"):
            return f"({args[0]} {nm} {args[1]})"
        return f"{nm}({', '.join(args)})"

    return rec(tree)
