"""Deterministic synthetic tree populations for tests and benchmarks.

The growth rule is the reference's own test generator
(/root/reference/test/tree_gen_utils.jl:27-91): start from a random leaf, repeatedly pick
a random leaf and replace it by a random unary/binary operator with fresh random leaves
(leaf = constant ``randn`` w.p. 1/2, else a uniform feature; P(binary) = nbin/(nuna+nbin)).
The reference stops at a node count (``gen_random_tree_fixed_size``); the benchmark configs
ask for a *depth*, so :func:`gen_tree_to_depth` applies the same rule until ``count_depth``
(leaf = 1, /root/reference/src/NodeUtils.jl:25-29) first reaches the target.

Trees are produced directly in wire form (include/dex_wire.h) for speed; PRNG is numpy
PCG64 seeded per tree, so populations are reproducible anywhere.
"""
from __future__ import annotations

import numpy as np

from .node import LEAF_CONST, LEAF_FEATURE, LEAF_PARAMETER, WIRE_DTYPE


class _T:
    """Growable tree in struct-of-lists form."""
    __slots__ = ("deg", "kind", "op", "feat", "val", "ch", "depth_of", "leaves")

    def __init__(self):
        self.deg, self.kind, self.op, self.feat, self.val, self.ch, self.depth_of = [], [], [], [], [], [], []
        self.leaves = []

    def new_leaf(self, rng, nfeatures, n_params, depth):
        i = len(self.deg)
        self.deg.append(0)
        self.op.append(0)
        self.ch.append(())
        self.depth_of.append(depth)
        if n_params > 0:
            r = rng.integers(3)
            kind = (LEAF_CONST, LEAF_FEATURE, LEAF_PARAMETER)[r]
        else:
            kind = LEAF_CONST if rng.random() < 0.5 else LEAF_FEATURE
        self.kind.append(kind)
        if kind == LEAF_CONST:
            self.val.append(float(rng.standard_normal()))
            self.feat.append(0)
        elif kind == LEAF_FEATURE:
            self.val.append(0.0)
            self.feat.append(int(rng.integers(nfeatures)))
        else:
            self.val.append(0.0)
            self.feat.append(int(rng.integers(n_params)))
        self.leaves.append(i)
        return i

    def append_random_op(self, rng, nuna, nbin, nfeatures, n_params, force_unary=False):
        """append_random_op (tree_gen_utils.jl:37-69); returns the new maximum depth reached."""
        li = int(rng.integers(len(self.leaves)))
        i = self.leaves[li]
        self.leaves[li] = self.leaves[-1]
        self.leaves.pop()
        d = self.depth_of[i]
        binary = (not force_unary) and (rng.random() < nbin / (nuna + nbin))
        if binary:
            self.deg[i] = 2
            self.op[i] = int(rng.integers(nbin))
            a = self.new_leaf(rng, nfeatures, n_params, d + 1)
            b = self.new_leaf(rng, nfeatures, n_params, d + 1)
            self.ch[i] = (a, b)
        else:
            self.deg[i] = 1
            self.op[i] = int(rng.integers(nuna))
            a = self.new_leaf(rng, nfeatures, n_params, d + 1)
            self.ch[i] = (a,)
        return d + 1

    def to_wire(self):
        n = len(self.deg)
        out = np.zeros(n, dtype=WIRE_DTYPE)
        order = []
        stack = [0]
        while stack:
            i = stack.pop()
            order.append(i)
            stack.extend(reversed(self.ch[i]))
        idx = np.asarray(order)
        out["degree"] = np.asarray(self.deg, dtype=np.uint8)[idx]
        out["kind"] = np.where(out["degree"] == 0, np.asarray(self.kind, dtype=np.uint8)[idx], 0)
        out["op"] = np.asarray(self.op, dtype=np.uint8)[idx]
        out["feature"] = np.asarray(self.feat, dtype=np.uint16)[idx]
        out["val"] = np.asarray(self.val, dtype=np.float64)[idx]
        return out


def gen_tree_to_depth(depth, nuna, nbin, nfeatures, rng, max_nodes=None, n_params=0, dtype=np.float32):
    """Wire array of a random tree whose count_depth first reaches ``depth``."""
    if max_nodes is None:
        max_nodes = 2 ** depth - 1
    t = _T()
    t.new_leaf(rng, nfeatures, n_params, 1)
    cur = 1
    while cur < depth and len(t.deg) + 2 <= max_nodes:
        cur = max(cur, t.append_random_op(rng, nuna, nbin, nfeatures, n_params))
    w = t.to_wire()
    if dtype == np.float32:
        w["val"] = w["val"].astype(np.float32).astype(np.float64)  # exactly representable
    return w


def gen_random_tree_fixed_size(node_count, nuna, nbin, nfeatures, rng, n_params=0, dtype=np.float32):
    """gen_random_tree_fixed_size (tree_gen_utils.jl:71-91) in wire form."""
    t = _T()
    t.new_leaf(rng, nfeatures, n_params, 1)
    while len(t.deg) < node_count:
        if len(t.deg) == node_count - 1:
            if nuna == 0:
                break
            t.append_random_op(rng, nuna, nbin, nfeatures, n_params, force_unary=True)
        else:
            t.append_random_op(rng, nuna, nbin, nfeatures, n_params)
    w = t.to_wire()
    if dtype == np.float32:
        w["val"] = w["val"].astype(np.float32).astype(np.float64)
    return w


def gen_population(n_trees, depth, nuna, nbin, nfeatures, seed=0, max_nodes=None, n_params=0,
                   dtype=np.float32, node_count=None):
    """(nodes, offsets) of ``n_trees`` trees; tree i uses ``default_rng([seed, i])``."""
    wires = []
    for i in range(n_trees):
        rng = np.random.default_rng([seed, i])
        if node_count is not None:
            wires.append(gen_random_tree_fixed_size(node_count, nuna, nbin, nfeatures, rng, n_params, dtype))
        else:
            wires.append(gen_tree_to_depth(depth, nuna, nbin, nfeatures, rng, max_nodes, n_params, dtype))
    offsets = np.zeros(n_trees + 1, dtype=np.int64)
    np.cumsum([len(w) for w in wires], out=offsets[1:])
    return np.concatenate(wires), offsets


# operator sets of SURVEY.md §8d
OPSET_A = {1: ("cos", "exp"), 2: ("+", "-", "/", "*")}          # benchmark/benchmarks.jl:32-36
OPSET_B = {1: ("sin", "cos", "exp", "abs"), 2: ("+", "-", "*", "/")}
