"""Loader for tests/golden/reference_known_answers.json (see make_known_answers.py)."""
import json
import os

import numpy as np

import dexb200
from dexb200 import Node, OperatorEnum

HERE = os.path.dirname(os.path.abspath(__file__))


def load_cases():
    with open(os.path.join(HERE, "golden", "reference_known_answers.json")) as f:
        return json.load(f)["cases"]


def _num(v, dtype):
    if isinstance(v, str):
        if v == "inf":
            return np.inf
        if v == "-inf":
            return -np.inf
        if v == "nan":
            return np.nan
        if v == "floatmax":
            return float(np.finfo(dtype).max)
        raise ValueError(v)
    return float(v)


def make_operators(case):
    spec = case["operators"]
    return OperatorEnum({int(d): tuple(names) for d, names in spec.items()})


def make_tree(spec, operators, dtype):
    if isinstance(spec, dict):  # explicit op index form
        ch = [make_tree(c, operators, dtype) for c in spec["children"]]
        n = Node(spec["op_index"], *ch)
        n.dtype = dtype
        return n
    if isinstance(spec, list):
        name, args = spec[0], spec[1:]
        ch = [make_tree(a, operators, dtype) for a in args]
        n = Node(operators.index_of(name, len(ch)), *ch)
        n.dtype = dtype
        return n
    if isinstance(spec, str) and spec[0] == "x" and spec[1:].isdigit():
        return Node(feature=int(spec[1:]), T=dtype)
    if isinstance(spec, str) and spec[0] == "p" and spec[1:].isdigit():
        return Node(parameter=int(spec[1:]), T=dtype)
    return Node(val=_num(spec, dtype), T=dtype)


def make_matrix(spec, dtype):
    if isinstance(spec, dict):
        rng = np.random.default_rng(spec["seed"])
        if "randn" in spec:
            a = rng.standard_normal(tuple(spec["randn"]))
        else:
            a = rng.random(tuple(spec["rand"]))
        a = a * spec.get("scale", 1.0) + spec.get("offset", 0.0)
        return a.astype(dtype)
    return np.array([[_num(v, dtype) for v in row] for row in spec], dtype=dtype)


def expected_y(case, X, P=None, cls=None):
    e = case["expect"]
    if "y" in e:
        return [None if v == "nonfinite" else _num(v, X.dtype.type) for v in e["y"]]
    if "formula" in e:
        X64 = X.astype(np.float64)
        P64 = None if P is None else np.asarray(P, dtype=np.float64)
        with np.errstate(all="ignore"):
            return np.asarray(eval(e["formula"], {"np": np, "X": X64, "P": P64, "cls": cls}),
                              dtype=np.float64)
    return None


def expected_grad(case, X):
    e = case["expect"]
    if "grad" in e:
        return np.array(e["grad"], dtype=np.float64)
    if "grad_formula" in e:
        X64 = X.astype(np.float64)
        with np.errstate(all="ignore"):
            return np.stack([np.asarray(eval(f, {"np": np, "X": X64}), dtype=np.float64) + 0 * X64[0]
                             for f in e["grad_formula"]])
    return None


CONTEXTS = {
    "default": dict(),
    "bumper": dict(bumper=True),
    "unfused": dict(use_fused=False),
    "no_early_exit": dict(early_exit=False),
    "bumper_no_early_exit": dict(bumper=True, early_exit=False),
}
