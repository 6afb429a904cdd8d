"""dexb200 — B200-native batched expression-tree evaluation behind the
DynamicExpressions.jl evaluation API (eval_tree_array / eval_grad_tree_array /
OperatorEnum / Expression / ParametricExpression).

Host mirror (pure Python, this package) -> C ABI (include/dexb200.h,
libdexb200.so built from csrc/ with nvcc for sm_100a) -> CUDA kernels.
There is no CPU fallback: every evaluation call runs the CUDA library and raises
if it cannot be loaded or no device is present.
"""
from .node import (Node, count_nodes, count_depth, count_constant_nodes, is_constant,
                   get_scalar_constants, set_scalar_constants, string_tree, to_wire,
                   to_wire_population, from_wire, WIRE_DTYPE)
from .operators import OperatorEnum, extend_operators, call, opcode_of, OPCODE_INFO, OPCODE_TABLE

__all__ = [
    "Node", "count_nodes", "count_depth", "count_constant_nodes", "is_constant",
    "get_scalar_constants", "set_scalar_constants", "string_tree", "to_wire",
    "to_wire_population", "from_wire", "WIRE_DTYPE", "OperatorEnum", "extend_operators", "call",
    "opcode_of", "OPCODE_INFO", "OPCODE_TABLE",
]


def __getattr__(name):
    # evaluation entry points are imported lazily so that pure-host utilities
    # (tree building, wire format) work without touching the CUDA library
    import importlib
    for mod in ("evaluate", "expression", "device", "sharded", "treegen"):
        try:
            m = importlib.import_module(f"dexb200.{mod}")
        except ModuleNotFoundError as e:
            if e.name == f"dexb200.{mod}":
                continue
            raise
        if hasattr(m, name):
            return getattr(m, name)
    raise AttributeError(f"module 'dexb200' has no attribute {name!r}")
