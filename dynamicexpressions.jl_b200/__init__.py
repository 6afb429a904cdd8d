"""dexb200 — B200-native batched expression-tree evaluation behind the
DynamicExpressions.jl evaluation API (eval_tree_array / eval_grad_tree_array /
OperatorEnum / Expression / ParametricExpression).

Host mirror (pure Python, this package) -> C ABI (include/dexb200.h,
libdexb200.so built from csrc/ with nvcc for sm_100a) -> CUDA kernels.
There is no CPU fallback: every evaluation call runs the CUDA library and raises
if it cannot be loaded or no device is present.

The native library is loaded lazily (first evaluation / ``device.lib()``), so tree
building and the wire format work on machines without it.
"""
from .node import (Node, count_nodes, count_depth, count_constant_nodes, is_constant,
                   get_scalar_constants, set_scalar_constants, string_tree, to_wire,
                   to_wire_population, from_wire, WIRE_DTYPE)
from .operators import OperatorEnum, extend_operators, call, opcode_of, OPCODE_INFO, OPCODE_TABLE

from . import device, treegen
from .device import Context, Population, DexError
from .evaluate import (EvalContext, EvalOptions, eval_tree_array, eval_trees_array,
                       eval_grad_tree_array, eval_grad_trees_array, eval_diff_tree_array,
                       call_tree, grad_tree, validate_input)
from .expression import (Expression, ParametricExpression, ParametricNode,
                         eval_parametric_trees_array)

__all__ = [
    "Node", "count_nodes", "count_depth", "count_constant_nodes", "is_constant",
    "get_scalar_constants", "set_scalar_constants", "string_tree", "to_wire",
    "to_wire_population", "from_wire", "WIRE_DTYPE", "OperatorEnum", "extend_operators", "call",
    "opcode_of", "OPCODE_INFO", "OPCODE_TABLE", "Context", "Population", "DexError",
    "EvalContext", "EvalOptions", "eval_tree_array", "eval_trees_array", "eval_grad_tree_array",
    "eval_grad_trees_array", "eval_diff_tree_array", "call_tree", "grad_tree", "validate_input",
    "Expression", "ParametricExpression", "ParametricNode", "eval_parametric_trees_array",
    "device", "treegen",
]
