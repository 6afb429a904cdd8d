// dex_tape.h — device tape format of libdexb200 (internal; not ABI).
//
// A tree (/root/reference/src/Node.jl:74-90) is flattened on the host into a tape
// of fixed-size instructions for an ACCUMULATOR MACHINE: every instruction writes
// the accumulator ACC; operands come from ACC, from a shared-memory ROW (a staged
// feature row of X or an operand-stack slot), from an inline constant, or from a
// per-sample parameter gather.  Leaves never cost an instruction of their own
// unless they must be materialised (root leaf, second constant operand, ternary
// accumulator operand) — this is the device analogue of the reference's fused
// 2-/3-node kernels (/root/reference/src/Evaluate.jl:693-993).
//
// Stack slots and feature rows are ABSOLUTE shared-memory row indices fixed at pack
// time: rows [0, max_stack) are the operand stack, rows [max_stack, max_stack+F)
// are the features, so the interpreter never maintains a stack pointer.
#pragma once
#include <cstdint>
#include <vector>
#include <string>

namespace dex {

// ---- evaluation tape -------------------------------------------------------------
// 16 bytes, fetched as one uint4 (x=w0, y=w1, z/w = constant).
//   w0 [ 7: 0] opcode   builtin opcode of include/dex_ops.def (IDENTITY doubles as LOAD)
//      [ 9: 8] srcA     SRC_*
//      [11:10] srcB     SRC_*   (ternary: third operand is always ACC)
//      [12]    PUSH     store ACC to row push_row BEFORE executing
//      [13]    CHK_OUT  result participates in the `complete` flag
//      [14]    CHK_A    operand A (a leaf) participates
//      [15]    CHK_B    operand B (a leaf) participates
//      [16]    ALWAYS   checks apply even when early_exit is off (constant-subtree
//                       folding, /root/reference/src/Evaluate.jl:1059-1067)
//      [17]    GUARD    unary: result = isfinite(arg) ? op(arg) : Inf
//                       (/root/reference/src/Evaluate.jl:722, 737, 754, 787)
//      [31:24] push_row
//   w1 [15: 0] rowA  (ROW: smem row; PARAM: parameter index)
//      [31:16] rowB
//   c  inline constant: float in .z (F32) or double in .z/.w (F64)
struct Instr {
    uint32_t w0;
    uint32_t w1;
    uint32_t c_lo;
    uint32_t c_hi;
};
static_assert(sizeof(Instr) == 16, "tape instruction must be 16 bytes");

enum : uint32_t { SRC_ACC = 0, SRC_ROW = 1, SRC_CONST = 2, SRC_PARAM = 3 };
enum : uint32_t {
    F_PUSH = 1u << 12,
    F_CHK_OUT = 1u << 13,
    F_CHK_A = 1u << 14,
    F_CHK_B = 1u << 15,
    F_ALWAYS = 1u << 16,
    F_GUARD = 1u << 17,
};
constexpr int MAX_STACK_ROWS = 250;  // push_row is 8 bits

// ---- gradient tape ---------------------------------------------------------------
// Unfused post-order stack machine (the reference's derivative evaluator has no
// fused kernels, /root/reference/src/EvaluateDerivative.jl:262-365): every leaf is
// a LOAD that pushes (value, one-hot gradient seed), every operator pops its
// operands and pushes (value, gradient).  16 bytes.
//   w0 [ 7: 0] opcode (0 = LOAD)   [ 9: 8] leaf kind (LOAD only: 0 const, 1 feature, 2 parameter)
//      [31:16] dst stack slot (operands of an n-ary op are slots dst .. dst+n-1)
//   w1 leaf: feature / parameter index, or constant ordinal (0-based, leaf order)
//   c  constant value
struct GInstr {
    uint32_t w0;
    uint32_t w1;
    uint32_t c_lo;
    uint32_t c_hi;
};

// ---- host-side description of a packed population ----------------------------------
struct OpTable {
    std::vector<int32_t> ops[3];  // builtin opcodes per degree
};

struct PackedPopulation {
    int dtype = 0;
    int pack_flags = 0;
    int64_t n_trees = 0;
    int64_t n_nodes = 0;
    int64_t n_constants = 0;
    int32_t max_stack = 0;       // eval tape stack rows
    int32_t max_gstack = 0;      // grad tape stack slots
    int32_t max_feature = -1;
    int32_t max_parameter = -1;
    std::vector<Instr> tape;               // all trees, concatenated
    std::vector<int64_t> tape_off;         // n_trees + 1
    std::vector<GInstr> gtape;
    std::vector<int64_t> gtape_off;        // n_trees + 1
    std::vector<int32_t> n_nodes_tree;     // count_nodes per tree
    std::vector<int32_t> n_const_tree;     // count_constant_nodes per tree
    std::vector<int64_t> const_off;        // n_trees + 1 (prefix of n_const_tree)
    // constant ordinal (global) -> instruction index in tape / gtape holding its value
    std::vector<int64_t> const_pos;
    std::vector<int64_t> gconst_pos;
};

// Flatten `n_trees` wire trees.  Returns 0 or a negative DEX_ERR_* code with a
// message (tree index included) in `err`.
struct WireNode;  // = dex_node
int flatten_population(const OpTable& ops, const void* nodes, const int64_t* offsets,
                       int64_t n_trees, int dtype, int pack_flags, PackedPopulation& out,
                       std::string& err);

}  // namespace dex
