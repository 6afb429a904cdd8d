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
// Stack slots, parameter rows and feature rows are ABSOLUTE shared-memory row indices
// fixed at pack time: rows [0, max_stack) are the operand stack, the next n_param_rows
// rows hold the current tree's per-sample parameters (ParametricExpression), the features
// follow, so the interpreter never maintains a stack pointer.
//
// Every instruction also carries a HANDLER id: the index of a code path in the
// interpreter that is specialised for (operator, operand sources), so that the hot
// operators need one indirect branch and no operand decoding.  Handler 0 is the
// generic path (any operator, any operand source, every flag).
#pragma once
#include <cstdint>
#include <memory>
#include <vector>
#include <string>

namespace dex {

// ---- evaluation tape -------------------------------------------------------------
// 16 bytes, fetched as one uint4 (x=w0, y=w1, z/w = constant).
//   w0 [ 5: 0] handler  HANDLER id (H_*), 0 = generic
//      [ 6]    copy of PUSH: the Float32 jump tables have one entry per (handler, PUSH)
//              so that the push costs nothing when it does not happen
//      [ 7]    copy of CHK_OUT (reserved for a table variant; see gen_interp_ptx.py)
//      [15: 8] opcode   builtin opcode of include/dex_ops.def (IDENTITY doubles as LOAD)
//      [17:16] srcA     SRC_*
//      [19:18] srcB     SRC_*   (ternary: third operand is always ACC)
//      [20]    PUSH     store ACC to row push_row BEFORE executing
//      [21]    CHK_OUT  result participates in the `complete` flag
//      [22]    CHK_A    operand A (a leaf) participates
//      [23]    CHK_B    operand B (a leaf) participates
//      [24]    ALWAYS   checks apply even when early_exit is off (constant-subtree
//                       folding, /root/reference/src/Evaluate.jl:1059-1067)
//      [25]    GUARD    unary: result = isfinite(arg) ? op(arg) : Inf
//                       (/root/reference/src/Evaluate.jl:722, 737, 754, 787)
//      [26]    SWAPPED  the flattener exchanged the operands of max / min (to reach an
//                       (ACC|ROW, ROW|CONST) handler form): values are symmetric, but the
//                       reference's partials (x > y, !(x > y)) break ties by operand ORDER, so
//                       the gradient interpreters apply them to the original order
//      CHK_A / CHK_B apply to whatever the operand is: a feature ROW or the inline constant
//      [31:27] push_row  stack slot the PUSH stores to (a tree of n nodes needs about
//                       log2(n) slots in Sethi-Ullman order, so five bits are ample)
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
    F_PUSH = 1u << 20,
    F_CHK_OUT = 1u << 21,
    F_CHK_A = 1u << 22,
    F_CHK_B = 1u << 23,
    F_ALWAYS = 1u << 24,
    F_GUARD = 1u << 25,
    F_SWAPPED = 1u << 26,
};
constexpr int MAX_STACK_ROWS = 31;   // push_row is 5 bits
constexpr int MAX_ROWS = 65535;      // row fields are 16 bits
constexpr int PUSH_ROW_SHIFT = 27;
#if defined(__CUDACC__)
#define DEX_HD __host__ __device__ __forceinline__
#else
#define DEX_HD inline
#endif
DEX_HD uint32_t row_a(uint32_t w1) { return w1 & 0xffffu; }
DEX_HD uint32_t row_b(uint32_t w1) { return w1 >> 16; }
DEX_HD uint32_t push_row(uint32_t w0) { return w0 >> PUSH_ROW_SHIFT; }

// Operators with specialised handlers.  Unary: operand in ACC (_A) or in a ROW (_R).
// Commutative binary: (ACC,ROW) (ACC,CONST) (ROW,ROW) (ROW,CONST) — the flattener swaps
// operands into these forms.  Non-commutative binary: all seven operand forms.
#define DEX_FAST_UNARY(X) \
    X(NEG) X(ABS) X(SQUARE) X(CUBE) X(INV) X(SQRT) X(EXP) X(LOG) X(SIN) X(COS) X(TANH) \
    X(RELU) X(SAFE_LOG) X(SAFE_SQRT)
#define DEX_FAST_BIN_COMM(X) X(ADD) X(MUL) X(MAX) X(MIN)
#define DEX_FAST_BIN_NC(X) X(SUB) X(DIV)

enum Handler : uint32_t {
    H_GENERIC = 0,
    H_LOAD_R,  // ACC = ROW
    H_LOAD_C,  // ACC = const
#define X(S) H_##S##_A, H_##S##_R,
    DEX_FAST_UNARY(X)
#undef X
#define X(S) H_##S##_AR, H_##S##_AC, H_##S##_RR, H_##S##_RC,
    DEX_FAST_BIN_COMM(X)
#undef X
#define X(S) H_##S##_AR, H_##S##_RA, H_##S##_AC, H_##S##_CA, H_##S##_RR, H_##S##_RC, H_##S##_CR,
    DEX_FAST_BIN_NC(X)
#undef X
    H_KEEP,    // ACC unchanged: with PUSH it stores ACC to a row that outlives the operand stack
               // discipline (a shared subexpression, dex_flatten.cpp)
    H__COUNT
};
constexpr uint32_t HANDLER_MASK = 63u, HANDLER_PUSH = 64u, HANDLER_CHK = 128u;
static_assert(H__COUNT <= HANDLER_PUSH - 1, "handler ids must fit six bits (one slot is reserved)");

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
    int32_t max_feature = -1;
    int32_t max_parameter = -1;
    int32_t n_param_rows = 0;    // shared-memory rows reserved for parameters (max_parameter + 1)
    int64_t n_generic = 0;       // instructions that take the generic handler
    int64_t n_checks = 0;        // validity checks left after elision
    std::vector<Instr> tape;               // all trees, concatenated
    std::vector<int64_t> tape_off;         // n_trees + 1
    std::vector<int32_t> tape_const_ord;   // per instruction: tree-local ordinal of its inline constant, -1 if none
    std::vector<int32_t> n_nodes_tree;     // count_nodes per tree
    std::vector<int32_t> n_const_tree;     // count_constant_nodes per tree
    std::vector<int64_t> const_off;        // n_trees + 1 (prefix of n_const_tree)
    // constant ordinal (global) -> instruction index in the tape holding its value; in a
    // folded image a constant that lives in the scalar tape is stored as -(1 + ctape index)
    std::vector<int64_t> const_pos;

    // ---- folded image (evaluation only) ------------------------------------------------
    // The reference evaluates every maximal constant subtree as a scalar and fills the result
    // (_eval_constant_tree, /root/reference/src/Evaluate.jl:347-354, 1059-1114).  Device
    // analogue: `folded` is a second flattening of the same trees in which such a subtree is
    // ONE inline constant of its consumer; its operators live in the scalar tape `ctape`,
    // which the prepass kernel runs once per evaluation call (one thread per tree) and whose
    // results it stores into the constant slots of `folded->tape`.
    std::vector<Instr> ctape;        // scalar instructions of all folded subtrees
    std::vector<int64_t> seg;        // per subtree: ctape begin, ctape end, target instruction
    std::vector<int64_t> seg_off;    // n_trees + 1: subtrees of tree t are seg_off[t]..seg_off[t+1]
    std::shared_ptr<PackedPopulation> folded;   // only set on the outer (unfolded) image
};

// Flatten `n_trees` wire trees.  Returns 0 or a negative DEX_ERR_* code with a
// message (tree index included) in `err`.
int flatten_population(const OpTable& ops, const void* nodes, const int64_t* offsets,
                       int64_t n_trees, int dtype, int pack_flags, PackedPopulation& out,
                       std::string& err);

}  // namespace dex
