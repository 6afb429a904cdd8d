// dex_fold.cuh — scalar evaluation of the folded constant subtrees of one tree (device).
// Shared by the evaluation prepass (dex_eval.cu) and the gradient prepass (dex_grad.cu).
#pragma once
#include "dex_ops.cuh"
#include "dex_tape.h"

namespace dex {

template <typename T> __device__ __forceinline__ T fold_const_of(const uint4& ins);
template <> __device__ __forceinline__ float fold_const_of<float>(const uint4& ins) { return __uint_as_float(ins.z); }
template <> __device__ __forceinline__ double fold_const_of<double>(const uint4& ins) { return __hiloint2double((int)ins.w, (int)ins.z); }

// Constant-subtree folding of tree t: runs the scalar segments of the tree (dex_tape.h,
// PackedPopulation::ctape) and stores each result into the inline-constant slot of the
// instruction that consumes it.  Same operator code as the sample loop.
//
// GRAD = false (evaluation): returns false when a value the reference's scalar walk checks is
// not finite (_eval_constant_tree returns ResultOk(.., false),
// /root/reference/src/Evaluate.jl:1059-1114) — whatever early_exit says.
//
// GRAD = true (eval_grad_tree_array with variable = Val(true), whose reference path has no
// constant folding and validates value AND gradient of every node,
// /root/reference/src/EvaluateDerivative.jl:238-243): every node value of the subtree must be
// finite, and so must every partial derivative — the subtree's gradient is a chain of
// `partial * 0` products (all seeds are zero), which is NaN exactly when some partial is not
// finite, and 0 otherwise.
template <typename T, bool GRAD>
__device__ bool fold_tree(Instr* tape, const Instr* ctape, const int64_t* seg, const int64_t* seg_off, int64_t t) {
    bool ok = true;
    T st[MAX_STACK_ROWS + 1];
    for (int64_t sg = seg_off[t]; sg < seg_off[t + 1]; ++sg) {
        const int64_t begin = seg[3 * sg], end = seg[3 * sg + 1], target = seg[3 * sg + 2];
        T acc = T(0);
        for (int64_t pc = begin; pc < end; ++pc) {
            const uint4 ins = *reinterpret_cast<const uint4*>(ctape + pc);
            const uint32_t w0 = ins.x;
            const T c = fold_const_of<T>(ins);
            if (w0 & F_PUSH) st[push_row(w0)] = acc;
            const uint32_t sa = (w0 >> 16) & 3u, sb = (w0 >> 18) & 3u;
            const T x = sa == SRC_ROW ? st[row_a(ins.y)] : (sa == SRC_CONST ? c : acc);
            const T y = sb == SRC_ROW ? st[row_b(ins.y)] : (sb == SRC_CONST ? c : acc);
            const T z = acc;
            const uint32_t op = (w0 >> 8) & 0xffu;
            if (GRAD) {
                if (!t_finite(x)) ok = false;
                if (op >= 64u && !t_finite(y)) ok = false;
                if (op >= 128u && !t_finite(z)) ok = false;
            } else {
                if ((w0 & F_CHK_A) && !t_finite(x)) ok = false;
                if ((w0 & F_CHK_B) && !t_finite(y)) ok = false;
            }
            T v, p0 = T(0), p1 = T(0), p2 = T(0);
            switch (op) {
#define U_CASE(SYM, VEXPR, GEXPR) \
    case DEX_OP_##SYM: { v = (VEXPR); if (GRAD) p0 = (GEXPR); } break;
                DEX_UNARY_OPS(U_CASE)
#undef U_CASE
#define B_CASE(SYM, VEXPR, GA, GB) \
    case DEX_OP_##SYM: { v = (VEXPR); if (GRAD) { p0 = (GA); p1 = (GB); } } break;
                DEX_BINARY_OPS(B_CASE)
#undef B_CASE
#define T_CASE(SYM, VEXPR, GA, GB, GZ) \
    case DEX_OP_##SYM: { v = (VEXPR); if (GRAD) { p0 = (GA); p1 = (GB); p2 = (GZ); } } break;
                DEX_TERNARY_OPS(T_CASE)
#undef T_CASE
                default: v = t_nan<T>(); break;
            }
            (void)y; (void)z;
            if (GRAD) {
                if (!t_finite(v) || !t_finite(p0) || !t_finite(p1) || !t_finite(p2)) ok = false;
            } else if ((w0 & F_CHK_OUT) && !t_finite(v)) ok = false;
            acc = v;
        }
        if (target >= 0) {
            uint32_t lo, hi;
            if (sizeof(T) == 4) { lo = __float_as_uint((float)acc); hi = 0; }
            else { lo = (uint32_t)__double2loint((double)acc); hi = (uint32_t)__double2hiint((double)acc); }
            tape[target].c_lo = lo;
            tape[target].c_hi = hi;
        }
    }
    return ok;
}

}  // namespace dex
