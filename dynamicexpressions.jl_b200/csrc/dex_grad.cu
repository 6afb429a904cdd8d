// dex_grad.cu — batched forward-mode derivative evaluation (sm_100a).
//
// Replaces eval_grad_tree_array / eval_diff_tree_array
// (/root/reference/src/EvaluateDerivative.jl:40-168, 193-404) for a whole population in
// one launch.  The reference allocates and zero-fills a (G x N) matrix at EVERY leaf
// (:376-380) and streams two of them through memory at every operator (:340-365); here the
// dual number (value, d/d theta_1..GC) of the accumulator lives in registers, the operand
// stack in shared memory, and nothing but X in and (value, gradient) out touches HBM.
//
// The kernel interprets the SAME fused tape as dex_eval.cu (leaves folded into their
// consumer, Sethi-Ullman order, absolute stack rows, handler ids): an operand is ACC, a
// stack slot, a feature row, or an inline constant, and a leaf's derivative is a one-hot (or
// zero) vector that is never materialised in memory.
//
// Mapping
//   grid.x  sample tiles of TILE = blockDim.x * K samples
//   grid.y  chunks of trees
//   thread  K = U * C samples as U 16-byte chunks (C = 4 floats / 2 doubles; chunk u of
//           thread t covers samples u*(blockDim.x*C) + t*C ..), ACC = (1 + GC) * K registers
//   smem    stack slot s component c: row s * (1 + GC) + c;  features behind the stack rows
//   GC      compile-time number of directions per pass (1..8); trees with more directions
//           (many constants) take several passes, recomputing the primal per pass
//
// One tape instruction runs in two stages, each entered through one warp-uniform switch:
//   stage 1 (by HANDLER id = operator x operand forms): fetch the operand values, compute the
//           value v and the partials p0, p1, note where each operand's derivative lives
//           (ACC registers / stack slot rows / one-hot leaf) and the coefficient class;
//   stage 2 (by coefficient class x operand kinds): d[g] = p0 * dA[g] + p1 * dB[g] for the GC
//           directions, fully unrolled and branch-free.  Classes: ADD (1, 1), SUB (1, -1),
//           VAR (partials that are finite whenever the node values are: *, max, min, exp,
//           sin, ...) and GEN (anything: /, sqrt, log, ...).  For ADD/SUB/VAR a one-hot leaf
//           contributes nothing to the directions it does not own, so those products are
//           skipped; for GEN they are computed, because an infinite partial times a zero
//           seed is NaN in the reference (`grad * d_cumulator`, :355-361) and must fail here
//           too (sqrt(x1) at x1 = 0).
//
// Semantics: every VALUE of every node, leaves included, must be finite for `complete`
// (:238-243) and so must every gradient component.  A non-finite derivative component can
// never become finite again on its way to the root (d_parent[g] = sum_i p_i * d_i[g]; a
// non-finite factor makes the product non-finite and a non-finite term makes the sum
// non-finite), so derivatives are checked once, at the root; values are checked at every
// node.  eval_diff never checks (:68-85) and takes the GEN class everywhere, so that its
// non-finite patterns are those of the reference's arithmetic.
//
// The C++ `step` below is the complete form of one tape instruction (all operators, both element
// types).  For Float32 the instruction loop of the 256- and 128-thread launches is the generated
// inline-PTX block of gen_grad_ptx.py (jump tables, in-place register updates, stage 2 inlined
// per operand-kind variant); it hands instructions without a native code path to `step`.
// Kernel modes (KM_*): gradient blocks stored, eval_diff, or the fused loss + gradient of the
// loss whose reduction replaces the stores.
#include "dex_kernels.h"
#include "dex_ops.cuh"
#include "dex_fold.cuh"
#include "../../include/dexb200.h"

#include <algorithm>

// 16-byte chunks per thread for Float64 (K = 2 * U samples)
#ifndef DEX_GRAD_U64
#define DEX_GRAD_U64 1
#endif
#ifndef DEX_GRAD_U
#define DEX_GRAD_U 1
#endif
#ifndef DEX_GRAD_PTX
#define DEX_GRAD_PTX 1
#endif
// shared memory per CTA above which the launcher halves the CTA (keeps >= 2 CTAs per SM)
#ifndef DEX_GRAD_SMEM_SOFT
#define DEX_GRAD_SMEM_SOFT (100 * 1024)
#endif
#ifndef DEX_GRAD_MIN_CTAS
#define DEX_GRAD_MIN_CTAS 2
#endif
#ifndef DEX_GRAD_THREADS
#define DEX_GRAD_THREADS 256
#endif

namespace dex {
namespace {

template <typename T> struct GK {
    const uint4* tape;
    const int64_t* tape_off;
    const int32_t* const_ord;   // per tape instruction: tree-local ordinal of its inline constant, or -1
    const int64_t* const_off;   // per tree: global ordinal base (for n_const)
    const int32_t* chunk_start;
    const T* X;                 // feature-major padded copy XT[f][ldx]
    T* out;
    T* grad;
    const int64_t* grad_off;
    uint8_t* ok;
    int64_t N, ldx, ldo, n_trees;
    int32_t F, max_stack, mode, direction;
    // fused loss mode (KM_LOSS): targets, optional weights, per-tile partial sums
    const T* y;
    const T* w;
    double* partial;          // [n_tiles][partial_stride]: n_trees losses, then the gradient entries
    int64_t partial_stride;
    // ParametricExpression: F counts ALL leaf rows (n_param_rows gathered rows, then the features)
    const T* params;
    const int32_t* classes;
    int32_t n_params, n_classes, n_param_rows;
};

template <typename T> __device__ __forceinline__ T gconst_of(const uint4& ins);
template <> __device__ __forceinline__ float gconst_of<float>(const uint4& ins) { return __uint_as_float(ins.z); }
template <> __device__ __forceinline__ double gconst_of<double>(const uint4& ins) { return __hiloint2double((int)ins.w, (int)ins.z); }

// ---- K-vectors of one thread ------------------------------------------------------------
template <typename T, int U> struct VK {
    static constexpr int C = 16 / (int)sizeof(T);
    static constexpr int K = U * C;
    static __device__ __forceinline__ void ld(T (&v)[K], const T* p, int CS) {
#pragma unroll
        for (int u = 0; u < U; ++u)
            *reinterpret_cast<uint4*>(&v[u * C]) = *reinterpret_cast<const uint4*>(p + u * CS);
    }
    static __device__ __forceinline__ void st(T* p, int CS, const T (&v)[K]) {
#pragma unroll
        for (int u = 0; u < U; ++u)
            *reinterpret_cast<uint4*>(p + u * CS) = *reinterpret_cast<const uint4*>(&v[u * C]);
    }
};

// element-wise arithmetic; every operation rounds once, like the reference's scalar code.
// Float32 uses Blackwell's packed FMUL2 / FADD2 / FFMA2 (two IEEE operations per instruction).
template <typename T, int K> struct VA {
    static __device__ __forceinline__ void mul(T (&d)[K], const T (&a)[K], const T (&b)[K]) {
#pragma unroll
        for (int k = 0; k < K; ++k) d[k] = a[k] * b[k];
    }
    static __device__ __forceinline__ void muls(T (&d)[K], const T (&a)[K], T s) {
#pragma unroll
        for (int k = 0; k < K; ++k) d[k] = a[k] * s;
    }
    static __device__ __forceinline__ void add(T (&d)[K], const T (&a)[K], const T (&b)[K]) {
#pragma unroll
        for (int k = 0; k < K; ++k) d[k] = a[k] + b[k];
    }
    // d = a * b + c, fused (the derivative combination p0 dA + p1 dB keeps one rounding less
    // than the reference's separate products; same form as the generated PTX loops)
    static __device__ __forceinline__ void fma(T (&d)[K], const T (&a)[K], const T (&b)[K], const T (&c)[K]) {
#pragma unroll
        for (int k = 0; k < K; ++k) d[k] = m_fma(a[k], b[k], c[k]);
    }
    static __device__ __forceinline__ void adds(T (&d)[K], const T (&a)[K], T s) {
#pragma unroll
        for (int k = 0; k < K; ++k) d[k] = a[k] + s;
    }
    static __device__ __forceinline__ void sub(T (&d)[K], const T (&a)[K], const T (&b)[K]) {
#pragma unroll
        for (int k = 0; k < K; ++k) d[k] = a[k] - b[k];
    }
    static __device__ __forceinline__ void neg(T (&d)[K], const T (&a)[K]) {
#pragma unroll
        for (int k = 0; k < K; ++k) d[k] = -a[k];
    }
    static __device__ __forceinline__ void check(T& nf, const T (&a)[K]) {
#pragma unroll
        for (int k = 0; k < K; ++k) nf = m_fma(a[k], T(0), nf);
    }
};
template <int K> struct VA<float, K> {
    static __device__ __forceinline__ float2 f2(const float* p) { return make_float2(p[0], p[1]); }
    static __device__ __forceinline__ void mul(float (&d)[K], const float (&a)[K], const float (&b)[K]) {
#pragma unroll
        for (int k = 0; k < K; k += 2) { const float2 r = __fmul2_rn(f2(a + k), f2(b + k)); d[k] = r.x; d[k + 1] = r.y; }
    }
    static __device__ __forceinline__ void muls(float (&d)[K], const float (&a)[K], float s) {
        const float2 ss = make_float2(s, s);
#pragma unroll
        for (int k = 0; k < K; k += 2) { const float2 r = __fmul2_rn(f2(a + k), ss); d[k] = r.x; d[k + 1] = r.y; }
    }
    static __device__ __forceinline__ void add(float (&d)[K], const float (&a)[K], const float (&b)[K]) {
#pragma unroll
        for (int k = 0; k < K; k += 2) { const float2 r = __fadd2_rn(f2(a + k), f2(b + k)); d[k] = r.x; d[k + 1] = r.y; }
    }
    static __device__ __forceinline__ void fma(float (&d)[K], const float (&a)[K], const float (&b)[K], const float (&c)[K]) {
#pragma unroll
        for (int k = 0; k < K; k += 2) { const float2 r = __ffma2_rn(f2(a + k), f2(b + k), f2(c + k)); d[k] = r.x; d[k + 1] = r.y; }
    }
    static __device__ __forceinline__ void adds(float (&d)[K], const float (&a)[K], float s) {
        const float2 ss = make_float2(s, s);
#pragma unroll
        for (int k = 0; k < K; k += 2) { const float2 r = __fadd2_rn(f2(a + k), ss); d[k] = r.x; d[k + 1] = r.y; }
    }
    // a - b as fma(b, -1, a): the product is exact, one rounding — identical to a + (-b)
    static __device__ __forceinline__ void sub(float (&d)[K], const float (&a)[K], const float (&b)[K]) {
        const float2 m1 = make_float2(-1.f, -1.f);
#pragma unroll
        for (int k = 0; k < K; k += 2) { const float2 r = __ffma2_rn(f2(b + k), m1, f2(a + k)); d[k] = r.x; d[k + 1] = r.y; }
    }
    static __device__ __forceinline__ void neg(float (&d)[K], const float (&a)[K]) {
#pragma unroll
        for (int k = 0; k < K; ++k) d[k] = -a[k];
    }
    static __device__ __forceinline__ void check(float& nf, const float (&a)[K]) {
        float2 acc = make_float2(nf, 0.f);
#pragma unroll
        for (int k = 0; k < K; k += 2) acc = __ffma2_rn(f2(a + k), make_float2(0.f, 0.f), acc);
        nf = acc.x + acc.y;
    }
};

// ---- operators: value and partials as compile-time functors --------------------------------
template <int OPC, typename T> struct G1;
template <int OPC, typename T> struct G2;
#define X1(SYM, VEXPR, GEXPR)                                                        \
    template <typename T> struct G1<DEX_OP_##SYM, T> {                               \
        static __device__ __forceinline__ void f(T x, T& vo, T& p0) {                \
            const T v = (VEXPR);                                                     \
            p0 = (GEXPR);                                                            \
            vo = v;                                                                  \
        }                                                                            \
    };
DEX_UNARY_OPS(X1)
#undef X1
#define X2(SYM, VEXPR, GA, GB)                                                       \
    template <typename T> struct G2<DEX_OP_##SYM, T> {                               \
        static __device__ __forceinline__ void f(T x, T y, T& vo, T& p0, T& p1) {    \
            const T v = (VEXPR);                                                     \
            p0 = (GA);                                                               \
            p1 = (GB);                                                               \
            vo = v;                                                                  \
        }                                                                            \
    };
DEX_BINARY_OPS(X2)
#undef X2
// x / y: the quotient and 1/y are IEEE divisions; d/dy = -(v/y) is taken as -(v * (1/y)), one
// multiplication instead of a third division (<= 1 ulp from the quotient; when 1/y overflows
// the tree fails through p0 either way).
template <> struct G2<DEX_OP_DIV, float> {
    static __device__ __forceinline__ void f(float x, float y, float& vo, float& p0, float& p1) {
        const float r = 1.0f / y;
        vo = x / y;
        p0 = r;
        p1 = -(vo * r);
    }
};

// where the derivative of an operand lives
enum : int { DK_ACC = 0, DK_SLOT = 1, DK_LEAF = 2 };
// coefficient classes (see the header comment)
enum : int { CL_ADD = 0, CL_SUB = 1, CL_VAR = 2, CL_GEN = 3 };      // binary: (p0, p1)
enum : int { UL_ONE = 0, UL_NEG = 1, UL_VAR = 2, UL_GEN = 3 };      // unary: p0
// stage-2 selector: binary cls*9 + ka*3 + kb in [0, 36); unary 36 + ucls*3 + ka in [36, 48)
constexpr int SEL_UNARY = 36, SEL_NONE = 48;

template <int OPC> struct BinClass { static constexpr int v = CL_GEN; };
template <> struct BinClass<DEX_OP_ADD> { static constexpr int v = CL_ADD; };
template <> struct BinClass<DEX_OP_SUB> { static constexpr int v = CL_SUB; };
template <> struct BinClass<DEX_OP_MUL> { static constexpr int v = CL_VAR; };
template <> struct BinClass<DEX_OP_MAX> { static constexpr int v = CL_VAR; };
template <> struct BinClass<DEX_OP_MIN> { static constexpr int v = CL_VAR; };
template <int OPC> struct UnClass { static constexpr int v = UL_GEN; };
template <> struct UnClass<DEX_OP_NEG> { static constexpr int v = UL_NEG; };
#define UVAR(S) template <> struct UnClass<DEX_OP_##S> { static constexpr int v = UL_VAR; };
UVAR(ABS) UVAR(SQUARE) UVAR(CUBE) UVAR(EXP) UVAR(SIN) UVAR(COS) UVAR(TANH) UVAR(RELU)
#undef UVAR

// dense index over the operators with an unrolled stage-1 code path; everything else is
// FO_GENERIC.  The handler id of a tape instruction (operator x operand forms) maps to it.
enum : int {
    FO_GENERIC = 0,
    FO_LOAD,
#define X(S) FO_##S,
    DEX_FAST_UNARY(X) DEX_FAST_BIN_COMM(X) DEX_FAST_BIN_NC(X)
#undef X
    FO__COUNT
};
struct FastOpTable { uint8_t v[64]; };
constexpr FastOpTable make_fast_op_table() {
    FastOpTable t{};
    t.v[H_LOAD_R] = FO_LOAD;
    t.v[H_LOAD_C] = FO_LOAD;
    t.v[H_KEEP] = FO_LOAD;      // identity on ACC (value and derivatives unchanged)
#define X(S) t.v[H_##S##_A] = FO_##S; t.v[H_##S##_R] = FO_##S;
    DEX_FAST_UNARY(X)
#undef X
#define X(S) t.v[H_##S##_AR] = FO_##S; t.v[H_##S##_AC] = FO_##S; t.v[H_##S##_RR] = FO_##S; t.v[H_##S##_RC] = FO_##S;
    DEX_FAST_BIN_COMM(X)
#undef X
#define X(S)                                                                                         \
    t.v[H_##S##_AR] = FO_##S; t.v[H_##S##_RA] = FO_##S; t.v[H_##S##_AC] = FO_##S; t.v[H_##S##_CA] = FO_##S; \
    t.v[H_##S##_RR] = FO_##S; t.v[H_##S##_RC] = FO_##S; t.v[H_##S##_CR] = FO_##S;
    DEX_FAST_BIN_NC(X)
#undef X
    return t;
}
__constant__ FastOpTable c_fast_op_table = make_fast_op_table();

// ---- stage 2 ----------------------------------------------------------------------------------
// dA[g] of a densely evaluated operand
template <typename T, int GC, int U, int KIND>
__device__ __forceinline__ void dsrc(T (&x)[VK<T, U>::K], const T (&adg)[VK<T, U>::K], const T* slot, int g,
                                     int idx, int TILE, int CS) {
    constexpr int K = VK<T, U>::K;
    if (KIND == DK_ACC) {
#pragma unroll
        for (int k = 0; k < K; ++k) x[k] = adg[k];
    } else if (KIND == DK_SLOT) {
        VK<T, U>::ld(x, slot + (size_t)(1 + g) * TILE, CS);
    } else {
        const T e = (g == idx) ? T(1) : T(0);
#pragma unroll
        for (int k = 0; k < K; ++k) x[k] = e;
    }
}

template <typename T, int GC, int U, int CLS, int KA, int KB>
__device__ __forceinline__ void combine_bin(T (&ad)[GC][VK<T, U>::K], const T (&p0)[VK<T, U>::K],
                                            const T (&p1)[VK<T, U>::K], const T* sa, const T* sb, int ia,
                                            int ib, int TILE, int CS) {
    constexpr int K = VK<T, U>::K;
    using A = VA<T, K>;
    constexpr bool spA = (KA == DK_LEAF) && (CLS != CL_GEN);   // one-hot operand, skippable zeros
    constexpr bool spB = (KB == DK_LEAF) && (CLS != CL_GEN);
#pragma unroll
    for (int g = 0; g < GC; ++g) {
        T a[K], b[K], d[K];
        if (!spA) dsrc<T, GC, U, KA>(a, ad[g], sa, g, ia, TILE, CS);
        if (!spB) dsrc<T, GC, U, KB>(b, ad[g], sb, g, ib, TILE, CS);
        if (CLS == CL_ADD) {
            if (!spA && !spB) A::add(d, a, b);
            else if (!spA) { A::adds(d, a, (g == ib) ? T(1) : T(0)); }
            else if (!spB) { A::adds(d, b, (g == ia) ? T(1) : T(0)); }
            else {
                const T e = ((g == ia) ? T(1) : T(0)) + ((g == ib) ? T(1) : T(0));
#pragma unroll
                for (int k = 0; k < K; ++k) d[k] = e;
            }
        } else if (CLS == CL_SUB) {
            if (!spA && !spB) A::sub(d, a, b);
            else if (!spA) { A::adds(d, a, (g == ib) ? T(-1) : T(0)); }
            else if (!spB) { A::neg(d, b); A::adds(d, d, (g == ia) ? T(1) : T(0)); }
            else {
                const T e = ((g == ia) ? T(1) : T(0)) - ((g == ib) ? T(1) : T(0));
#pragma unroll
                for (int k = 0; k < K; ++k) d[k] = e;
            }
        } else if (CLS == CL_VAR) {
            if (!spA && !spB) { T t[K]; A::mul(t, p1, b); A::fma(d, p0, a, t); }
            else if (!spA) { A::mul(d, p0, a); if (g == ib) A::add(d, d, p1); }
            else if (!spB) { A::mul(d, p1, b); if (g == ia) A::add(d, p0, d); }
            else {
#pragma unroll
                for (int k = 0; k < K; ++k) d[k] = T(0);
                if (g == ia) A::add(d, d, p0);
                if (g == ib) A::add(d, d, p1);
            }
        } else {   // CL_GEN: every product is formed, also with the zeros of a one-hot
            T t[K];
            if (KA == DK_LEAF && KB != DK_LEAF) { A::mul(t, p0, a); A::fma(d, p1, b, t); }
            else { A::mul(t, p1, b); A::fma(d, p0, a, t); }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) ad[g][k] = d[k];
    }
}

template <typename T, int GC, int U, int CLS, int KA>
__device__ __forceinline__ void combine_un(T (&ad)[GC][VK<T, U>::K], const T (&p0)[VK<T, U>::K], const T* sa,
                                           int ia, int TILE, int CS) {
    constexpr int K = VK<T, U>::K;
    using A = VA<T, K>;
#pragma unroll
    for (int g = 0; g < GC; ++g) {
        T a[K], d[K];
        if (KA == DK_LEAF && CLS != UL_GEN) {
            const bool hit = (g == ia);
#pragma unroll
            for (int k = 0; k < K; ++k)
                d[k] = CLS == UL_ONE ? (hit ? T(1) : T(0)) : CLS == UL_NEG ? (hit ? T(-1) : T(0)) : (hit ? p0[k] : T(0));
        } else {
            dsrc<T, GC, U, KA>(a, ad[g], sa, g, ia, TILE, CS);
            if (CLS == UL_ONE) {
#pragma unroll
                for (int k = 0; k < K; ++k) d[k] = a[k];
            } else if (CLS == UL_NEG) A::neg(d, a);
            else A::mul(d, p0, a);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) ad[g][k] = d[k];
    }
}

#if DEX_GRAD_PTX
#ifndef DEX_GRAD_INC
#define DEX_GRAD_INC "dex_grad_f32.inc"     // generated for DEX_GRAD_THREADS threads per CTA
#endif
#include DEX_GRAD_INC
// runs the generated loop for (GC, U, NT) when it exists; otherwise leaves pc untouched and the
// C++ `step` executes the whole tape
template <typename T, int GC, int U, int NT, typename... A>
__device__ __forceinline__ void ptx_loop(A&&... a) {
    if constexpr (sizeof(T) == 4) {
        if constexpr (GradLoopF32<GC, U, NT>::exists) GradLoopF32<GC, U, NT>::run(static_cast<A&&>(a)...);
    } else {
        if constexpr (GradLoopF64<GC, U, NT>::exists) GradLoopF64<GC, U, NT>::run(static_cast<A&&>(a)...);
    }
}
#endif

// Kernel modes.  KM_GRAD: eval_grad_tree_array, value rows and (G x N) gradient blocks are stored.
// KM_DIFF: eval_diff_tree_array (one direction, no validity checks, GEN class everywhere).
// KM_LOSS: fused weighted squared-error loss and its gradient (what constant optimisation
// consumes: sum_j w_j (tree(x_j) - y_j)^2 and d/d theta of it, the contraction of the
// reference's pullback, /root/reference/src/ChainRules.jl:56-77) — neither the value rows nor
// the (G x N) gradient ever leave the SM; per tile one double per tree and gradient entry.
enum : int { KM_GRAD = 0, KM_DIFF = 1, KM_LOSS = 2 };

template <typename T, int GC, int U, int KMODE>
__global__ void __launch_bounds__(DEX_GRAD_THREADS, DEX_GRAD_MIN_CTAS) grad_kernel(const GK<T> a) {
    constexpr bool DIFF = KMODE == KM_DIFF;
    constexpr bool LOSS = KMODE == KM_LOSS;
    using V = VK<T, U>;
    using A = VA<T, V::K>;
    constexpr int C = V::C;
    constexpr int K = V::K;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* rows = reinterpret_cast<T*>(smem_raw);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int TILE = nthr * K;
    const int CS = nthr * C;
    const int S = a.max_stack;
    T* xs = rows + (size_t)S * (1 + GC) * TILE;   // feature rows
    const int64_t s0 = (int64_t)blockIdx.x * TILE;

    // stage the feature rows of this tile (XT is tile-padded: always in range, 16 B aligned);
    // they sit behind the parameter rows, which every tree fills for itself
    const int NPR = a.n_param_rows;
    for (int idx = tid; idx < (a.F - NPR) * (TILE / C); idx += nthr) {
        const int f = idx / (TILE / C), q = idx - f * (TILE / C);
        *reinterpret_cast<uint4*>(xs + (size_t)(NPR + f) * TILE + q * C) =
            __ldg(reinterpret_cast<const uint4*>(a.X + (size_t)f * a.ldx + s0 + q * C));
    }
    __syncthreads();

    T* my = rows + tid * C;
    const T* myx = xs + tid * C;
    // fused loss: this thread's targets and weights (0 for the padded tail of the last tile)
    T yv[LOSS ? K : 1], wv[LOSS ? K : 1];
    __shared__ double red[LOSS ? 2 : 1][LOSS ? DEX_GRAD_THREADS / 32 : 1][LOSS ? GC + 1 : 1];
    int red_buf = 0;
    if (LOSS) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            int64_t gs = s0 + (int64_t)(k / C) * CS + (int64_t)tid * C + (k % C);
            const bool in = gs < a.N;
            if (!in) gs = a.N - 1;
            yv[k] = __ldg(a.y + gs);
            wv[k] = in ? (a.w ? __ldg(a.w + gs) : T(1)) : T(0);
        }
    }
    const int t0 = a.chunk_start[blockIdx.y], t1 = a.chunk_start[blockIdx.y + 1];
    const int mode = a.mode;
    const bool full_tile = (s0 + TILE <= a.N);
    const int NEVER = -(1 << 20);   // one-hot index that matches no direction
    // tree-invariant conditions, hoisted out of the tree loop
    const bool out_vec = !LOSS && full_tile && (a.ldo % C) == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0;
    const bool grad_al = !LOSS && !DIFF && full_tile && (reinterpret_cast<uintptr_t>(a.grad) & 15) == 0;
    const bool mode_feat = mode == DEX_GRAD_FEATURES, mode_const = mode == DEX_GRAD_CONSTANTS;
    const int F = a.F;

    // Per-tree metadata is loaded one tree ahead (the loads of tree t + 1 are issued while tree t
    // runs) and the first instruction of a tree arrives as the prefetch of its predecessor's last
    // one (tapes are contiguous), so no tree starts with a chain of dependent global loads.
    int64_t off = a.tape_off[t0], off_next = a.tape_off[t0 + 1];
    int64_t co = a.const_off[t0], co_next = a.const_off[t0 + 1];
    int64_t go = (!DIFF && a.grad_off) ? a.grad_off[t0] : 0;
    uint4 ins = __ldg(a.tape + off);
    for (int t = t0; t < t1; ++t) {
        const int n = (int)(off_next - off);
        const uint4* ip = a.tape + off;
        const int32_t* ordp = a.const_ord + off;
        const int nconst = (int)(co_next - co);
        const int64_t goff = go;
        // next tree (the tables carry slack past their last entry, dex_api.cu upload())
        const int64_t off_next2 = a.tape_off[t + 2];
        const int64_t co_next2 = a.const_off[t + 2];
        if (!DIFF && a.grad_off) go = a.grad_off[t + 1 < (int)a.n_trees ? t + 1 : t];
        const int G = DIFF ? 1 : mode_feat ? F : mode_const ? nconst : F + nconst;
        T nf = T(0);
        if (NPR > 0) {
            // ParametricExpression: this tree's per-sample parameters parameters[p, classes[j]]
            // (src/ParametricExpression.jl:380-384) into the parameter rows; each thread fills and
            // later reads only its own columns, so no barrier
            const T* ptree = a.params + (size_t)t * a.n_params * a.n_classes;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                int64_t gs = s0 + (int64_t)(k / C) * CS + (int64_t)tid * C + (k % C);
                if (gs >= a.N) gs = a.N - 1;
                const int cl = __ldg(a.classes + gs) * a.n_params;
                for (int p = 0; p < NPR; ++p)
                    const_cast<T*>(myx)[(size_t)p * TILE + (k / C) * CS + (k % C)] = __ldg(ptree + cl + p);
            }
        }
        const int npass = G <= GC ? 1 : (G + GC - 1) / GC;
        for (int pass = 0; pass < npass; ++pass) {
            const int g0 = pass * GC;
            // one-hot index of feature f: f + foff; of the constant with ordinal o: o + coff
            const int foff = DIFF ? -a.direction : (mode_const ? NEVER : -g0);
            const int coff = DIFF ? NEVER : mode_const ? -g0 : mode_feat ? NEVER : F - g0;
            T av[K], ad[GC][K];   // accumulator dual
#pragma unroll
            for (int k = 0; k < K; ++k) av[k] = T(0);
#pragma unroll
            for (int g = 0; g < GC; ++g)
#pragma unroll
                for (int k = 0; k < K; ++k) ad[g][k] = T(0);

            // one tape instruction, C++ form (every handler, generic operators included)
            auto step = [&](const uint4& ins, const int pc) {
                const uint32_t w0 = ins.x;
                const T c = gconst_of<T>(ins);
                if (w0 & F_PUSH) {
                    T* dst = my + (size_t)push_row(w0) * (1 + GC) * TILE;
                    V::st(dst, CS, av);
#pragma unroll
                    for (int g = 0; g < GC; ++g) V::st(dst + (size_t)(1 + g) * TILE, CS, ad[g]);
                }
                T x[K], y[K], vo[K], p0[K], p1[K];
                const T* sa = my;
                const T* sb = my;
                int ia = NEVER, ib = NEVER, ka = DK_LEAF, kb = DK_LEAF, sel = SEL_NONE;
                const uint32_t op = (w0 >> 8) & 0xffu;

                // ---- stage 0: operand values; where their derivatives live -------------------
                // ROW is a stack slot (dual in shared memory) or a feature leaf (one-hot
                // derivative, value checked: grad_deg0_eval :387-399)
                auto fetch = [&](uint32_t src, uint32_t row, T (&v)[K], const T*& sp, int& idx, int& kind) {
                    if (src == SRC_ACC) {
#pragma unroll
                        for (int k = 0; k < K; ++k) v[k] = av[k];
                        kind = DK_ACC;
                    } else if (src == SRC_ROW) {
                        if ((int)row < S) {
                            sp = my + (size_t)row * (1 + GC) * TILE;
                            V::ld(v, sp, CS);
                            kind = DK_SLOT;
                        } else {
                            const int f = (int)row - S;
                            V::ld(v, myx + (size_t)f * TILE, CS);
                            idx = f + foff;
                            if (!DIFF) A::check(nf, v);
                        }
                    } else {   // inline constant
#pragma unroll
                        for (int k = 0; k < K; ++k) v[k] = c;
                        if (!DIFF) {
                            nf = m_fma(c, T(0), nf);
                            if (mode != DEX_GRAD_FEATURES) idx = __ldg(ordp + pc) + coff;
                        }
                    }
                };
                fetch((w0 >> 16) & 3u, row_a(ins.y), x, sa, ia, ka);
                if (op >= 64u) fetch((w0 >> 18) & 3u, row_b(ins.y), y, sb, ib, kb);

                // ---- stage 1: value and partials, by operator ----------------------------------
                // c_fast_op maps the handler id (operator x operand forms, csrc/dex_tape.h) to a
                // dense index over the operators that have an unrolled code path here
                switch ((int)c_fast_op_table.v[w0 & HANDLER_MASK]) {
                    case FO_LOAD: {
#pragma unroll
                        for (int k = 0; k < K; ++k) { vo[k] = x[k]; p0[k] = T(1); }
                        sel = SEL_UNARY + (DIFF ? UL_GEN : UL_ONE) * 3 + ka;
                    } break;
#define UN_CASE(S)                                                                          \
    case FO_##S: {                                                                          \
        _Pragma("unroll") for (int k = 0; k < K; ++k) G1<DEX_OP_##S, T>::f(x[k], vo[k], p0[k]); \
        sel = SEL_UNARY + (DIFF ? UL_GEN : UnClass<DEX_OP_##S>::v) * 3 + ka;                \
    } break;
                    DEX_FAST_UNARY(UN_CASE)
#undef UN_CASE
#define BIN_CASE(S)                                                                         \
    case FO_##S: {                                                                          \
        _Pragma("unroll") for (int k = 0; k < K; ++k) G2<DEX_OP_##S, T>::f(x[k], y[k], vo[k], p0[k], p1[k]); \
        sel = (DIFF ? CL_GEN : BinClass<DEX_OP_##S>::v) * 9 + ka * 3 + kb;                  \
    } break;
                    DEX_FAST_BIN_COMM(BIN_CASE)
                    DEX_FAST_BIN_NC(BIN_CASE)
#undef BIN_CASE
                    // ---- generic: any operator -------------------------------------------------
                    default: {
                        const int deg = op >= 128u ? 3 : (op >= 64u ? 2 : 1);
                        T p2[K];
                        if (op == (uint32_t)DEX_OP_POW) {
                            // `^` is the one generic operator that symbolic-regression operator sets
                            // use all the time: unrolled, in registers (expressions of dex_ops.cuh)
#pragma unroll
                            for (int k = 0; k < K; ++k) {
                                const T xx = x[k], yy = y[k];
                                const T v = m_pow(xx, yy);
                                vo[k] = v;
                                p0[k] = yy * m_pow(xx, yy - T(1));
                                p1[k] = (xx == T(0) && yy > T(0)) ? T(0) : v * m_log(m_fabs(xx));
                                p2[k] = T(0);
                            }
                        } else {
                            // rolled over the samples, operands in local memory: this rarely taken
                            // path stays small and keeps the hot code in the instruction cache
                            T lx[K], ly[K], lz[K], lv[K], l0[K], l1[K], l2[K];
#pragma unroll
                            for (int k = 0; k < K; ++k) { lx[k] = x[k]; ly[k] = deg >= 2 ? y[k] : T(0); lz[k] = av[k]; }
#pragma unroll 1
                            for (int k = 0; k < K; ++k) {
                                const T xx = lx[k], yy = ly[k], zz = lz[k];
                                T v_, q0 = T(0), q1 = T(0), q2 = T(0);
                                switch (op) {
#define U_CASE(SYM, VEXPR, GEXPR) \
    case DEX_OP_##SYM: { const T x = xx; const T v = (VEXPR); q0 = (GEXPR); v_ = v; } break;
                                    DEX_UNARY_OPS(U_CASE)
#undef U_CASE
#define B_CASE(SYM, VEXPR, GA, GB) \
    case DEX_OP_##SYM: { const T x = xx, y = yy; const T v = (VEXPR); q0 = (GA); q1 = (GB); v_ = v; } break;
                                    DEX_BINARY_OPS(B_CASE)
#undef B_CASE
#define T_CASE(SYM, VEXPR, GA, GB, GZ) \
    case DEX_OP_##SYM: { const T x = xx, y = yy, z = zz; const T v = (VEXPR); q0 = (GA); q1 = (GB); q2 = (GZ); v_ = v; (void)z; } break;
                                    DEX_TERNARY_OPS(T_CASE)
#undef T_CASE
                                    default: v_ = q0 = q1 = q2 = t_nan<T>(); break;
                                }
                                lv[k] = v_; l0[k] = q0; l1[k] = q1; l2[k] = q2;
                            }
#pragma unroll
                            for (int k = 0; k < K; ++k) { vo[k] = lv[k]; p0[k] = l0[k]; p1[k] = l1[k]; p2[k] = l2[k]; }
                        }
                        if (deg == 1) sel = SEL_UNARY + UL_GEN * 3 + ka;
                        else if (deg == 2) sel = CL_GEN * 9 + ka * 3 + kb;
                        else {
                            // ternary: d = p0 dA + p1 dB + p2 dACC (third operand is always ACC)
#pragma unroll
                            for (int g = 0; g < GC; ++g) {
                                T da[K], db[K], d[K], tt[K];
                                if (ka == DK_ACC) dsrc<T, GC, U, DK_ACC>(da, ad[g], sa, g, ia, TILE, CS);
                                else if (ka == DK_SLOT) dsrc<T, GC, U, DK_SLOT>(da, ad[g], sa, g, ia, TILE, CS);
                                else dsrc<T, GC, U, DK_LEAF>(da, ad[g], sa, g, ia, TILE, CS);
                                if (kb == DK_ACC) dsrc<T, GC, U, DK_ACC>(db, ad[g], sb, g, ib, TILE, CS);
                                else if (kb == DK_SLOT) dsrc<T, GC, U, DK_SLOT>(db, ad[g], sb, g, ib, TILE, CS);
                                else dsrc<T, GC, U, DK_LEAF>(db, ad[g], sb, g, ib, TILE, CS);
                                A::mul(d, p0, da);
                                A::mul(tt, p1, db);
                                A::add(d, d, tt);
                                A::mul(tt, p2, ad[g]);
                                A::add(d, d, tt);
#pragma unroll
                                for (int k = 0; k < K; ++k) ad[g][k] = d[k];
                            }
                            sel = SEL_NONE;
                        }
                    } break;
                }

                // max / min whose operands the flattener exchanged: the reference's partials
                // (x > y, !(x > y)) give a tie to the SECOND operand of the original order, which
                // is operand A here (/root/reference/src/EvaluateDerivative.jl:340-365 via Zygote)
                if (w0 & F_SWAPPED) {
                    const bool is_max = op == (uint32_t)DEX_OP_MAX;
#pragma unroll
                    for (int k = 0; k < K; ++k)
                        if (x[k] == y[k]) { p0[k] = is_max ? T(1) : T(0); p1[k] = is_max ? T(0) : T(1); }
                }

                // ---- stage 2: derivative combination ---------------------------------------------
                switch (sel) {
#define CB(CLS, KA, KB) \
    case (CLS) * 9 + (KA) * 3 + (KB): combine_bin<T, GC, U, CLS, KA, KB>(ad, p0, p1, sa, sb, ia, ib, TILE, CS); break;
#define CB_ALL(CLS)                                                                                   \
    CB(CLS, DK_ACC, DK_SLOT) CB(CLS, DK_ACC, DK_LEAF) CB(CLS, DK_SLOT, DK_ACC) CB(CLS, DK_SLOT, DK_SLOT) \
    CB(CLS, DK_SLOT, DK_LEAF) CB(CLS, DK_LEAF, DK_ACC) CB(CLS, DK_LEAF, DK_SLOT) CB(CLS, DK_LEAF, DK_LEAF)
                    CB_ALL(CL_ADD) CB_ALL(CL_SUB) CB_ALL(CL_VAR) CB_ALL(CL_GEN)
                    CB(CL_GEN, DK_ACC, DK_ACC)   // generic path only (never emitted by the flattener)
#undef CB_ALL
#undef CB
#define CU1(CLS, KA) \
    case SEL_UNARY + (CLS) * 3 + (KA): combine_un<T, GC, U, CLS, KA>(ad, p0, sa, ia, TILE, CS); break;
#define CU_ALL(CLS) CU1(CLS, DK_ACC) CU1(CLS, DK_SLOT) CU1(CLS, DK_LEAF)
                    CU_ALL(UL_ONE) CU_ALL(UL_NEG) CU_ALL(UL_VAR) CU_ALL(UL_GEN)
#undef CU_ALL
#undef CU1
                    default: break;
                }
#pragma unroll
                for (int k = 0; k < K; ++k) av[k] = vo[k];
                if (!DIFF) A::check(nf, vo);
            };
            // single call site of `step` (so that it is inlined and the dual accumulator stays in
            // registers): the Float32 256- and 128-thread launches run the instruction loop as one inline-PTX
            // block with jump-table dispatch (gen_grad_ptx.py), which returns at the end of the tape
            // or at the first instruction it does not implement natively; `step` executes that one.
            // (Float64: the same generated structure with one double per 64-bit register; + - * / neg
            // square cube inv sqrt native, the rest handed to `step`)
            constexpr bool HAS_PTX = DEX_GRAD_PTX && !DIFF;
            if (pass > 0) ins = __ldg(ip);
            int pc = 0;
            while (pc < n) {
#if DEX_GRAD_PTX
                if constexpr (HAS_PTX) {
                    // the loops are generated per CTA size (row strides are immediates)
                    if (nthr == 256 || nthr == 128) {
                        const uint32_t my_s = (uint32_t)__cvta_generic_to_shared(my);
                        T nfa[2] = {nf, T(0)};
                        if (nthr == 256)
                            ptx_loop<T, GC, U, 256>(pc, av, ad, nfa, ins, ip, n, my_s, S, S * GC, foff - S, coff, ordp,
                                                    mode != DEX_GRAD_FEATURES ? 1 : 0);
                        else
                            ptx_loop<T, GC, U, 128>(pc, av, ad, nfa, ins, ip, n, my_s, S, S * GC, foff - S, coff, ordp,
                                                    mode != DEX_GRAD_FEATURES ? 1 : 0);
                        nf = nfa[0] + nfa[1];
                        if (pc < n) ins = __ldg(ip + pc);   // early exit: `ins` is two instructions ahead
                    }
                }
#endif
                if (pc < n) {
                    const uint4 nxt = __ldg(ip + pc + 1);
                    step(ins, pc);
                    ins = nxt;
                    ++pc;
                }
            }
            // root derivative check (see the header comment: non-finite components propagate)
            if (!DIFF) {   // padded directions (g0 + g >= G) are not part of the gradient
                T racc[2] = {nf, T(0)};
                if (G - g0 >= GC) {
#pragma unroll
                    for (int g = 0; g < GC; ++g)
#pragma unroll
                        for (int k = 0; k < K; ++k) racc[k & 1] = m_fma(ad[g][k], T(0), racc[k & 1]);
                } else {
#pragma unroll
                    for (int g = 0; g < GC; ++g)
                        if (g0 + g < G) {
#pragma unroll
                            for (int k = 0; k < K; ++k) racc[k & 1] = m_fma(ad[g][k], T(0), racc[k & 1]);
                        }
                }
                nf = racc[0] + racc[1];
            }
            if constexpr (LOSS) {
                // ---- fused loss: r_j = 2 w_j (v_j - y_j);  loss += w_j (v_j - y_j)^2;
                //      dloss/dtheta_g += r_j * d_j[g]   — reduced over the tile, in double -----------
                double acc[GC + 1];
#pragma unroll
                for (int g = 0; g <= GC; ++g) acc[g] = 0.0;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const double d = (double)av[k] - (double)yv[k];
                    const double wd = (double)wv[k] * d;
                    acc[GC] += wd * d;
#pragma unroll
                    for (int g = 0; g < GC; ++g) acc[g] += (2.0 * wd) * (double)ad[g][k];
                }
#pragma unroll
                for (int g = 0; g <= GC; ++g)
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc[g] += __shfl_xor_sync(0xffffffffu, acc[g], o);
                if ((tid & 31) == 0) {
#pragma unroll
                    for (int g = 0; g <= GC; ++g) red[red_buf][tid >> 5][g] = acc[g];
                }
                __syncthreads();   // one barrier per (tree, pass): red[] is double buffered
                if (tid <= GC) {
                    double sum = 0.0;
                    for (int wi = 0; wi < (nthr >> 5); ++wi) sum += red[red_buf][wi][tid];
                    double* prow = a.partial + (size_t)blockIdx.x * a.partial_stride;
                    if (tid == GC) {
                        if (pass == 0) prow[t] = sum;
                    } else if (g0 + tid < G) {
                        prow[a.n_trees + goff + g0 + tid] = sum;
                    }
                }
                red_buf ^= 1;
            } else {
            // ---- outputs of this pass -----------------------------------------------------
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t sbase = s0 + (int64_t)u * CS + (int64_t)tid * C;
                if (pass == 0) {
                    T* o = a.out + (size_t)t * a.ldo + sbase;
                    if (out_vec)
                        __stcs(reinterpret_cast<uint4*>(o), *reinterpret_cast<const uint4*>(&av[u * C]));
                    else {
#pragma unroll
                        for (int k = 0; k < C; ++k) if (sbase + k < a.N) o[k] = av[u * C + k];
                    }
                }
                if (DIFF) {   // eval_diff: one derivative row per tree, laid out like `out`
                    T* o = a.grad + (size_t)t * a.ldo + sbase;
#pragma unroll
                    for (int k = 0; k < C; ++k) if (sbase + k < a.N) o[k] = ad[0][u * C + k];
                } else if (G > 0) {
                    // (G x N) column-major block: element (g, s) at s * G + g
                    T* gout = a.grad + goff + sbase * G + g0;
                    const int gc = min(GC, G - g0);
                    // 16-byte alignment of gout given an aligned base: (goff + sbase * G + g0) % C == 0,
                    // and sbase is a multiple of C, g0 = 0 when G == GC
                    if (G == GC && grad_al && (goff % C) == 0) {
                        // the thread's C samples x GC directions are C*GC contiguous elements
                        T flat[C * GC];
#pragma unroll
                        for (int k = 0; k < C; ++k)
#pragma unroll
                            for (int g = 0; g < GC; ++g) flat[k * GC + g] = ad[g][u * C + k];
#pragma unroll
                        for (int q = 0; q < C * GC; q += C)      // streaming: 1.5 GB per C3 launch, never re-read
                            __stcs(reinterpret_cast<uint4*>(gout + q), *reinterpret_cast<const uint4*>(flat + q));
                    } else {
#pragma unroll
                        for (int k = 0; k < C; ++k)
                            if (sbase + k < a.N) {
#pragma unroll
                                for (int g = 0; g < GC; ++g)
                                    if (g < gc) gout[(size_t)k * G + g] = ad[g][u * C + k];
                            }
                    }
                }
            }
        }
            }
        if (!DIFF) {
            const bool bad = nf != nf;
            if (__any_sync(0xffffffffu, bad) && (tid & 31) == 0) a.ok[t] = 0;
        }
        off = off_next; off_next = off_next2;
        co = co_next; co_next = co_next2;
    }
}

// Pre-pass: feature-major tile-padded copy of X; presets ok[]; when the launch runs the
// constant-folded tape (seg_off != null) the first n_trees threads also evaluate the folded
// subtrees of one tree each, with the gradient path's validity rule (dex_fold.cuh).
template <typename T>
__global__ void gtranspose_pad_kernel(const T* __restrict__ X, int64_t ldx, int F, int64_t N,
                                      T* __restrict__ XT, int64_t Npad, uint8_t* ok, int64_t n_trees,
                                      Instr* tape, const Instr* ctape, const int64_t* seg,
                                      const int64_t* seg_off, const uint8_t* fold_ok) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < Npad) {
        const T* col = X + (s < N ? s : N - 1) * ldx;
        for (int f = 0; f < F; ++f) XT[(size_t)f * Npad + s] = __ldg(col + f);
    }
    if (s < n_trees)
        ok[s] = fold_ok ? fold_ok[s] : ((!seg_off || fold_tree<T, true>(tape, ctape, seg, seg_off, s)) ? 1 : 0);
}

constexpr size_t G_SMEM_LIMIT = 227 * 1024;
constexpr int GRAD_U = DEX_GRAD_U;
constexpr int GRAD_U64 = DEX_GRAD_U64;

struct GradShape { int threads; int GC; size_t smem; int64_t tile; };

GradShape pick_shape(int dtype, int F, int max_stack, int Gmax, bool loss = false) {
    (void)loss;
    const size_t es = dtype == DEX_F32 ? 4 : 8;
    const int K = dtype == DEX_F32 ? 4 * GRAD_U : 2 * GRAD_U64;
    GradShape s;
    s.threads = DEX_GRAD_THREADS;
    s.GC = std::max(1, std::min(Gmax, 8));
    if (s.GC == 7) s.GC = 8;   // instantiated: 1..6, 8
    auto bytes = [&](int th, int gc) {
        return ((size_t)max_stack * (1 + gc) + (size_t)F) * (size_t)th * K * es;
    };
    while (bytes(s.threads, s.GC) > DEX_GRAD_SMEM_SOFT && s.threads > 32) s.threads >>= 1;
    while (bytes(s.threads, s.GC) > G_SMEM_LIMIT && s.GC > 1) s.GC = s.GC > 4 ? 4 : s.GC > 2 ? 2 : 1;   // all instantiated
    s.smem = std::max<size_t>(bytes(s.threads, s.GC), 16);
    s.tile = (int64_t)s.threads * K;
    return s;
}

template <typename T, int GC, int KMODE>
cudaError_t launch_one(const GK<T>& a, const GradShape& sh, int64_t n_tiles, int n_chunks, cudaStream_t stream) {
    auto kern = grad_kernel<T, GC, (sizeof(T) == 8 ? GRAD_U64 : GRAD_U), KMODE>;
    cudaError_t err = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), sh.smem);
    if (err != cudaSuccess) return err;
    dim3 grid((unsigned)n_tiles, (unsigned)n_chunks);
    kern<<<grid, sh.threads, sh.smem, stream>>>(a);
    return cudaGetLastError();
}

template <typename T>
GK<T> make_args(const GradArgs& g, const int32_t* chunk_start, int64_t Npad) {
    GK<T> a;
    a.tape = reinterpret_cast<const uint4*>(g.tape);
    a.tape_off = g.tape_off;
    a.const_ord = g.const_ord;
    a.const_off = g.const_off;
    a.chunk_start = chunk_start;
    a.X = static_cast<const T*>(g.xt);
    a.out = static_cast<T*>(g.out);
    a.grad = static_cast<T*>(g.grad);
    a.grad_off = g.grad_off;
    a.ok = g.ok;
    a.N = g.N; a.ldx = Npad; a.ldo = g.ldo; a.n_trees = g.n_trees;
    a.F = g.F; a.max_stack = g.max_stack; a.mode = g.mode; a.direction = g.direction;
    a.y = static_cast<const T*>(g.y); a.w = static_cast<const T*>(g.w);
    a.partial = g.partial; a.partial_stride = g.partial_stride;
    a.params = static_cast<const T*>(g.params); a.classes = g.classes;
    a.n_params = g.n_params; a.n_classes = g.n_classes; a.n_param_rows = g.n_param_rows;
    return a;
}

// second stage of the fused loss: out[i] = scale * sum over tiles of partial[tile][i], in tile order
// (deterministic); scale = 1 / sum of weights (read from *wsum when given) or 1 / N
__global__ void loss_grad_reduce_kernel(const double* partial, int64_t n_tiles, int64_t stride, int64_t n_trees,
                                        double inv_n, const double* wsum, double* loss, double* grad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= stride) return;
    double s = 0.0;
    for (int64_t tl = 0; tl < n_tiles; ++tl) s += partial[tl * stride + i];
    const double scale = wsum ? 1.0 / *wsum : inv_n;
    if (i < n_trees) loss[i] = s * scale;
    else grad[i - n_trees] = s * scale;
}

// sum of the weights, one block, fixed order (deterministic)
template <typename T>
__global__ void weight_sum_kernel(const T* w, int64_t n, double* out) {
    __shared__ double sh[256];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 256) s += (double)w[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sh[0];
}

}  // namespace

cudaError_t launch_loss_grad_reduce(const double* partial, int64_t n_tiles, int64_t stride, int64_t n_trees,
                                    double inv_n, const double* wsum, double* loss, double* grad,
                                    cudaStream_t stream) {
    if (stride == 0) return cudaSuccess;
    loss_grad_reduce_kernel<<<(unsigned)((stride + 255) / 256), 256, 0, stream>>>(partial, n_tiles, stride, n_trees,
                                                                               inv_n, wsum, loss, grad);
    return cudaGetLastError();
}

cudaError_t launch_weight_sum(int dtype, const void* w, int64_t n, double* out, cudaStream_t stream) {
    if (dtype == DEX_F32) weight_sum_kernel<float><<<1, 256, 0, stream>>>(static_cast<const float*>(w), n, out);
    else weight_sum_kernel<double><<<1, 256, 0, stream>>>(static_cast<const double*>(w), n, out);
    return cudaGetLastError();
}

size_t grad_xt_bytes(int dtype, int F, int max_stack, int Gmax, int64_t N, bool loss) {
    const GradShape sh = pick_shape(dtype, F, max_stack, std::max(Gmax, 1), loss);
    const int64_t n_tiles = (N + sh.tile - 1) / sh.tile;
    return (size_t)std::max<int64_t>(n_tiles * sh.tile, 1) * (size_t)std::max(F, 1) * (dtype == DEX_F32 ? 4 : 8);
}

int64_t grad_num_tiles(int dtype, int F, int max_stack, int Gmax, int64_t N, bool loss) {
    const GradShape sh = pick_shape(dtype, F, max_stack, std::max(Gmax, 1), loss);
    return (N + sh.tile - 1) / sh.tile;
}

cudaError_t launch_grad_ex(const GradArgs& g, const int32_t* chunk_start, int n_chunks, int Gmax,
                           cudaStream_t stream, int* launches) {
    if (g.n_trees == 0 || g.N == 0) return cudaSuccess;
    const bool loss = g.partial != nullptr;
    const GradShape sh = pick_shape(g.dtype, g.F, g.max_stack, std::max(Gmax, 1), loss);
    if (sh.smem > G_SMEM_LIMIT) return cudaErrorInvalidConfiguration;
    const int64_t n_tiles = (g.N + sh.tile - 1) / sh.tile;
    const int64_t Npad = n_tiles * sh.tile;
    const int64_t cover = std::max<int64_t>(Npad, g.n_trees);
    const int FX = g.F - g.n_param_rows;     // rows of the caller's X
    if (g.dtype == DEX_F32)
        gtranspose_pad_kernel<float><<<(unsigned)((cover + 255) / 256), 256, 0, stream>>>(
            static_cast<const float*>(g.X), g.ldx, FX, g.N, static_cast<float*>(g.xt), Npad, g.ok, g.n_trees,
            const_cast<Instr*>(g.tape), g.ctape, g.seg, g.seg_off, g.fold_ok);
    else
        gtranspose_pad_kernel<double><<<(unsigned)((cover + 255) / 256), 256, 0, stream>>>(
            static_cast<const double*>(g.X), g.ldx, FX, g.N, static_cast<double*>(g.xt), Npad, g.ok, g.n_trees,
            const_cast<Instr*>(g.tape), g.ctape, g.seg, g.seg_off, g.fold_ok);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    if (launches) *launches += 1;
    const bool diff = g.mode < 0;
    if (loss) {
        if (g.dtype == DEX_F32) {
            const GK<float> a = make_args<float>(g, chunk_start, Npad);
            switch (sh.GC) {
                case 1: err = launch_one<float, 1, KM_LOSS>(a, sh, n_tiles, n_chunks, stream); break;
                case 2: err = launch_one<float, 2, KM_LOSS>(a, sh, n_tiles, n_chunks, stream); break;
                case 3: err = launch_one<float, 3, KM_LOSS>(a, sh, n_tiles, n_chunks, stream); break;
                case 4: err = launch_one<float, 4, KM_LOSS>(a, sh, n_tiles, n_chunks, stream); break;
                case 5: err = launch_one<float, 5, KM_LOSS>(a, sh, n_tiles, n_chunks, stream); break;
                case 6: err = launch_one<float, 6, KM_LOSS>(a, sh, n_tiles, n_chunks, stream); break;
                default: err = launch_one<float, 8, KM_LOSS>(a, sh, n_tiles, n_chunks, stream); break;
            }
        } else {
            const GK<double> a = make_args<double>(g, chunk_start, Npad);
            switch (sh.GC) {
                case 1: err = launch_one<double, 1, KM_LOSS>(a, sh, n_tiles, n_chunks, stream); break;
                case 2: err = launch_one<double, 2, KM_LOSS>(a, sh, n_tiles, n_chunks, stream); break;
                case 3: err = launch_one<double, 3, KM_LOSS>(a, sh, n_tiles, n_chunks, stream); break;
                case 4: err = launch_one<double, 4, KM_LOSS>(a, sh, n_tiles, n_chunks, stream); break;
                case 5: err = launch_one<double, 5, KM_LOSS>(a, sh, n_tiles, n_chunks, stream); break;
                case 6: err = launch_one<double, 6, KM_LOSS>(a, sh, n_tiles, n_chunks, stream); break;
                default: err = launch_one<double, 8, KM_LOSS>(a, sh, n_tiles, n_chunks, stream); break;
            }
        }
    } else if (g.dtype == DEX_F32) {
        const GK<float> a = make_args<float>(g, chunk_start, Npad);
        if (diff) err = launch_one<float, 1, KM_DIFF>(a, sh, n_tiles, n_chunks, stream);
        else switch (sh.GC) {
            case 1: err = launch_one<float, 1, KM_GRAD>(a, sh, n_tiles, n_chunks, stream); break;
            case 2: err = launch_one<float, 2, KM_GRAD>(a, sh, n_tiles, n_chunks, stream); break;
            case 3: err = launch_one<float, 3, KM_GRAD>(a, sh, n_tiles, n_chunks, stream); break;
            case 4: err = launch_one<float, 4, KM_GRAD>(a, sh, n_tiles, n_chunks, stream); break;
            case 5: err = launch_one<float, 5, KM_GRAD>(a, sh, n_tiles, n_chunks, stream); break;
            case 6: err = launch_one<float, 6, KM_GRAD>(a, sh, n_tiles, n_chunks, stream); break;
            default: err = launch_one<float, 8, KM_GRAD>(a, sh, n_tiles, n_chunks, stream); break;
        }
    } else {
        const GK<double> a = make_args<double>(g, chunk_start, Npad);
        if (diff) err = launch_one<double, 1, KM_DIFF>(a, sh, n_tiles, n_chunks, stream);
        else switch (sh.GC) {
            case 1: err = launch_one<double, 1, KM_GRAD>(a, sh, n_tiles, n_chunks, stream); break;
            case 2: err = launch_one<double, 2, KM_GRAD>(a, sh, n_tiles, n_chunks, stream); break;
            case 3: err = launch_one<double, 3, KM_GRAD>(a, sh, n_tiles, n_chunks, stream); break;
            case 4: err = launch_one<double, 4, KM_GRAD>(a, sh, n_tiles, n_chunks, stream); break;
            case 5: err = launch_one<double, 5, KM_GRAD>(a, sh, n_tiles, n_chunks, stream); break;
            case 6: err = launch_one<double, 6, KM_GRAD>(a, sh, n_tiles, n_chunks, stream); break;
            default: err = launch_one<double, 8, KM_GRAD>(a, sh, n_tiles, n_chunks, stream); break;
        }
    }
    if (err == cudaSuccess && launches) *launches += 1;
    return err;
}

}  // namespace dex
