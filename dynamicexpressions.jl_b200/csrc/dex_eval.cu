// dex_eval.cu — the batched tape interpreter (sm_100a).
//
// One launch evaluates a whole population: replaces the serial comprehension
// `[eval_tree_array(tree, X, operators) for tree in trees]`
// (/root/reference/benchmark/benchmarks.jl:76-91) and, inside it, every streaming loop
// kernel of /root/reference/src/Evaluate.jl:366-404, 693-993.  No intermediate array ever
// leaves the SM: per (tree, sample) the only HBM traffic is the staged X column and the
// result element.
//
// Mapping
//   grid.x  sample tiles of TILE = blockDim.x * K samples; the (F x TILE) slab of the
//           column-major X is staged ONCE per CTA into shared memory, feature-major
//   grid.y  chunks of trees (balanced by tape length on the host)
//   thread  K = U * C samples held as U 16-byte chunks (C = 4 floats / 2 doubles):
//           chunk u of thread t covers samples u*(blockDim.x*C) + t*C .. +C-1, so every
//           operand fetch is a conflict-free LDS.128 and every result store a fully
//           coalesced STG.128.  Accumulator, operands and validity accumulators live in
//           registers; the operand stack is rows of the same shared array.
//   CTA     walks the tapes of its chunk; the tape pointer depends only on blockIdx and
//           loop counters, so instruction fetch/decode/branch are warp-uniform and run on
//           the uniform datapath (no divergence on the op switch).
// The kernel is instruction-issue bound (ncu: >90 % of peak issue rate), so the design
// goal is the fewest SASS instructions per tape instruction: one indirect branch to a
// handler specialised for (operator, operand sources), K samples per dispatch, validity
// checks elided on the host where a later check subsumes them.
// No tensor cores: the path is elementwise, not a contraction.
#include "dex_kernels.h"
// Float64 sin / cos behind a call here (inline they cost the evaluation kernel 5 %: C2-f64 1.19 -> 1.25 ms);
// the gradient kernel inlines them (C3-f64 4.36 -> 4.23 ms)
#define DEX_F64_SINCOS_INLINE 0
#include "dex_ops.cuh"
#include "dex_fold.cuh"

#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

// samples per thread = DEX_EVAL_U chunks of 16 bytes (8 floats / 4 doubles for U = 2)
#ifndef DEX_EVAL_U
#define DEX_EVAL_U 2
#endif
#ifndef DEX_PTX_INTERP
#define DEX_PTX_INTERP 1
#endif
#ifndef DEX_MIN_CTAS
#define DEX_MIN_CTAS 3
#endif
#ifndef DEX_PTX_INTERP_F64
#define DEX_PTX_INTERP_F64 1
#endif
#ifndef DEX_PACKED_FP32
#define DEX_PACKED_FP32 1
#endif
#ifndef DEX_MAX_THREADS
#define DEX_MAX_THREADS 256
#endif
// residual of the fused loss: the exact difference of the two values in double (default), or the
// difference rounded to T first (what a caller computing `pred .- y` in T would square)
// resident CTAs per SM the wide-input evaluation kernel is compiled for: 4 (64 registers, 20 bytes of
// spills) beats 3 on wide inputs (C4 shard 14.8 -> 13.7 ms); its fused-loss form spills more and stays at 3
#ifndef DEX_GX_MIN_CTAS
#define DEX_GX_MIN_CTAS 4
#endif
#ifndef DEX_LOSS_RESIDUAL_IN_T
#define DEX_LOSS_RESIDUAL_IN_T 0
#endif
#if DEX_LOSS_RESIDUAL_IN_T
#define DEX_LOSS_RESIDUAL(v, y) ((double)((v) - (y)))
#else
#define DEX_LOSS_RESIDUAL(v, y) ((double)(v) - (double)(y))
#endif
#ifndef DEX_SYNC_TREE
#define DEX_SYNC_TREE 0
#endif


namespace dex {

namespace {

template <typename T> struct ChunkOf;
template <> struct ChunkOf<float> { static constexpr int C = 4; };
template <> struct ChunkOf<double> { static constexpr int C = 2; };

// operators as compile-time functors: Op1<DEX_OP_COS, T>::f(x)
template <int OPC, typename T> struct Op1;
template <int OPC, typename T> struct Op2;
#define X1(SYM, VEXPR, GEXPR) \
    template <typename T> struct Op1<DEX_OP_##SYM, T> { static __device__ __forceinline__ T f(T x) { return (VEXPR); } };
DEX_UNARY_OPS(X1)
#undef X1
#define X2(SYM, VEXPR, G0, G1) \
    template <typename T> struct Op2<DEX_OP_##SYM, T> { static __device__ __forceinline__ T f(T x, T y) { return (VEXPR); } };
DEX_BINARY_OPS(X2)
#undef X2

// ---- vector-level operator application ------------------------------------------------
// out[k] = op(x[k]) / op(x[k], y[k]) over the K samples of a thread.  The generic form is the
// scalar functor unrolled K times.  For Float32 the cheap arithmetic operators and sin/cos use
// Blackwell's packed FP32 instructions (FADD2 / FMUL2 / FFMA2 via __fadd2_rn / __fmul2_rn /
// __ffma2_rn: two IEEE-754 single-precision operations per issued instruction, bit-identical
// to the scalar forms) — the kernel is issue-bound, so halving the instruction count of the
// arithmetic is a direct win.
template <int OPC, typename T, int K> struct VOp1 {
    static __device__ __forceinline__ void f(T* out, const T* x) {
#pragma unroll
        for (int k = 0; k < K; ++k) out[k] = Op1<OPC, T>::f(x[k]);
    }
};
template <int OPC, typename T, int K> struct VOp2 {
    static __device__ __forceinline__ void f(T* out, const T* x, const T* y) {
#pragma unroll
        for (int k = 0; k < K; ++k) out[k] = Op2<OPC, T>::f(x[k], y[k]);
    }
};
#if DEX_PACKED_FP32
__device__ __forceinline__ float2 f2(const float* p) { return make_float2(p[0], p[1]); }
template <int K> struct VOp2<DEX_OP_ADD, float, K> {
    static __device__ __forceinline__ void f(float* out, const float* x, const float* y) {
#pragma unroll
        for (int k = 0; k < K; k += 2) { const float2 r = __fadd2_rn(f2(x + k), f2(y + k)); out[k] = r.x; out[k + 1] = r.y; }
    }
};
template <int K> struct VOp2<DEX_OP_SUB, float, K> {
    static __device__ __forceinline__ void f(float* out, const float* x, const float* y) {
#pragma unroll
        for (int k = 0; k < K; k += 2) {
            const float2 r = __fadd2_rn(f2(x + k), make_float2(-y[k], -y[k + 1]));
            out[k] = r.x; out[k + 1] = r.y;
        }
    }
};
template <int K> struct VOp2<DEX_OP_MUL, float, K> {
    static __device__ __forceinline__ void f(float* out, const float* x, const float* y) {
#pragma unroll
        for (int k = 0; k < K; k += 2) { const float2 r = __fmul2_rn(f2(x + k), f2(y + k)); out[k] = r.x; out[k + 1] = r.y; }
    }
};
// packed sin/cos: the same algorithm as fast_sincosf (dex_ops.cuh) on pairs, operation for
// operation (bit-identical results).
template <int QADD, int K> __device__ __forceinline__ void sincos_packed(float* out, const float* x) {
    float big = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) big = fmaxf(big, fabsf(x[k]));
    if (!(big <= 105615.0f)) {   // some sample needs Payne-Hanek (or is NaN/Inf): scalar path for all
#pragma unroll
        for (int k = 0; k < K; ++k) out[k] = QADD ? m_cos(x[k]) : m_sin(x[k]);
        return;
    }
    const float2 MAGIC = make_float2(12582912.0f, 12582912.0f), NMAGIC = make_float2(-12582912.0f, -12582912.0f);
    const float2 INVPI = make_float2(0.318309886183790672f, 0.318309886183790672f);
#pragma unroll
    for (int k = 0; k < K; k += 2) {
        const float2 xx = f2(x + k);
        float2 m, q;
        if (QADD) {
            const float2 u = __ffma2_rn(xx, INVPI, make_float2(-0.5f, -0.5f));
            m = __fadd2_rn(u, MAGIC);
            q = __ffma2_rn(__fadd2_rn(m, NMAGIC), make_float2(2.0f, 2.0f), make_float2(1.0f, 1.0f));
        } else {
            m = __ffma2_rn(xx, INVPI, MAGIC);
            const float2 j = __fadd2_rn(m, NMAGIC);
            q = __fadd2_rn(j, j);
        }
        float2 r = __ffma2_rn(q, make_float2(-1.5707962513e+00f, -1.5707962513e+00f), xx);
        r = __ffma2_rn(q, make_float2(-7.5497894159e-08f, -7.5497894159e-08f), r);
        r = __ffma2_rn(q, make_float2(-5.3903029534e-15f, -5.3903029534e-15f), r);
        const float2 z = __fmul2_rn(r, r);
        float2 sp = __ffma2_rn(z, make_float2(0x1.5dbce6p-19f, 0x1.5dbce6p-19f), make_float2(-0x1.9f6feep-13f, -0x1.9f6feep-13f));
        sp = __ffma2_rn(sp, z, make_float2(0x1.110ed4p-7f, 0x1.110ed4p-7f));
        sp = __ffma2_rn(sp, z, make_float2(-0x1.55554cp-3f, -0x1.55554cp-3f));
        sp = __ffma2_rn(__fmul2_rn(sp, z), r, r);
        const unsigned p0 = (__float_as_uint(m.x) & 1u) ^ (QADD ? 1u : 0u), p1 = (__float_as_uint(m.y) & 1u) ^ (QADD ? 1u : 0u);
        out[k] = __uint_as_float(__float_as_uint(sp.x) ^ (p0 << 31));
        out[k + 1] = __uint_as_float(__float_as_uint(sp.y) ^ (p1 << 31));
    }
}
template <int K> struct VOp1<DEX_OP_SIN, float, K> {
    static __device__ __forceinline__ void f(float* out, const float* x) { sincos_packed<0, K>(out, x); }
};
template <int K> struct VOp1<DEX_OP_COS, float, K> {
    static __device__ __forceinline__ void f(float* out, const float* x) { sincos_packed<1, K>(out, x); }
};
#endif

// operand positions of the specialised binary handlers whose feature ROW may carry a check flag
// (dex_flatten.cpp pick_handler): the divisor of /, either operand of max / min
template <int OPC> struct RowChk { static constexpr bool a = OPC == DEX_OP_MAX || OPC == DEX_OP_MIN;
                                   static constexpr bool b = a || OPC == DEX_OP_DIV; };

template <typename T> __device__ __forceinline__ T const_of(const uint4& ins);
template <> __device__ __forceinline__ float const_of<float>(const uint4& ins) { return __uint_as_float(ins.z); }
template <> __device__ __forceinline__ double const_of<double>(const uint4& ins) { return __hiloint2double((int)ins.w, (int)ins.z); }

template <typename T> struct KArgs {
    const uint4* tape;
    const int64_t* tape_off;
    const int32_t* chunk_start;
    const T* X;
    T* out;
    uint8_t* ok;
    const T* params;
    const int32_t* classes;
    const T* y;
    const T* w;
    double* loss_partial;
    int64_t N, ldx, ldo, n_trees;
    int32_t F, max_stack, n_param_rows, early_exit, n_params, n_classes;
    // re-align the warps of a CTA at every tree: with short tapes (~11 instructions per tree) the
    // eight warps then share tape lines in L1 and handler code in the instruction cache (C6: -3 %);
    // with long tapes or the parametric gather the barrier costs more than it saves (C4: +3 %)
    int32_t sync_tree;
    // wide inputs (GX kernels): rows [0, smem_rows) live in shared memory (the stack rows and the
    // first smem_rows - max_stack features), the remaining feature rows are read from the
    // feature-major global copy through L1 — a third CTA fits per SM
    int32_t smem_rows;
};

// A row vector of one thread: U chunks of C elements.
template <typename T, int U> struct Vec {
    static constexpr int C = ChunkOf<T>::C;
    static constexpr int K = U * C;
    T v[K];
};

template <typename T, int U>
__device__ __forceinline__ void ld_row(Vec<T, U>& r, const T* row_base, int chunk_stride) {
    // row_base already includes the thread offset t*C
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const uint4 q = *reinterpret_cast<const uint4*>(row_base + u * chunk_stride);
        *reinterpret_cast<uint4*>(&r.v[u * Vec<T, U>::C]) = q;
    }
}
template <typename T, int U>
__device__ __forceinline__ void st_row(T* row_base, int chunk_stride, const Vec<T, U>& r) {
#pragma unroll
    for (int u = 0; u < U; ++u)
        *reinterpret_cast<uint4*>(row_base + u * chunk_stride) = *reinterpret_cast<const uint4*>(&r.v[u * Vec<T, U>::C]);
}

// validity accumulator: nf stays +-0 while every checked value is finite and turns NaN
// forever once one is not (v * 0 is NaN for Inf and NaN)  — is_valid, ValueInterface.jl:5-9
template <typename T, int U>
__device__ __forceinline__ void check(T (&nf)[2], const Vec<T, U>& r) {
#pragma unroll
    for (int k = 0; k < Vec<T, U>::K; ++k) nf[k & 1] = m_fma(r.v[k], T(0), nf[k & 1]);
}
#if DEX_PACKED_FP32
template <int U>
__device__ __forceinline__ void check(float (&nf)[2], const Vec<float, U>& r) {
    float2 a = make_float2(nf[0], nf[1]);
#pragma unroll
    for (int k = 0; k < Vec<float, U>::K; k += 2) a = __ffma2_rn(f2(&r.v[k]), make_float2(0.f, 0.f), a);
    nf[0] = a.x; nf[1] = a.y;
}
#endif

// result = isfinite(argument) ? result : Inf, per sample (the GUARD flag; early_exit = false only)
template <typename T, int U>
__device__ __forceinline__ void guard_inf(Vec<T, U>& r, const Vec<T, U>& x) {
#pragma unroll
    for (int k = 0; k < Vec<T, U>::K; ++k)
        if (!t_finite(x.v[k])) r.v[k] = t_inf<T>();
}

// NT: CTA size fixed at compile time (the full-size 256-thread launch: row and chunk strides
// become immediates of the shared-memory accesses) or 0 = read blockDim.x.
template <typename T, int U, bool FAST, bool PARAM, bool LOSS, int NT = 0, bool GX = false>
__global__ void __launch_bounds__(DEX_MAX_THREADS, (U == 1 && sizeof(T) == 4) ? 4 : ((GX && !LOSS && !PARAM) ? DEX_GX_MIN_CTAS : DEX_MIN_CTAS)) eval_kernel(const KArgs<T> a) {
    using V = Vec<T, U>;
    constexpr int C = V::C;
    constexpr int K = V::K;
    static_assert(!GX || (NT == 256 && sizeof(T) == 4 && U == 2), "the GX loop hard-codes the chunk stride of 256 threads");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* rows = reinterpret_cast<T*>(smem_raw);
    const int tid = threadIdx.x;
    const int nthr = NT ? NT : (int)blockDim.x;
    const int TILE = nthr * K;          // samples per CTA = row stride in elements
    const int CS = nthr * C;            // chunk stride in elements
    const int64_t s0 = (int64_t)blockIdx.x * TILE;

    // ---- stage the X slab feature-major: xs[f][s] = X[f, s0 + s] -------------------
    // a.X is the feature-major, tile-padded copy XT[f][Npad] written by transpose_pad_kernel
    // (tail columns replay the last valid sample), so every feature row of this tile is one
    // contiguous, 16-byte aligned run of TILE elements: one elected thread issues F bulk
    // async copies (TMA, cp.async.bulk -> SASS UBLKCP) that complete on an mbarrier.
    __shared__ __align__(8) unsigned long long stage_bar;
    {
        T* xs = rows + (size_t)(a.max_stack + a.n_param_rows) * TILE;
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&stage_bar);
        const uint32_t row_bytes = (uint32_t)TILE * (uint32_t)sizeof(T);
        const int FS = GX ? a.smem_rows - a.max_stack - a.n_param_rows : a.F;   // features staged
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                         "r"(row_bytes * (uint32_t)FS)
                         : "memory");
            for (int f = 0; f < FS; ++f) {
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(xs + (size_t)f * TILE);
                const T* src = a.X + (size_t)f * a.ldx + s0;
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                    "l"(src), "r"(row_bytes), "r"(bar)
                    : "memory");
            }
        }
        __syncthreads();  // the barrier is initialised before anybody polls it
        if (FS > 0) {
            uint32_t done = 0;
            while (!done) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                    "selp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(done)
                    : "r"(bar)
                    : "memory");
            }
        }
    }
    // sample index of element k of this thread: (k / C) * CS + tid * C + k % C
    int cls[PARAM ? K : 1];
    if (PARAM) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            int64_t gs = s0 + (k / C) * CS + tid * C + (k % C);
            if (gs >= a.N) gs = a.N - 1;
            cls[k] = __ldg(a.classes + gs) * a.n_params;
        }
    }
    T yv[LOSS ? K : 1], wv[LOSS ? K : 1];
    if (LOSS) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            int64_t gs = s0 + (k / C) * CS + tid * C + (k % C);
            const bool in = gs < a.N;
            if (!in) gs = a.N - 1;
            yv[k] = __ldg(a.y + gs);
            wv[k] = in ? (a.w ? __ldg(a.w + gs) : T(1)) : T(0);
        }
    }
    __syncthreads();

    T* my = rows + tid * C;  // this thread's first chunk inside row 0
    // GX: this thread's first chunk of "row 0" in the global copy (feature f is row
    // max_stack + n_param_rows + f; the base is moved back by that many rows)
    const T* xg = GX ? a.X + s0 + tid * C - (int64_t)(a.max_stack + a.n_param_rows) * a.ldx : nullptr;
    const int t0 = a.chunk_start[blockIdx.y], t1 = a.chunk_start[blockIdx.y + 1];
    const bool early = FAST ? true : (a.early_exit != 0);
    const bool full_tile_samples = s0 + TILE <= a.N;
    const bool full_tile = full_tile_samples && ((a.ldo % C) == 0) &&
                           ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0);

    // The tape offsets are loaded one tree ahead, and the first instruction of a tree arrives as
    // the prefetch of its predecessor's last one (tapes are contiguous, the buffer carries slack):
    // no tree starts with a chain of dependent global loads.
    int64_t off = a.tape_off[t0], off_next = a.tape_off[t0 + 1];
    uint4 ins0 = __ldg(a.tape + off);   // instruction at the current pc of the PTX loop
    for (int t = t0; t < t1; ++t) {
        // every sync_tree-th tree (a power of two; 0 = never) the warps of the CTA are re-aligned: they
        // share tape lines in L1 and handler code in the instruction cache
        if (DEX_SYNC_TREE || (!PARAM && a.sync_tree && (t & (a.sync_tree - 1)) == 0)) __syncthreads();
        const int n = (int)(off_next - off);
        const uint4* ip = a.tape + off;
        const int64_t off_next2 = a.tape_off[t + 2];   // slack behind the table: dex_api.cu upload()
        const T* ptree = PARAM ? a.params + (size_t)t * a.n_params * a.n_classes : nullptr;
        if (PARAM) {
            // ParametricExpression: gather this tree's per-sample parameters
            // parameters[p, classes[j]] (src/ParametricExpression.jl:380-384) into the parameter
            // rows; each thread fills and later reads only its own columns, so no barrier
            for (int p = 0; p < a.n_param_rows; ++p) {
                V pv;
#pragma unroll
                for (int k = 0; k < K; ++k) pv.v[k] = __ldg(ptree + cls[k] + p);
                st_row<T, U>(my + (size_t)(a.max_stack + p) * TILE, CS, pv);
            }
        }
        V acc;
        T nf[2] = {T(0), T(0)};
#pragma unroll
        for (int k = 0; k < K; ++k) acc.v[k] = T(0);

        // one tape instruction, C++ form (all handlers + the generic path)
        auto step = [&](const uint4& ins) {
            const uint32_t w0 = ins.x;
            const T* ra = my + (size_t)row_a(ins.y) * TILE;
            const T* rb = my + (size_t)row_b(ins.y) * TILE;
            if constexpr (GX) {
                if ((int)row_a(ins.y) >= a.smem_rows) ra = xg + (int64_t)row_a(ins.y) * a.ldx;
                if ((int)row_b(ins.y) >= a.smem_rows) rb = xg + (int64_t)row_b(ins.y) * a.ldx;
            }
            const T c = const_of<T>(ins);
            V cv;  // the inline constant broadcast over the K samples
#pragma unroll
            for (int k = 0; k < K; ++k) cv.v[k] = c;
            if (w0 & F_PUSH) st_row<T, U>(my + (size_t)push_row(w0) * TILE, CS, acc);

            // early_exit = false (FAST == false) runs the same specialised handlers: the flag
            // checks then apply only to ALWAYS instructions, and the two fused unary kernels of the
            // reference substitute Inf where their inner value is invalid (GUARD,
            // /root/reference/src/Evaluate.jl:722, 737, 754, 787)
            const uint32_t h = w0 & HANDLER_MASK;
            const bool con = FAST ? true : (early || (w0 & F_ALWAYS));   // checks are on
#define HANDLER_END break;
            switch (h) {
                // ---- specialised handlers: one indirect branch, no operand decoding ----
                case H_LOAD_R: {
                    ld_row<T, U>(acc, ra, CS);
                    if (con && (w0 & F_CHK_A)) check<T, U>(nf, acc);
                } HANDLER_END
                case H_KEEP: {   // the PUSH above was the point
                } HANDLER_END
                case H_LOAD_C: {
#pragma unroll
                    for (int k = 0; k < K; ++k) acc.v[k] = c;
                    if (con && (w0 & F_CHK_A)) nf[0] = m_fma(c, T(0), nf[0]);
                } HANDLER_END
#define UNARY_HANDLERS(S)                                                          \
    case H_##S##_A: {                                                              \
        V x;                                                                       \
        if (!FAST) x = acc;                                                        \
        VOp1<DEX_OP_##S, T, K>::f(acc.v, acc.v);                                   \
        if (!FAST && (w0 & F_GUARD)) guard_inf<T, U>(acc, x);                      \
    } HANDLER_END                                                                  \
    case H_##S##_R: {                                                              \
        V x;                                                                       \
        ld_row<T, U>(x, ra, CS);                                                   \
        if (con && (w0 & F_CHK_A)) check<T, U>(nf, x);                             \
        VOp1<DEX_OP_##S, T, K>::f(acc.v, x.v);                                     \
        if (!FAST && (w0 & F_GUARD)) guard_inf<T, U>(acc, x);                      \
    } HANDLER_END
                DEX_FAST_UNARY(UNARY_HANDLERS)
#undef UNARY_HANDLERS
#define BIN_AR(S)                                                                  \
    case H_##S##_AR: {                                                             \
        V y;                                                                       \
        ld_row<T, U>(y, rb, CS);                                                   \
        if (RowChk<DEX_OP_##S>::b && con && (w0 & F_CHK_B)) check<T, U>(nf, y);           \
        VOp2<DEX_OP_##S, T, K>::f(acc.v, acc.v, y.v);                              \
    } HANDLER_END
#define BIN_RA(S)                                                                  \
    case H_##S##_RA: {                                                             \
        V x;                                                                       \
        ld_row<T, U>(x, ra, CS);                                                   \
        if (RowChk<DEX_OP_##S>::a && con && (w0 & F_CHK_A)) check<T, U>(nf, x);           \
        VOp2<DEX_OP_##S, T, K>::f(acc.v, x.v, acc.v);                              \
    } HANDLER_END
#define BIN_AC(S)                                                                  \
    case H_##S##_AC: {                                                             \
        if (con && (w0 & F_CHK_B)) nf[0] = m_fma(c, T(0), nf[0]);                        \
        VOp2<DEX_OP_##S, T, K>::f(acc.v, acc.v, cv.v);                             \
    } HANDLER_END
#define BIN_CA(S)                                                                  \
    case H_##S##_CA: {                                                             \
        if (con && (w0 & F_CHK_A)) nf[0] = m_fma(c, T(0), nf[0]);                        \
        VOp2<DEX_OP_##S, T, K>::f(acc.v, cv.v, acc.v);                             \
    } HANDLER_END
#define BIN_RR(S)                                                                  \
    case H_##S##_RR: {                                                             \
        V x, y;                                                                    \
        ld_row<T, U>(x, ra, CS);                                                   \
        ld_row<T, U>(y, rb, CS);                                                   \
        if (RowChk<DEX_OP_##S>::a && con && (w0 & F_CHK_A)) check<T, U>(nf, x);           \
        if (RowChk<DEX_OP_##S>::b && con && (w0 & F_CHK_B)) check<T, U>(nf, y);           \
        VOp2<DEX_OP_##S, T, K>::f(acc.v, x.v, y.v);                                \
    } HANDLER_END
#define BIN_RC(S)                                                                  \
    case H_##S##_RC: {                                                             \
        V x;                                                                       \
        ld_row<T, U>(x, ra, CS);                                                   \
        if (RowChk<DEX_OP_##S>::a && con && (w0 & F_CHK_A)) check<T, U>(nf, x);           \
        if (con && (w0 & F_CHK_B)) nf[0] = m_fma(c, T(0), nf[0]);                        \
        VOp2<DEX_OP_##S, T, K>::f(acc.v, x.v, cv.v);                               \
    } HANDLER_END
#define BIN_CR(S)                                                                  \
    case H_##S##_CR: {                                                             \
        V y;                                                                       \
        ld_row<T, U>(y, rb, CS);                                                   \
        if (con && (w0 & F_CHK_A)) nf[0] = m_fma(c, T(0), nf[0]);                        \
        if (RowChk<DEX_OP_##S>::b && con && (w0 & F_CHK_B)) check<T, U>(nf, y);           \
        VOp2<DEX_OP_##S, T, K>::f(acc.v, cv.v, y.v);                               \
    } HANDLER_END
#define COMM_HANDLERS(S) BIN_AR(S) BIN_AC(S) BIN_RR(S) BIN_RC(S)
#define NC_HANDLERS(S) BIN_AR(S) BIN_RA(S) BIN_AC(S) BIN_CA(S) BIN_RR(S) BIN_RC(S) BIN_CR(S)
                DEX_FAST_BIN_COMM(COMM_HANDLERS)
                DEX_FAST_BIN_NC(NC_HANDLERS)
#undef COMM_HANDLERS
#undef NC_HANDLERS
#undef BIN_AR
#undef BIN_RA
#undef BIN_AC
#undef BIN_CA
#undef BIN_RR
#undef BIN_RC
#undef BIN_CR
                // ---- generic handler: any operator, any operand source, every flag --------
                default: {
                    V va, vb;
                    {
                        const uint32_t src = (w0 >> 16) & 3u;
                        if (src == SRC_ROW) ld_row<T, U>(va, ra, CS);
                        else if (src == SRC_CONST) {
#pragma unroll
                            for (int k = 0; k < K; ++k) va.v[k] = c;
                        } else if (PARAM && src == SRC_PARAM) {
#pragma unroll
                            for (int k = 0; k < K; ++k) va.v[k] = __ldg(ptree + cls[k] + row_a(ins.y));
                        } else va = acc;
                    }
                    {
                        const uint32_t src = (w0 >> 18) & 3u;
                        if (src == SRC_ROW) ld_row<T, U>(vb, rb, CS);
                        else if (src == SRC_CONST) {
#pragma unroll
                            for (int k = 0; k < K; ++k) vb.v[k] = c;
                        } else if (PARAM && src == SRC_PARAM) {
#pragma unroll
                            for (int k = 0; k < K; ++k) vb.v[k] = __ldg(ptree + cls[k] + row_b(ins.y));
                        } else vb = acc;
                    }
                    const bool chk = early || (w0 & F_ALWAYS);
                    if (chk && (w0 & F_CHK_A)) check<T, U>(nf, va);
                    if (chk && (w0 & F_CHK_B)) check<T, U>(nf, vb);
                    // FAST kernels keep this rarely taken path small (one rolled loop over the K
                    // samples, operands in local memory) so the hot handlers stay in the
                    // instruction cache; the early_exit-off kernel runs every instruction here
                    // and gets the fully unrolled, register-resident form.
                    V r;
                    const uint32_t opc = (w0 >> 8) & 0xffu;
                    const bool guard = !FAST && (w0 & F_GUARD);
                    if (opc == DEX_OP_POW || opc == DEX_OP_POW_ABS) {
                        // the one generic operator that symbolic-regression operator sets use all
                        // the time: K independent library sequences in registers instead of the
                        // rolled loop (no local memory, no per-sample opcode dispatch)
                        if (opc == DEX_OP_POW) {
#pragma unroll
                            for (int k = 0; k < K; ++k) r.v[k] = m_pow(va.v[k], vb.v[k]);
                        } else {
#pragma unroll
                            for (int k = 0; k < K; ++k) r.v[k] = m_exp(vb.v[k] * m_log(m_fabs(va.v[k])));
                        }
                        if (guard) {
#pragma unroll
                            for (int k = 0; k < K; ++k) if (!t_finite(va.v[k])) r.v[k] = t_inf<T>();
                        }
                    } else {
                        T la[K], lb[K], lz[K], lr[K];
#pragma unroll
                        for (int k = 0; k < K; ++k) { la[k] = va.v[k]; lb[k] = vb.v[k]; lz[k] = acc.v[k]; }
                        // the opcode dispatch runs ONCE per instruction; each case is a rolled loop over
                        // the samples (the reference's fused unary kernels substitute Inf where the inner
                        // value is invalid — `guard`, only observable when early_exit is off)
#define GEN_LOOP(EXPR)                                                              \
    _Pragma("unroll 1") for (int k = 0; k < K; ++k) {                               \
        const T x = la[k], y = lb[k], z = lz[k];                                    \
        T v = (EXPR);                                                               \
        if (guard && !t_finite(x)) v = t_inf<T>();                                  \
        lr[k] = v;                                                                  \
        (void)y; (void)z;                                                           \
    }
                        switch (opc) {
#define U_CASE(SYM, VEXPR, GEXPR) case DEX_OP_##SYM: GEN_LOOP(VEXPR) break;
                            DEX_UNARY_OPS(U_CASE)
#undef U_CASE
#define B_CASE(SYM, VEXPR, G0, G1) case DEX_OP_##SYM: GEN_LOOP(VEXPR) break;
                            DEX_BINARY_OPS(B_CASE)
#undef B_CASE
#define T_CASE(SYM, VEXPR, G0, G1, G2) case DEX_OP_##SYM: GEN_LOOP(VEXPR) break;
                            DEX_TERNARY_OPS(T_CASE)
#undef T_CASE
                            default: GEN_LOOP(t_nan<T>()) break;
                        }
#undef GEN_LOOP
#pragma unroll
                        for (int k = 0; k < K; ++k) r.v[k] = lr[k];
                    }
                    acc = r;
                    if (!chk) return;  // CHK_OUT below is unconditional for FAST
                } break;
            }
            if (con && (w0 & F_CHK_OUT)) check<T, U>(nf, acc);
        };
#undef HANDLER_END

        if constexpr (DEX_PTX_INTERP && sizeof(T) == 4 && (U == 2 || (U == 1 && FAST))) {
            // Float32 hot path: the instruction loop as one inline-PTX block with a real jump
            // table (gen_interp_ptx.py, one block per U).  It returns at the end of the tape or at
            // the first instruction it does not implement natively, which `step` then executes.
            const uint32_t my_s = (uint32_t)__cvta_generic_to_shared(my);
            const uint32_t tile_b = (uint32_t)TILE * 4u, cs_b = (uint32_t)CS * 4u;
            int pc = 0;
            float* av = reinterpret_cast<float*>(acc.v);
            float* nfv = reinterpret_cast<float*>(nf);
            while (pc < n) {
                if constexpr (U == 2 && FAST && GX) {
                    asm volatile(
#include "dex_interp_f32_gx.inc"
                        : "+r"(pc), "+f"(av[0]), "+f"(av[1]), "+f"(av[2]), "+f"(av[3]), "+f"(av[4]), "+f"(av[5]),
                          "+f"(av[6]), "+f"(av[7]), "+f"(nfv[0]), "+f"(nfv[1]), "+r"(ins0.x), "+r"(ins0.y),
                          "+r"(ins0.z), "+r"(ins0.w)
                        : "l"(ip), "r"(n), "r"(my_s), "r"(tile_b), "r"(cs_b), "l"(__cvta_generic_to_global(k_inv_pio4)),
                          "r"((uint32_t)a.smem_rows), "l"(__cvta_generic_to_global(xg)), "r"((uint32_t)a.ldx * 4u)
                        : "memory");
                } else if constexpr (U == 2 && FAST) {
                    asm volatile(
#include "dex_interp_f32.inc"
                        : "+r"(pc), "+f"(av[0]), "+f"(av[1]), "+f"(av[2]), "+f"(av[3]), "+f"(av[4]), "+f"(av[5]),
                          "+f"(av[6]), "+f"(av[7]), "+f"(nfv[0]), "+f"(nfv[1]), "+r"(ins0.x), "+r"(ins0.y),
                          "+r"(ins0.z), "+r"(ins0.w)
                        : "l"(ip), "r"(n), "r"(my_s), "r"(tile_b), "r"(cs_b), "l"(__cvta_generic_to_global(k_inv_pio4))
                        : "memory");
                } else if constexpr (U == 2 && GX) {
                    asm volatile(
#include "dex_interp_f32_noexit_gx.inc"
                        : "+r"(pc), "+f"(av[0]), "+f"(av[1]), "+f"(av[2]), "+f"(av[3]), "+f"(av[4]), "+f"(av[5]),
                          "+f"(av[6]), "+f"(av[7]), "+f"(nfv[0]), "+f"(nfv[1]), "+r"(ins0.x), "+r"(ins0.y),
                          "+r"(ins0.z), "+r"(ins0.w)
                        : "l"(ip), "r"(n), "r"(my_s), "r"(tile_b), "r"(cs_b), "l"(__cvta_generic_to_global(k_inv_pio4)),
                          "r"((uint32_t)a.smem_rows), "l"(__cvta_generic_to_global(xg)), "r"((uint32_t)a.ldx * 4u)
                        : "memory");
                } else if constexpr (U == 2) {
                    // early_exit = false: checks only where ALWAYS is set, GUARD substitution, no skipping
                    asm volatile(
#include "dex_interp_f32_noexit.inc"
                        : "+r"(pc), "+f"(av[0]), "+f"(av[1]), "+f"(av[2]), "+f"(av[3]), "+f"(av[4]), "+f"(av[5]),
                          "+f"(av[6]), "+f"(av[7]), "+f"(nfv[0]), "+f"(nfv[1]), "+r"(ins0.x), "+r"(ins0.y),
                          "+r"(ins0.z), "+r"(ins0.w)
                        : "l"(ip), "r"(n), "r"(my_s), "r"(tile_b), "r"(cs_b), "l"(__cvta_generic_to_global(k_inv_pio4))
                        : "memory");
                } else {
                    asm volatile(
#include "dex_interp_f32_u1.inc"
                        : "+r"(pc), "+f"(av[0]), "+f"(av[1]), "+f"(av[2]), "+f"(av[3]), "+f"(nfv[0]), "+f"(nfv[1]),
                          "+r"(ins0.x), "+r"(ins0.y), "+r"(ins0.z), "+r"(ins0.w)
                        : "l"(ip), "r"(n), "r"(my_s), "r"(tile_b), "r"(cs_b), "l"(__cvta_generic_to_global(k_inv_pio4))
                        : "memory");
                }
                if (pc < n) {   // handed over: ins0 is already two instructions ahead
                    step(__ldg(ip + pc));
                    ++pc;
                    ins0 = __ldg(ip + pc);
                }
            }
        } else if constexpr (DEX_PTX_INTERP_F64 && sizeof(T) == 8 && U == 2 && FAST) {
            // Float64: the same loop with scalar double arithmetic (gen_interp_f64_ptx.py): loads, + - * /
            // max min and the cheap unary handlers are native, the transcendental ones are handed to
            // `step` (the library's double-precision sequences)
            const uint32_t my_s = (uint32_t)__cvta_generic_to_shared(my);
            const uint32_t tile_b = (uint32_t)TILE * 8u, cs_b = (uint32_t)CS * 8u;
            int pc = 0;
            double* av = reinterpret_cast<double*>(acc.v);
            double* nfv = reinterpret_cast<double*>(nf);
            while (pc < n) {
                asm volatile(
#include "dex_interp_f64.inc"
                    : "+r"(pc), "+d"(av[0]), "+d"(av[1]), "+d"(av[2]), "+d"(av[3]), "+d"(nfv[0]), "+d"(nfv[1]),
                      "+r"(ins0.x), "+r"(ins0.y), "+r"(ins0.z), "+r"(ins0.w)
                    : "l"(ip), "r"(n), "r"(my_s), "r"(tile_b), "r"(cs_b)
                    : "memory");
                if (pc < n) {   // handed over: ins0 is already two instructions ahead
                    step(__ldg(ip + pc));
                    ++pc;
                    ins0 = __ldg(ip + pc);
                }
            }
        } else {
            uint4 ins = __ldg(ip);
            bool bail = false;
            for (int pc = 0; pc < n; ++pc) {
                uint4 nxt = ins;
                if (pc + 1 < n) nxt = __ldg(ip + pc + 1);  // prefetch the next instruction
                // pull the tape line two lines (16 instructions) ahead into L1: with thousands of
                // trees per CTA the tape no longer stays L1-resident and an L2 miss per line would
                // otherwise be exposed (tapes of consecutive trees are contiguous; the buffer is
                // padded, so running past this tree's end just prefetches the next tree)
                if ((pc & 7) == 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(ip + pc + 16));
                step(ins);
                // early exit proper (/root/reference/src/Evaluate.jl:26-32): once a sample of this warp
                // has tripped a check the tree is incomplete whatever follows and its row unspecified
                // (decided one instruction late, so that the vote's latency hides behind a handler)
                if constexpr (FAST) {
                    if (bail) break;
                    bail = (ins.x & F_CHK_OUT) && __any_sync(0xffffffffu, !(nf[0] + nf[1] == T(0)));
                }
                ins = nxt;
            }
        }

        // ---- result row segment ----------------------------------------------------
        if (!LOSS) {
            T* o = a.out + (size_t)t * a.ldo + s0 + (size_t)tid * C;
            if (full_tile) {
#pragma unroll
                for (int u = 0; u < U; ++u)
                    __stcs(reinterpret_cast<float4*>(o + u * CS), *reinterpret_cast<const float4*>(&acc.v[u * C]));
            } else {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int64_t s = (k / C) * CS + tid * C + (k % C);
                    if (s0 + s < a.N) a.out[(size_t)t * a.ldo + s0 + s] = acc.v[k];
                }
            }
        }
        const bool bad = (nf[0] != nf[0]) || (nf[1] != nf[1]);
        const bool warp_bad = __any_sync(0xffffffffu, bad);
        if (!LOSS) {
            // stored above
        } else if (FAST && warp_bad) {
            // an incomplete tree has no loss (its row is unspecified under early exit): NaN, no epilogue
            if ((tid & 31) == 0)
                a.loss_partial[((size_t)blockIdx.x * (DEX_MAX_THREADS / 32) + (tid >> 5)) * a.n_trees + t] =
                    __longlong_as_double(0x7ff8000000000000LL);
        } else {
            // fused loss: sum_j w_j (v_j - y_j)^2 over this warp's samples, in double (the caller's
            // yardstick is the float64 reduction of the float32 values).  One partial per (tile, warp,
            // tree) goes straight to global memory — no shared memory, no barrier; the second-stage
            // kernel adds the partials in a fixed order (deterministic).
            // Four independent accumulation chains per thread (a single chain of K dependent DFMAs
            // is pure latency in front of the warp reduction and the barrier).
            double lp[4] = {0.0, 0.0, 0.0, 0.0};
            if (a.w || !full_tile_samples) {   // w_j given, or 1 with 0 on the padded tail of the last tile
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const double d = DEX_LOSS_RESIDUAL(acc.v[k], yv[k]);
                    lp[k & 3] = fma((double)wv[k] * d, d, lp[k & 3]);
                }
            } else {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const double d = DEX_LOSS_RESIDUAL(acc.v[k], yv[k]);
                    lp[k & 3] = fma(d, d, lp[k & 3]);
                }
            }
            double ls = (lp[0] + lp[1]) + (lp[2] + lp[3]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ls += __shfl_xor_sync(0xffffffffu, ls, o);
            if ((tid & 31) == 0)
                a.loss_partial[((size_t)blockIdx.x * (DEX_MAX_THREADS / 32) + (tid >> 5)) * a.n_trees + t] = ls;
        }
        if (warp_bad && (tid & 31) == 0) a.ok[t] = 0;
        off = off_next;
        off_next = off_next2;
    }
}

// Pre-pass: XT[f][s] = X[f, min(s, N-1)] for s < Npad — the feature-major, tile-padded image
// of the caller's column-major X that the interpreter stages with bulk async copies.  The
// first n_trees threads also fold the constant subtrees of one tree each and preset ok[] to
// the outcome (1 unless a folded constant is invalid).  X is tiny next to the results (F*N vs
// P*N elements), the scalar tape tiny next to the sample loop.
template <typename T>
__global__ void transpose_pad_kernel(const T* __restrict__ X, int64_t ldx, int F, int64_t N,
                                     T* __restrict__ XT, int64_t Npad, uint8_t* ok, int64_t n_trees,
                                     Instr* tape, const Instr* ctape, const int64_t* seg,
                                     const int64_t* seg_off, const uint8_t* fold_ok) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < Npad) {
        const T* col = X + (s < N ? s : N - 1) * ldx;
        for (int f = 0; f < F; ++f) XT[(size_t)f * Npad + s] = __ldg(col + f);
    }
    if (s < n_trees)
        ok[s] = fold_ok ? fold_ok[s] : ((!seg_off || fold_tree<T, false>(tape, ctape, seg, seg_off, s)) ? 1 : 0);
}

template <typename T>
__global__ void scatter_constants_kernel(Instr* tape, Instr* scalar_tape, const int64_t* pos, const T* values, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T v = values[i];
    uint32_t lo, hi;
    if (sizeof(T) == 4) { lo = __float_as_uint((float)v); hi = 0; }
    else { lo = (uint32_t)__double2loint((double)v); hi = (uint32_t)__double2hiint((double)v); }
    const int64_t p = pos[i];
    Instr* dst = p >= 0 ? tape + p : (scalar_tape ? scalar_tape - (p + 1) : nullptr);
    if (dst) { dst->c_lo = lo; dst->c_hi = hi; }
}

constexpr size_t SMEM_LIMIT = 227 * 1024;

template <typename T, int U>
cudaError_t launch_typed(const EvalArgs& e, cudaStream_t stream, int threads, size_t smem,
                         int64_t n_tiles) {
    KArgs<T> a;
    a.tape = reinterpret_cast<const uint4*>(e.tape);
    a.tape_off = e.tape_off;
    a.chunk_start = e.chunk_start;
    a.X = static_cast<const T*>(e.X);
    a.out = static_cast<T*>(e.out);
    a.ok = e.ok;
    a.params = static_cast<const T*>(e.params);
    a.classes = e.classes;
    a.y = static_cast<const T*>(e.y);
    a.w = static_cast<const T*>(e.w);
    a.loss_partial = e.loss_partial;
    a.N = e.N; a.ldx = e.ldx; a.ldo = e.ldo; a.n_trees = e.n_trees;
    a.F = e.F; a.max_stack = e.max_stack; a.n_param_rows = e.n_param_rows; a.early_exit = e.early_exit;
    a.n_params = e.n_params; a.n_classes = e.n_classes;
    a.sync_tree = e.sync_tree;
    a.smem_rows = e.smem_rows;
    dim3 grid((unsigned)n_tiles, (unsigned)e.n_chunks);
    const bool param = e.params != nullptr, loss = e.y != nullptr, fast = e.early_exit != 0;
    void (*kern)(const KArgs<T>);
    if constexpr (sizeof(T) == 4 && U == 2) {
        if (e.smem_rows > 0) {   // eval_num_tiles has checked: Float32, early exit, 256 threads, no parameter rows
            kern = !fast  ? eval_kernel<T, U, false, false, false, DEX_MAX_THREADS, true>
                   : param ? eval_kernel<T, U, true, true, false, DEX_MAX_THREADS, true>
                   : loss ? eval_kernel<T, U, true, false, true, DEX_MAX_THREADS, true>
                          : eval_kernel<T, U, true, false, false, DEX_MAX_THREADS, true>;
            cudaError_t err = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem);
            if (err != cudaSuccess) return err;
            kern<<<grid, threads, smem, stream>>>(a);
            return cudaGetLastError();
        }
    }
    if (fast && threads == DEX_MAX_THREADS && sizeof(T) == 4) {
        constexpr int NT = sizeof(T) == 4 ? DEX_MAX_THREADS : 0;
        kern = loss ? (param ? eval_kernel<T, U, true, true, true, NT> : eval_kernel<T, U, true, false, true, NT>)
                    : (param ? eval_kernel<T, U, true, true, false, NT> : eval_kernel<T, U, true, false, false, NT>);
    } else if (fast) {
        kern = loss ? (param ? eval_kernel<T, U, true, true, true> : eval_kernel<T, U, true, false, true>)
                    : (param ? eval_kernel<T, U, true, true, false> : eval_kernel<T, U, true, false, false>);
    } else {
        // early_exit off: GUARD substitution, ALWAYS-only checks, nothing skipped (Float32: the
        // _noexit form of the PTX loop)
        kern = loss ? (param ? eval_kernel<T, U, false, true, true> : eval_kernel<T, U, false, false, true>)
                    : (param ? eval_kernel<T, U, false, true, false> : eval_kernel<T, U, false, false, false>);
    }
    cudaError_t err = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem);
    if (err != cudaSuccess) return err;
    kern<<<grid, threads, smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace

// samples per thread: U chunks of 16 bytes.  U = 2 (8 floats) amortises the dispatch over twice
// the arithmetic.  A U = 1 variant of the PTX loop exists (half the shared memory per CTA, 63
// registers -> 4 CTAs = 32 warps per SM instead of 16-24) for inputs with many rows, but the extra
// dispatches cost more than the occupancy gains: C4 shard (10 features + 4 stack rows) 19.1 -> 20.6
// ms, C2 0.30 -> 0.36 ms, C6 40.9 -> 51.7 ms.  It is kept behind DEXB200_EVAL_U=1 for experiments and
// for shapes whose rows do not fit otherwise.  Float64 always uses U = 2 (4 doubles).
constexpr int EVAL_U = DEX_EVAL_U;
constexpr size_t U1_ROWS = (size_t)1 << 30;      // rows from which a Float32 launch takes U = 1: never

static int eval_pick_u(int dtype, size_t rows) {
    if (const char* env = getenv("DEXB200_EVAL_U")) {   // tuning knob for experiments
        const int v = atoi(env);
        if (v == 1 || v == 2) return dtype == DEX_F32 ? v : EVAL_U;
    }
    return (dtype == DEX_F32 && rows >= U1_ROWS) ? 1 : EVAL_U;
}

size_t eval_xt_bytes(int dtype, int32_t F, int32_t max_stack, int64_t N, int wide) {
    int threads;
    size_t smem;
    const int64_t n_tiles = eval_num_tiles(dtype, F, max_stack, N, &threads, &smem, wide);
    const int u = eval_pick_u(dtype, (size_t)F + (size_t)max_stack);
    const int64_t tile = (int64_t)threads * (dtype == DEX_F32 ? 4 : 2) * u;
    return (size_t)std::max<int64_t>(n_tiles * tile, 1) * (size_t)std::max(F, 1) * (dtype == DEX_F32 ? 4 : 8);
}

// The GX kernels (Float32, early exit, no parameter rows: `wide` = EVAL_WIDE_STORE / EVAL_WIDE_LOSS) keep
// only the first rows in shared memory (the stack and a few features) and read the other feature rows from
// the feature-major global copy through L1.  A row of a 2 048-sample tile costs 8 KB of shared memory: with
// more than 9 rows only two CTAs used to fit where the registers allow three, beyond 14 the block had to
// shrink.  With the shared memory out of the way the store form is compiled for FOUR resident CTAs (64
// registers, 8 bytes of spills) and serves every input, narrow ones included:
//   C4 shard (10 features + 4 stack rows): all rows in shared memory, 2 CTAs/SM 16.5 ms -> 12.97 ms
//     (4 rows kept; 6 rows 13.09; 8 rows = 3 CTAs 13.6)
//   C2 (5 features + 3 stack rows, 3 CTAs/SM before) 0.2778 -> 0.2578 ms in bench.py (3..6 rows kept:
//     within 1 %);  C6 37.85 -> 37.16 ms with 6 rows kept (3..4 rows: 37.7)
// The fused-loss form needs more registers (four CTAs: 72 bytes of spills, 46 ms against 38.5 on C6-loss):
// it stays at three CTAs and is used for wide inputs only.  The parametric form (its class indices take
// eight more registers) is compiled for three CTAs as well; with the stack and parameter rows only in
// shared memory three fit where C5's 11 rows allowed two.
constexpr int GX_SMEM_ROWS_STORE = 6;   // 4 x (6 x 8 KB + 1 KB reserved) = 196 KB of the SM's 228 KB
constexpr int GX_SMEM_ROWS_LOSS = 8;    // 3 x (8 x 8 KB + 1 KB reserved) = 195 KB

int64_t eval_num_tiles(int dtype, int32_t F, int32_t max_stack, int64_t N, int* threads_out,
                       size_t* smem_out, int wide, int* smem_rows_out) {
    const size_t es = dtype == DEX_F32 ? 4 : 8;
    const size_t all_rows = (size_t)F + (size_t)max_stack;
    const int K = (dtype == DEX_F32 ? 4 : 2) * eval_pick_u(dtype, all_rows);
    static const bool gx_off = getenv("DEXB200_NO_GX") != nullptr;
    const bool wide_ok = wide != EVAL_WIDE_NO;
    int gx_rows = std::max(wide == EVAL_WIDE_LOSS ? GX_SMEM_ROWS_LOSS : GX_SMEM_ROWS_STORE, max_stack);
    if (const char* env = getenv("DEXB200_GX_ROWS")) gx_rows = std::max(max_stack, std::min(atoi(env), 9));   // tuning knob
    bool gx = wide_ok && !gx_off && dtype == DEX_F32 && K == 8 && all_rows > 0 && max_stack <= 14 &&
              (wide != EVAL_WIDE_LOSS || (all_rows > 9 && (int64_t)all_rows > gx_rows)) &&
              N >= (int64_t)DEX_MAX_THREADS * K && N < ((int64_t)1 << 30) - DEX_MAX_THREADS * K;
    gx_rows = (int)std::min<size_t>((size_t)gx_rows, all_rows);
    int threads = 256;
    bool forced = false;
    if (const char* env = getenv("DEXB200_THREADS")) {   // tuning knob for experiments
        const int v = atoi(env);
        if (v >= 32 && v <= DEX_MAX_THREADS && v % 32 == 0) { threads = v; forced = true; gx = gx && v == DEX_MAX_THREADS; }
    }
    const size_t rows = gx ? (size_t)gx_rows : all_rows;
    // keep >= 2 CTAs resident per SM when possible; shrink the block if the rows do not fit
    const size_t budget = forced ? SMEM_LIMIT : SMEM_LIMIT / 2;
    while (threads > 32 && rows * (size_t)threads * K * es > budget) threads >>= 1;
    if (N < (int64_t)threads * K) {  // tiny inputs: do not stage more columns than exist
        while (threads > 32 && (int64_t)(threads / 2) * K >= N) threads >>= 1;
    }
    size_t smem = rows * (size_t)threads * K * es;
    if (smem == 0) smem = 16;
    if (threads_out) *threads_out = threads;
    if (smem_out) *smem_out = smem;
    if (smem_rows_out) *smem_rows_out = gx ? gx_rows : 0;
    const int64_t tile = (int64_t)threads * K;
    return (N + tile - 1) / tile;
}

cudaError_t launch_eval(const EvalArgs& e, cudaStream_t stream, int sm_count, int* launches) {
    (void)sm_count;
    int threads;
    size_t smem;
    int smem_rows = 0;
    const int wide = eval_wide_mode(e.early_exit != 0, e.params != nullptr, e.y != nullptr);
    const int64_t n_tiles = eval_num_tiles(e.dtype, e.F, e.max_stack + e.n_param_rows, e.N, &threads, &smem, wide, &smem_rows);
    if (smem > SMEM_LIMIT) return cudaErrorInvalidConfiguration;
    if (e.n_trees == 0 || e.N == 0) return cudaSuccess;
    const int u = eval_pick_u(e.dtype, (size_t)e.F + (size_t)e.max_stack + (size_t)e.n_param_rows);
    const int64_t tile = (int64_t)threads * (e.dtype == DEX_F32 ? 4 : 2) * u;
    const int64_t Npad = n_tiles * tile;
    const int64_t cover = std::max<int64_t>(Npad, e.n_trees);
    cudaError_t err = cudaSuccess;
    if (!e.skip_prepass) {
        if (e.dtype == DEX_F32)
            transpose_pad_kernel<float><<<(unsigned)((cover + 255) / 256), 256, 0, stream>>>(
                static_cast<const float*>(e.X), e.ldx, e.F, e.N, static_cast<float*>(e.xt), Npad, e.ok, e.n_trees,
                const_cast<Instr*>(e.tape), e.ctape, e.seg, e.seg_off, e.fold_ok);
        else
            transpose_pad_kernel<double><<<(unsigned)((cover + 255) / 256), 256, 0, stream>>>(
                static_cast<const double*>(e.X), e.ldx, e.F, e.N, static_cast<double*>(e.xt), Npad, e.ok, e.n_trees,
                const_cast<Instr*>(e.tape), e.ctape, e.seg, e.seg_off, e.fold_ok);
        err = cudaGetLastError();
        if (err != cudaSuccess) return err;
        if (launches) *launches += 1;
    }
    EvalArgs k = e;   // the interpreter reads the staged copy
    k.X = e.xt;
    k.ldx = Npad;
    k.smem_rows = smem_rows;   // > 0: the wide-input kernel (eval_num_tiles)
    err = e.dtype == DEX_F64 ? launch_typed<double, EVAL_U>(k, stream, threads, smem, n_tiles)
          : u == 1           ? launch_typed<float, 1>(k, stream, threads, smem, n_tiles)
                             : launch_typed<float, EVAL_U>(k, stream, threads, smem, n_tiles);
    if (err == cudaSuccess && launches) *launches += 1;
    return err;
}

template <typename T, bool GRAD>
__global__ void fold_kernel(Instr* tape, const Instr* ctape, const int64_t* seg, const int64_t* seg_off,
                            int64_t n_trees, uint8_t* fold_ok) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_trees) fold_ok[t] = fold_tree<T, GRAD>(tape, ctape, seg, seg_off, t) ? 1 : 0;
}

// The attribute belongs to the (device, function), not to the calling thread: the record of what
// has been granted is process-wide and the limit is only ever raised — a thread that needs less
// must not lower it under a launch that another thread has sized for more.
cudaError_t ensure_dynamic_smem(const void* kernel, size_t bytes) {
    static std::mutex m;
    static std::map<std::pair<int, const void*>, size_t> granted;
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    std::lock_guard<std::mutex> lk(m);
    size_t& have = granted[std::make_pair(dev, kernel)];
    if (bytes <= have) return cudaSuccess;
    err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (err == cudaSuccess) have = bytes;
    return err;
}

cudaError_t launch_fold(int dtype, bool grad_rule, Instr* tape, const Instr* ctape, const int64_t* seg,
                        const int64_t* seg_off, int64_t n_trees, uint8_t* fold_ok, cudaStream_t stream) {
    if (n_trees == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((n_trees + 63) / 64);
    if (dtype == DEX_F32) {
        if (grad_rule) fold_kernel<float, true><<<blocks, 64, 0, stream>>>(tape, ctape, seg, seg_off, n_trees, fold_ok);
        else fold_kernel<float, false><<<blocks, 64, 0, stream>>>(tape, ctape, seg, seg_off, n_trees, fold_ok);
    } else {
        if (grad_rule) fold_kernel<double, true><<<blocks, 64, 0, stream>>>(tape, ctape, seg, seg_off, n_trees, fold_ok);
        else fold_kernel<double, false><<<blocks, 64, 0, stream>>>(tape, ctape, seg, seg_off, n_trees, fold_ok);
    }
    return cudaGetLastError();
}

cudaError_t launch_scatter_constants(int dtype, Instr* tape, Instr* scalar_tape, const int64_t* pos,
                                     const void* values, int64_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (dtype == DEX_F32)
        scatter_constants_kernel<float><<<blocks, 256, 0, stream>>>(tape, scalar_tape, pos, static_cast<const float*>(values), n);
    else
        scatter_constants_kernel<double><<<blocks, 256, 0, stream>>>(tape, scalar_tape, pos, static_cast<const double*>(values), n);
    return cudaGetLastError();
}

}  // namespace dex
