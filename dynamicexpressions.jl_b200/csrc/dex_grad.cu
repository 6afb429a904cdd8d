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
// consumer, Sethi-Ullman order, absolute stack rows): an operand is ACC, a stack slot, a
// feature row, or an inline constant, and a leaf's derivative is a one-hot (or zero) vector
// that is never materialised in memory.
//
// Mapping
//   grid.x  sample tiles of blockDim.x * K samples (K = 16 bytes / sizeof(T))
//   grid.y  chunks of trees
//   thread  K consecutive samples; ACC = (1 + GC) * K registers
//   smem    stack slot s component c: row s * (1 + GC) + c;  features behind the stack rows
//   GC      compile-time number of directions per pass (1..8); trees with more directions
//           (many constants) take several passes, recomputing the primal per pass
// Semantics: every value and every gradient component of every node, leaves included, must
// be finite for `complete` (:238-243) — including products with the zeros of a one-hot
// (Inf * 0 = NaN is a failure in the reference, and here); eval_diff never checks (:68-85).
#include "dex_kernels.h"
#include "dex_ops.cuh"
#include "../../include/dexb200.h"

#include <algorithm>

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
    int64_t N, ldx, ldo;
    int32_t F, max_stack, mode, direction;
};

template <typename T> __device__ __forceinline__ T gconst_of(const uint4& ins);
template <> __device__ __forceinline__ float gconst_of<float>(const uint4& ins) { return __uint_as_float(ins.z); }
template <> __device__ __forceinline__ double gconst_of<double>(const uint4& ins) { return __hiloint2double((int)ins.w, (int)ins.z); }

template <typename T, int K> __device__ __forceinline__ void ldv(T (&v)[K], const T* p) {
    *reinterpret_cast<uint4*>(v) = *reinterpret_cast<const uint4*>(p);
}
template <typename T, int K> __device__ __forceinline__ void stv(T* p, const T (&v)[K]) {
    *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(v);
}

// how the derivative of an operand is obtained
enum : int { DK_ACC = 0, DK_SLOT = 1, DK_LEAF = 2 };

// d = p * x   and   d = d + p * x   over the K samples of a thread.  Products and the sum are
// rounded separately, like the reference's `grad_1 * d_1 + grad_2 * d_2`; Float32 uses
// Blackwell's packed FMUL2 / FADD2 / FFMA2 (two IEEE operations per instruction).
template <typename T, int K> struct DOps {
    static __device__ __forceinline__ void mul(T (&d)[K], const T (&p)[K], const T (&x)[K]) {
#pragma unroll
        for (int k = 0; k < K; ++k) d[k] = p[k] * x[k];
    }
    static __device__ __forceinline__ void mul_add(T (&d)[K], const T (&p)[K], const T (&x)[K]) {
#pragma unroll
        for (int k = 0; k < K; ++k) d[k] = d[k] + p[k] * x[k];
    }
    static __device__ __forceinline__ void check(T& nf, const T (&d)[K]) {
#pragma unroll
        for (int k = 0; k < K; ++k) nf = m_fma(d[k], T(0), nf);
    }
};
template <int K> struct DOps<float, K> {
    static __device__ __forceinline__ void mul(float (&d)[K], const float (&p)[K], const float (&x)[K]) {
#pragma unroll
        for (int k = 0; k < K; k += 2) {
            const float2 r = __fmul2_rn(make_float2(p[k], p[k + 1]), make_float2(x[k], x[k + 1]));
            d[k] = r.x; d[k + 1] = r.y;
        }
    }
    static __device__ __forceinline__ void mul_add(float (&d)[K], const float (&p)[K], const float (&x)[K]) {
#pragma unroll
        for (int k = 0; k < K; k += 2) {
            const float2 m = __fmul2_rn(make_float2(p[k], p[k + 1]), make_float2(x[k], x[k + 1]));
            const float2 r = __fadd2_rn(make_float2(d[k], d[k + 1]), m);
            d[k] = r.x; d[k + 1] = r.y;
        }
    }
    static __device__ __forceinline__ void check(float& nf, const float (&d)[K]) {
        float2 acc = make_float2(nf, 0.f);
#pragma unroll
        for (int k = 0; k < K; k += 2) acc = __ffma2_rn(make_float2(d[k], d[k + 1]), make_float2(0.f, 0.f), acc);
        nf = acc.x + acc.y;
    }
};

// derivative row g of one operand
template <typename T, int GC, int K, int KIND>
__device__ __forceinline__ void dsrc(T (&x)[K], const T (&ad)[GC][K], const T* drow, int idx, int TILE, int g) {
    if (KIND == DK_ACC) {
#pragma unroll
        for (int k = 0; k < K; ++k) x[k] = ad[g][k];
    } else if (KIND == DK_SLOT) {
        ldv<T, K>(x, drow + (size_t)(1 + g) * TILE);
    } else {   // one-hot / zero seed of a leaf
        const T e = (g == idx) ? T(1) : T(0);
#pragma unroll
        for (int k = 0; k < K; ++k) x[k] = e;
    }
}

// ad[g] <- p0 * dA[g] (+ p1 * dB[g] (+ p2 * ad[g]))  for every direction of the pass
template <typename T, int GC, int K, int DEG, int KA, int KB>
__device__ __forceinline__ void combine(T (&ad)[GC][K], const T (&p)[3][K], const T* const (&drow)[3],
                                        const int (&idx)[3], int TILE, T& nf, bool chk) {
#pragma unroll
    for (int g = 0; g < GC; ++g) {
        T d[K], x[K];
        dsrc<T, GC, K, KA>(x, ad, drow[0], idx[0], TILE, g);
        DOps<T, K>::mul(d, p[0], x);
        if (DEG >= 2) {
            dsrc<T, GC, K, KB>(x, ad, drow[1], idx[1], TILE, g);
            DOps<T, K>::mul_add(d, p[1], x);
        }
        if (DEG >= 3) {   // third operand is always ACC
            dsrc<T, GC, K, DK_ACC>(x, ad, drow[2], idx[2], TILE, g);
            DOps<T, K>::mul_add(d, p[2], x);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) ad[g][k] = d[k];
        if (chk) DOps<T, K>::check(nf, d);
    }
}

template <typename T, int GC>
__global__ void __launch_bounds__(128) grad_kernel(const GK<T> a) {
    constexpr int K = 16 / (int)sizeof(T);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* rows = reinterpret_cast<T*>(smem_raw);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int TILE = nthr * K;
    const int S = a.max_stack;
    T* xs = rows + (size_t)S * (1 + GC) * TILE;   // feature rows
    const int64_t s0 = (int64_t)blockIdx.x * TILE;

    // stage the feature rows of this tile (XT is tile-padded: always in range, 16 B aligned)
    for (int idx = tid; idx < a.F * nthr; idx += nthr) {
        const int f = idx / nthr, t = idx - f * nthr;
        *reinterpret_cast<uint4*>(xs + (size_t)f * TILE + t * K) =
            __ldg(reinterpret_cast<const uint4*>(a.X + (size_t)f * a.ldx + s0 + t * K));
    }
    __syncthreads();

    T* my = rows + tid * K;
    const T* myx = xs + tid * K;
    const int t0 = a.chunk_start[blockIdx.y], t1 = a.chunk_start[blockIdx.y + 1];
    const int mode = a.mode;
    const bool full_tile = (s0 + TILE <= a.N);

    for (int t = t0; t < t1; ++t) {
        const int64_t off = a.tape_off[t];
        const int n = (int)(a.tape_off[t + 1] - off);
        const uint4* ip = a.tape + off;
        const int32_t* ordp = a.const_ord + off;
        const int nconst = (int)(a.const_off[t + 1] - a.const_off[t]);
        const int G = mode < 0 ? 1 : mode == DEX_GRAD_FEATURES ? a.F
                    : mode == DEX_GRAD_CONSTANTS ? nconst : a.F + nconst;
        T nf = T(0);
        const int npass = G > 0 ? (G + GC - 1) / GC : 1;
        for (int pass = 0; pass < npass; ++pass) {
            const int g0 = pass * GC;
            T av[K], ad[GC][K];   // accumulator dual
#pragma unroll
            for (int k = 0; k < K; ++k) av[k] = T(0);
#pragma unroll
            for (int g = 0; g < GC; ++g)
#pragma unroll
                for (int k = 0; k < K; ++k) ad[g][k] = T(0);

            for (int pc = 0; pc < n; ++pc) {
                const uint4 ins = __ldg(ip + pc);
                const uint32_t w0 = ins.x;
                const uint32_t op = (w0 >> 8) & 0xffu;
                const T c = gconst_of<T>(ins);
                if (w0 & F_PUSH) {
                    T* dst = my + (size_t)push_row(w0) * (1 + GC) * TILE;
                    stv<T, K>(dst, av);
#pragma unroll
                    for (int g = 0; g < GC; ++g) stv<T, K>(dst + (size_t)(1 + g) * TILE, ad[g]);
                }
                // ---- operands: value, derivative source, one-hot index --------------------
                T xv[3][K];
                int dk[3], idx[3];
                const T* drow[3];
                const int deg = op >= 128u ? 3 : (op >= 64u ? 2 : 1);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const uint32_t src = (w0 >> (16 + 2 * i)) & 3u;
                    const int row = (int)((ins.y >> (16 * i)) & 0xffffu);
                    dk[i] = DK_LEAF; idx[i] = -1; drow[i] = my;
                    if (i >= deg) {
#pragma unroll
                        for (int k = 0; k < K; ++k) xv[i][k] = T(0);
                    } else if (src == SRC_ACC) {
#pragma unroll
                        for (int k = 0; k < K; ++k) xv[i][k] = av[k];
                        dk[i] = DK_ACC;
                    } else if (src == SRC_ROW && row < S) {
                        drow[i] = my + (size_t)row * (1 + GC) * TILE;
                        ldv<T, K>(xv[i], drow[i]);
                        dk[i] = DK_SLOT;
                    } else if (src == SRC_ROW) {   // feature leaf: grad_deg0_eval :387-399
                        const int f = row - S;
                        ldv<T, K>(xv[i], myx + (size_t)f * TILE);
                        if (mode == DEX_GRAD_FEATURES || mode == DEX_GRAD_BOTH) idx[i] = f - g0;
                        else if (mode < 0 && f == a.direction) idx[i] = 0;
                    } else {                        // constant leaf
#pragma unroll
                        for (int k = 0; k < K; ++k) xv[i][k] = c;
                        const int ord = __ldg(ordp + pc);
                        if (mode == DEX_GRAD_CONSTANTS) idx[i] = ord - g0;
                        else if (mode == DEX_GRAD_BOTH) idx[i] = a.F + ord - g0;
                    }
                }
                // third operand of a ternary operator is always ACC
                dk[2] = DK_ACC; idx[2] = -1; drow[2] = my;
#pragma unroll
                for (int k = 0; k < K; ++k) xv[2][k] = av[k];
                if (mode >= 0) {   // leaf values take part in the validity check (:238-243)
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        if (i < deg && dk[i] == DK_LEAF) {
#pragma unroll
                            for (int k = 0; k < K; ++k) nf = m_fma(xv[i][k], T(0), nf);
                        }
                }
                // ---- value and partials -------------------------------------------------------
                T vo[K], p[3][K];
                switch (op) {
#define U_CASE(SYM, VEXPR, GEXPR)                                                       \
    case DEX_OP_##SYM: {                                                                \
        _Pragma("unroll") for (int k = 0; k < K; ++k) {                                 \
            const T x = xv[0][k];                                                       \
            const T v = (VEXPR);                                                        \
            p[0][k] = (GEXPR);                                                          \
            vo[k] = v; p[1][k] = T(0); p[2][k] = T(0);                                  \
        }                                                                               \
    } break;
                    DEX_UNARY_OPS(U_CASE)
#undef U_CASE
#define B_CASE(SYM, VEXPR, G0, G1)                                                      \
    case DEX_OP_##SYM: {                                                                \
        _Pragma("unroll") for (int k = 0; k < K; ++k) {                                 \
            const T x = xv[0][k], y = xv[1][k];                                         \
            const T v = (VEXPR);                                                        \
            p[0][k] = (G0); p[1][k] = (G1);                                             \
            vo[k] = v; p[2][k] = T(0);                                                  \
        }                                                                               \
    } break;
                    DEX_BINARY_OPS(B_CASE)
#undef B_CASE
#define T_CASE(SYM, VEXPR, G0, G1, G2)                                                  \
    case DEX_OP_##SYM: {                                                                \
        _Pragma("unroll") for (int k = 0; k < K; ++k) {                                 \
            const T x = xv[0][k], y = xv[1][k], z = xv[2][k];                           \
            const T v = (VEXPR);                                                        \
            p[0][k] = (G0); p[1][k] = (G1); p[2][k] = (G2);                             \
            vo[k] = v; (void)z;                                                         \
        }                                                                               \
    } break;
                    DEX_TERNARY_OPS(T_CASE)
#undef T_CASE
                    default: {
#pragma unroll
                        for (int k = 0; k < K; ++k) { vo[k] = t_nan<T>(); p[0][k] = p[1][k] = p[2][k] = t_nan<T>(); }
                    } break;
                }
                // ---- d[g] = sum_i p_i * d_i[g]   (grad_degn_eval :355-361), left to right ------
                // the operand kinds are hoisted out of the direction loop: one branch-free,
                // fully unrolled instance of `combine` per (kind A, kind B) pair
                const bool chk = mode >= 0;
                if (deg == 1) {
                    switch (dk[0]) {
                        case DK_ACC: combine<T, GC, K, 1, DK_ACC, DK_ACC>(ad, p, drow, idx, TILE, nf, chk); break;
                        case DK_SLOT: combine<T, GC, K, 1, DK_SLOT, DK_ACC>(ad, p, drow, idx, TILE, nf, chk); break;
                        default: combine<T, GC, K, 1, DK_LEAF, DK_ACC>(ad, p, drow, idx, TILE, nf, chk); break;
                    }
                } else {
#define COMBINE_PAIR(DEG)                                                                               \
    switch (dk[0] * 3 + dk[1]) {                                                                        \
        case DK_ACC * 3 + DK_SLOT: combine<T, GC, K, DEG, DK_ACC, DK_SLOT>(ad, p, drow, idx, TILE, nf, chk); break;   \
        case DK_ACC * 3 + DK_LEAF: combine<T, GC, K, DEG, DK_ACC, DK_LEAF>(ad, p, drow, idx, TILE, nf, chk); break;   \
        case DK_SLOT * 3 + DK_ACC: combine<T, GC, K, DEG, DK_SLOT, DK_ACC>(ad, p, drow, idx, TILE, nf, chk); break;   \
        case DK_SLOT * 3 + DK_SLOT: combine<T, GC, K, DEG, DK_SLOT, DK_SLOT>(ad, p, drow, idx, TILE, nf, chk); break; \
        case DK_SLOT * 3 + DK_LEAF: combine<T, GC, K, DEG, DK_SLOT, DK_LEAF>(ad, p, drow, idx, TILE, nf, chk); break; \
        case DK_LEAF * 3 + DK_ACC: combine<T, GC, K, DEG, DK_LEAF, DK_ACC>(ad, p, drow, idx, TILE, nf, chk); break;   \
        case DK_LEAF * 3 + DK_SLOT: combine<T, GC, K, DEG, DK_LEAF, DK_SLOT>(ad, p, drow, idx, TILE, nf, chk); break; \
        case DK_LEAF * 3 + DK_LEAF: combine<T, GC, K, DEG, DK_LEAF, DK_LEAF>(ad, p, drow, idx, TILE, nf, chk); break; \
        default: combine<T, GC, K, DEG, DK_ACC, DK_ACC>(ad, p, drow, idx, TILE, nf, chk); break;        \
    }
                    if (deg == 2) { COMBINE_PAIR(2) } else { COMBINE_PAIR(3) }
#undef COMBINE_PAIR
                }
#pragma unroll
                for (int k = 0; k < K; ++k) av[k] = vo[k];
                if (mode >= 0) {
#pragma unroll
                    for (int k = 0; k < K; ++k) nf = m_fma(vo[k], T(0), nf);
                }
            }
            // ---- outputs of this pass -----------------------------------------------------
            const int64_t sbase = s0 + (int64_t)tid * K;
            if (pass == 0) {
                T* o = a.out + (size_t)t * a.ldo + sbase;
                if (full_tile && (a.ldo % K) == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0) stv<T, K>(o, av);
                else {
#pragma unroll
                    for (int k = 0; k < K; ++k) if (sbase + k < a.N) o[k] = av[k];
                }
            }
            if (mode < 0) {   // eval_diff: one derivative row per tree, laid out like `out`
                T* o = a.grad + (size_t)t * a.ldo + sbase;
#pragma unroll
                for (int k = 0; k < K; ++k) if (sbase + k < a.N) o[k] = ad[0][k];
            } else if (G > 0) {
                // (G x N) column-major block: element (g, s) at s * G + g
                T* gout = a.grad + a.grad_off[t] + sbase * G + g0;
                const int gc = min(GC, G - g0);
                if (gc == GC && G == GC && full_tile && ((reinterpret_cast<uintptr_t>(gout) & 15) == 0)) {
                    // the thread's K samples x GC directions are K*GC contiguous elements
                    T flat[K * GC];
#pragma unroll
                    for (int k = 0; k < K; ++k)
#pragma unroll
                        for (int g = 0; g < GC; ++g) flat[k * GC + g] = ad[g][k];
#pragma unroll
                    for (int q = 0; q < K * GC; q += K)
                        *reinterpret_cast<uint4*>(gout + q) = *reinterpret_cast<const uint4*>(flat + q);
                } else {
#pragma unroll
                    for (int k = 0; k < K; ++k)
                        if (sbase + k < a.N) {
#pragma unroll
                            for (int g = 0; g < GC; ++g)
                                if (g < gc) gout[(size_t)k * G + g] = ad[g][k];
                        }
                }
            }
        }
        if (mode >= 0) {
            const bool bad = nf != nf;
            if (__any_sync(0xffffffffu, bad) && (tid & 31) == 0) a.ok[t] = 0;
        }
    }
}

template <typename T>
__global__ void gtranspose_pad_kernel(const T* __restrict__ X, int64_t ldx, int F, int64_t N,
                                      T* __restrict__ XT, int64_t Npad, uint8_t* ok, int64_t n_trees) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_trees) ok[s] = 1;
    if (s >= Npad) return;
    const T* col = X + (s < N ? s : N - 1) * ldx;
    for (int f = 0; f < F; ++f) XT[(size_t)f * Npad + s] = __ldg(col + f);
}

constexpr size_t G_SMEM_LIMIT = 227 * 1024;

struct GradShape { int threads; int GC; size_t smem; int64_t tile; };

GradShape pick_shape(int dtype, int F, int max_stack, int Gmax) {
    const size_t es = dtype == DEX_F32 ? 4 : 8;
    const int K = dtype == DEX_F32 ? 4 : 2;
    GradShape s;
    s.threads = 128;
    s.GC = std::max(1, std::min(Gmax, 8));
    if (dtype == DEX_F64) {   // instantiated: 1, 2, 4, 8
        s.GC = s.GC <= 1 ? 1 : s.GC <= 2 ? 2 : s.GC <= 4 ? 4 : 8;
    } else if (s.GC == 7) {
        s.GC = 8;             // instantiated: 1..6, 8
    }
    auto bytes = [&](int th, int gc) {
        return ((size_t)max_stack * (1 + gc) + (size_t)F) * (size_t)th * K * es;
    };
    while (bytes(s.threads, s.GC) > 100 * 1024 && s.threads > 32) s.threads >>= 1;
    while (bytes(s.threads, s.GC) > G_SMEM_LIMIT && s.GC > 1) s.GC = s.GC > 4 ? 4 : s.GC > 2 ? 2 : 1;
    s.smem = std::max<size_t>(bytes(s.threads, s.GC), 16);
    s.tile = (int64_t)s.threads * K;
    return s;
}

template <typename T, int GC>
cudaError_t launch_one(const GK<T>& a, const GradShape& sh, int64_t n_tiles, int n_chunks, cudaStream_t stream) {
    cudaError_t err = cudaFuncSetAttribute(grad_kernel<T, GC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh.smem);
    if (err != cudaSuccess) return err;
    dim3 grid((unsigned)n_tiles, (unsigned)n_chunks);
    grad_kernel<T, GC><<<grid, sh.threads, sh.smem, stream>>>(a);
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
    a.N = g.N; a.ldx = Npad; a.ldo = g.ldo;
    a.F = g.F; a.max_stack = g.max_stack; a.mode = g.mode; a.direction = g.direction;
    return a;
}

}  // namespace

size_t grad_xt_bytes(int dtype, int F, int max_stack, int Gmax, int64_t N) {
    const GradShape sh = pick_shape(dtype, F, max_stack, std::max(Gmax, 1));
    const int64_t n_tiles = (N + sh.tile - 1) / sh.tile;
    return (size_t)std::max<int64_t>(n_tiles * sh.tile, 1) * (size_t)std::max(F, 1) * (dtype == DEX_F32 ? 4 : 8);
}

int64_t grad_num_tiles(int dtype, int F, int max_stack, int Gmax, int64_t N) {
    const GradShape sh = pick_shape(dtype, F, max_stack, std::max(Gmax, 1));
    return (N + sh.tile - 1) / sh.tile;
}

cudaError_t launch_grad_ex(const GradArgs& g, const int32_t* chunk_start, int n_chunks, int Gmax,
                           cudaStream_t stream, int* launches) {
    if (g.n_trees == 0 || g.N == 0) return cudaSuccess;
    const GradShape sh = pick_shape(g.dtype, g.F, g.max_stack, std::max(Gmax, 1));
    if (sh.smem > G_SMEM_LIMIT) return cudaErrorInvalidConfiguration;
    const int64_t n_tiles = (g.N + sh.tile - 1) / sh.tile;
    const int64_t Npad = n_tiles * sh.tile;
    const int64_t cover = std::max<int64_t>(Npad, g.n_trees);
    if (g.dtype == DEX_F32)
        gtranspose_pad_kernel<float><<<(unsigned)((cover + 255) / 256), 256, 0, stream>>>(
            static_cast<const float*>(g.X), g.ldx, g.F, g.N, static_cast<float*>(g.xt), Npad, g.ok, g.n_trees);
    else
        gtranspose_pad_kernel<double><<<(unsigned)((cover + 255) / 256), 256, 0, stream>>>(
            static_cast<const double*>(g.X), g.ldx, g.F, g.N, static_cast<double*>(g.xt), Npad, g.ok, g.n_trees);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    if (launches) *launches += 1;
    if (g.dtype == DEX_F32) {
        const GK<float> a = make_args<float>(g, chunk_start, Npad);
        switch (sh.GC) {
            case 1: err = launch_one<float, 1>(a, sh, n_tiles, n_chunks, stream); break;
            case 2: err = launch_one<float, 2>(a, sh, n_tiles, n_chunks, stream); break;
            case 3: err = launch_one<float, 3>(a, sh, n_tiles, n_chunks, stream); break;
            case 4: err = launch_one<float, 4>(a, sh, n_tiles, n_chunks, stream); break;
            case 5: err = launch_one<float, 5>(a, sh, n_tiles, n_chunks, stream); break;
            case 6: err = launch_one<float, 6>(a, sh, n_tiles, n_chunks, stream); break;
            default: err = launch_one<float, 8>(a, sh, n_tiles, n_chunks, stream); break;
        }
    } else {
        const GK<double> a = make_args<double>(g, chunk_start, Npad);
        switch (sh.GC) {
            case 1: err = launch_one<double, 1>(a, sh, n_tiles, n_chunks, stream); break;
            case 2: err = launch_one<double, 2>(a, sh, n_tiles, n_chunks, stream); break;
            case 4: err = launch_one<double, 4>(a, sh, n_tiles, n_chunks, stream); break;
            default: err = launch_one<double, 8>(a, sh, n_tiles, n_chunks, stream); break;
        }
    }
    if (err == cudaSuccess && launches) *launches += 1;
    return err;
}

}  // namespace dex
