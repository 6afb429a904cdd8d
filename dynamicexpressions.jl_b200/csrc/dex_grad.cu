// dex_grad.cu — batched forward-mode derivative evaluation (sm_100a).
//
// Replaces eval_grad_tree_array / eval_diff_tree_array
// (/root/reference/src/EvaluateDerivative.jl:40-168, 193-404) for a whole population in
// one launch.  The reference allocates and zero-fills a (G x N) matrix at EVERY leaf
// (:376-380) and streams two of them through memory at every operator (:340-365); here
// the dual numbers (value, d/d theta_1..GC) of one sample live in one thread's column of
// a shared-memory stack and never touch HBM: the only traffic is X in, value + gradient out.
//
// Mapping
//   grid.x  sample tiles (one sample per thread)          grid.y  chunks of trees
//   the gradient directions of a tree are processed in passes of GC <= 8 directions
//   (the primal is recomputed per pass), so shared memory stays bounded for trees with
//   many constants.
//   stack slot s, component c (0 = value, 1.. = directions), thread t:
//        stk[(s * (1 + GC) + c) * RS + t],  RS = blockDim.x + 1   (conflict-free both for
//        the per-thread walk and for the transposed, coalesced gradient store)
// Validity (`complete`): every value and every gradient component of every node,
// leaves included, must be finite (:238-243); eval_diff never checks (:68-85).
#include "dex_kernels.h"
#include "dex_ops.cuh"
#include "../../include/dexb200.h"

#include <algorithm>

namespace dex {
namespace {

template <typename T> struct GK {
    const uint4* gtape;
    const int64_t* gtape_off;
    const int64_t* const_off;
    const int32_t* chunk_start;
    const T* X;
    T* out;
    T* grad;
    const int64_t* grad_off;
    uint8_t* ok;
    int64_t N, ldx, ldo;
    int32_t F, max_gstack, mode, direction, GC;
};

template <typename T> __device__ __forceinline__ T gconst_of(const uint4& ins);
template <> __device__ __forceinline__ float gconst_of<float>(const uint4& ins) { return __uint_as_float(ins.z); }
template <> __device__ __forceinline__ double gconst_of<double>(const uint4& ins) { return __hiloint2double((int)ins.w, (int)ins.z); }

template <typename T>
__global__ void __launch_bounds__(128) grad_kernel(const GK<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* stk = reinterpret_cast<T*>(smem_raw);
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int RS = nthr + 1;
    const int GC = a.GC;
    const int SLOT = (1 + GC) * RS;
    T* xs = stk + (size_t)a.max_gstack * SLOT;  // F rows of nthr
    const int64_t s0 = (int64_t)blockIdx.x * nthr;
    {
        const int F = a.F, total = F * nthr;
        int s = tid / F, f = tid - s * F;
        const int ds = nthr / F, df = nthr - ds * F;
        for (int idx = tid; idx < total; idx += nthr) {
            int64_t gs = s0 + s;
            if (gs >= a.N) gs = a.N - 1;
            xs[(size_t)f * nthr + s] = __ldg(a.X + gs * a.ldx + f);
            s += ds; f += df;
            if (f >= F) { f -= F; ++s; }
        }
    }
    __syncthreads();
    const int64_t gs_mine = s0 + tid;
    const bool in_range = gs_mine < a.N;
    const int t0 = a.chunk_start[blockIdx.y], t1 = a.chunk_start[blockIdx.y + 1];
    const int mode = a.mode;

    for (int t = t0; t < t1; ++t) {
        const int64_t off = a.gtape_off[t];
        const int n = (int)(a.gtape_off[t + 1] - off);
        const uint4* ip = a.gtape + off;
        const int nconst = (int)(a.const_off[t + 1] - a.const_off[t]);
        const int G = mode < 0 ? 1 : mode == DEX_GRAD_FEATURES ? a.F
                    : mode == DEX_GRAD_CONSTANTS ? nconst : a.F + nconst;
        T nf = T(0);
        const int npass = G > 0 ? (G + GC - 1) / GC : 1;
        for (int pass = 0; pass < npass; ++pass) {
            const int g0 = pass * GC;
            const int gc = min(GC, G - g0);  // live directions in this pass (may be <= 0 when G == 0)
            for (int pc = 0; pc < n; ++pc) {
                const uint4 ins = __ldg(ip + pc);
                const uint32_t op = ins.x & 0xffu;
                T* S = stk + (size_t)(ins.x >> 16) * SLOT + tid;
                if (op == 0) {  // LOAD leaf: value + one-hot seed (grad_deg0_eval :367-404)
                    const uint32_t kind = (ins.x >> 8) & 3u;
                    T v;
                    int index = -1;  // global gradient row seeded by this leaf
                    if (kind == DEX_LEAF_CONST) {
                        v = gconst_of<T>(ins);
                        if (mode == DEX_GRAD_CONSTANTS) index = (int)ins.y;
                        else if (mode == DEX_GRAD_BOTH) index = a.F + (int)ins.y;
                    } else {
                        v = xs[(size_t)ins.y * nthr + tid];
                        if (mode == DEX_GRAD_FEATURES || mode == DEX_GRAD_BOTH) index = (int)ins.y;
                        else if (mode < 0 && (int)ins.y == a.direction) index = 0;
                    }
                    S[0] = v;
                    nf = m_fma(v, T(0), nf);
                    const int local = index - g0;
                    for (int g = 0; g < gc; ++g) S[(1 + g) * RS] = (g == local) ? T(1) : T(0);
                    continue;
                }
                T v, p0, p1 = T(0), p2 = T(0);
                int deg;
                {
                    const T x = S[0];
                    T y = T(0), z = T(0);
                    if (op >= 64u) y = S[SLOT];
                    if (op >= 128u) z = S[2 * SLOT];
                    deg = op >= 128u ? 3 : (op >= 64u ? 2 : 1);
                    switch (op) {
#define U_CASE(SYM, VEXPR, GEXPR) case DEX_OP_##SYM: { v = (VEXPR); p0 = (GEXPR); } break;
                        DEX_UNARY_OPS(U_CASE)
#undef U_CASE
#define B_CASE(SYM, VEXPR, G0, G1) case DEX_OP_##SYM: { v = (VEXPR); p0 = (G0); p1 = (G1); } break;
                        DEX_BINARY_OPS(B_CASE)
#undef B_CASE
#define T_CASE(SYM, VEXPR, G0, G1, G2) case DEX_OP_##SYM: { v = (VEXPR); p0 = (G0); p1 = (G1); p2 = (G2); } break;
                        DEX_TERNARY_OPS(T_CASE)
#undef T_CASE
                        default: v = t_nan<T>(); p0 = t_nan<T>(); break;
                    }
                }
                S[0] = v;
                nf = m_fma(v, T(0), nf);
                // d[k] = sum_i partial_i * d_i[k]   (grad_degn_eval :355-361)
                if (deg == 1) {
                    for (int g = 0; g < gc; ++g) {
                        const T d = p0 * S[(1 + g) * RS];
                        S[(1 + g) * RS] = d;
                        nf = m_fma(d, T(0), nf);
                    }
                } else if (deg == 2) {
                    for (int g = 0; g < gc; ++g) {
                        const T d = p0 * S[(1 + g) * RS] + p1 * S[SLOT + (1 + g) * RS];
                        S[(1 + g) * RS] = d;
                        nf = m_fma(d, T(0), nf);
                    }
                } else {
                    for (int g = 0; g < gc; ++g) {
                        const T d = (p0 * S[(1 + g) * RS] + p1 * S[SLOT + (1 + g) * RS]) +
                                    p2 * S[2 * SLOT + (1 + g) * RS];
                        S[(1 + g) * RS] = d;
                        nf = m_fma(d, T(0), nf);
                    }
                }
            }
            // ---- outputs of this pass: slot 0 ------------------------------------------
            if (pass == 0 && in_range) a.out[(size_t)t * a.ldo + gs_mine] = stk[tid];
            if (gc > 0) {
                __syncthreads();  // slot 0 columns are read across threads below
                if (mode < 0) {
                    // eval_diff: one derivative row per tree, laid out like `out`
                    if (in_range) a.grad[(size_t)t * a.ldo + gs_mine] = stk[RS + tid];
                } else {
                    // (G x N) column-major block: element (g0+g, s0+s) at s*G + g0 + g.
                    T* gout = a.grad + a.grad_off[t] + s0 * G + g0;
                    const int total = nthr * gc;
                    int s = tid / gc, g = tid - s * gc;
                    const int ds = nthr / gc, dg = nthr - ds * gc;
                    for (int idx = tid; idx < total; idx += nthr) {
                        if (s0 + s < a.N) gout[(size_t)s * G + g] = stk[(1 + g) * RS + s];
                        s += ds; g += dg;
                        if (g >= gc) { g -= gc; ++s; }
                    }
                }
                __syncthreads();  // before the next pass / tree overwrites slot 0
            }
        }
        if (mode >= 0) {
            const bool bad = nf != nf;
            if (__any_sync(0xffffffffu, bad) && (tid & 31) == 0) a.ok[t] = 0;
        }
    }
}

__global__ void gfill_u8_kernel(uint8_t* p, int64_t n, uint8_t v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

constexpr size_t G_SMEM_LIMIT = 227 * 1024;

struct GradShape { int threads; int GC; size_t smem; };

GradShape pick_shape(int dtype, int F, int max_gstack, int Gmax) {
    const size_t es = dtype == DEX_F32 ? 4 : 8;
    GradShape s;
    s.threads = 128;
    s.GC = std::max(1, std::min(Gmax, 8));
    auto bytes = [&](int th, int gc) {
        return ((size_t)max_gstack * (1 + gc) * (th + 1) + (size_t)F * th) * es;
    };
    while (bytes(s.threads, s.GC) > 96 * 1024) {
        if (s.GC > 2) s.GC = (s.GC + 1) / 2;
        else if (s.threads > 32) s.threads >>= 1;
        else if (s.GC > 1) s.GC = 1;
        else break;
    }
    s.smem = bytes(s.threads, s.GC);
    return s;
}

template <typename T>
cudaError_t launch_grad_typed(const GradArgs& g, const int32_t* chunk_start, int n_chunks,
                              const GradShape& sh, int64_t n_tiles, cudaStream_t stream) {
    GK<T> a;
    a.gtape = reinterpret_cast<const uint4*>(g.gtape);
    a.gtape_off = g.gtape_off;
    a.const_off = g.const_off;
    a.chunk_start = chunk_start;
    a.X = static_cast<const T*>(g.X);
    a.out = static_cast<T*>(g.out);
    a.grad = static_cast<T*>(g.grad);
    a.grad_off = g.grad_off;
    a.ok = g.ok;
    a.N = g.N; a.ldx = g.ldx; a.ldo = g.ldo;
    a.F = g.F; a.max_gstack = g.max_gstack; a.mode = g.mode; a.direction = g.direction; a.GC = sh.GC;
    cudaError_t err = cudaFuncSetAttribute(grad_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh.smem);
    if (err != cudaSuccess) return err;
    dim3 grid((unsigned)n_tiles, (unsigned)n_chunks);
    grad_kernel<T><<<grid, sh.threads, sh.smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_grad_ex(const GradArgs& a, const int32_t* chunk_start, int n_chunks, int Gmax,
                           cudaStream_t stream, int* launches) {
    if (a.n_trees == 0 || a.N == 0) return cudaSuccess;
    const GradShape sh = pick_shape(a.dtype, a.F, a.max_gstack, std::max(Gmax, 1));
    if (sh.smem > G_SMEM_LIMIT) return cudaErrorInvalidConfiguration;
    const int64_t n_tiles = (a.N + sh.threads - 1) / sh.threads;
    gfill_u8_kernel<<<(unsigned)((a.n_trees + 255) / 256), 256, 0, stream>>>(a.ok, a.n_trees, 1);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    if (launches) *launches += 1;
    err = a.dtype == DEX_F32 ? launch_grad_typed<float>(a, chunk_start, n_chunks, sh, n_tiles, stream)
                             : launch_grad_typed<double>(a, chunk_start, n_chunks, sh, n_tiles, stream);
    if (err == cudaSuccess && launches) *launches += 1;
    return err;
}

int64_t grad_num_tiles(int dtype, int F, int max_gstack, int Gmax, int64_t N) {
    const GradShape sh = pick_shape(dtype, F, max_gstack, std::max(Gmax, 1));
    return (N + sh.threads - 1) / sh.threads;
}

}  // namespace dex
