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
//   thread  K consecutive samples (K * sizeof(T) = 16 bytes => one LDS.128 per operand,
//           one STG.128 per result); accumulator, operands and validity accumulators
//           live in registers; the operand stack is rows of the same shared array
//   CTA     walks the tapes of its chunk; the tape pointer depends only on blockIdx and
//           loop counters, so instruction fetch/decode/branch are warp-uniform
//           (no divergence on the op switch).
// No tensor cores: the path is elementwise, not a contraction.
#include "dex_kernels.h"
#include "dex_ops.cuh"

#include <algorithm>

namespace dex {

namespace {

template <typename T> struct KOf;
template <> struct KOf<float> { static constexpr int K = 4; };
template <> struct KOf<double> { static constexpr int K = 2; };

template <typename T, int K> __device__ __forceinline__ void ld_row(T (&v)[K], const T* p) {
    static_assert(sizeof(T) * K == 16, "row vectors are 16 bytes");
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    *reinterpret_cast<uint4*>(v) = u;
}
template <typename T, int K> __device__ __forceinline__ void st_row(T* p, const T (&v)[K]) {
    *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(v);
}

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
    int32_t F, max_stack, early_exit, n_params, n_classes;
};

template <typename T, bool PARAM, bool LOSS>
__global__ void __launch_bounds__(256) eval_kernel(const KArgs<T> a) {
    constexpr int K = KOf<T>::K;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* rows = reinterpret_cast<T*>(smem_raw);
    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    const int TILE = nthr * K;
    const int64_t s0 = (int64_t)blockIdx.x * TILE;

    // ---- stage the X slab feature-major: xs[f][s] = X[f, s0 + s] -------------------
    {
        T* xs = rows + (size_t)a.max_stack * TILE;
        const int F = a.F;
        const int total = F * TILE;
        int s = tid / F, f = tid - s * F;
        const int ds = nthr / F, df = nthr - ds * F;
        for (int idx = tid; idx < total; idx += nthr) {
            int64_t gs = s0 + s;
            if (gs >= a.N) gs = a.N - 1;  // tail lanes replay the last valid sample
            xs[(size_t)f * TILE + s] = __ldg(a.X + gs * a.ldx + f);
            s += ds;
            f += df;
            if (f >= F) { f -= F; ++s; }
        }
    }
    int cls[K];
    if (PARAM) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            int64_t gs = s0 + (int64_t)tid * K + k;
            if (gs >= a.N) gs = a.N - 1;
            cls[k] = __ldg(a.classes + gs) * a.n_params;
        }
    }
    T yv[K], wv[K];
    if (LOSS) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            int64_t gs = s0 + (int64_t)tid * K + k;
            const bool in = gs < a.N;
            if (!in) gs = a.N - 1;
            yv[k] = __ldg(a.y + gs);
            wv[k] = in ? (a.w ? __ldg(a.w + gs) : T(1)) : T(0);
        }
    }
    __syncthreads();

    T* my = rows + tid * K;  // this thread's 16-byte column inside every row
    const int t0 = a.chunk_start[blockIdx.y], t1 = a.chunk_start[blockIdx.y + 1];
    const bool early = a.early_exit != 0;
    const bool full_tile = (s0 + TILE <= a.N) && ((a.ldo % K) == 0) &&
                           ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0);

    for (int t = t0; t < t1; ++t) {
        const int64_t off = a.tape_off[t];
        const int n = (int)(a.tape_off[t + 1] - off);
        const uint4* ip = a.tape + off;
        const T* ptree = PARAM ? a.params + (size_t)t * a.n_params * a.n_classes : nullptr;
        T acc[K], nf[K];
#pragma unroll
        for (int k = 0; k < K; ++k) { acc[k] = T(0); nf[k] = T(0); }

        uint4 ins = __ldg(ip);
        for (int pc = 0; pc < n; ++pc) {
            uint4 nxt = ins;
            if (pc + 1 < n) nxt = __ldg(ip + pc + 1);  // prefetch the next instruction
            const uint32_t w0 = ins.x;
            if (w0 & F_PUSH) st_row<T, K>(my + (size_t)(w0 >> 24) * TILE, acc);
            const T c = const_of<T>(ins);
            T va[K], vb[K], r[K];
            // operand A
            {
                const uint32_t src = (w0 >> 8) & 3u, row = ins.y & 0xffffu;
                if (src == SRC_ROW) ld_row<T, K>(va, my + (size_t)row * TILE);
                else if (src == SRC_CONST) {
#pragma unroll
                    for (int k = 0; k < K; ++k) va[k] = c;
                } else if (PARAM && src == SRC_PARAM) {
#pragma unroll
                    for (int k = 0; k < K; ++k) va[k] = __ldg(ptree + cls[k] + row);
                } else {
#pragma unroll
                    for (int k = 0; k < K; ++k) va[k] = acc[k];
                }
            }
            // operand B
            {
                const uint32_t src = (w0 >> 10) & 3u, row = ins.y >> 16;
                if (src == SRC_ROW) ld_row<T, K>(vb, my + (size_t)row * TILE);
                else if (src == SRC_CONST) {
#pragma unroll
                    for (int k = 0; k < K; ++k) vb[k] = c;
                } else if (PARAM && src == SRC_PARAM) {
#pragma unroll
                    for (int k = 0; k < K; ++k) vb[k] = __ldg(ptree + cls[k] + row);
                } else {
#pragma unroll
                    for (int k = 0; k < K; ++k) vb[k] = acc[k];
                }
            }
            const bool chk = early || (w0 & F_ALWAYS);
            if (chk && (w0 & F_CHK_A)) {
#pragma unroll
                for (int k = 0; k < K; ++k) nf[k] = m_fma(va[k], T(0), nf[k]);
            }
            if (chk && (w0 & F_CHK_B)) {
#pragma unroll
                for (int k = 0; k < K; ++k) nf[k] = m_fma(vb[k], T(0), nf[k]);
            }
            switch (w0 & 0xffu) {
#define U_CASE(SYM, VEXPR, GEXPR)                                   \
    case DEX_OP_##SYM: {                                            \
        _Pragma("unroll") for (int k = 0; k < K; ++k) {             \
            const T x = va[k];                                      \
            r[k] = (VEXPR);                                         \
        }                                                           \
    } break;
                DEX_UNARY_OPS(U_CASE)
#undef U_CASE
#define B_CASE(SYM, VEXPR, G0, G1)                                  \
    case DEX_OP_##SYM: {                                            \
        _Pragma("unroll") for (int k = 0; k < K; ++k) {             \
            const T x = va[k], y = vb[k];                           \
            r[k] = (VEXPR);                                         \
        }                                                           \
    } break;
                DEX_BINARY_OPS(B_CASE)
#undef B_CASE
#define T_CASE(SYM, VEXPR, G0, G1, G2)                              \
    case DEX_OP_##SYM: {                                            \
        _Pragma("unroll") for (int k = 0; k < K; ++k) {             \
            const T x = va[k], y = vb[k], z = acc[k];               \
            r[k] = (VEXPR);                                         \
        }                                                           \
    } break;
                DEX_TERNARY_OPS(T_CASE)
#undef T_CASE
                default: {
#pragma unroll
                    for (int k = 0; k < K; ++k) r[k] = t_nan<T>();
                } break;
            }
            if (w0 & F_GUARD) {
#pragma unroll
                for (int k = 0; k < K; ++k)
                    if (!t_finite(va[k])) r[k] = t_inf<T>();
            }
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] = r[k];
            if (chk && (w0 & F_CHK_OUT)) {
#pragma unroll
                for (int k = 0; k < K; ++k) nf[k] = m_fma(r[k], T(0), nf[k]);
            }
            ins = nxt;
        }

        // ---- result row segment ----------------------------------------------------
        if (!LOSS) {
            T* o = a.out + (size_t)t * a.ldo + s0 + (size_t)tid * K;
            if (full_tile) {
                __stcs(reinterpret_cast<float4*>(o), *reinterpret_cast<const float4*>(acc));
            } else {
#pragma unroll
                for (int k = 0; k < K; ++k)
                    if (s0 + (int64_t)tid * K + k < a.N) o[k] = acc[k];
            }
        } else {
            double ls = 0.0;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const double d = (double)acc[k] - (double)yv[k];
                ls += (double)wv[k] * d * d;
            }
            // deterministic block reduction -> loss_partial[tile][tree]
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ls += __shfl_xor_sync(0xffffffffu, ls, o);
            __shared__ double red[8];
            __syncthreads();
            if ((tid & 31) == 0) red[tid >> 5] = ls;
            __syncthreads();
            if (tid == 0) {
                double s = 0.0;
                for (int wdx = 0; wdx < (nthr + 31) / 32; ++wdx) s += red[wdx];
                a.loss_partial[(size_t)blockIdx.x * a.n_trees + t] = s;
            }
        }
        bool bad = false;
#pragma unroll
        for (int k = 0; k < K; ++k) bad |= (nf[k] != nf[k]);
        if (__any_sync(0xffffffffu, bad) && (tid & 31) == 0) a.ok[t] = 0;
    }
}

__global__ void fill_u8_kernel(uint8_t* p, int64_t n, uint8_t v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

template <typename T>
__global__ void scatter_constants_kernel(Instr* tape, const int64_t* pos, GInstr* gtape,
                                         const int64_t* gpos, const T* values, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T v = values[i];
    uint32_t lo, hi;
    if (sizeof(T) == 4) { lo = __float_as_uint((float)v); hi = 0; }
    else { lo = (uint32_t)__double2loint((double)v); hi = (uint32_t)__double2hiint((double)v); }
    if (pos[i] >= 0) { tape[pos[i]].c_lo = lo; tape[pos[i]].c_hi = hi; }
    if (gpos[i] >= 0) { gtape[gpos[i]].c_lo = lo; gtape[gpos[i]].c_hi = hi; }
}

__global__ void loss_reduce_kernel(const double* partial, int64_t n_tiles, int64_t n_trees,
                                   double denom_inv, double* loss) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_trees) return;
    double s = 0.0;
    for (int64_t i = 0; i < n_tiles; ++i) s += partial[i * n_trees + t];
    loss[t] = s * denom_inv;
}

constexpr size_t SMEM_LIMIT = 227 * 1024;

template <typename T>
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
    a.F = e.F; a.max_stack = e.max_stack; a.early_exit = e.early_exit;
    a.n_params = e.n_params; a.n_classes = e.n_classes;
    dim3 grid((unsigned)n_tiles, (unsigned)e.n_chunks);
    const bool param = e.params != nullptr, loss = e.y != nullptr;
    void (*kern)(const KArgs<T>) =
        loss ? (param ? eval_kernel<T, true, true> : eval_kernel<T, false, true>)
             : (param ? eval_kernel<T, true, false> : eval_kernel<T, false, false>);
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    kern<<<grid, threads, smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace

int64_t eval_num_tiles(int dtype, int32_t F, int32_t max_stack, int64_t N, int* threads_out,
                       size_t* smem_out) {
    const int K = dtype == DEX_F32 ? 4 : 2;
    const size_t es = dtype == DEX_F32 ? 4 : 8;
    const size_t rows = (size_t)F + (size_t)max_stack;
    int threads = 256;
    // keep >= 2 CTAs resident per SM when possible; shrink the block if the rows do not fit
    while (threads > 32 && rows * (size_t)threads * K * es > SMEM_LIMIT / 2) threads >>= 1;
    if (N < (int64_t)threads * K) {  // tiny inputs: do not stage more columns than exist
        while (threads > 32 && (int64_t)(threads / 2) * K >= N) threads >>= 1;
    }
    size_t smem = rows * (size_t)threads * K * es;
    if (smem == 0) smem = 16;
    if (threads_out) *threads_out = threads;
    if (smem_out) *smem_out = smem;
    const int64_t tile = (int64_t)threads * K;
    return (N + tile - 1) / tile;
}

cudaError_t launch_eval(const EvalArgs& e, cudaStream_t stream, int sm_count, int* launches) {
    (void)sm_count;
    int threads;
    size_t smem;
    const int64_t n_tiles = eval_num_tiles(e.dtype, e.F, e.max_stack, e.N, &threads, &smem);
    if (smem > SMEM_LIMIT) return cudaErrorInvalidConfiguration;
    if (e.n_trees == 0 || e.N == 0) return cudaSuccess;
    fill_u8_kernel<<<(unsigned)((e.n_trees + 255) / 256), 256, 0, stream>>>(e.ok, e.n_trees, 1);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    if (launches) *launches += 1;
    err = e.dtype == DEX_F32 ? launch_typed<float>(e, stream, threads, smem, n_tiles)
                             : launch_typed<double>(e, stream, threads, smem, n_tiles);
    if (err == cudaSuccess && launches) *launches += 1;
    return err;
}

cudaError_t launch_scatter_constants(int dtype, Instr* tape, const int64_t* pos, GInstr* gtape,
                                     const int64_t* gpos, const void* values, int64_t n,
                                     cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (dtype == DEX_F32)
        scatter_constants_kernel<float><<<blocks, 256, 0, stream>>>(tape, pos, gtape, gpos, static_cast<const float*>(values), n);
    else
        scatter_constants_kernel<double><<<blocks, 256, 0, stream>>>(tape, pos, gtape, gpos, static_cast<const double*>(values), n);
    return cudaGetLastError();
}

cudaError_t launch_loss_reduce(const double* partial, int64_t n_tiles, int64_t n_trees,
                               double denom_inv, double* loss, cudaStream_t stream) {
    if (n_trees == 0) return cudaSuccess;
    loss_reduce_kernel<<<(unsigned)((n_trees + 255) / 256), 256, 0, stream>>>(partial, n_tiles, n_trees, denom_inv, loss);
    return cudaGetLastError();
}

}  // namespace dex
