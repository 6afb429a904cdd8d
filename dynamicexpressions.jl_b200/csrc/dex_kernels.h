// dex_kernels.h — launch interface between the C ABI (dex_api.cu) and the CUDA
// kernels (dex_eval.cu, dex_grad.cu).  Internal.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "dex_tape.h"

namespace dex {

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) costs about a microsecond of driver time: issue it
// only when a (device, kernel) pair needs more than it was last given.  Thread-local: a context is
// single-threaded and different host threads may drive different devices.
cudaError_t ensure_dynamic_smem(const void* kernel, size_t bytes);

struct EvalArgs {
    int dtype;                  // DEX_F32 / DEX_F64
    const Instr* tape;          // device
    const int64_t* tape_off;    // device, n_trees + 1
    // constant-subtree folding (null seg_off => nothing to fold): the prepass runs the scalar
    // tape and stores the results into the constant slots of `tape`
    const Instr* ctape;         // device
    const int64_t* seg;         // device, 3 per folded subtree: ctape begin, end, target in `tape`
    const int64_t* seg_off;     // device, n_trees + 1
    // folding already done for the current constants (launch_fold): per-tree outcome; when set,
    // the prepass copies it into ok[] instead of running the scalar tape
    const uint8_t* fold_ok;
    int64_t n_trees;
    const int32_t* chunk_start; // device, n_chunks + 1 (tree index ranges, balanced by tape length)
    int32_t n_chunks;
    int32_t max_stack;          // stack rows in front of the parameter and feature rows
    int32_t n_param_rows;       // rows between the stack rows and the feature rows
    const void* X;              // device, column-major F x N, leading dimension ldx
    void* xt;                   // device scratch >= eval_xt_bytes(): feature-major padded copy of X
    int32_t F;
    int64_t N;
    int64_t ldx;
    void* out;                  // device, n_trees x N row-major (row stride ldo); may be null in loss mode
    int64_t ldo;
    uint8_t* ok;                // device, n_trees (pre-set to 1 by the launcher)
    int32_t early_exit;
    // ParametricExpression (null params => plain evaluation)
    const void* params;         // device, per tree (n_params x n_classes) column-major
    int32_t n_params;
    int32_t n_classes;
    const int32_t* classes;     // device, N
    // fused loss (null => store results)
    const void* y;              // device, N
    const void* w;              // device, N or null
    double* loss_partial;       // device, n_tiles x n_trees partial sums (deterministic 2-stage)
    int32_t sync_tree;          // barrier per tree (short tapes: see dex_eval.cu KArgs)
    int32_t skip_prepass;       // the transposed copy of X and ok[] are already prepared (slice > 0)
    // launch shape chosen by the launcher
    int32_t threads;
    // chosen by launch_eval: rows kept in shared memory by the wide-input kernel (0 = all of them)
    int32_t smem_rows;
};

// Chooses the block size / shared memory, presets ok[], launches.  Returns cudaError_t.
cudaError_t launch_eval(const EvalArgs& a, cudaStream_t stream, int sm_count, int* launches);
// which GX kernel a launch may take (dex_eval.cu eval_num_tiles): none (Float64, early_exit = false), the
// store form, the fused-loss form or the parametric form
enum { EVAL_WIDE_NO = 0, EVAL_WIDE_STORE = 1, EVAL_WIDE_LOSS = 2, EVAL_WIDE_PARAM = 3 };
inline int eval_wide_mode(bool early_exit, bool has_params, bool loss) {
    if (!early_exit) return (has_params || loss) ? EVAL_WIDE_NO : EVAL_WIDE_STORE;
    if (has_params && loss) return EVAL_WIDE_NO;
    return has_params ? EVAL_WIDE_PARAM : loss ? EVAL_WIDE_LOSS : EVAL_WIDE_STORE;
}
size_t eval_xt_bytes(int dtype, int32_t F, int32_t max_stack, int64_t N, int wide);
// number of sample tiles launch_eval will use for (dtype, F, max_stack, N)
// `max_stack` here = stack rows + parameter rows (every row in front of the features)
// `wide`: the wide-input kernel the launch may take (it keeps only *smem_rows_out rows in shared memory;
// 0 = every row is in shared memory)
int64_t eval_num_tiles(int dtype, int32_t F, int32_t max_stack, int64_t N, int* threads_out,
                       size_t* smem_out, int wide, int* smem_rows_out = nullptr);

struct GradArgs {
    int dtype;
    const Instr* tape;          // device: the evaluation tape
    const int64_t* tape_off;    // device, n_trees + 1
    // constant-folded launch (d/dX only; null seg_off => `tape` is the unfolded image): the
    // prepass runs the scalar tape with the gradient path's validity rule (dex_fold.cuh)
    const Instr* ctape;
    const int64_t* seg;
    const int64_t* seg_off;
    const uint8_t* fold_ok;     // as in EvalArgs (gradient rule)
    const int32_t* const_ord;   // device, per tape instruction: tree-local constant ordinal or -1
    const int64_t* const_off;   // device, n_trees + 1 (constant ordinal base per tree)
    int64_t n_trees;
    int32_t max_stack;
    const void* X;
    void* xt;                   // device scratch >= grad_xt_bytes()
    int32_t F;
    int64_t N;
    int64_t ldx;
    int32_t mode;               // DEX_GRAD_*; -1 = eval_diff along `direction`
    int32_t direction;
    void* out;                  // n_trees x N
    int64_t ldo;
    void* grad;                 // per tree (G_t x N) column-major at grad_off[t]; diff: n_trees x N rows
    const int64_t* grad_off;    // device, n_trees + 1 (unused for diff)
    uint8_t* ok;
    // fused loss + gradient of the loss (null partial => store value rows and gradient blocks):
    // grad_off then indexes the per-tree gradient VECTORS (dex_grad_offsets with nsamples = 1)
    const void* y;              // device, N targets
    const void* w;              // device, N weights or null
    double* partial;            // device, n_tiles x partial_stride
    int64_t partial_stride;     // n_trees + total gradient entries
    // ParametricExpression (null params => plain trees): the per-sample parameter rows
    // parameters[p, classes[j]] are leaf rows in FRONT of the features, exactly the
    // vcat(indexed_parameters, X) of /root/reference/src/ParametricExpression.jl:380-385, so
    // d/d(parameter row) are the first n_param_rows feature directions
    const void* params;         // device, per tree (n_params x n_classes) column-major
    const int32_t* classes;     // device, N
    int32_t n_params;
    int32_t n_classes;
    int32_t n_param_rows;
};
// chunk_start: device table of n_chunks + 1 tree indices; Gmax: largest gradient count of any tree
cudaError_t launch_grad_ex(const GradArgs& a, const int32_t* chunk_start, int n_chunks, int Gmax,
                           cudaStream_t stream, int* launches);
int64_t grad_num_tiles(int dtype, int F, int max_stack, int Gmax, int64_t N, bool loss = false);
size_t grad_xt_bytes(int dtype, int F, int max_stack, int Gmax, int64_t N, bool loss = false);
cudaError_t launch_loss_grad_reduce(const double* partial, int64_t n_tiles, int64_t stride, int64_t n_trees,
                                    double inv_n, const double* wsum, double* loss, double* grad,
                                    cudaStream_t stream);
cudaError_t launch_weight_sum(int dtype, const void* w, int64_t n, double* out, cudaStream_t stream);

// Runs the scalar tape of every tree once (one thread per tree): stores the folded constants into
// `tape` and the per-tree outcome into fold_ok.  grad_rule: dex_fold.cuh.  The result stays valid
// until the constants change.
cudaError_t launch_fold(int dtype, bool grad_rule, Instr* tape, const Instr* ctape, const int64_t* seg,
                        const int64_t* seg_off, int64_t n_trees, uint8_t* fold_ok, cudaStream_t stream);

// tiny helpers
// pos[i] >= 0: tape[pos[i]]; pos[i] < 0: scalar_tape[-(1 + pos[i])] (folded image)
cudaError_t launch_scatter_constants(int dtype, Instr* tape, Instr* scalar_tape, const int64_t* pos,
                                     const void* values, int64_t n, cudaStream_t stream);
}  // namespace dex
