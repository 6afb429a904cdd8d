/* dexb200.h — C ABI of libdexb200.so: the B200-native batched replacement for the
 * evaluation hot path of DynamicExpressions.jl.
 *
 * This is the drop-in boundary.  The reference is pure Julia; what its FFI for
 * this path would bind (a package extension in the style of
 * /root/reference/ext/DynamicExpressionsBumperExt.jl that overloads
 * eval_tree_array / eval_grad_tree_array and `ccall`s a shared library) is exactly
 * the set of entry points below.  INTEGRATION.md shows the Julia-side binding.
 *
 * Conventions
 *   - plain C: opaque handles, raw pointers, sizes; no C++/torch types.
 *   - every function returns DEX_OK (0) or a negative DEX_ERR_* code and never
 *     throws or exits; dex_last_error(ctx) has the message of the last failure.
 *   - indices inside the ABI are 0-BASED (the Julia shim subtracts 1).
 *   - X is column-major (nfeatures x nsamples) with leading dimension ldx, i.e.
 *     exactly the memory of the reference's `cX::Matrix{T}`
 *     (/root/reference/src/Evaluate.jl:124-128: sample stride = nfeatures).
 *   - results are (n_trees x nsamples) row-major with row stride ldo: row t is the
 *     `Vector{T}` the reference returns for tree t.
 *   - `_dev` pointers are device memory on the context's device, owned by the
 *     caller; `_host` pointers are host memory.  The library owns only its context
 *     scratch and the packed populations.
 *   - a context is single-threaded (one per host thread, like the reference's
 *     per-task Bumper slab); different contexts are independent => re-entrant.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     returns DEX_ERR_CUDA.
 */
#ifndef DEXB200_H
#define DEXB200_H

#include <stdint.h>
#include "dex_wire.h"

#ifdef __cplusplus
extern "C" {
#endif

#define DEXB200_ABI_VERSION 1

typedef struct dex_ctx dex_ctx;
typedef struct dex_optable dex_optable;
typedef struct dex_population dex_population;

enum {
    DEX_OK = 0,
    DEX_ERR_INVALID = -1,     /* bad argument / malformed tree (index in message) */
    DEX_ERR_NOMEM = -2,
    DEX_ERR_CUDA = -3,        /* CUDA runtime failure or no device */
    DEX_ERR_UNSUPPORTED = -4, /* operator without device implementation, tree too deep, ... */
    DEX_ERR_RANGE = -5        /* feature / parameter / class index out of range */
};

/* Evaluation policy = the flags of the reference's EvalContext
 * (/root/reference/src/Evaluate.jl:156-181).                                        */
enum {
    DEX_EVAL_EARLY_EXIT = 1, /* early_exit=Val(true): `complete` is false when any checked
                                intermediate is non-finite (reference default).  As in the
                                reference (`@return_on_nonfinite_array`, Evaluate.jl:26-32),
                                evaluation of such a tree stops early: the result (and gradient)
                                rows of a tree whose flag is 0 are UNSPECIFIED.  Without the flag
                                every row is computed to the end, non-finite values included.   */
    DEX_EVAL_SKIP_INCOMPLETE = 2, /* dex_eval_host only, with EARLY_EXIT: the rows of trees whose flag
                                is 0 are not transferred to the host at all (they are unspecified
                                anyway, and the device->host link is what that entry point is
                                bound by); out_host keeps whatever it held in those rows.      */
    DEX_EVAL_DEFAULT = 1
};

/* Pack-time policy: which validity checks the reference path performs depends on how
 * it fuses nodes (see DESIGN.md "completion flag").                                  */
enum {
    DEX_PACK_FUSED = 1,   /* use_fused=Val(true), the reference default                   */
    DEX_PACK_BUMPER = 2,  /* semantics of the Bumper evaluator
                             (/root/reference/ext/DynamicExpressionsBumperExt.jl:11-89)    */
    DEX_PACK_DEFAULT = 1
};
/* ParametricExpression populations whose gradients will be taken: OR this into pack_flags with
 * n = size(parameters, 1), so that every parameter of the expression has its row (and its
 * gradient direction) even when the trees do not use all of them.                       */
#define DEX_PACK_PARAM_ROWS(n) (((int)(n) & 0xffff) << 8)

/* gradient modes: `variable` of eval_grad_tree_array
 * (/root/reference/src/EvaluateDerivative.jl:193-228)                                */
enum { DEX_GRAD_CONSTANTS = 0, DEX_GRAD_FEATURES = 1, DEX_GRAD_BOTH = 2 };

/* ---- library ---------------------------------------------------------------------- */
int dex_abi_version(void);
const char* dex_strerror(int code);
/* number of CUDA devices visible (0 when there is no GPU / driver) */
int dex_device_count(void);
/* builtin opcode for a Julia function name at a given arity, or -1
 * (`(degree, op_idx) -> opcode` is what replaces OperatorEnum's tuple of functions,
 * /root/reference/src/OperatorEnum.jl:14-49)                                         */
int dex_opcode_from_name(const char* name, int degree);
const char* dex_opcode_name(int opcode);
int dex_opcode_degree(int opcode);

/* ---- context ---------------------------------------------------------------------- */
/* device < 0: host-only context (packing / validation only, no CUDA calls).           */
int dex_ctx_create(int device, dex_ctx** out);
int dex_ctx_destroy(dex_ctx* ctx);
/* stream: a cudaStream_t; NULL is CUDA's legacy default stream.  All device work of later
 * calls is enqueued on it and is asynchronous w.r.t. the host, except the *_host
 * convenience entry points, which synchronise.  A new context uses a private
 * non-blocking stream; dex_ctx_use_own_stream returns to it.                          */
int dex_ctx_set_stream(dex_ctx* ctx, void* stream);
int dex_ctx_use_own_stream(dex_ctx* ctx);
int dex_ctx_synchronize(dex_ctx* ctx);
const char* dex_last_error(const dex_ctx* ctx);

/* ---- operators -------------------------------------------------------------------- */
/* opcodes[degree_offsets[d-1] .. degree_offsets[d]) are the builtin opcodes of
 * operators[d] in order; max_degree <= DEX_MAX_DEGREE.                               */
int dex_optable_create(const int32_t* opcodes, const int32_t* degree_offsets, int max_degree,
                       dex_optable** out);
int dex_optable_destroy(dex_optable* t);

/* ---- populations ------------------------------------------------------------------ */
/* Flatten n_trees wire trees (nodes[offsets[t] .. offsets[t+1])) into device tapes.
 * dtype: DEX_F32 / DEX_F64.  Validates arity, operator indices and leaf kinds
 * (feature / parameter ranges are checked at evaluation time against the actual X).
 * replaces: the per-call recursive walk of /root/reference/src/Evaluate.jl:337-364.   */
int dex_population_pack(dex_ctx* ctx, const dex_optable* ops, const dex_node* nodes,
                        const int64_t* offsets, int64_t n_trees, int dtype, int pack_flags,
                        dex_population** out);
int dex_population_destroy(dex_population* pop);

typedef struct dex_population_info {
    int64_t n_trees;
    int64_t n_nodes;          /* sum of count_nodes(tree) — the node-ops/sample of the metric */
    int64_t n_instructions;   /* fused evaluation tape length                                   */
    int64_t n_constants;      /* sum of count_constant_nodes(tree)                              */
    int32_t max_stack;        /* operand-stack rows the evaluation tape needs                   */
    int32_t max_feature;      /* largest feature index used (0-based), -1 if none               */
    int32_t max_parameter;    /* largest parameter index used (0-based), -1 if none             */
    int32_t dtype;
    int64_t n_generic;        /* instructions executed by the generic (non-specialised) handler */
    int64_t n_checks;         /* validity checks left after host-side elision                   */
    /* the image dex_eval* run: constant subtrees (Evaluate.jl:347-354) folded into scalars   */
    int64_t n_folded_instructions;  /* sample-loop instructions after folding                  */
    int64_t n_scalar_instructions;  /* instructions of the folded subtrees (run once per call)  */
    int64_t n_folded_subtrees;
    int32_t folded_max_stack;       /* operand-stack rows of the folded image                   */
    int32_t reserved0;
} dex_population_info;
int dex_population_get_info(const dex_population* pop, dex_population_info* info);
/* per-tree count_constant_nodes (/root/reference/src/NodeUtils.jl:43-51); counts[n_trees] */
int dex_population_constant_counts(const dex_population* pop, int32_t* counts);
/* get/set_scalar_constants (/root/reference/src/NodeUtils.jl:99-143) for the whole
 * population without re-packing: values are in tree order, then depth-first
 * left-to-right leaf order inside a tree; n_values must equal info.n_constants.
 * Element type = the population dtype.                                               */
int dex_population_get_constants(dex_ctx* ctx, const dex_population* pop, void* values_host,
                                 int64_t n_values);
int dex_population_set_constants(dex_ctx* ctx, dex_population* pop, const void* values_host,
                                 int64_t n_values);

/* ---- evaluation ------------------------------------------------------------------- */
/* Batched eval_tree_array (/root/reference/src/Evaluate.jl:279-309), one launch for the
 * population: out_dev[t*ldo + j] = tree_t(X[:, j]);  ok_dev[t] = `complete` flag.      */
int dex_eval(dex_ctx* ctx, const dex_population* pop, const void* X_dev, int32_t nfeatures,
             int64_t nsamples, int64_t ldx, void* out_dev, int64_t ldo, uint8_t* ok_dev,
             int eval_flags);

/* ParametricExpression evaluation (/root/reference/src/ParametricExpression.jl:371-390)
 * without materialising parameters[:, classes] or the vcat: params_dev holds, per tree,
 * an (n_params x n_classes) column-major block; classes_dev[j] in [0, n_classes).       */
int dex_eval_parametric(dex_ctx* ctx, const dex_population* pop, const void* X_dev,
                        int32_t nfeatures, int64_t nsamples, int64_t ldx, const void* params_dev,
                        int32_t n_params, int32_t n_classes, const int32_t* classes_dev,
                        void* out_dev, int64_t ldo, uint8_t* ok_dev, int eval_flags);

/* Batched eval_grad_tree_array (/root/reference/src/EvaluateDerivative.jl:193-404).
 * Tree t writes a (G_t x nsamples) column-major block (gradient index fastest) at
 * grad_dev + grad_offsets[t] (element offsets, host array of n_trees+1 entries);
 * G_t = nfeatures | n_constants(t) | nfeatures + n_constants(t) by mode.              */
int dex_eval_grad(dex_ctx* ctx, const dex_population* pop, const void* X_dev, int32_t nfeatures,
                  int64_t nsamples, int64_t ldx, int mode, void* out_dev, int64_t ldo,
                  void* grad_dev, const int64_t* grad_offsets_host, uint8_t* ok_dev);
/* Batched eval_grad_tree_array of ParametricExpressions: the reference differentiates
 *   eval_tree_array(convert(Node, ex), vcat(parameters[:, classes], X), operators)
 * (/root/reference/src/ParametricExpression.jl:305-350, 380-389 through the pullback of
 * /root/reference/src/ChainRules.jl:56-77), in which the per-sample parameter rows are the FIRST
 * n_params "features".  Same here, without materialising that matrix: the feature directions are
 * [d/d parameter row 0 .. n_params-1, d/dX row 0 .. nfeatures-1], then the constants; G_t =
 * n_params + nfeatures | n_constants(t) | both, by mode.  grad_offsets_host =
 * dex_grad_offsets(pop, n_params + nfeatures, nsamples, mode).  The population must have been
 * packed with DEX_PACK_PARAM_ROWS(n_params) unless its trees use every parameter.
 * d loss / d parameters[p, c] is the sum over the samples of class c of dY[j] * grad[p, j].     */
int dex_eval_grad_parametric(dex_ctx* ctx, const dex_population* pop, const void* X_dev, int32_t nfeatures,
                             int64_t nsamples, int64_t ldx, const void* params_dev, int32_t n_params,
                             int32_t n_classes, const int32_t* classes_dev, int mode, void* out_dev,
                             int64_t ldo, void* grad_dev, const int64_t* grad_offsets_host, uint8_t* ok_dev);
/* fills offsets[n_trees+1] for the layout above */
int dex_grad_offsets(const dex_population* pop, int32_t nfeatures, int64_t nsamples, int mode,
                     int64_t* offsets_host);

/* Batched eval_diff_tree_array (/root/reference/src/EvaluateDerivative.jl:40-168):
 * derivative along feature `direction` (0-based); never reports failure (ok = 1).      */
int dex_eval_diff(dex_ctx* ctx, const dex_population* pop, const void* X_dev, int32_t nfeatures,
                  int64_t nsamples, int64_t ldx, int32_t direction, void* out_dev,
                  void* dout_dev, int64_t ldo, uint8_t* ok_dev);

/* Fused loss reduction (what callers of the path consume):
 * loss_dev[t] = sum_j w_j (tree_t(X[:,j]) - y[j])^2 / sum_j w_j   (weights may be NULL),
 * without writing the (n_trees x nsamples) result matrix.  loss is float64.            */
int dex_eval_loss(dex_ctx* ctx, const dex_population* pop, const void* X_dev, int32_t nfeatures,
                  int64_t nsamples, int64_t ldx, const void* y_dev, const void* weights_dev,
                  double* loss_dev, uint8_t* ok_dev, int eval_flags);

/* Fused loss AND its gradient (what constant optimisation consumes — the
 * contraction the reference's pullback performs, /root/reference/src/ChainRules.jl:56-77, with
 * dY = d loss / d y, used by /root/reference/ext/DynamicExpressionsOptimExt.jl and
 * test/test_optim.jl:44-52):
 *   loss_dev[t]                 = sum_j w_j (tree_t(X[:,j]) - y[j])^2 / sum_j w_j
 *   grad_dev[offsets[t] + g]    = d loss_dev[t] / d theta_g      (g as in dex_eval_grad, `mode`)
 * Neither the (n_trees x nsamples) values nor the (G x nsamples) gradients are written: the
 * reduction happens in the interpreter, deterministically (per-tile partial sums in float64,
 * summed in tile order).  grad_offsets_host = dex_grad_offsets(pop, nfeatures, 1, mode).
 * weights_dev may be NULL (w_j = 1).  `ok` as in dex_eval_grad.                              */
int dex_eval_loss_grad(dex_ctx* ctx, const dex_population* pop, const void* X_dev, int32_t nfeatures,
                       int64_t nsamples, int64_t ldx, const void* y_dev, const void* weights_dev, int mode,
                       double* loss_dev, double* grad_dev, const int64_t* grad_offsets_host,
                       uint8_t* ok_dev);

/* ---- host-buffer convenience (the reference-facing call: host arrays in, host arrays
 * out; copies are issued on the context stream through pinned staging buffers and the
 * call returns after the results have landed)                                          */
int dex_eval_host(dex_ctx* ctx, const dex_population* pop, const void* X_host, int32_t nfeatures,
                  int64_t nsamples, int64_t ldx, void* out_host, int64_t ldo, uint8_t* ok_host,
                  int eval_flags);

/* The same over several devices of ONE process (the reference is
 * single-device: callers write `[eval_tree_array(t, X, ops) for t in trees]`).  ctxs[d] / pops[d]:
 * one context and one packed copy of the SAME trees per device.  Device d evaluates the contiguous
 * column block [N d / R, N (d+1) / R) of X and its rows land in place in out_host; all devices are
 * enqueued before any is waited for.  ok_host[t] = complete on every shard.                    */
int dex_shard_eval_host(dex_ctx* const* ctxs, const dex_population* const* pops, int32_t n_devices,
                        const void* X_host, int32_t nfeatures, int64_t nsamples, int64_t ldx, void* out_host,
                        int64_t ldo, uint8_t* ok_host, int eval_flags);

/* ... and with the result left on ONE device (SURVEY.md §8b `dex_shard_eval` + `dex_gather` in one
 * call).  X_devs[d]: device d's column block [N d / R, N (d+1) / R) of X, resident on device d,
 * column-major with leading dimension ldx.  out_root_dev / ok_root_dev live on the device of
 * ctxs[root]: every device's interpreter kernel stores its block of the (n_trees x nsamples) result
 * straight into that matrix through peer memory (NVLink), so the gather overlaps the arithmetic and
 * nothing is staged; ok_root_dev[t] = complete on every shard.  Asynchronous: ordered behind the work
 * already enqueued on ctxs[root]'s stream, complete in that stream's order.  The multi-PROCESS form
 * of the same scheme uses dex_ipc_export / dex_ipc_open below.                                   */
int dex_shard_eval(dex_ctx* const* ctxs, const dex_population* const* pops, int32_t n_devices,
                   const void* const* X_devs, int32_t nfeatures, int64_t nsamples, int64_t ldx,
                   void* out_root_dev, int64_t ldo, uint8_t* ok_root_dev, int32_t root, int eval_flags);

/* plain copies on the context's stream for hosts without a CUDA binding of their own (a Julia
 * extension holding dex_device_alloc'ed buffers as raw pointers): to_device is asynchronous (the
 * source must stay valid until the next synchronising call), to_host returns when the bytes
 * have landed                                                                                */
int dex_copy_to_device(dex_ctx* ctx, void* dst_dev, const void* src_host, int64_t bytes);
int dex_copy_to_host(dex_ctx* ctx, void* dst_host, const void* src_dev, int64_t bytes);

/* pinned (page-locked) host buffers, so the copies of the *_host entry points run at full
 * PCIe speed and asynchronously */
int dex_host_alloc(void** out, int64_t bytes);
int dex_host_free(void* p);

/* ---- peer memory: fused evaluate + gather over NVLink ---------------------
 * The sample axis shards across GPUs, one process per GPU
 * (/root/reference has no multi-device path; callers write
 * `[eval_tree_array(t, X, ops) for t in trees]`, benchmark/benchmarks.jl:76-91).  When the caller
 * wants the reference's single (n_trees x nsamples) result on ONE device, no collective is
 * needed after the kernel: rank r passes `out_dev = root_buffer + first_column_r` with
 * `ldo = nsamples_total` to dex_eval, and the interpreter's result stores land directly in the
 * root GPU's memory through NVLink peer mapping, overlapped with the arithmetic tile by tile.
 * These three calls provide the mapping across processes (CUDA IPC):
 *   dex_device_alloc   plain cudaMalloc'ed buffer on the context's device (IPC-exportable:
 *                      not a sub-allocation of a caching allocator)
 *   dex_ipc_export     64-byte handle of such a buffer, to be sent to the peer processes
 *   dex_ipc_open       maps a peer's buffer into this process (enables peer access);
 *                      the returned pointer is valid as `out_dev` / `ok_dev` of dex_eval*
 *   dex_ipc_close      unmaps it                                                        */
#define DEX_IPC_HANDLE_BYTES 64
int dex_device_alloc(dex_ctx* ctx, void** out, int64_t bytes);
int dex_device_free(dex_ctx* ctx, void* p);
int dex_ipc_export(dex_ctx* ctx, const void* dev_ptr, uint8_t* handle64);
int dex_ipc_open(dex_ctx* ctx, const uint8_t* handle64, void** out);
int dex_ipc_close(dex_ctx* ctx, void* p);

/* ---- introspection (tests, benchmarks) -------------------------------------------------- */
/* kernels launched through this context so far */
int64_t dex_ctx_launch_count(const dex_ctx* ctx);
/* launch geometry dex_eval / dex_eval_loss choose for a population (no device needed): threads per CTA,
 * dynamic shared memory per CTA in bytes, sample tiles, and how many rows of a tile live in shared memory
 * (0 = all of them; > 0 = the wide-input layout: the remaining feature rows are read from the global copy
 * through L1).  eval_flags as for dex_eval; `loss` != 0 describes dex_eval_loss; `parametric` != 0
 * dex_eval_parametric.  Returns DEX_OK or DEX_ERR_INVALID. */
int dex_eval_launch_info(const dex_population* pop, int32_t nfeatures, int64_t nsamples, int eval_flags,
                         int loss, int parametric, int32_t* threads, int64_t* smem_bytes, int64_t* n_tiles,
                         int32_t* smem_rows);
/* name of an interpreter handler id ("ADD_AR", "COS_R", ...; csrc/dex_tape.h), NULL if none */
const char* dex_handler_name(int handler);
/* copies the host image of the evaluation tape (16-byte instructions, csrc/dex_tape.h);
 * returns the instruction count; offsets (n_trees+1) may be NULL */
int64_t dex_population_copy_tape(const dex_population* pop, void* instrs, int64_t capacity,
                                 int64_t* offsets);
/* host image of what dex_eval* execute: the folded tape (n_folded_instructions, offsets
 * n_trees+1), the scalar tape (n_scalar_instructions) and its segment table — per folded
 * subtree three int64: scalar-tape begin, end, and the folded-tape instruction whose inline
 * constant receives the result; seg_offsets (n_trees+1) delimits the subtrees of each tree.
 * Any pointer may be NULL. */
int dex_population_copy_folded(const dex_population* pop, void* instrs, int64_t* offsets,
                               void* scalar_instrs, int64_t* segs, int64_t* seg_offsets);

#ifdef __cplusplus
}
#endif
#endif /* DEXB200_H */
