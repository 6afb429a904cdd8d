/* dex_oracle.h — CPU oracle for the expression-tree evaluation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * CPU algorithm (DynamicExpressions.jl v2.9.2).  It is linked/called only by
 * tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline /
 * `--impl reference` legs.  The product (libdexb200.so) never links, loads or
 * calls anything in this directory.
 *
 * Parity status: PINNED against the reference's own closed-form known-answer
 * tests (tests/golden/reference_known_answers.json, transcribed from
 * /root/reference/test/ and /root/reference/docs/, see tests/golden/README.md).
 * The reference itself (Julia) cannot run in this environment (no julia
 * binary, no network), so there are no reference-generated vectors; the
 * reference's tests are closed-form formulas, which is what we check against.
 *
 * Functions and the reference code they follow:
 *   dexo_eval_tree_array        src/Evaluate.jl:279-309 (entry), :337-364 (recursion),
 *                               :428-651 (dispatch / fused-kernel choice),
 *                               :693-993 (loop kernels), :1002-1067 (constant folding),
 *                               ext/DynamicExpressionsBumperExt.jl:11-89 (bumper=1)
 *   dexo_eval_diff_tree_array   src/EvaluateDerivative.jl:40-168
 *   dexo_eval_grad_tree_array   src/EvaluateDerivative.jl:193-404,
 *                               constant numbering src/NodeUtils.jl:184-201
 *   dexo_eval_parametric        src/ParametricExpression.jl:305-324, 371-390
 *   is_valid / is_valid_array   src/ValueInterface.jl:5-9
 *
 * Trees arrive in the wire format of include/dex_wire.h (preorder dex_node
 * arrays, 0-based op / feature indices); operators as a per-degree table of
 * builtin opcodes from include/dex_ops.def.
 */
#ifndef DEX_ORACLE_H
#define DEX_ORACLE_H

#include <stdint.h>
#include "../include/dex_wire.h"

#ifdef __cplusplus
extern "C" {
#endif

/* operator table: ops[d-1][i] = builtin opcode of operators[d][i+1] (Julia indexing) */
typedef struct dexo_optable {
    int32_t nops[DEX_MAX_DEGREE];
    const int32_t* ops[DEX_MAX_DEGREE];
} dexo_optable;

/* EvalContext flags (src/Evaluate.jl:156-181) */
enum {
    DEXO_EARLY_EXIT = 1, /* early_exit = Val(true) (reference default)  */
    DEXO_USE_FUSED = 2,  /* use_fused  = Val(true) (reference default)  */
    DEXO_BUMPER = 4,     /* bumper     = Val(true): unfused post-order evaluator */
    DEXO_ELEMENTWISE = 8 /* NOT a reference mode: array validity = all elements finite
                            (instead of isfinite(sum)); isolates sum-overflow cases      */
};

/* gradient modes (src/EvaluateDerivative.jl:200-202) */
enum { DEXO_GRAD_CONSTANTS = 0, DEXO_GRAD_FEATURES = 1, DEXO_GRAD_BOTH = 2,
       DEXO_GRAD_ELEMENTWISE = 8 /* or-ed into mode: same meaning as DEXO_ELEMENTWISE */ };

/* All entry points: dtype = DEX_F32 / DEX_F64; X is column-major F x N with
 * leading dimension ldx (>= F).  Return 0 on success, <0 on malformed input.
 * *ok receives the reference's `complete` flag.                               */

int dexo_eval_tree_array(const dex_node* nodes, int64_t n_nodes, const dexo_optable* ops,
                         int dtype, const void* X, int32_t nfeatures, int64_t nsamples,
                         int64_t ldx, int flags, void* out, uint8_t* ok);

int dexo_eval_diff_tree_array(const dex_node* nodes, int64_t n_nodes, const dexo_optable* ops,
                              int dtype, const void* X, int32_t nfeatures, int64_t nsamples,
                              int64_t ldx, int32_t direction /*0-based*/, void* out, void* dout,
                              uint8_t* ok);

/* grad is (G x N) column-major, gradient index fastest.  *n_grad_out = G.      */
int dexo_eval_grad_tree_array(const dex_node* nodes, int64_t n_nodes, const dexo_optable* ops,
                              int dtype, const void* X, int32_t nfeatures, int64_t nsamples,
                              int64_t ldx, int mode, void* out, void* grad, int64_t grad_capacity,
                              int32_t* n_grad_out, uint8_t* ok);

/* parameters: (n_params x n_classes) column-major; classes: length N, 0-based. */
int dexo_eval_parametric(const dex_node* nodes, int64_t n_nodes, const dexo_optable* ops,
                         int dtype, const void* X, int32_t nfeatures, int64_t nsamples,
                         int64_t ldx, const void* parameters, int32_t n_params,
                         int32_t n_classes, const int32_t* classes, int flags, void* out,
                         uint8_t* ok);

/* Population forms (a serial or OpenMP loop over trees — the comprehension of
 * benchmark/benchmarks.jl:76-91).  offsets has n_trees+1 entries into nodes.
 * out is (n_trees x N) row-major.  nthreads <= 0 means all cores.
 * params (may be NULL): per tree a (n_params x n_classes) block.               */
int dexo_eval_population(const dex_node* nodes, const int64_t* offsets, int64_t n_trees,
                         const dexo_optable* ops, int dtype, const void* X, int32_t nfeatures,
                         int64_t nsamples, int64_t ldx, int flags, int nthreads, void* out,
                         uint8_t* ok);

/* conditioning yardstick for the Float64 parity tests: same algorithm, 80-bit intermediates */
int dexo_eval_population_f80(const dex_node* nodes, const int64_t* offsets, int64_t n_trees,
                             const dexo_optable* ops, const double* X, int32_t nfeatures,
                             int64_t nsamples, int64_t ldx, int flags, int nthreads, double* out,
                             uint8_t* ok);

int dexo_eval_grad_population(const dex_node* nodes, const int64_t* offsets, int64_t n_trees,
                              const dexo_optable* ops, int dtype, const void* X,
                              int32_t nfeatures, int64_t nsamples, int64_t ldx, int mode,
                              int nthreads, void* out, void* grad, const int64_t* grad_offsets,
                              uint8_t* ok);

int dexo_eval_parametric_population(const dex_node* nodes, const int64_t* offsets,
                                    int64_t n_trees, const dexo_optable* ops, int dtype,
                                    const void* X, int32_t nfeatures, int64_t nsamples,
                                    int64_t ldx, const void* parameters, int32_t n_params,
                                    int32_t n_classes, const int32_t* classes, int flags,
                                    int nthreads, void* out, uint8_t* ok);

/* count_constant_nodes (src/NodeUtils.jl:43-51) for a wire tree */
int32_t dexo_count_constants(const dex_node* nodes, int64_t n_nodes);
/* scalar operator application, for op-table unit tests */
double dexo_apply_f64(int opcode, double a, double b, double c);
float dexo_apply_f32(int opcode, float a, float b, float c);
/* partial derivatives (the analytic stand-in for Zygote's scalar rules,
 * ext/DynamicExpressionsZygoteExt.jl:7-15) */
void dexo_partials_f64(int opcode, double a, double b, double c, double* g);
int dexo_max_threads(void);

/* TEST-ONLY: move every transcendental unary result by n ulps (0 = off); see dex_oracle_ops.inc */
void dexo_set_ulp_nudge(int n);

#ifdef __cplusplus
}
#endif
#endif
