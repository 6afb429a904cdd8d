/* dex_oracle.c — CPU oracle (TEST INFRASTRUCTURE ONLY; see dex_oracle.h for the
 * scope statement, the reference file:line map and the parity-pinning note).
 *
 * Build: make -C oracle   (gcc -O3 -ffp-contract=off -fopenmp; no fast-math)
 */
#define _GNU_SOURCE
#include "dex_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* OPERATOR_LIMIT_BEFORE_SLOWDOWN, src/Evaluate.jl:14 */
#define DEXO_OPERATOR_LIMIT 15

/* ---- per-tree structure derived from the preorder wire array ---------------- */
typedef struct tree_info {
    const dex_node* nodes;
    int64_t n;
    int32_t* size;     /* subtree size of node i (so children are found by skipping) */
    uint8_t* isconst;  /* is_constant(subtree i): no feature leaves, src/NodeUtils.jl:73 */
} tree_info;

static void tree_info_free(tree_info* t) {
    free(t->size);
    free(t->isconst);
    t->size = NULL;
    t->isconst = NULL;
}

/* returns index one past the subtree rooted at i, or -1 if malformed */
static int64_t tree_scan(tree_info* t, int64_t i, const dexo_optable* ops, int32_t F, int depth) {
    if (i >= t->n || depth > 100000) return -1;
    const dex_node* nd = &t->nodes[i];
    if (nd->degree > DEX_MAX_DEGREE) return -1;
    if (nd->degree == 0) {
        if (nd->kind > DEX_LEAF_PARAMETER) return -1;
        if (nd->kind == DEX_LEAF_FEATURE && F >= 0 && nd->feature >= F) return -1;
        t->size[i] = 1;
        t->isconst[i] = nd->kind == DEX_LEAF_CONST;
        return i + 1;
    }
    if (ops && nd->op >= ops->nops[nd->degree - 1]) return -1;
    int64_t j = i + 1;
    uint8_t allc = 1;
    for (int k = 0; k < nd->degree; ++k) {
        int64_t c = j;
        j = tree_scan(t, j, ops, F, depth + 1);
        if (j < 0) return -1;
        allc &= t->isconst[c];
    }
    t->size[i] = (int32_t)(j - i);
    t->isconst[i] = allc;
    return j;
}

static int tree_info_init(tree_info* t, const dex_node* nodes, int64_t n, const dexo_optable* ops,
                          int32_t F) {
    if (n <= 0) return -1;
    t->nodes = nodes;
    t->n = n;
    t->size = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    t->isconst = (uint8_t*)malloc((size_t)n);
    if (!t->size || !t->isconst) { tree_info_free(t); return -2; }
    if (tree_scan(t, 0, ops, F, 0) != n) { tree_info_free(t); return -1; }
    return 0;
}

/* constant numbering: depth-first, children left to right = preorder order of
 * constant leaves (index_constant_nodes, src/NodeUtils.jl:184-201 +
 * call_mapreducer src/base.jl:123-158) */
static int32_t number_constants(const dex_node* nodes, int64_t n, int32_t* index) {
    int32_t k = 0;
    for (int64_t i = 0; i < n; ++i) {
        int isc = nodes[i].degree == 0 && nodes[i].kind == DEX_LEAF_CONST;
        if (index) index[i] = isc ? k : -1;
        k += isc;
    }
    return k;
}

int32_t dexo_count_constants(const dex_node* nodes, int64_t n_nodes) {
    return number_constants(nodes, n_nodes, NULL);
}

/* DEXO_ELEMENTWISE: validity of an array = every element finite, instead of the
 * reference's isfinite(sum(x)).  The two differ only when a sum of finite values
 * overflows (e.g. 3000 samples of 1e36 in Float32) — the one documented deviation of the
 * device path (DESIGN.md "completion flag"); the tests use this mode to tell that case
 * apart from a real disagreement. */
static __thread int g_elementwise = 0;

/* ---- instantiate for Float32 and Float64 ------------------------------------ */
/* see dex_oracle_ops.inc (apply1): test-only conditioning yardstick */
static int dexo_ulp_nudge = 0;
void dexo_set_ulp_nudge(int n) { dexo_ulp_nudge = n; }

#define T float
#define S(name) name##_f32
#define M(fn) fn##f
#include "dex_oracle_impl.inc"
#undef T
#undef S
#undef M

#define T double
#define S(name) name##_f64
#define M(fn) fn
#include "dex_oracle_impl.inc"
#undef T
#undef S
#undef M

/* extended precision (x87 80-bit): NOT a reference path — the yardstick the parity tests use
 * to measure how sensitive a tree is to intermediate rounding (see dexo_eval_population_f80) */
#define T long double
#define S(name) name##_f80
#define M(fn) fn##l
#include "dex_oracle_impl.inc"
#undef T
#undef S
#undef M

/* ---- entry points ------------------------------------------------------------ */
int dexo_eval_tree_array(const dex_node* nodes, int64_t n_nodes, const dexo_optable* ops,
                         int dtype, const void* X, int32_t F, int64_t N, int64_t ldx, int flags,
                         void* out, uint8_t* ok) {
    tree_info tr;
    int rc = tree_info_init(&tr, nodes, n_nodes, ops, F);
    if (rc) return rc;
    if (dtype == DEX_F32) {
        ctx_f32 c;
        ctx_init_f32(&c, &tr, ops, (const float*)X, F, N, ldx, flags);
        eval_entry_f32(&c, flags, (float*)out, ok);
        ctx_free_f32(&c);
    } else {
        ctx_f64 c;
        ctx_init_f64(&c, &tr, ops, (const double*)X, F, N, ldx, flags);
        eval_entry_f64(&c, flags, (double*)out, ok);
        ctx_free_f64(&c);
    }
    tree_info_free(&tr);
    return 0;
}

static int grad_common(const dex_node* nodes, int64_t n_nodes, const dexo_optable* ops, int dtype,
                       const void* X, int32_t F, int64_t N, int64_t ldx, int mode,
                       int32_t direction, void* out, void* grad, int64_t grad_capacity,
                       int32_t* n_grad_out, uint8_t* ok) {
    tree_info tr;
    g_elementwise = (mode & DEXO_GRAD_ELEMENTWISE) != 0;
    if (mode >= 0) mode &= 3;
    int rc = tree_info_init(&tr, nodes, n_nodes, ops, F);
    if (rc) return rc;
    int32_t* cidx = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_nodes);
    int32_t nconst = number_constants(nodes, n_nodes, cidx);
    int32_t G = mode < 0 ? 1
              : mode == DEXO_GRAD_FEATURES ? F
              : mode == DEXO_GRAD_CONSTANTS ? nconst : F + nconst; /* :204-210 */
    if (n_grad_out) *n_grad_out = G;
    if (grad_capacity >= 0 && (int64_t)G * N > grad_capacity) {
        free(cidx); tree_info_free(&tr); return -4;
    }
    if (dtype == DEX_F32) {
        ctx_f32 c;
        ctx_init_f32(&c, &tr, ops, (const float*)X, F, N, ldx, DEXO_EARLY_EXIT);
        grad_entry_f32(&c, mode, direction, cidx, G, (float*)out, (float*)grad, ok);
        ctx_free_f32(&c);
    } else {
        ctx_f64 c;
        ctx_init_f64(&c, &tr, ops, (const double*)X, F, N, ldx, DEXO_EARLY_EXIT);
        grad_entry_f64(&c, mode, direction, cidx, G, (double*)out, (double*)grad, ok);
        ctx_free_f64(&c);
    }
    free(cidx);
    tree_info_free(&tr);
    return 0;
}

int dexo_eval_diff_tree_array(const dex_node* nodes, int64_t n_nodes, const dexo_optable* ops,
                              int dtype, const void* X, int32_t F, int64_t N, int64_t ldx,
                              int32_t direction, void* out, void* dout, uint8_t* ok) {
    return grad_common(nodes, n_nodes, ops, dtype, X, F, N, ldx, -1, direction, out, dout, -1,
                       NULL, ok);
}

int dexo_eval_grad_tree_array(const dex_node* nodes, int64_t n_nodes, const dexo_optable* ops,
                              int dtype, const void* X, int32_t F, int64_t N, int64_t ldx,
                              int mode, void* out, void* grad, int64_t grad_capacity,
                              int32_t* n_grad_out, uint8_t* ok) {
    if (mode < 0 || (mode & 3) > 2) return -1;
    return grad_common(nodes, n_nodes, ops, dtype, X, F, N, ldx, mode, 0, out, grad,
                       grad_capacity, n_grad_out, ok);
}

int dexo_eval_parametric(const dex_node* nodes, int64_t n_nodes, const dexo_optable* ops,
                         int dtype, const void* X, int32_t F, int64_t N, int64_t ldx,
                         const void* parameters, int32_t n_params, int32_t n_classes,
                         const int32_t* classes, int flags, void* out, uint8_t* ok) {
    /* validate parameter indices up front */
    for (int64_t k = 0; k < n_nodes; ++k)
        if (nodes[k].degree == 0 && nodes[k].kind == DEX_LEAF_PARAMETER &&
            nodes[k].feature >= n_params)
            return -1;
    if (dtype == DEX_F32)
        return parametric_entry_f32(nodes, n_nodes, ops, (const float*)X, F, N, ldx,
                                    (const float*)parameters, n_params, n_classes, classes, flags,
                                    (float*)out, ok);
    return parametric_entry_f64(nodes, n_nodes, ops, (const double*)X, F, N, ldx,
                                (const double*)parameters, n_params, n_classes, classes, flags,
                                (double*)out, ok);
}

/* ---- population loops (benchmark/benchmarks.jl:76-91 comprehension) ---------- */
int dexo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
static int pick_threads(int nthreads) {
    /* an explicit count wins (torchrun exports OMP_NUM_THREADS=1); <= 0 means the OpenMP default */
    return nthreads > 0 ? nthreads : dexo_max_threads();
}

int dexo_eval_population(const dex_node* nodes, const int64_t* offsets, int64_t n_trees,
                         const dexo_optable* ops, int dtype, const void* X, int32_t F, int64_t N,
                         int64_t ldx, int flags, int nthreads, void* out, uint8_t* ok) {
    int err = 0;
    size_t es = dtype == DEX_F32 ? 4 : 8;
    int nt = pick_threads(nthreads);
    (void)nt;
#pragma omp parallel num_threads(nt)
    {
        /* one arena per thread, reused across trees (ArrayBuffer + reset_index!) */
        ctx_f32 c32; ctx_f64 c64;
        memset(&c32, 0, sizeof(c32)); memset(&c64, 0, sizeof(c64));
#pragma omp for schedule(dynamic, 1)
        for (int64_t t = 0; t < n_trees; ++t) {
            tree_info tr;
            int rc = tree_info_init(&tr, nodes + offsets[t], offsets[t + 1] - offsets[t], ops, F);
            if (rc) { err = rc; ok[t] = 0; continue; }
            char* o = (char*)out + (size_t)t * (size_t)N * es;
            if (dtype == DEX_F32) {
                c32.tr = &tr; c32.ops = ops; c32.X = (const float*)X; c32.F = F; c32.N = N;
                c32.ldx = ldx; c32.early_exit = (flags & DEXO_EARLY_EXIT) != 0;
                c32.use_fused = (flags & DEXO_USE_FUSED) != 0;
                eval_entry_f32(&c32, flags, (float*)o, &ok[t]);
            } else {
                c64.tr = &tr; c64.ops = ops; c64.X = (const double*)X; c64.F = F; c64.N = N;
                c64.ldx = ldx; c64.early_exit = (flags & DEXO_EARLY_EXIT) != 0;
                c64.use_fused = (flags & DEXO_USE_FUSED) != 0;
                eval_entry_f64(&c64, flags, (double*)o, &ok[t]);
            }
            tree_info_free(&tr);
        }
        ctx_free_f32(&c32);
        ctx_free_f64(&c64);
    }
    return err;
}

int dexo_eval_grad_population(const dex_node* nodes, const int64_t* offsets, int64_t n_trees,
                              const dexo_optable* ops, int dtype, const void* X, int32_t F,
                              int64_t N, int64_t ldx, int mode, int nthreads, void* out,
                              void* grad, const int64_t* grad_offsets, uint8_t* ok) {
    int err = 0;
    size_t es = dtype == DEX_F32 ? 4 : 8;
    int nt = pick_threads(nthreads);
    (void)nt;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
    for (int64_t t = 0; t < n_trees; ++t) {
        int rc = dexo_eval_grad_tree_array(
            nodes + offsets[t], offsets[t + 1] - offsets[t], ops, dtype, X, F, N, ldx, mode,
            (char*)out + (size_t)t * (size_t)N * es, (char*)grad + (size_t)grad_offsets[t] * es,
            grad_offsets[t + 1] - grad_offsets[t], NULL, &ok[t]);
        if (rc) { err = rc; ok[t] = 0; }
    }
    return err;
}

int dexo_eval_parametric_population(const dex_node* nodes, const int64_t* offsets,
                                    int64_t n_trees, const dexo_optable* ops, int dtype,
                                    const void* X, int32_t F, int64_t N, int64_t ldx,
                                    const void* parameters, int32_t n_params, int32_t n_classes,
                                    const int32_t* classes, int flags, int nthreads, void* out,
                                    uint8_t* ok) {
    int err = 0;
    size_t es = dtype == DEX_F32 ? 4 : 8;
    int nt = pick_threads(nthreads);
    (void)nt;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
    for (int64_t t = 0; t < n_trees; ++t) {
        const char* p = (const char*)parameters + (size_t)t * (size_t)n_params * (size_t)n_classes * es;
        int rc = dexo_eval_parametric(nodes + offsets[t], offsets[t + 1] - offsets[t], ops, dtype,
                                      X, F, N, ldx, p, n_params, n_classes, classes, flags,
                                      (char*)out + (size_t)t * (size_t)N * es, &ok[t]);
        if (rc) { err = rc; ok[t] = 0; }
    }
    return err;
}

/* Float64 inputs, every intermediate in 80-bit long double, Float64 outputs.  Comparing this
 * with the Float64 evaluation of the same tree tells how much of a disagreement between two
 * correct Float64 implementations is explained by rounding amplification alone. */
int dexo_eval_population_f80(const dex_node* nodes, const int64_t* offsets, int64_t n_trees,
                             const dexo_optable* ops, const double* X, int32_t F, int64_t N,
                             int64_t ldx, int flags, int nthreads, double* out, uint8_t* ok) {
    int err = 0;
    int nt = pick_threads(nthreads);
    (void)nt;
    long double* XL = (long double*)malloc(sizeof(long double) * (size_t)F * (size_t)(N > 0 ? N : 1));
    if (!XL) return -2;
    for (int64_t j = 0; j < N; ++j)
        for (int32_t f = 0; f < F; ++f) XL[f + (int64_t)F * j] = X[f + ldx * j];
#pragma omp parallel num_threads(nt)
    {
        ctx_f80 c;
        memset(&c, 0, sizeof(c));
        long double* tmp = (long double*)malloc(sizeof(long double) * (size_t)(N > 0 ? N : 1));
#pragma omp for schedule(dynamic, 1)
        for (int64_t t = 0; t < n_trees; ++t) {
            tree_info tr;
            int rc = tree_info_init(&tr, nodes + offsets[t], offsets[t + 1] - offsets[t], ops, F);
            if (rc) { err = rc; ok[t] = 0; continue; }
            c.tr = &tr; c.ops = ops; c.X = XL; c.F = F; c.N = N; c.ldx = F;
            c.early_exit = (flags & DEXO_EARLY_EXIT) != 0;
            c.use_fused = (flags & DEXO_USE_FUSED) != 0;
            eval_entry_f80(&c, flags, tmp, &ok[t]);
            for (int64_t j = 0; j < N; ++j) out[(size_t)t * (size_t)N + j] = (double)tmp[j];
            tree_info_free(&tr);
        }
        free(tmp);
        ctx_free_f80(&c);
    }
    free(XL);
    return err;
}

/* ---- scalar access for operator-table tests ---------------------------------- */
static int degree_of(int opcode) {
    switch (opcode) {
#define DEX_OP(SYM, code, degree, name, aliases) case code: return degree;
#include "../include/dex_ops.def"
#undef DEX_OP
        default: return 0;
    }
}
double dexo_apply_f64(int opcode, double a, double b, double c) {
    int d = degree_of(opcode);
    return d == 1 ? apply1_f64(opcode, a) : d == 2 ? apply2_f64(opcode, a, b)
         : d == 3 ? apply3_f64(opcode, a, b, c) : NAN;
}
float dexo_apply_f32(int opcode, float a, float b, float c) {
    int d = degree_of(opcode);
    return d == 1 ? apply1_f32(opcode, a) : d == 2 ? apply2_f32(opcode, a, b)
         : d == 3 ? apply3_f32(opcode, a, b, c) : NAN;
}
void dexo_partials_f64(int opcode, double a, double b, double c, double* g) {
    int d = degree_of(opcode);
    g[0] = g[1] = g[2] = 0.0;
    if (d == 1) partials1_f64(opcode, a, g);
    else if (d == 2) partials2_f64(opcode, a, b, g);
    else if (d == 3) partials3_f64(opcode, a, b, c, g);
}
