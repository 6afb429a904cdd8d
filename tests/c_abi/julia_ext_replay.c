/* Replays, from PLAIN C, the exact sequence of libdexb200 calls that ext/DynamicExpressionsB200Ext.jl
 * makes — the Julia extension cannot be executed here (no `julia` in this environment), so this
 * harness is its stand-in: same argument order, same struct layout, same index conversions, same
 * interpretation of the result memory as Julia column-major arrays.
 *
 *   1. layout of `DexNode` (the Julia struct) == `dex_node` (include/dex_wire.h)
 *   2. context(), optable(operators) by function NAME (dex_opcode_from_name), population():
 *      flatten! (parent first, children left to right, 1-based -> 0-based) + dex_population_pack
 *   3. eval_tree_array(trees, B200Matrix(X), operators)        -> dex_eval_host
 *   4. eval_grad_tree_array(...; variable = Val(:both))         -> dex_device_alloc, dex_copy_to_device,
 *      dex_grad_offsets, dex_eval_grad, dex_copy_to_host
 *   5. eval_diff_tree_array(tree, X, operators, direction)      -> dex_eval_diff
 *   6. eval_tree_array(::ParametricExpression, X, classes)      -> DEX_PACK_PARAM_ROWS, dex_eval_parametric
 *   7. eval_loss_and_grad(...)                                  -> dex_eval_loss_grad
 *   8. set_constants!                                           -> dex_population_set_constants
 *   9. two contexts on the device, dex_shard_eval_host (the single-process multi-device entry point)
 *  10. the same with the result gathered on the root device through peer stores: dex_shard_eval
 * Expected values are the closed forms of the reference's own tests / docs (cited below).
 * Exit code 0 = all checks pass; 3 = no CUDA device (every compute call must fail with
 * DEX_ERR_CUDA: no CPU fallback); 1 = failure.                                                   */
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dexb200.h"

/* ---- 1. the Julia struct DexNode: fieldoffsets 0,1,2,3,4,6,8; sizeof 16 ------------------------- */
_Static_assert(sizeof(dex_node) == 16, "DexNode is 16 bytes");
_Static_assert(offsetof(dex_node, degree) == 0 && offsetof(dex_node, kind) == 1 && offsetof(dex_node, op) == 2, "u8 fields");
_Static_assert(offsetof(dex_node, feature) == 4 && offsetof(dex_node, val) == 8, "feature::UInt16 at 4, val::Float64 at 8");
_Static_assert(DEX_IPC_HANDLE_BYTES == 64, "NTuple{64,UInt8}");

static dex_ctx* ctx = NULL;
#define CHECK(call)                                                                                  \
    do {                                                                                             \
        int rc_ = (call);                                                                            \
        if (rc_ != DEX_OK) {                                                                         \
            fprintf(stderr, "%s:%d %s -> %d (%s) %s\n", __FILE__, __LINE__, #call, rc_, dex_strerror(rc_), \
                    ctx ? dex_last_error(ctx) : "");                                                 \
            return 1;                                                                                \
        }                                                                                            \
    } while (0)
#define EXPECT(cond, ...)                                       \
    do {                                                        \
        if (!(cond)) {                                          \
            fprintf(stderr, "%s:%d FAILED: ", __FILE__, __LINE__); \
            fprintf(stderr, __VA_ARGS__);                       \
            fprintf(stderr, "\n");                              \
            return 1;                                           \
        }                                                       \
    } while (0)

/* ---- the image of a Julia Node{T,2} / ParametricNode (src/Node.jl:74-90): 1-BASED indices -------- */
typedef struct jl_node {
    int degree, constant, is_parameter;
    double val;
    int feature, parameter, op; /* 1-based */
    struct jl_node* children[2];
} jl_node;
static jl_node pool[256];
static int n_pool = 0;
static jl_node* leaf_const(double v) { jl_node* n = &pool[n_pool++]; memset(n, 0, sizeof *n); n->constant = 1; n->val = v; return n; }
static jl_node* leaf_feature(int f) { jl_node* n = &pool[n_pool++]; memset(n, 0, sizeof *n); n->feature = f; return n; }
static jl_node* leaf_parameter(int p) { jl_node* n = &pool[n_pool++]; memset(n, 0, sizeof *n); n->is_parameter = 1; n->parameter = p; return n; }
static jl_node* unary(int op, jl_node* a) { jl_node* n = &pool[n_pool++]; memset(n, 0, sizeof *n); n->degree = 1; n->op = op; n->children[0] = a; return n; }
static jl_node* binary(int op, jl_node* a, jl_node* b) { jl_node* n = unary(op, a); n->degree = 2; n->children[1] = b; return n; }

/* flatten!(out, tree): parent first, children left to right (= tree_mapreduce order = constant numbering) */
static void flatten(dex_node* out, int64_t* n, const jl_node* t) {
    dex_node r;
    memset(&r, 0, sizeof r);
    r.degree = (uint8_t)t->degree;
    if (t->degree == 0) {
        if (t->constant) { r.kind = DEX_LEAF_CONST; r.val = t->val; }
        else if (t->is_parameter) { r.kind = DEX_LEAF_PARAMETER; r.feature = (uint16_t)(t->parameter - 1); }
        else { r.kind = DEX_LEAF_FEATURE; r.feature = (uint16_t)(t->feature - 1); }
        out[(*n)++] = r;
        return;
    }
    r.op = (uint8_t)(t->op - 1);
    out[(*n)++] = r;
    for (int i = 0; i < t->degree; ++i) flatten(out, n, t->children[i]);
}

static int close_to(double a, double b, double rtol) { return fabs(a - b) <= rtol * fmax(1.0, fabs(b)); }

int main(void) {
    const int have_gpu = dex_device_count() > 0;
    CHECK(dex_ctx_create(have_gpu ? 0 : -1, &ctx));

    /* ---- 2. optable(operators): OperatorEnum(1 => (cos, sin), 2 => (+, -, *, /)) by function name */
    const char* una[] = {"cos", "sin"};
    const char* bin[] = {"+", "-", "*", "/"};
    int32_t codes[6], offs[4] = {0, 2, 6, 6};
    for (int i = 0; i < 2; ++i) codes[i] = dex_opcode_from_name(una[i], 1);
    for (int i = 0; i < 4; ++i) codes[2 + i] = dex_opcode_from_name(bin[i], 2);
    for (int i = 0; i < 6; ++i) EXPECT(codes[i] >= 0, "operator %d has no device implementation", i);
    EXPECT(dex_opcode_from_name("my_custom_function", 1) < 0, "unknown operators must be rejected at table-build time");
    dex_optable* ops = NULL;
    CHECK(dex_optable_create(codes, offs, 3, &ops));
    enum { COS = 1, SIN = 2, ADD = 1, SUB = 2, MUL = 3, DIV = 4 };

    /* trees[1] = x1 * cos(x2 - 3.2)            README.md:30-39
     * trees[2] = 0.5 * x1 + cos(x2 - 0.2)      docs/src/eval.md:171-216, test/test_enzyme.jl:49-79 */
    jl_node* t1 = binary(MUL, leaf_feature(1), unary(COS, binary(SUB, leaf_feature(2), leaf_const(3.2))));
    jl_node* t2 = binary(ADD, binary(MUL, leaf_const(0.5), leaf_feature(1)), unary(COS, binary(SUB, leaf_feature(2), leaf_const(0.2))));
    dex_node nodes[64];
    int64_t offsets[3] = {0, 0, 0}, nn = 0;
    flatten(nodes, &nn, t1); offsets[1] = nn;
    flatten(nodes, &nn, t2); offsets[2] = nn;
    EXPECT(nn == 6 + 8, "flattened %lld records", (long long)nn);
    dex_population* pop = NULL;
    CHECK(dex_population_pack(ctx, ops, nodes, offsets, 2, DEX_F64, DEX_PACK_FUSED, &pop));
    int32_t counts[2];
    CHECK(dex_population_constant_counts(pop, counts));
    EXPECT(counts[0] == 1 && counts[1] == 2, "count_constant_nodes");

    /* X = [1 2 3; 4 5 6] (docs/src/eval.md:40-47) as a Julia Matrix{Float64}: column-major 2 x 3 */
    enum { F = 2, N = 3, P = 2 };
    double X[F * N] = {1, 4, 2, 5, 3, 6};

    /* ---- 3. eval_tree_array(trees, B200Matrix(X), operators): out :: Matrix (N x P) --------------- */
    double out[N * P];
    uint8_t ok[P];
    int rc = dex_eval_host(ctx, pop, X, F, N, F, out, N, ok, DEX_EVAL_EARLY_EXIT | DEX_EVAL_SKIP_INCOMPLETE);   /* eval_flags(nothing) */
    if (!have_gpu) {
        printf("no CUDA device: dex_eval_host -> %d (%s)\n", rc, dex_strerror(rc));
        return rc == DEX_ERR_CUDA ? 3 : 1;
    }
    CHECK(rc);
    for (int j = 0; j < N; ++j) {
        const double x1 = X[j * F], x2 = X[j * F + 1];
        EXPECT(close_to(out[0 * N + j], x1 * cos(x2 - 3.2), 1e-14), "tree 1 sample %d", j);
        EXPECT(close_to(out[1 * N + j], 0.5 * x1 + cos(x2 - 0.2), 1e-14), "tree 2 sample %d", j);
    }
    EXPECT(ok[0] == 1 && ok[1] == 1, "complete flags");

    /* ---- 4. eval_grad_tree_array(trees, X, operators; variable = Val(:both)) ----------------------- */
    int64_t goffs[P + 1];
    CHECK(dex_grad_offsets(pop, F, N, DEX_GRAD_BOTH, goffs));
    EXPECT(goffs[1] == (F + 1) * N && goffs[2] - goffs[1] == (F + 2) * N, "G = nfeatures + n_constants, features first");
    void *dX = NULL, *dO = NULL, *dG = NULL, *dK = NULL;
    CHECK(dex_device_alloc(ctx, &dX, sizeof X));
    CHECK(dex_device_alloc(ctx, &dO, sizeof out));
    CHECK(dex_device_alloc(ctx, &dG, goffs[P] * sizeof(double)));
    CHECK(dex_device_alloc(ctx, &dK, P));
    CHECK(dex_copy_to_device(ctx, dX, X, sizeof X));
    CHECK(dex_eval_grad(ctx, pop, dX, F, N, F, DEX_GRAD_BOTH, dO, N, dG, goffs, (uint8_t*)dK));
    double grad[(F + 1) * N + (F + 2) * N];
    CHECK(dex_copy_to_host(ctx, out, dO, sizeof out));
    CHECK(dex_copy_to_host(ctx, grad, dG, sizeof grad));
    CHECK(dex_copy_to_host(ctx, ok, dK, P));
    for (int j = 0; j < N; ++j) {
        const double x1 = X[j * F], x2 = X[j * F + 1];
        /* tree 2's block: reshape(view(grad, offs[2]+1:offs[3]), :, N) is (4 x N), gradient index fastest */
        const double* g = grad + goffs[1] + (size_t)j * 4;
        EXPECT(close_to(g[0], 0.5, 1e-14), "d/dx1 = 0.5 (docs/src/eval.md:213-216)");
        EXPECT(close_to(g[1], -sin(x2 - 0.2), 1e-14), "d/dx2 = -sin(x2 - 0.2)");
        EXPECT(close_to(g[2], x1, 1e-14), "d/dc1 = x1 (test/test_enzyme.jl:49-79)");
        EXPECT(close_to(g[3], sin(x2 - 0.2), 1e-14), "d/dc2 = sin(x2 - 0.2)");
        const double* h = grad + goffs[0] + (size_t)j * 3;
        EXPECT(close_to(h[0], cos(x2 - 3.2), 1e-14) && close_to(h[1], -x1 * sin(x2 - 3.2), 1e-14) &&
               close_to(h[2], x1 * sin(x2 - 3.2), 1e-14), "tree 1 gradient");
    }
    EXPECT(ok[0] == 1 && ok[1] == 1, "gradient complete flags");
    /* the second docs value: 0.611858 0.996165 0.464602 */
    EXPECT(fabs(grad[goffs[1] + 1] - 0.611858) < 1e-6 && fabs(grad[goffs[1] + 4 + 1] - 0.996165) < 1e-6, "docs/src/eval.md:216");

    /* ---- 5. eval_diff_tree_array(tree, X, operators, direction = 2) ------------------------------- */
    dex_population* pop1 = NULL;
    CHECK(dex_population_pack(ctx, ops, nodes, offsets, 1, DEX_F64, DEX_PACK_FUSED, &pop1));
    void* dD = NULL;
    CHECK(dex_device_alloc(ctx, &dD, N * sizeof(double)));
    CHECK(dex_eval_diff(ctx, pop1, dX, F, N, F, 2 - 1, dO, dD, N, (uint8_t*)dK));
    double dout[N];
    CHECK(dex_copy_to_host(ctx, dout, dD, sizeof dout));
    for (int j = 0; j < N; ++j) EXPECT(close_to(dout[j], -X[j * F] * sin(X[j * F + 1] - 3.2), 1e-14), "eval_diff");

    /* ---- 6. ParametricExpression: sin(x) + y + p1 * p2, parameters = [1 1 .8; 2 3 5],
     *         X = [0 pi/2 pi 1.2; 0 0 1.5 .1], classes = [1, 1, 2, 3] -> [2, 3, 4.5, 5.032039085967226]
     *         (test/test_parametric_expression.jl:103-128) --------------------------------------------- */
    jl_node* tp = binary(ADD, binary(ADD, unary(SIN, leaf_feature(1)), leaf_feature(2)),
                         binary(MUL, leaf_parameter(1), leaf_parameter(2)));
    dex_node pnodes[16];
    int64_t poffs[2] = {0, 0}, pn = 0;
    flatten(pnodes, &pn, tp);
    poffs[1] = pn;
    dex_population* ppop = NULL;
    enum { NP = 2, NC = 3, PN = 4 };
    CHECK(dex_population_pack(ctx, ops, pnodes, poffs, 1, DEX_F64, DEX_PACK_FUSED | DEX_PACK_PARAM_ROWS(NP), &ppop));
    const double pi = 3.14159265358979323846;
    double PX[F * PN] = {0, 0, pi / 2, 0, pi, 1.5, 1.2, 0.1};
    double params[NP * NC] = {1, 2, 1, 3, 0.8, 5};        /* Array{T,3}(n_params, n_classes, P), column-major */
    int32_t cls0[PN] = {0, 0, 1, 2};                        /* Int32.(classes .- 1) */
    void *dPX = NULL, *dP = NULL, *dC = NULL, *dPO = NULL;
    CHECK(dex_device_alloc(ctx, &dPX, sizeof PX));
    CHECK(dex_device_alloc(ctx, &dP, sizeof params));
    CHECK(dex_device_alloc(ctx, &dC, sizeof cls0));
    CHECK(dex_device_alloc(ctx, &dPO, PN * sizeof(double)));
    CHECK(dex_copy_to_device(ctx, dPX, PX, sizeof PX));
    CHECK(dex_copy_to_device(ctx, dP, params, sizeof params));
    CHECK(dex_copy_to_device(ctx, dC, cls0, sizeof cls0));
    CHECK(dex_eval_parametric(ctx, ppop, dPX, F, PN, F, dP, NP, NC, (const int32_t*)dC, dPO, PN, (uint8_t*)dK, DEX_EVAL_EARLY_EXIT));
    double pout[PN];
    const double pwant[PN] = {2, 3, 4.5, 5.032039085967226};
    CHECK(dex_copy_to_host(ctx, pout, dPO, sizeof pout));
    for (int j = 0; j < PN; ++j) EXPECT(close_to(pout[j], pwant[j], 1e-14), "parametric sample %d: %.17g", j, pout[j]);
    /* ... and its derivatives: the parameter rows are the first feature directions */
    int64_t pgo[2];
    CHECK(dex_grad_offsets(ppop, NP + F, PN, DEX_GRAD_FEATURES, pgo));
    void* dPG = NULL;
    CHECK(dex_device_alloc(ctx, &dPG, pgo[1] * sizeof(double)));
    CHECK(dex_eval_grad_parametric(ctx, ppop, dPX, F, PN, F, dP, NP, NC, (const int32_t*)dC, DEX_GRAD_FEATURES, dPO, PN, dPG, pgo, (uint8_t*)dK));
    double pg[(NP + F) * PN];
    CHECK(dex_copy_to_host(ctx, pg, dPG, sizeof pg));
    for (int j = 0; j < PN; ++j) {
        const double p1 = params[cls0[j] * NP], p2 = params[cls0[j] * NP + 1];
        const double* g = pg + (size_t)j * (NP + F);
        EXPECT(close_to(g[0], p2, 1e-14) && close_to(g[1], p1, 1e-14) && close_to(g[2], cos(PX[j * F]), 1e-14) &&
               close_to(g[3], 1.0, 1e-14), "parametric gradient sample %d", j);
    }

    /* ---- 7. eval_loss_and_grad(trees, X, y, operators; variable = Val(false)) ------------------------ */
    double y[N] = {0.3, -1.0, 2.0}, loss[P], lgrad[3];
    int64_t lo[P + 1];
    CHECK(dex_grad_offsets(pop, F, 1, DEX_GRAD_CONSTANTS, lo));
    EXPECT(lo[1] == 1 && lo[2] == 3, "loss gradient offsets");
    void *dY = NULL, *dL = NULL, *dLG = NULL;
    CHECK(dex_device_alloc(ctx, &dY, sizeof y));
    CHECK(dex_device_alloc(ctx, &dL, sizeof loss));
    CHECK(dex_device_alloc(ctx, &dLG, sizeof lgrad));
    CHECK(dex_copy_to_device(ctx, dY, y, sizeof y));
    CHECK(dex_eval_loss_grad(ctx, pop, dX, F, N, F, dY, NULL, DEX_GRAD_CONSTANTS, (double*)dL, (double*)dLG, lo, (uint8_t*)dK));
    CHECK(dex_copy_to_host(ctx, loss, dL, sizeof loss));
    CHECK(dex_copy_to_host(ctx, lgrad, dLG, sizeof lgrad));
    double wl = 0, wg0 = 0, wg1 = 0;
    for (int j = 0; j < N; ++j) {
        const double x1 = X[j * F], x2 = X[j * F + 1], r = 0.5 * x1 + cos(x2 - 0.2) - y[j];
        wl += r * r / N; wg0 += 2 * r * x1 / N; wg1 += 2 * r * sin(x2 - 0.2) / N;
    }
    EXPECT(close_to(loss[1], wl, 1e-13) && close_to(lgrad[1], wg0, 1e-13) && close_to(lgrad[2], wg1, 1e-13), "fused loss and gradient");

    /* ---- 8. set_constants!: (0.5, 0.2) -> (2.0, 1.0) without re-packing ----------------------------- */
    double consts[3];
    CHECK(dex_population_get_constants(ctx, pop, consts, 3));
    EXPECT(consts[0] == 3.2 && consts[1] == 0.5 && consts[2] == 0.2, "get_scalar_constants: tree order, leaf order");
    consts[1] = 2.0; consts[2] = 1.0;
    CHECK(dex_population_set_constants(ctx, pop, consts, 3));
    CHECK(dex_eval_host(ctx, pop, X, F, N, F, out, N, ok, DEX_EVAL_EARLY_EXIT));
    for (int j = 0; j < N; ++j) EXPECT(close_to(out[1 * N + j], 2.0 * X[j * F] + cos(X[j * F + 1] - 1.0), 1e-14), "after set_constants!");

    /* ---- 9. one process, several "devices" (here two contexts on device 0): dex_shard_eval_host ---- */
    dex_ctx* ctx2 = NULL;
    CHECK(dex_ctx_create(0, &ctx2));
    dex_population* pop2 = NULL;
    CHECK(dex_population_pack(ctx2, ops, nodes, offsets, 2, DEX_F64, DEX_PACK_FUSED, &pop2));
    CHECK(dex_population_set_constants(ctx2, pop2, consts, 3));
    enum { NS = 1001 };
    static double XS[F * NS], OS[P * NS], OR[P * NS];
    uint8_t oks[P], okr[P];
    srand(1);
    for (int i = 0; i < F * NS; ++i) XS[i] = 4.0 * rand() / RAND_MAX - 2.0;
    dex_ctx* ctxs[2] = {ctx, ctx2};
    const dex_population* pops[2] = {pop, pop2};
    CHECK(dex_shard_eval_host(ctxs, pops, 2, XS, F, NS, F, OS, NS, oks, DEX_EVAL_EARLY_EXIT));
    CHECK(dex_eval_host(ctx, pop, XS, F, NS, F, OR, NS, okr, DEX_EVAL_EARLY_EXIT));
    EXPECT(memcmp(OS, OR, sizeof OS) == 0 && memcmp(oks, okr, P) == 0, "sharded == unsharded, bit for bit");

    /* ---- 10. ... with the result gathered on the root DEVICE: dex_shard_eval (every context's kernel
     *          stores its column block straight into the root's matrix) ------------------------------ */
    {
        const int64_t n0 = NS / 2, n1 = NS - n0;           /* blocks [0, NS/2) and [NS/2, NS) */
        void *dX0 = NULL, *dX1 = NULL, *dOR = NULL, *dKR = NULL;
        CHECK(dex_device_alloc(ctx, &dX0, (int64_t)(F * n0 * sizeof(double))));
        CHECK(dex_device_alloc(ctx2, &dX1, (int64_t)(F * n1 * sizeof(double))));
        CHECK(dex_device_alloc(ctx, &dOR, (int64_t)sizeof OS));
        CHECK(dex_device_alloc(ctx, &dKR, P));
        CHECK(dex_copy_to_device(ctx, dX0, XS, (int64_t)(F * n0 * sizeof(double))));
        CHECK(dex_copy_to_device(ctx2, dX1, XS + F * n0, (int64_t)(F * n1 * sizeof(double))));
        CHECK(dex_ctx_synchronize(ctx2));
        const void* xdevs[2] = {dX0, dX1};
        CHECK(dex_shard_eval(ctxs, pops, 2, xdevs, F, NS, F, dOR, NS, (uint8_t*)dKR, 0, DEX_EVAL_EARLY_EXIT));
        memset(OS, 0, sizeof OS);
        CHECK(dex_copy_to_host(ctx, OS, dOR, (int64_t)sizeof OS));      /* root stream order: complete */
        CHECK(dex_copy_to_host(ctx, oks, dKR, P));
        EXPECT(memcmp(OS, OR, sizeof OS) == 0 && memcmp(oks, okr, P) == 0, "device-side sharded gather == unsharded");
        dex_device_free(ctx, dX0); dex_device_free(ctx2, dX1); dex_device_free(ctx, dOR); dex_device_free(ctx, dKR);
    }

    printf("julia_ext_replay: all checks passed (launches=%lld)\n", (long long)dex_ctx_launch_count(ctx));
    dex_device_free(ctx, dX); dex_device_free(ctx, dO); dex_device_free(ctx, dG); dex_device_free(ctx, dK);
    dex_device_free(ctx, dD); dex_device_free(ctx, dPX); dex_device_free(ctx, dP); dex_device_free(ctx, dC);
    dex_device_free(ctx, dPO); dex_device_free(ctx, dPG); dex_device_free(ctx, dY); dex_device_free(ctx, dL);
    dex_device_free(ctx, dLG);
    dex_population_destroy(pop); dex_population_destroy(pop1); dex_population_destroy(ppop); dex_population_destroy(pop2);
    dex_optable_destroy(ops);
    dex_ctx_destroy(ctx2);
    dex_ctx_destroy(ctx);
    return 0;
}
