/* The README example of DynamicExpressions.jl (/root/reference/README.md:30-39) through the
 * C ABI of libdexb200 from PLAIN C — no Python, no torch: the drop-in boundary as a Julia
 * `ccall` shim would use it.
 *
 *     tree = x1 * cos(x2 - 3.2),  X :: 2 x 100 Float64  ->  X[1,:] .* cos.(X[2,:] .- 3.2)
 *
 * Exit code 0: results match the closed form (1e-12) and `complete` is true.
 * Exit code 3: no CUDA device — every compute entry point must fail with DEX_ERR_CUDA
 *              (there is no CPU fallback); anything else is exit code 1.                    */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dexb200.h"

#define CHECK(call)                                                                       \
    do {                                                                                  \
        int rc_ = (call);                                                                 \
        if (rc_ != DEX_OK) {                                                              \
            fprintf(stderr, "%s -> %d (%s) %s\n", #call, rc_, dex_strerror(rc_),          \
                    ctx ? dex_last_error(ctx) : "");                                      \
            return 1;                                                                     \
        }                                                                                 \
    } while (0)

int main(void) {
    dex_ctx* ctx = NULL;
    const int have_gpu = dex_device_count() > 0;
    CHECK(dex_ctx_create(have_gpu ? 0 : -1, &ctx));

    /* OperatorEnum(1 => (cos,), 2 => (+, -, *)) as builtin opcodes */
    int32_t opcodes[4] = {dex_opcode_from_name("cos", 1), dex_opcode_from_name("+", 2),
                          dex_opcode_from_name("-", 2), dex_opcode_from_name("*", 2)};
    int32_t degree_offsets[4] = {0, 1, 4, 4};
    dex_optable* ops = NULL;
    CHECK(dex_optable_create(opcodes, degree_offsets, 3, &ops));

    /* preorder: *(x1, cos(-(x2, 3.2)));  op / feature indices are 0-based inside the ABI */
    dex_node nodes[6];
    memset(nodes, 0, sizeof(nodes));
    nodes[0].degree = 2; nodes[0].op = 2;                       /* *   */
    nodes[1].degree = 0; nodes[1].kind = DEX_LEAF_FEATURE; nodes[1].feature = 0;
    nodes[2].degree = 1; nodes[2].op = 0;                       /* cos */
    nodes[3].degree = 2; nodes[3].op = 1;                       /* -   */
    nodes[4].degree = 0; nodes[4].kind = DEX_LEAF_FEATURE; nodes[4].feature = 1;
    nodes[5].degree = 0; nodes[5].kind = DEX_LEAF_CONST; nodes[5].val = 3.2;
    int64_t offsets[2] = {0, 6};
    dex_population* pop = NULL;
    CHECK(dex_population_pack(ctx, ops, nodes, offsets, 1, DEX_F64, DEX_PACK_DEFAULT, &pop));

    enum { F = 2, N = 100 };
    static double X[F * N], out[N];
    uint8_t ok = 0;
    srand(0);
    for (int i = 0; i < F * N; ++i) X[i] = 4.0 * rand() / RAND_MAX - 2.0;   /* column-major F x N */

    int rc = dex_eval_host(ctx, pop, X, F, N, F, out, N, &ok, DEX_EVAL_DEFAULT);
    if (!have_gpu) {
        printf("no CUDA device: dex_eval_host -> %d (%s)\n", rc, dex_strerror(rc));
        return rc == DEX_ERR_CUDA ? 3 : 1;
    }
    if (rc != DEX_OK) { fprintf(stderr, "dex_eval_host -> %d %s\n", rc, dex_last_error(ctx)); return 1; }
    double worst = 0.0;
    for (int j = 0; j < N; ++j) {
        const double want = X[j * F + 0] * cos(X[j * F + 1] - 3.2);
        const double err = fabs(out[j] - want);
        if (err > worst) worst = err;
    }
    printf("complete=%d worst abs err %.3e launches=%lld\n", (int)ok, worst, (long long)dex_ctx_launch_count(ctx));
    dex_population_destroy(pop);
    dex_optable_destroy(ops);
    dex_ctx_destroy(ctx);
    return (ok == 1 && worst < 1e-12) ? 0 : 1;
}
