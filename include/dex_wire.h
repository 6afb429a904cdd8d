/* dex_wire.h — the language-neutral tree wire format shared by every entry point
 * of libdexb200 (include/dexb200.h) and by the CPU oracle (oracle/).
 *
 * A tree is a PREORDER array of dex_node: a node is followed by the complete
 * subtrees of its children, left to right.  This is the direct image of the
 * reference's node struct (/root/reference/src/Node.jl:74-90:
 *   degree::UInt8, constant::Bool, val::T, feature::UInt16, op::UInt8, children)
 * and, for parametric trees, of ParametricNode
 * (/root/reference/src/ParametricExpression.jl:52-74: is_parameter, parameter).
 * A host shim produces it with one parent-first, children-left-to-right walk —
 * the order of the reference's tree_mapreduce (/root/reference/src/base.jl:123-158)
 * — which is also the order in which constants are numbered by
 * index_constant_nodes (/root/reference/src/NodeUtils.jl:184-201).
 *
 * All indices are 0-BASED inside the ABI (Julia shims subtract 1 from
 * `feature`, `op` and `parameter`).
 */
#ifndef DEX_WIRE_H
#define DEX_WIRE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    DEX_LEAF_CONST = 0,     /* leaf: val                                   */
    DEX_LEAF_FEATURE = 1,   /* leaf: X[feature, :]                         */
    DEX_LEAF_PARAMETER = 2  /* leaf: parameters[feature, classes[:]]       */
};

typedef struct dex_node {
    uint8_t degree;    /* 0 = leaf, 1..3 = operator arity                   */
    uint8_t kind;      /* leaves only: DEX_LEAF_*                            */
    uint8_t op;        /* operators only: 0-based index into operators[degree] */
    uint8_t reserved0;
    uint16_t feature;  /* 0-based feature (or parameter) row                 */
    uint16_t reserved1;
    double val;        /* constant value (Float32 trees: exactly representable) */
} dex_node;            /* 16 bytes */

/* element types */
enum { DEX_F32 = 0, DEX_F64 = 1 };

/* maximum operator arity the library handles (reference: type parameter D) */
#define DEX_MAX_DEGREE 3

/* opcode enum generated from dex_ops.def */
#define DEX_OP(SYM, code, degree, name, aliases) DEX_OP_##SYM = code,
enum dex_opcode {
#include "dex_ops.def"
    DEX_OP__END = 160
};
#undef DEX_OP

#ifdef __cplusplus
}
#endif
#endif /* DEX_WIRE_H */
