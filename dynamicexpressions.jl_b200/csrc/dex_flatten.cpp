// dex_flatten.cpp — host-side flattening of wire trees into device tapes.
//
// Replaces the per-evaluation recursive tree walk of the reference
// (/root/reference/src/Evaluate.jl:337-364 and the dispatch functions :428-651):
// the walk is done ONCE here, at pack time, and its decisions are baked into the
// tape: evaluation order, operand locations, and — because the reference's
// `complete` flag depends on which of its kernels touches a value — the set of
// values that take part in the validity check (see DESIGN.md "completion flag").
#include "dex_tape.h"

#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <limits>
#include <thread>

#include "../../include/dexb200.h"

namespace dex {
namespace {

constexpr int OPERATOR_LIMIT_BEFORE_SLOWDOWN = 15;  // src/Evaluate.jl:14
constexpr int MAX_RECURSION = 20000;

struct Flattener {
    const OpTable& ops;
    const dex_node* nd = nullptr;  // current tree
    int64_t n = 0;
    int dtype;
    bool fused, bumper;
    bool fold = false;             // emit the folded image (constant subtrees -> scalar tape)
    bool in_fold = false;          // currently generating a scalar segment
    std::vector<int32_t> size;     // subtree sizes
    std::vector<uint8_t> isconst;  // subtree has no feature/parameter leaf
    std::vector<int32_t> need;     // stack slots needed with ACC free
    std::vector<int32_t> cord;     // constant ordinal of leaf i (tree-local), -1 otherwise
    // shared subexpressions (evaluation image only): nodes of the chosen repeated subtree carry its
    // group id at their root; the first occurrence in evaluation order computes it and KEEPs the
    // value in row cse_row, the others load it (GraphNode sharing, /root/reference/src/Node.jl:137-166,
    // which the reference's evaluators expand: /root/reference/ext/DynamicExpressionsBumperExt.jl:42)
    std::vector<int32_t> cse_group;   // per node: 0 = root of an occurrence of the shared subtree, -1 otherwise
    bool cse_done = false;            // the shared value sits in its row
    int cse_row = -1;
    PackedPopulation& out;
    std::string& err;
    int64_t tree_index = 0;
    int64_t const_base = 0;  // global ordinal of this tree's first constant
    int max_slot = 0;        // slots used by the tree being emitted

    Flattener(const OpTable& o, int dt, int pack_flags, bool fold_, PackedPopulation& p, std::string& e)
        : ops(o), dtype(dt), fused((pack_flags & DEX_PACK_FUSED) != 0),
          bumper((pack_flags & DEX_PACK_BUMPER) != 0), fold(fold_), out(p), err(e) {}

    int fail(int code, const std::string& msg) {
        err = "tree " + std::to_string(tree_index) + ": " + msg;
        return code;
    }

    // ---- structural scan + validation ------------------------------------------------
    int64_t scan(int64_t i, int depth, int& rc) {
        if (i >= n) { rc = fail(DEX_ERR_INVALID, "truncated tree (node " + std::to_string(i) + ")"); return -1; }
        if (depth > MAX_RECURSION) { rc = fail(DEX_ERR_UNSUPPORTED, "tree deeper than " + std::to_string(MAX_RECURSION)); return -1; }
        const dex_node& x = nd[i];
        if (x.degree > DEX_MAX_DEGREE) { rc = fail(DEX_ERR_INVALID, "node degree " + std::to_string(x.degree) + " > " + std::to_string(DEX_MAX_DEGREE)); return -1; }
        if (x.degree == 0) {
            if (x.kind > DEX_LEAF_PARAMETER) { rc = fail(DEX_ERR_INVALID, "bad leaf kind"); return -1; }
            size[i] = 1;
            isconst[i] = x.kind == DEX_LEAF_CONST;
            need[i] = 0;
            if (x.kind == DEX_LEAF_FEATURE) out.max_feature = std::max<int32_t>(out.max_feature, x.feature);
            if (x.kind == DEX_LEAF_PARAMETER) out.max_parameter = std::max<int32_t>(out.max_parameter, x.feature);
            return i + 1;
        }
        if (x.op >= ops.ops[x.degree - 1].size()) {
            rc = fail(DEX_ERR_INVALID, "node " + std::to_string(i) + " has op index " + std::to_string(x.op) +
                                           " but only " + std::to_string(ops.ops[x.degree - 1].size()) +
                                           " operators of degree " + std::to_string(x.degree) + " were passed");
            return -1;
        }
        int64_t j = i + 1;
        int64_t ch[DEX_MAX_DEGREE];
        uint8_t allc = 1;
        for (int k = 0; k < x.degree; ++k) {
            ch[k] = j;
            j = scan(j, depth + 1, rc);
            if (j < 0) return -1;
            allc &= isconst[ch[k]];
        }
        size[i] = (int32_t)(j - i);
        isconst[i] = allc;
        if (x.degree == 1) {
            need[i] = need[ch[0]];
        } else if (x.degree == 2) {
            // a folded constant subtree is an inline constant of its consumer, like a leaf
            const bool fl = fold && !bumper;
            bool ll = nd[ch[0]].degree == 0 || (fl && isconst[ch[0]]);
            bool rl = nd[ch[1]].degree == 0 || (fl && isconst[ch[1]]);
            if (ll && rl) need[i] = 0;
            else if (ll) need[i] = need[ch[1]];
            else if (rl) need[i] = need[ch[0]];
            else {
                int a = std::max(need[ch[0]], need[ch[1]]), b = std::min(need[ch[0]], need[ch[1]]);
                need[i] = std::max(a, b + 1);
            }
        } else {
            need[i] = std::max({need[ch[0]], need[ch[1]] + 1, need[ch[2]] + 2});
        }
        return j;
    }

    int opcode(int64_t i) const { return ops.ops[nd[i].degree - 1][nd[i].op]; }
    bool leaf(int64_t i) const { return nd[i].degree == 0; }
    int64_t child(int64_t i, int k) const {
        int64_t c = i + 1;
        for (int t = 0; t < k; ++t) c += size[c];
        return c;
    }

    // ---- evaluation tape ---------------------------------------------------------------
    struct Opnd {
        uint32_t src = SRC_ACC;
        uint32_t row = 0;   // stack slot (SRC_ROW, is_feature=false), feature idx, or param idx
        bool is_feature = false;
        bool is_param = false;  // a parameter row (gathered per tree into shared memory), rebased like features
        double c = 0.0;
        int32_t cord = -1;
        bool chk = false;
        int64_t fold = -1;  // index (in out.seg) of the scalar segment that produces this constant
    };
    // one instruction before lowering
    struct EIns {
        int op;
        Opnd a, b;
        uint32_t flags;   // F_CHK_OUT | F_ALWAYS | F_GUARD
        int push_slot;
    };
    std::vector<EIns> cur;  // instructions of the tree being flattened
    int last = -1;          // index in `cur` of the instruction that produced the latest value

    // `feature_checked`: whether the reference kernel that consumes this leaf checks it
    Opnd leaf_operand(int64_t i, bool feature_checked, bool const_mode) const {
        Opnd o;
        const dex_node& x = nd[i];
        if (x.kind == DEX_LEAF_CONST) {
            o.src = SRC_CONST;
            o.c = x.val;
            o.cord = cord[i];
            o.chk = true;  // @return_on_nonfinite_val / array check / Bumper isfinite(v)
        } else if (x.kind == DEX_LEAF_FEATURE) {
            o.src = SRC_ROW;
            o.is_feature = true;
            o.row = x.feature;
            o.chk = bumper ? false : (feature_checked || const_mode);
        } else {
            // a ParametricExpression parameter: the kernel gathers parameters[p, classes[:]] of
            // the current tree into a shared-memory row, so the operand is an ordinary ROW
            o.src = SRC_ROW;
            o.is_param = true;
            o.row = x.feature;
            o.chk = bumper ? false : (feature_checked || const_mode);
        }
        return o;
    }
    static Opnd acc() { return Opnd(); }
    static Opnd slot(int s) {
        Opnd o;
        o.src = SRC_ROW;
        o.row = (uint32_t)s;
        return o;
    }

    void put_const(Instr& ins, double c) const {
        if (dtype == DEX_F32) {
            float f = (float)c;
            std::memcpy(&ins.c_lo, &f, 4);
            ins.c_hi = 0;
        } else {
            uint64_t u;
            std::memcpy(&u, &c, 8);
            ins.c_lo = (uint32_t)u;
            ins.c_hi = (uint32_t)(u >> 32);
        }
    }

    // ---- check elision ------------------------------------------------------------------
    // transparent(op, k): a non-finite k-th operand ALWAYS yields a non-finite result
    // (per sample).  A value consumed through a transparent operand position of a checked
    // instruction does not need a check of its own: its non-finiteness reaches the
    // consumer's check (by induction, ultimately the root's).  The `complete` flag is
    // unchanged; only redundant work is removed.
    static bool transparent(int op, int k) {
        switch (op) {
            case DEX_OP_NEG: case DEX_OP_ABS: case DEX_OP_ABS2: case DEX_OP_SQUARE: case DEX_OP_CUBE:
            case DEX_OP_SQRT: case DEX_OP_CBRT: case DEX_OP_LOG: case DEX_OP_LOG2: case DEX_OP_LOG10:
            case DEX_OP_LOG1P: case DEX_OP_SIN: case DEX_OP_COS: case DEX_OP_TAN: case DEX_OP_ASIN:
            case DEX_OP_ACOS: case DEX_OP_SINH: case DEX_OP_COSH: case DEX_OP_ASINH: case DEX_OP_ACOSH:
            case DEX_OP_ATANH: case DEX_OP_ROUND: case DEX_OP_FLOOR: case DEX_OP_CEIL: case DEX_OP_TRUNC:
            case DEX_OP_IDENTITY: case DEX_OP_SAFE_LOG: case DEX_OP_SAFE_LOG2: case DEX_OP_SAFE_LOG10:
            case DEX_OP_SAFE_LOG1P: case DEX_OP_SAFE_SQRT: case DEX_OP_SAFE_ACOSH: case DEX_OP_COS2:
                return true;
            case DEX_OP_ADD: case DEX_OP_SUB: case DEX_OP_MUL: return true;   // Inf-Inf, Inf*0 -> NaN
            case DEX_OP_DIV: return k == 0;                                   // x/Inf = 0 hides the denominator
            case DEX_OP_MOD: return k == 0;
            case DEX_OP_COPYSIGN: return k == 0;
            case DEX_OP_FMA: case DEX_OP_MULADD: case DEX_OP_ADD3: case DEX_OP_MUL3: return true;
            default: return false;  // exp(-Inf)=0, 1/Inf=0, max(-Inf,1)=1, Inf^0=1, tanh, atan, ...
        }
    }
    // the value produced by cur[child] is consumed as operand k of cur[parent]
    void elide_result(int child, int parent, int k) {
        if (child < 0 || !elide) return;
        EIns& c = cur[(size_t)child];
        const EIns& p = cur[(size_t)parent];
        if (!(p.flags & F_CHK_OUT)) return;
        if ((c.flags & F_ALWAYS) && !(p.flags & F_ALWAYS)) return;  // root of a constant subtree
        if (!transparent(p.op, k)) return;
        if (c.op == DEX_OP_IDENTITY && c.b.src == SRC_ACC && c.a.src != SRC_ACC && !(c.flags & F_CHK_OUT))
            c.a.chk = false;  // a LOAD: the check sits on its operand
        c.flags &= ~F_CHK_OUT;
    }
    void elide_operand(Opnd& o, int op, int k, uint32_t pflags) const {
        if (elide && o.src != SRC_ACC && (pflags & F_CHK_OUT) && transparent(op, k)) o.chk = false;
    }

    // Emits one instruction into `cur`; returns its index.
    int emit(int op, Opnd a, Opnd b, uint32_t flags, int push_slot) {
        int deg = op >= 128 ? 3 : (op >= 64 ? 2 : 1);
        elide_operand(a, op, 0, flags);
        if (deg >= 2) elide_operand(b, op, 1, flags);
        EIns e{op, a, b, flags, push_slot};
        if (push_slot >= 0) max_slot = std::max(max_slot, push_slot + 1);
        cur.push_back(e);
        last = (int)cur.size() - 1;
        return last;
    }

    uint32_t out_flags(bool const_mode, bool guard) const {
        uint32_t f = F_CHK_OUT;
        if (const_mode) f |= F_ALWAYS;
        if (guard) f |= F_GUARD;
        return f;
    }

    // ---- constant-subtree folding ---------------------------------------------------------
    bool foldable(int64_t i, bool allow_fold = true) const {
        return fold && !bumper && !in_fold && allow_fold && nd[i].degree > 0 && isconst[i];
    }
    // Generates the scalar segment of constant subtree i and returns the inline-constant
    // operand that stands for it.  Its validity (every value the reference's scalar walk
    // checks) is reported by the fold pass per tree, not by a check in the sample loop.
    int fold_operand(int64_t i, int rec, Opnd& o) {
        std::vector<EIns> saved;
        saved.swap(cur);
        const int saved_last = last, saved_slots = max_slot;
        max_slot = 0;
        in_fold = true;
        int rc = gen(i, -1, 0, true, false, rec + 1);
        in_fold = false;
        if (!rc) {
            const int64_t begin = (int64_t)out.ctape.size();
            lower(cur, true);
            o = Opnd();
            o.src = SRC_CONST;
            o.c = std::numeric_limits<double>::quiet_NaN();   // overwritten by the fold pass
            o.fold = (int64_t)out.seg.size() / 3;
            out.seg.push_back(begin);
            out.seg.push_back((int64_t)out.ctape.size());
            out.seg.push_back(-1);
            scalar_slots = std::max(scalar_slots, max_slot);
        }
        cur.swap(saved);
        last = saved_last;
        max_slot = saved_slots;
        return rc;
    }
    int scalar_slots = 0;
    // folded subtree -> ACC
    int emit_load_fold(int64_t i, int push_slot, int rec) {
        Opnd a;
        if (int rc = fold_operand(i, rec, a)) return rc;
        const bool keep = elide;
        elide = false;
        emit(DEX_OP_IDENTITY, a, acc(), 0u, push_slot);
        elide = keep;
        return DEX_OK;
    }

    // Materialise a leaf into ACC (LOAD = IDENTITY with the leaf as operand A).
    int emit_load(int64_t i, bool feature_checked, bool const_mode, int push_slot) {
        Opnd a = leaf_operand(i, feature_checked, const_mode);
        const bool keep = elide;
        elide = false;  // the check of a LOAD lives on its operand
        int idx = emit(DEX_OP_IDENTITY, a, acc(), const_mode ? F_ALWAYS : 0u, push_slot);
        elide = keep;
        return idx;
    }

    // Emit code leaving the value of operator node i in ACC (and its producing instruction
    // index in `last`).
    //   push_slot  >= 0: ACC holds a live value that must be saved to that slot by the
    //              first instruction emitted here
    //   depth      first free stack slot (after the pending push)
    //   const_mode node lies in a constant subtree (folded by the reference)
    //   unchecked_leaves  leaves of THIS node are consumed by a fused unary kernel
    //              (deg1_l2_ll0_lr0 / deg1_l1_ll0) and are therefore not checked
    int gen(int64_t i, int push_slot, int depth, bool const_mode, bool unchecked_leaves, int rec,
            bool allow_fold = true) {
        if (rec > MAX_RECURSION) return fail(DEX_ERR_UNSUPPORTED, "tree too deep");
        if (depth >= MAX_STACK_ROWS) return fail(DEX_ERR_UNSUPPORTED, "operand stack deeper than " + std::to_string(MAX_STACK_ROWS));
        if (foldable(i, allow_fold) && !const_mode) return emit_load_fold(i, push_slot, rec);
        if (cse_row >= 0 && cse_group[(size_t)i] == 0) {
            if (cse_done) {     // a later occurrence: the value (already validated) comes from its row
                const bool keep = elide;
                elide = false;
                emit(DEX_OP_IDENTITY, slot(cse_row), acc(), 0u, push_slot);
                elide = keep;
                return DEX_OK;
            }
            cse_group[(size_t)i] = -1;                     // generate it normally ...
            if (int rc0 = gen(i, push_slot, depth, const_mode, unchecked_leaves, rec + 1, allow_fold)) return rc0;
            cse_group[(size_t)i] = 0;
            const bool keep = elide;
            elide = false;
            emit(DEX_OP_IDENTITY, acc(), acc(), 0u, cse_row);   // ... and KEEP it (PUSH to its row)
            elide = keep;
            cse_done = true;
            return DEX_OK;
        }
        const dex_node& x = nd[i];
        const int op = opcode(i);
        // constant subtrees are folded by _eval_tree_array (src/Evaluate.jl:347-354) — but the
        // branch of deg2_branch0_eval is evaluated inside the fused kernel, never folded
        if (!bumper && !const_mode && isconst[i] && allow_fold) const_mode = true;
        const bool fused1 = fused && !bumper && ops.ops[0].size() <= (size_t)OPERATOR_LIMIT_BEFORE_SLOWDOWN;
        const bool fused2 = fused && !bumper && ops.ops[1].size() <= (size_t)OPERATOR_LIMIT_BEFORE_SLOWDOWN;
        int rc;
        if (x.degree == 1) {
            int64_t c = i + 1;
            if (leaf(c)) {
                // generic dispatch_deg1_eval branch (:641-646): the leaf goes through
                // _eval_tree_array and is checked — unless this node is the inner op of
                // deg1_l1_ll0_eval, whose leaf is not.
                Opnd a = leaf_operand(c, !unchecked_leaves, const_mode);
                emit(op, a, acc(), out_flags(const_mode, false), push_slot);
                return DEX_OK;
            }
            bool guard = false, inner_unchecked = false;
            if (fused1 && !const_mode) {
                const dex_node& ch = nd[c];
                if (ch.degree == 2 && leaf(child(c, 0)) && leaf(child(c, 1))) { guard = true; inner_unchecked = true; }      // :624-632
                else if (ch.degree == 1 && leaf(c + 1)) { guard = true; inner_unchecked = true; }                              // :633-640
            }
            if ((rc = gen(c, push_slot, depth, const_mode, inner_unchecked, rec + 1))) return rc;
            const int ci = last;
            const int p = emit(op, acc(), acc(), out_flags(const_mode, guard), -1);
            elide_result(ci, p, 0);
            return DEX_OK;
        }
        if (x.degree == 2) {
            int64_t l = i + 1, r = l + size[l];
            bool ll = leaf(l), rl = leaf(r);
            if (ll && rl) {
                // deg2_l0_r0_eval (:874-933): feature leaves unchecked when fused
                bool chk = !(fused2 || unchecked_leaves);
                Opnd a = leaf_operand(l, chk, const_mode), b = leaf_operand(r, chk, const_mode);
                if (a.src == SRC_CONST && b.src == SRC_CONST) {
                    const int li = emit_load(l, chk, const_mode, push_slot);
                    const int p = emit(op, acc(), b, out_flags(const_mode, false), -1);
                    elide_result(li, p, 0);
                } else {
                    emit(op, a, b, out_flags(const_mode, false), push_slot);
                }
                return DEX_OK;
            }
            if (rl) {  // op(branch, leaf)
                bool chk = !fused2;  // deg2_r0_eval (:966-993) does not check the leaf
                const dex_node& L = nd[l];
                bool branch0 = fused2 && L.degree == 2 && leaf(child(l, 0)) && leaf(child(l, 1));
                if (branch0) chk = true;  // branch0 :left, x3 checked (:799)
                if (foldable(l, !branch0) && !const_mode && nd[r].kind != DEX_LEAF_CONST) {
                    Opnd a;
                    if ((rc = fold_operand(l, rec, a))) return rc;
                    emit(op, a, leaf_operand(r, chk, const_mode), out_flags(const_mode, false), push_slot);
                    return DEX_OK;
                }
                if ((rc = gen(l, push_slot, depth, const_mode, false, rec + 1, !branch0))) return rc;
                const int ci = last;
                const int p = emit(op, acc(), leaf_operand(r, chk, const_mode), out_flags(const_mode, false), -1);
                elide_result(ci, p, 0);
                return DEX_OK;
            }
            if (ll) {  // op(leaf, branch)
                bool chk = !fused2;
                const dex_node& R = nd[r];
                bool branch0 = fused2 && R.degree == 2 && leaf(child(r, 0)) && leaf(child(r, 1));
                if (branch0) chk = true;  // branch0 :right, x1 checked (:810)
                if (foldable(r, !branch0) && !const_mode && nd[l].kind != DEX_LEAF_CONST) {
                    Opnd b;
                    if ((rc = fold_operand(r, rec, b))) return rc;
                    emit(op, leaf_operand(l, chk, const_mode), b, out_flags(const_mode, false), push_slot);
                    return DEX_OK;
                }
                if ((rc = gen(r, push_slot, depth, const_mode, false, rec + 1, !branch0))) return rc;
                const int ci = last;
                const int p = emit(op, leaf_operand(l, chk, const_mode), acc(), out_flags(const_mode, false), -1);
                elide_result(ci, p, 1);
                return DEX_OK;
            }
            // both children are operators; one of them may be a folded constant subtree, which
            // the other one's consumer takes as its inline constant
            if (!const_mode && (foldable(l) != foldable(r))) {
                const bool lf = foldable(l);
                if ((rc = gen(lf ? r : l, push_slot, depth, const_mode, false, rec + 1))) return rc;
                const int ci = last;
                Opnd k;
                if ((rc = fold_operand(lf ? l : r, rec, k))) return rc;
                const int p = lf ? emit(op, k, acc(), out_flags(const_mode, false), -1)
                                 : emit(op, acc(), k, out_flags(const_mode, false), -1);
                elide_result(ci, p, lf ? 1 : 0);
                return DEX_OK;
            }
            // evaluate the one needing more stack first
            bool left_first = need[l] >= need[r];
            int64_t first = left_first ? l : r, second = left_first ? r : l;
            if ((rc = gen(first, push_slot, depth, const_mode, false, rec + 1))) return rc;
            const int fi = last;
            if ((rc = gen(second, depth, depth + 1, const_mode, false, rec + 1))) return rc;
            const int si = last;
            int p;
            if (left_first) p = emit(op, slot(depth), acc(), out_flags(const_mode, false), -1);
            else p = emit(op, acc(), slot(depth), out_flags(const_mode, false), -1);
            elide_result(fi, p, left_first ? 0 : 1);
            elide_result(si, p, left_first ? 1 : 0);
            return DEX_OK;
        }
        // degree 3 (dispatch_degn_eval :428-467): every child goes through
        // _eval_tree_array, so leaf children are checked.  Operands A and B may be a
        // leaf or a stack slot; C is always ACC.
        int64_t c[3] = {child(i, 0), child(i, 1), child(i, 2)};
        Opnd o[2];
        int prod[3] = {-1, -1, -1};  // producing instruction of each child (when not a direct leaf)
        int pending = push_slot;     // push to attach to the next ACC-overwriting instruction
        int acc_holds = -1;          // which of c[0], c[1] currently lives in ACC
        int d = depth;
        bool const_ab = leaf(c[0]) && leaf(c[1]) && nd[c[0]].kind == DEX_LEAF_CONST && nd[c[1]].kind == DEX_LEAF_CONST;
        for (int k = 0; k < 3; ++k) {
            bool direct = k < 2 && leaf(c[k]) && !(k == 0 && const_ab);
            if (direct) {
                o[k] = leaf_operand(c[k], true, const_mode);
                continue;
            }
            int ps = pending;
            if (acc_holds >= 0) {  // save the earlier operand into slot d
                ps = d;
                o[acc_holds] = slot(d);
                ++d;
                if (d >= MAX_STACK_ROWS) return fail(DEX_ERR_UNSUPPORTED, "operand stack too deep");
            }
            if (leaf(c[k])) prod[k] = emit_load(c[k], true, const_mode, ps);
            else {
                if ((rc = gen(c[k], ps, d, const_mode, false, rec + 1))) return rc;
                prod[k] = last;
            }
            pending = -1;
            acc_holds = k < 2 ? k : -1;
        }
        // encode: A = o[0], B = o[1], C = ACC
        const int p = emit(op, o[0], o[1], out_flags(const_mode, false), -1);
        for (int k = 0; k < 3; ++k) elide_result(prod[k], p, k);
        return DEX_OK;
    }

    // ---- shared subexpressions --------------------------------------------------------------
    // Picks the repeated subtree of this tree that saves the most instructions, if any: operator subtrees of >= 4 nodes that are not constant, compared record by
    // record.  One group per tree and only where its row fits the usual three stack rows: an extra
    // shared-memory row for the whole population would cost more occupancy than the reuse saves.
    static constexpr int CSE_MIN_NODES = 4;
    static constexpr int CSE_ROW_BUDGET = 3;
    static bool same_record(const dex_node& a, const dex_node& b) {
        return a.degree == b.degree && a.kind == b.kind && a.op == b.op && a.feature == b.feature &&
               std::memcmp(&a.val, &b.val, sizeof(double)) == 0;
    }
    void find_shared_subtree() {
        cse_row = -1;
        cse_done = false;
        if (!fold || bumper || !fused || n < 2 * CSE_MIN_NODES + 1) return;
        if (need[0] + 1 > CSE_ROW_BUDGET) return;
        // hash of every subtree, children folded in preorder (a subtree is a contiguous record range)
        std::vector<uint64_t> h((size_t)n);
        for (int64_t i = n - 1; i >= 0; --i) {
            const dex_node& x = nd[i];
            uint64_t v;
            std::memcpy(&v, &x.val, 8);
            uint64_t k = 0x9e3779b97f4a7c15ull ^ ((uint64_t)x.degree << 56) ^ ((uint64_t)x.kind << 48) ^
                         ((uint64_t)x.op << 40) ^ ((uint64_t)x.feature << 16) ^ (x.degree == 0 ? v * 0xff51afd7ed558ccdull : 0);
            int64_t c = i + 1;
            for (int d = 0; d < x.degree; ++d) {
                k = (k ^ h[(size_t)c]) * 0xc4ceb9fe1a85ec53ull;
                k ^= k >> 29;
                c += size[(size_t)c];
            }
            h[(size_t)i] = k;
        }
        std::vector<std::pair<uint64_t, int64_t>> cand;
        for (int64_t i = 1; i < n; ++i)
            if (nd[i].degree > 0 && size[(size_t)i] >= CSE_MIN_NODES && !isconst[(size_t)i]) cand.emplace_back(h[(size_t)i], i);
        std::sort(cand.begin(), cand.end());
        int64_t best = -1, best_gain = 0;
        std::vector<int64_t> best_occ, occ;
        for (size_t a = 0; a < cand.size();) {
            size_t b = a;
            while (b < cand.size() && cand[b].first == cand[a].first) ++b;
            if (b - a >= 2) {
                // occurrences identical to the first one, not nested in each other
                const int64_t i0 = cand[a].second, sz = size[(size_t)i0];
                occ.assign(1, i0);
                for (size_t k = a + 1; k < b; ++k) {
                    const int64_t j = cand[k].second;
                    if (size[(size_t)j] != sz || j < occ.back() + sz) continue;
                    bool eq = true;
                    for (int64_t q = 0; q < sz && eq; ++q) eq = same_record(nd[i0 + q], nd[j + q]);
                    if (eq) occ.push_back(j);
                }
                // instructions saved: every later occurrence costs one load instead of (about) one
                // instruction per operator node, and the first one an extra KEEP
                int64_t nops = 0;
                for (int64_t q = 0; q < sz; ++q) nops += nd[i0 + q].degree > 0;
                const int64_t gain = ((int64_t)occ.size() - 1) * (nops - 1) - 1;
                if (occ.size() >= 2 && gain > best_gain) { best_gain = gain; best = i0; best_occ = occ; }
            }
            a = b;
        }
        if (best < 0) return;
        cse_group.assign((size_t)n, -1);
        for (int64_t j : best_occ) cse_group[(size_t)j] = 0;
        cse_row = need[0];           // above every row the operand stack of this tree can use
    }

    // ---- lowering: EIns -> device encoding -------------------------------------------------
    static uint32_t pick_handler(const EIns& e, bool a_chk_row, bool b_chk_row) {
        const uint32_t sa = e.a.src, sb = e.b.src;
        if (sa == SRC_PARAM || sb == SRC_PARAM) return H_GENERIC;
        const int deg = e.op >= 128 ? 3 : (e.op >= 64 ? 2 : 1);
        if (deg == 1) {
            if (e.op == DEX_OP_IDENTITY) {
                if (sa == SRC_ROW) return H_LOAD_R;
                if (sa == SRC_CONST) return H_LOAD_C;
                return H_KEEP;
            }
            if (sa == SRC_CONST) return H_GENERIC;
            switch (e.op) {
#define X(S) case DEX_OP_##S: return sa == SRC_ACC ? H_##S##_A : H_##S##_R;
                DEX_FAST_UNARY(X)
#undef X
                default: return H_GENERIC;
            }
        }
        if (deg == 2) {
            // a checked feature operand: only positions that can hide a non-finite value keep the
            // flag (the divisor of /, either operand of max / min, ...); the specialised handlers
            // of / max min test it, any other operator takes the generic handler
            if (a_chk_row && !(e.op == DEX_OP_MAX || e.op == DEX_OP_MIN)) return H_GENERIC;
            if (b_chk_row && !(e.op == DEX_OP_DIV || e.op == DEX_OP_MAX || e.op == DEX_OP_MIN)) return H_GENERIC;
            switch (e.op) {
#define X(S)                                                              \
    case DEX_OP_##S:                                                      \
        if (sa == SRC_ACC && sb == SRC_ROW) return H_##S##_AR;            \
        if (sa == SRC_ACC && sb == SRC_CONST) return H_##S##_AC;          \
        if (sa == SRC_ROW && sb == SRC_ROW) return H_##S##_RR;            \
        if (sa == SRC_ROW && sb == SRC_CONST) return H_##S##_RC;          \
        return H_GENERIC;
                DEX_FAST_BIN_COMM(X)
#undef X
#define X(S)                                                              \
    case DEX_OP_##S:                                                      \
        if (sa == SRC_ACC && sb == SRC_ROW) return H_##S##_AR;            \
        if (sa == SRC_ROW && sb == SRC_ACC) return H_##S##_RA;            \
        if (sa == SRC_ACC && sb == SRC_CONST) return H_##S##_AC;          \
        if (sa == SRC_CONST && sb == SRC_ACC) return H_##S##_CA;          \
        if (sa == SRC_ROW && sb == SRC_ROW) return H_##S##_RR;            \
        if (sa == SRC_ROW && sb == SRC_CONST) return H_##S##_RC;          \
        if (sa == SRC_CONST && sb == SRC_ROW) return H_##S##_CR;          \
        return H_GENERIC;
                DEX_FAST_BIN_NC(X)
#undef X
                default: return H_GENERIC;
            }
        }
        return H_GENERIC;
    }
    static bool commutative(int op) {
        return op == DEX_OP_ADD || op == DEX_OP_MUL || op == DEX_OP_MAX || op == DEX_OP_MIN;
    }

    static bool fast_unary(int op) {
        switch (op) {
#define X(S) case DEX_OP_##S: return true;
            DEX_FAST_UNARY(X)
#undef X
            default: return false;
        }
    }

    // scalar == true: a folded constant subtree; its instructions go to the scalar tape, which
    // a one-thread interpreter runs (generic decode, no rows to rebase)
    void lower(const std::vector<EIns>& list, bool scalar) {
        // unary operator on an inline constant (only inside constant subtrees): split into
        // LOAD_C + OP_A so that both halves take specialised handlers
        std::vector<EIns> split;
        split.reserve(list.size());
        for (const EIns& e : list) {
            if (!scalar && e.op < 64 && e.op != DEX_OP_IDENTITY && e.a.src == SRC_CONST && fast_unary(e.op)) {
                EIns ld{DEX_OP_IDENTITY, e.a, acc(), e.flags & F_ALWAYS, e.push_slot};
                EIns op{e.op, acc(), acc(), e.flags, -1};
                split.push_back(ld);
                split.push_back(op);
            } else {
                split.push_back(e);
            }
        }
        std::vector<Instr>& tape = scalar ? out.ctape : out.tape;
        for (EIns e : split) {
            // commutative operators: bring the operands into (ACC|ROW, ROW|CONST) order
            bool swapped = false;
            if (commutative(e.op)) {
                auto rank = [](const Opnd& o) { return o.src == SRC_ACC ? 0 : o.src == SRC_ROW ? 1 : o.src == SRC_PARAM ? 1 : 2; };
                if (rank(e.a) > rank(e.b)) { std::swap(e.a, e.b); swapped = true; }
            }
            Instr ins{};
            const bool a_chk_row = e.a.chk && e.a.src != SRC_CONST && e.a.src != SRC_ACC;
            const bool b_chk_row = e.b.chk && e.b.src != SRC_CONST && e.b.src != SRC_ACC;
            const uint32_t h = scalar ? (uint32_t)H_GENERIC : pick_handler(e, a_chk_row, b_chk_row);
            ins.w0 = h | ((uint32_t)e.op << 8) | (e.a.src << 16) | (e.b.src << 18) | e.flags;
            // max / min partials break ties by operand order (dex_tape.h, SWAPPED)
            if (swapped && (e.op == DEX_OP_MAX || e.op == DEX_OP_MIN)) ins.w0 |= F_SWAPPED;
            if (e.a.chk) ins.w0 |= F_CHK_A;
            if (e.b.chk) ins.w0 |= F_CHK_B;
            if (e.flags & F_CHK_OUT) ins.w0 |= HANDLER_CHK;
            if (e.push_slot >= 0) ins.w0 |= F_PUSH | HANDLER_PUSH | ((uint32_t)e.push_slot << PUSH_ROW_SHIFT);
            // feature and parameter rows are rebased behind the stack rows once max_stack is known
            uint32_t ra = e.a.row, rb = e.b.row;
            ins.w1 = (ra & 0xffffu) | (rb << 16);
            const int64_t idx = (int64_t)tape.size();
            const int64_t pos = scalar ? -(1 + idx) : idx;
            for (const Opnd* o : {&e.a, &e.b}) {
                if (o->src != SRC_CONST) continue;
                put_const(ins, o->c);
                if (o->cord >= 0) out.const_pos[const_base + o->cord] = pos;
                if (o->fold >= 0) out.seg[(size_t)o->fold * 3 + 2] = idx;   // the fold pass stores here
            }
            if (!scalar) {
                rebase.push_back((uint8_t)((e.a.is_feature ? 1 : 0) | (e.b.is_feature ? 2 : 0) |
                                           (e.a.is_param ? 4 : 0) | (e.b.is_param ? 8 : 0)));
                out.tape_const_ord.push_back(e.a.src == SRC_CONST ? e.a.cord : e.b.src == SRC_CONST ? e.b.cord : -1);
                if (h == H_GENERIC) ++out.n_generic;
                out.n_checks += ((ins.w0 & F_CHK_OUT) ? 1 : 0) + (e.a.chk ? 1 : 0) + (e.b.chk ? 1 : 0);
            }
            tape.push_back(ins);
        }
    }
    std::vector<uint8_t> rebase;  // per tape instruction: bit0 rowA is a feature, bit1 rowB is a feature
    bool elide = true;

    int run(const dex_node* nodes, const int64_t* offsets, int64_t n_trees) {
        if (int rc = run_trees(nodes, offsets, n_trees)) return rc;
        return finish(out, rebase, err);
    }

    // the per-tree part: everything except the final row rebasing, which needs the stack depth
    // of the WHOLE population (so that partial results of several threads can be merged first)
    int run_trees(const dex_node* nodes, const int64_t* offsets, int64_t n_trees) {
        {   // one allocation per array instead of a growth series (a tape has at most one
            // instruction per node; about half of the nodes are leaves)
            const size_t nn = n_trees > 0 ? (size_t)(offsets[n_trees] - offsets[0]) : 0;
            out.tape.reserve(nn);
            out.tape_const_ord.reserve(nn);
            rebase.reserve(nn);
            out.const_pos.reserve(nn / 2 + 16);
            out.tape_off.reserve((size_t)n_trees + 1);
            out.const_off.reserve((size_t)n_trees + 1);
            out.seg_off.reserve((size_t)n_trees + 1);
            out.n_nodes_tree.reserve((size_t)n_trees);
            out.n_const_tree.reserve((size_t)n_trees);
            cur.reserve(256);
        }
        out.tape_off.assign(1, 0);
        out.const_off.assign(1, 0);
        out.seg_off.assign(1, 0);
        for (int64_t t = 0; t < n_trees; ++t) {
            tree_index = t;
            nd = nodes + offsets[t];
            n = offsets[t + 1] - offsets[t];
            if (n <= 0) return fail(DEX_ERR_INVALID, "empty tree");
            size.assign((size_t)n, 0);
            isconst.assign((size_t)n, 0);
            need.assign((size_t)n, 0);
            cord.assign((size_t)n, -1);
            int rc = DEX_OK;
            int64_t end = scan(0, 0, rc);
            if (end < 0) return rc;
            if (end != n) return fail(DEX_ERR_INVALID, "tree has " + std::to_string(n) + " records but its root subtree spans " + std::to_string(end));
            // constant ordinals: preorder = depth-first left-to-right leaf order
            // (index_constant_nodes, /root/reference/src/NodeUtils.jl:184-201)
            int32_t nc = 0;
            for (int64_t i = 0; i < n; ++i)
                if (nd[i].degree == 0 && nd[i].kind == DEX_LEAF_CONST) cord[i] = nc++;
            const_base = out.n_constants;
            out.const_pos.resize((size_t)(const_base + nc), -1);
            max_slot = 0;
            cur.clear();
            last = -1;
            find_shared_subtree();
            const size_t seg_mark = out.seg.size(), ctape_mark = out.ctape.size();
            const int scalar_mark = scalar_slots;
            if (nd[0].degree == 0) {
                // a bare leaf: deg0_eval then the final is_valid_array (:304-308); Bumper
                // checks constants only (ext/...BumperExt.jl:29)
                emit_load(0, true, false, -1);
            } else if ((rc = gen(0, -1, 0, false, false, 0))) {
                return rc;
            }
            if (cse_row >= 0) {
                // the shared value's row must be out of the operand stack's reach; if the stack ever
                // pushed there (it cannot, by the Sethi-Ullman bound), flatten again without sharing
                bool clash = false;
                for (const EIns& e : cur)
                    if (e.push_slot == cse_row && !(e.op == DEX_OP_IDENTITY && e.a.src == SRC_ACC)) clash = true;
                if (clash) {
                    cse_row = -1;
                    max_slot = 0;
                    cur.clear();
                    last = -1;
                    out.seg.resize(seg_mark);          // drop the scalar segments of the first attempt
                    out.ctape.resize(ctape_mark);
                    scalar_slots = scalar_mark;
                    if ((rc = gen(0, -1, 0, false, false, 0))) return rc;
                }
            }
            lower(cur, false);
            out.seg_off.push_back((int64_t)out.seg.size() / 3);
            out.max_stack = std::max(out.max_stack, max_slot);
            out.n_constants += nc;
            out.n_nodes += n;
            out.n_nodes_tree.push_back((int32_t)n);
            out.n_const_tree.push_back(nc);
            out.const_off.push_back(out.n_constants);
            out.tape_off.push_back((int64_t)out.tape.size());
        }
        out.n_trees = n_trees;
        return DEX_OK;
    }

    static int finish(PackedPopulation& out, const std::vector<uint8_t>& rebase, std::string& err) {
        // row layout: [0, max_stack) operand stack, then one row per parameter, then the features
        // rows reserved for parameters: those the trees use, or the count the caller announced
        // (DEX_PACK_PARAM_ROWS: gradient blocks have one row per parameter of the expression)
        out.n_param_rows = std::max<int32_t>(out.max_parameter + 1, (out.pack_flags >> 8) & 0xffff);
        const uint32_t pbase = (uint32_t)out.max_stack;
        const uint32_t base = pbase + (uint32_t)out.n_param_rows;
        if (out.max_feature >= 0 && (int64_t)out.max_feature + base > MAX_ROWS) {
            err = "feature index " + std::to_string(out.max_feature) + " + stack rows exceed the device row limit " +
                  std::to_string(MAX_ROWS);
            return DEX_ERR_UNSUPPORTED;
        }
        for (size_t k = 0; k < out.tape.size(); ++k) {
            Instr& ins = out.tape[k];
            uint32_t ra = row_a(ins.w1), rb = row_b(ins.w1);
            if (rebase[k] & 1) ra += base;
            if (rebase[k] & 2) rb += base;
            if (rebase[k] & 4) ra += pbase;
            if (rebase[k] & 8) rb += pbase;
            ins.w1 = ra | (rb << 16);
        }
        return DEX_OK;
    }
};

// appends the partial result `p` (trees [t0, t1) flattened on their own) to `out`
void append_part(PackedPopulation& out, std::vector<uint8_t>& rebase, const PackedPopulation& p,
                 const std::vector<uint8_t>& p_rebase) {
    const int64_t tape_base = (int64_t)out.tape.size(), ctape_base = (int64_t)out.ctape.size();
    const int64_t seg_base = (int64_t)out.seg.size() / 3, const_base = out.n_constants;
    out.tape.insert(out.tape.end(), p.tape.begin(), p.tape.end());
    rebase.insert(rebase.end(), p_rebase.begin(), p_rebase.end());
    out.tape_const_ord.insert(out.tape_const_ord.end(), p.tape_const_ord.begin(), p.tape_const_ord.end());
    out.n_nodes_tree.insert(out.n_nodes_tree.end(), p.n_nodes_tree.begin(), p.n_nodes_tree.end());
    out.n_const_tree.insert(out.n_const_tree.end(), p.n_const_tree.begin(), p.n_const_tree.end());
    for (size_t k = 1; k < p.tape_off.size(); ++k) out.tape_off.push_back(p.tape_off[k] + tape_base);
    for (size_t k = 1; k < p.const_off.size(); ++k) out.const_off.push_back(p.const_off[k] + const_base);
    for (size_t k = 1; k < p.seg_off.size(); ++k) out.seg_off.push_back(p.seg_off[k] + seg_base);
    for (int64_t v : p.const_pos) out.const_pos.push_back(v >= 0 ? v + tape_base : v - ctape_base);
    out.ctape.insert(out.ctape.end(), p.ctape.begin(), p.ctape.end());
    for (size_t k = 0; k + 2 < p.seg.size(); k += 3) {
        out.seg.push_back(p.seg[k] + ctape_base);
        out.seg.push_back(p.seg[k + 1] + ctape_base);
        out.seg.push_back(p.seg[k + 2] >= 0 ? p.seg[k + 2] + tape_base : p.seg[k + 2]);
    }
    out.n_trees += p.n_trees;
    out.n_nodes += p.n_nodes;
    out.n_constants += p.n_constants;
    out.n_generic += p.n_generic;
    out.n_checks += p.n_checks;
    out.max_stack = std::max(out.max_stack, p.max_stack);
    out.max_feature = std::max(out.max_feature, p.max_feature);
    out.max_parameter = std::max(out.max_parameter, p.max_parameter);
}

// One image (unfolded or folded) of the population.  Trees are independent, so large
// populations are flattened by several threads over contiguous tree ranges (balanced by node
// count) and the partial tapes concatenated; a failing range is redone serially so that the
// error message names the tree exactly as the one-thread walk would.
int flatten_image(const OpTable& ops, const dex_node* nodes, const int64_t* offsets, int64_t n_trees, int dtype,
                  int pack_flags, bool fold, int nthreads, PackedPopulation& out, std::string& err) {
    out = PackedPopulation();
    out.dtype = dtype;
    out.pack_flags = pack_flags;
    const int64_t total_nodes = n_trees > 0 ? offsets[n_trees] - offsets[0] : 0;
    nthreads = (int)std::min<int64_t>(nthreads, std::max<int64_t>(1, total_nodes / 2048));
    if (nthreads <= 1) {
        Flattener f(ops, dtype, pack_flags, fold, out, err);
        return f.run(nodes, offsets, n_trees);
    }
    struct Part {
        PackedPopulation pop;
        std::vector<uint8_t> rebase;
        std::string err;
        int rc = DEX_OK;
        int64_t t0 = 0, t1 = 0;
    };
    std::vector<Part> parts((size_t)nthreads);
    {   // contiguous ranges with about total_nodes / nthreads records each
        int64_t t = 0;
        for (int k = 0; k < nthreads; ++k) {
            parts[(size_t)k].t0 = t;
            const int64_t goal = offsets[0] + total_nodes * (k + 1) / nthreads;
            while (t < n_trees && (offsets[t + 1] <= goal || k == nthreads - 1)) ++t;
            parts[(size_t)k].t1 = t;
        }
        parts.back().t1 = n_trees;
    }
    std::vector<std::thread> workers;
    for (Part& p : parts) {
        workers.emplace_back([&ops, nodes, offsets, dtype, pack_flags, fold, &p]() {
            p.pop.dtype = dtype;
            p.pop.pack_flags = pack_flags;
            Flattener f(ops, dtype, pack_flags, fold, p.pop, p.err);
            p.rc = f.run_trees(nodes, offsets + p.t0, p.t1 - p.t0);
            p.rebase.swap(f.rebase);
        });
    }
    for (std::thread& w : workers) w.join();
    for (const Part& p : parts)
        if (p.rc != DEX_OK) {   // exact message (global tree index): the one-thread walk
            out = PackedPopulation();
            out.dtype = dtype;
            out.pack_flags = pack_flags;
            Flattener f(ops, dtype, pack_flags, fold, out, err);
            return f.run(nodes, offsets, n_trees);
        }
    std::vector<uint8_t> rebase;
    out.tape_off.assign(1, 0);
    out.const_off.assign(1, 0);
    out.seg_off.assign(1, 0);
    for (const Part& p : parts) append_part(out, rebase, p.pop, p.rebase);
    return Flattener::finish(out, rebase, err);
}

}  // namespace

int flatten_population(const OpTable& ops, const void* nodes, const int64_t* offsets,
                       int64_t n_trees, int dtype, int pack_flags, PackedPopulation& out,
                       std::string& err) {
    int nthreads = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
    if (const char* env = getenv("DEXB200_PACK_THREADS")) nthreads = std::max(1, atoi(env));
    const dex_node* nd = reinterpret_cast<const dex_node*>(nodes);
    // second image for evaluation: constant subtrees folded into the scalar tape.  The two images
    // are independent: the folded one is built by its own thread (group) at the same time.
    auto folded = std::make_shared<PackedPopulation>();
    std::string err2;
    int rc2 = DEX_OK;
    const int half = std::max(1, nthreads / 2);
    if (nthreads > 1) {
        std::thread second([&]() { rc2 = flatten_image(ops, nd, offsets, n_trees, dtype, pack_flags, true, half, *folded, err2); });
        const int rc = flatten_image(ops, nd, offsets, n_trees, dtype, pack_flags, false, half, out, err);
        second.join();
        if (rc) return rc;
    } else {
        const int rc = flatten_image(ops, nd, offsets, n_trees, dtype, pack_flags, false, 1, out, err);
        if (rc) return rc;
        rc2 = flatten_image(ops, nd, offsets, n_trees, dtype, pack_flags, true, 1, *folded, err2);
    }
    if (rc2) { err = err2; return rc2; }
    out.folded = folded;
    return DEX_OK;
}

}  // namespace dex
