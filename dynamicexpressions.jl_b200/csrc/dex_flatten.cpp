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
    std::vector<int32_t> size;     // subtree sizes
    std::vector<uint8_t> isconst;  // subtree has no feature/parameter leaf
    std::vector<int32_t> need;     // stack slots needed with ACC free
    std::vector<int32_t> cord;     // constant ordinal of leaf i (tree-local), -1 otherwise
    PackedPopulation& out;
    std::string& err;
    int64_t tree_index = 0;
    int64_t const_base = 0;  // global ordinal of this tree's first constant
    int max_slot = 0;        // slots used by the tree being emitted
    int max_gslot = 0;

    Flattener(const OpTable& o, int dt, int pack_flags, PackedPopulation& p, std::string& e)
        : ops(o), dtype(dt), fused((pack_flags & DEX_PACK_FUSED) != 0),
          bumper((pack_flags & DEX_PACK_BUMPER) != 0), out(p), err(e) {}

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
            bool ll = nd[ch[0]].degree == 0, rl = nd[ch[1]].degree == 0;
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
        double c = 0.0;
        int32_t cord = -1;
        bool chk = false;
    };

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
            o.src = SRC_PARAM;
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

    // Emits one instruction.  Feature rows are stored as feature index here and are
    // rebased to absolute rows (max_stack + f) once the population's max_stack is known.
    void emit(int op, const Opnd& a, const Opnd& b, uint32_t flags, int push_slot) {
        Instr ins{};
        ins.w0 = (uint32_t)op | (a.src << 8) | (b.src << 10) | flags;
        if (a.chk) ins.w0 |= F_CHK_A;
        if (b.chk) ins.w0 |= F_CHK_B;
        if (push_slot >= 0) {
            ins.w0 |= F_PUSH | ((uint32_t)push_slot << 24);
            max_slot = std::max(max_slot, push_slot + 1);
        }
        // bit 15 of each row field marks "feature row, rebase later"
        uint32_t ra = a.row | (a.is_feature ? 0x8000u : 0u);
        uint32_t rb = b.row | (b.is_feature ? 0x8000u : 0u);
        ins.w1 = ra | (rb << 16);
        int64_t idx = (int64_t)out.tape.size();
        if (a.src == SRC_CONST) { put_const(ins, a.c); if (a.cord >= 0) out.const_pos[const_base + a.cord] = idx; }
        if (b.src == SRC_CONST) { put_const(ins, b.c); if (b.cord >= 0) out.const_pos[const_base + b.cord] = idx; }
        out.tape.push_back(ins);
    }

    uint32_t out_flags(bool const_mode, bool guard) const {
        uint32_t f = F_CHK_OUT;
        if (const_mode) f |= F_ALWAYS;
        if (guard) f |= F_GUARD;
        return f;
    }

    // Materialise a leaf into ACC (LOAD = IDENTITY with the leaf as operand A).
    void emit_load(int64_t i, bool feature_checked, bool const_mode, int push_slot) {
        Opnd a = leaf_operand(i, feature_checked, const_mode);
        emit(DEX_OP_IDENTITY, a, acc(), const_mode ? F_ALWAYS : 0u, push_slot);
    }

    // Emit code leaving the value of operator node i in ACC.
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
            emit(op, acc(), acc(), out_flags(const_mode, guard), -1);
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
                    emit(DEX_OP_IDENTITY, a, acc(), const_mode ? F_ALWAYS : 0u, push_slot);
                    emit(op, acc(), b, out_flags(const_mode, false), -1);
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
                if ((rc = gen(l, push_slot, depth, const_mode, false, rec + 1, !branch0))) return rc;
                emit(op, acc(), leaf_operand(r, chk, const_mode), out_flags(const_mode, false), -1);
                return DEX_OK;
            }
            if (ll) {  // op(leaf, branch)
                bool chk = !fused2;
                const dex_node& R = nd[r];
                bool branch0 = fused2 && R.degree == 2 && leaf(child(r, 0)) && leaf(child(r, 1));
                if (branch0) chk = true;  // branch0 :right, x1 checked (:810)
                if ((rc = gen(r, push_slot, depth, const_mode, false, rec + 1, !branch0))) return rc;
                emit(op, leaf_operand(l, chk, const_mode), acc(), out_flags(const_mode, false), -1);
                return DEX_OK;
            }
            // both children are operators: evaluate the one needing more stack first
            bool left_first = need[l] >= need[r];
            int64_t first = left_first ? l : r, second = left_first ? r : l;
            if ((rc = gen(first, push_slot, depth, const_mode, false, rec + 1))) return rc;
            if ((rc = gen(second, depth, depth + 1, const_mode, false, rec + 1))) return rc;
            if (left_first) emit(op, slot(depth), acc(), out_flags(const_mode, false), -1);
            else emit(op, acc(), slot(depth), out_flags(const_mode, false), -1);
            return DEX_OK;
        }
        // degree 3 (dispatch_degn_eval :428-467): every child goes through
        // _eval_tree_array, so leaf children are checked.  Operands A and B may be a
        // leaf or a stack slot; C is always ACC.
        int64_t c[3] = {child(i, 0), child(i, 1), child(i, 2)};
        Opnd o[2];
        bool have[2] = {false, false};
        int pending = push_slot;   // push to attach to the next ACC-overwriting instruction
        int acc_holds = -1;        // which of c[0], c[1] currently lives in ACC
        int d = depth;
        bool const_ab = leaf(c[0]) && leaf(c[1]) && nd[c[0]].kind == DEX_LEAF_CONST && nd[c[1]].kind == DEX_LEAF_CONST;
        for (int k = 0; k < 3; ++k) {
            bool direct = k < 2 && leaf(c[k]) && !(k == 0 && const_ab);
            if (direct) {
                o[k] = leaf_operand(c[k], true, const_mode);
                have[k] = true;
                continue;
            }
            int ps = pending;
            if (acc_holds >= 0) {  // save the earlier operand into slot d
                ps = d;
                o[acc_holds] = slot(d);
                have[acc_holds] = true;
                ++d;
                if (d >= MAX_STACK_ROWS) return fail(DEX_ERR_UNSUPPORTED, "operand stack too deep");
            }
            if (leaf(c[k])) emit_load(c[k], true, const_mode, ps);
            else if ((rc = gen(c[k], ps, d, const_mode, false, rec + 1))) return rc;
            pending = -1;
            acc_holds = k < 2 ? k : -1;
        }
        (void)have;
        // encode: A = o[0], B = o[1], C = ACC
        emit(op, o[0], o[1], out_flags(const_mode, false), -1);
        return DEX_OK;
    }

    // ---- gradient tape -----------------------------------------------------------------
    void gemit(uint32_t w0, uint32_t w1, double c) {
        GInstr g{};
        g.w0 = w0;
        g.w1 = w1;
        Instr tmp{};
        put_const(tmp, c);
        g.c_lo = tmp.c_lo;
        g.c_hi = tmp.c_hi;
        out.gtape.push_back(g);
    }
    int ggen(int64_t i, int dst, int rec) {
        if (rec > MAX_RECURSION) return fail(DEX_ERR_UNSUPPORTED, "tree too deep");
        if (dst >= 65535) return fail(DEX_ERR_UNSUPPORTED, "gradient stack too deep");
        max_gslot = std::max(max_gslot, dst + 1);
        const dex_node& x = nd[i];
        if (x.degree == 0) {
            uint32_t w1 = x.kind == DEX_LEAF_CONST ? (uint32_t)cord[i] : (uint32_t)x.feature;
            if (x.kind == DEX_LEAF_CONST) out.gconst_pos[const_base + cord[i]] = (int64_t)out.gtape.size();
            gemit(0u | ((uint32_t)x.kind << 8) | ((uint32_t)dst << 16), w1, x.val);
            return DEX_OK;
        }
        int64_t c = i + 1;
        for (int k = 0; k < x.degree; ++k) {
            int rc = ggen(c, dst + k, rec + 1);
            if (rc) return rc;
            c += size[c];
        }
        gemit((uint32_t)opcode(i) | ((uint32_t)dst << 16), 0, 0.0);
        return DEX_OK;
    }

    int run(const dex_node* nodes, const int64_t* offsets, int64_t n_trees) {
        out.tape_off.assign(1, 0);
        out.gtape_off.assign(1, 0);
        out.const_off.assign(1, 0);
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
            out.gconst_pos.resize((size_t)(const_base + nc), -1);
            max_slot = 0;
            max_gslot = 0;
            if (nd[0].degree == 0) {
                // a bare leaf: deg0_eval then the final is_valid_array (:304-308); Bumper
                // checks constants only (ext/...BumperExt.jl:29)
                emit_load(0, true, false, -1);
            } else if ((rc = gen(0, -1, 0, false, false, 0))) {
                return rc;
            }
            if ((rc = ggen(0, 0, 0))) return rc;
            out.max_stack = std::max(out.max_stack, max_slot);
            out.max_gstack = std::max(out.max_gstack, max_gslot);
            out.n_constants += nc;
            out.n_nodes += n;
            out.n_nodes_tree.push_back((int32_t)n);
            out.n_const_tree.push_back(nc);
            out.const_off.push_back(out.n_constants);
            out.tape_off.push_back((int64_t)out.tape.size());
            out.gtape_off.push_back((int64_t)out.gtape.size());
        }
        // rebase feature rows behind the stack rows
        const uint32_t base = (uint32_t)out.max_stack;
        for (Instr& ins : out.tape) {
            uint32_t ra = ins.w1 & 0xffffu, rb = ins.w1 >> 16;
            if (ra & 0x8000u) ra = (ra & 0x7fffu) + base;
            if (rb & 0x8000u) rb = (rb & 0x7fffu) + base;
            ins.w1 = ra | (rb << 16);
        }
        out.n_trees = n_trees;
        return DEX_OK;
    }
};

}  // namespace

int flatten_population(const OpTable& ops, const void* nodes, const int64_t* offsets,
                       int64_t n_trees, int dtype, int pack_flags, PackedPopulation& out,
                       std::string& err) {
    out = PackedPopulation();
    out.dtype = dtype;
    out.pack_flags = pack_flags;
    Flattener f(ops, dtype, pack_flags, out, err);
    int rc = f.run(reinterpret_cast<const dex_node*>(nodes), offsets, n_trees);
    if (rc) return rc;
    if (out.max_feature >= 0x7fff) { err = "feature index " + std::to_string(out.max_feature) + " exceeds the device limit 32766"; return DEX_ERR_UNSUPPORTED; }
    return DEX_OK;
}

}  // namespace dex
