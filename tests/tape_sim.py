"""numpy interpreter for the device evaluation tape (csrc/dex_tape.h) — TEST ONLY.

It executes exactly what csrc/dex_eval.cu executes — including the choice between a
specialised handler (operands implied by the handler id) and the generic handler
(operands decoded from the source fields) — with numpy ufuncs as the arithmetic, so that
the host-side flattener (csrc/dex_flatten.cpp: evaluation order, stack slots, check flags,
check elision, handler lowering) can be checked against the CPU oracle in this GPU-less
container.  It is not part of the product and is never imported by it.
"""
import numpy as np

from dexb200 import device as D
from oracle.oracle import _np_ops

SRC_ACC, SRC_ROW, SRC_CONST, SRC_PARAM = 0, 1, 2, 3
F_PUSH, F_CHK_OUT, F_CHK_A, F_CHK_B, F_ALWAYS, F_GUARD = (1 << 20, 1 << 21, 1 << 22, 1 << 23, 1 << 24, 1 << 25)


def handler_name(h):
    s = D.lib().dex_handler_name(int(h))
    return s.decode() if s else None


def run_tape(ins, X, max_stack, opcode_info, dtype, early_exit=True, params=None, classes0=None,
             n_param_rows=0):
    """ins: uint32[n, 4] of one tree; X: (F, N).  Returns (out[N], ok).
    Row layout (csrc/dex_tape.h): stack rows, parameter rows, feature rows."""
    u, b, t = _np_ops()
    sym2code = {v[0]: k for k, v in opcode_info.items()}
    F, N = X.shape
    rows = np.zeros((max_stack + n_param_rows + F, N), dtype=dtype)
    rows[max_stack + n_param_rows:] = X
    for p_ in range(n_param_rows):
        rows[max_stack + p_] = np.asarray(params, dtype=dtype)[p_, classes0]
    acc = np.zeros(N, dtype=dtype)
    ok = True

    def const_of(w):
        if dtype == np.float32:
            return np.array([w[2]], dtype=np.uint32).view(np.float32)[0]
        return np.array([w[2], w[3]], dtype=np.uint32).view(np.float64)[0]

    def bad(v):
        return not np.isfinite(v).all()

    for w in ins:
        w0, w1 = int(w[0]), int(w[1])
        rowA, rowB, push_row = w1 & 0xFFFF, w1 >> 16, w0 >> 27
        # the jump-table variant bits mirror the flags (csrc/dex_tape.h)
        assert bool(w0 & 64) == bool(w0 & F_PUSH) and bool(w0 & 128) == bool(w0 & F_CHK_OUT)
        if w0 & F_PUSH:
            rows[push_row] = acc
        c = const_of(w)
        cvec = np.full(N, c, dtype=dtype)
        h = (w0 & 0x3F) if early_exit else 0          # early_exit off => generic kernel
        code = (w0 >> 8) & 0xFF
        if h != 0:
            name = handler_name(h)
            if name == "KEEP":          # ACC unchanged (the PUSH above stored it): a shared subexpression
                assert opcode_info[code][0] == "IDENTITY" and (w0 & F_PUSH) and not (w0 & F_CHK_OUT)
                continue
            opname, pat = name.rsplit("_", 1)
            if opname == "LOAD":
                va = rows[rowA].copy() if pat == "R" else cvec
                if pat == "R" and (w0 & F_CHK_A) and bad(va):
                    ok = False
                if pat == "C" and (w0 & F_CHK_A) and bad(cvec):
                    ok = False
                assert opcode_info[code][0] == "IDENTITY"
                r = va
            else:
                hcode = sym2code[opname]
                assert hcode == code, (name, opcode_info[code])
                deg = opcode_info[code][1]
                if deg == 1:
                    va = acc.copy() if pat == "A" else rows[rowA].copy()
                    assert ((w0 >> 16) & 3) == (SRC_ACC if pat == "A" else SRC_ROW)
                    if pat == "R" and (w0 & F_CHK_A) and bad(va):
                        ok = False
                    with np.errstate(all="ignore"):
                        r = u[opname](va)
                else:
                    srcs = {"A": SRC_ACC, "R": SRC_ROW, "C": SRC_CONST}
                    assert ((w0 >> 16) & 3) == srcs[pat[0]] and ((w0 >> 18) & 3) == srcs[pat[1]]
                    # a checked feature ROW is specialised only where the handlers test it
                    assert not (w0 & F_CHK_A and pat[0] == "R") or opname in ("MAX", "MIN")
                    assert not (w0 & F_CHK_B and pat[1] == "R") or opname in ("DIV", "MAX", "MIN")
                    assert not (w0 & F_CHK_A and pat[0] == "A") and not (w0 & F_CHK_B and pat[1] == "A")
                    va = {"A": acc, "R": rows[rowA], "C": cvec}[pat[0]].copy()
                    vb = {"A": acc, "R": rows[rowB], "C": cvec}[pat[1]].copy()
                    if (w0 & F_CHK_A) and bad(va):
                        ok = False
                    if (w0 & F_CHK_B) and bad(vb):
                        ok = False
                    with np.errstate(all="ignore"):
                        r = b[opname](va, vb)
            acc = np.asarray(r, dtype=dtype)
            if (w0 & F_CHK_OUT) and bad(acc):
                ok = False
            continue

        def fetch(src, row):
            if src == SRC_ROW:
                return rows[row].copy()
            if src == SRC_CONST:
                return cvec.copy()
            if src == SRC_PARAM:
                return np.asarray(params, dtype=dtype)[row, classes0]
            return acc.copy()

        va = fetch((w0 >> 16) & 3, rowA)
        vb = fetch((w0 >> 18) & 3, rowB)
        chk = early_exit or bool(w0 & F_ALWAYS)
        if chk and (w0 & F_CHK_A) and bad(va):
            ok = False
        if chk and (w0 & F_CHK_B) and bad(vb):
            ok = False
        sym, deg, _ = opcode_info[code]
        with np.errstate(all="ignore"):
            if deg == 1:
                r = u[sym](va)
            elif deg == 2:
                r = b[sym](va, vb)
            else:
                r = t[sym](va, vb, acc)
        r = np.asarray(r, dtype=dtype)
        if (not early_exit) and (w0 & F_GUARD):
            r = np.where(np.isfinite(va), r, np.inf).astype(dtype)
        acc = r
        if chk and (w0 & F_CHK_OUT) and bad(r):
            ok = False
    return acc, ok


def run_folded(img, t, X, max_stack, opcode_info, dtype, early_exit=True, params=None, classes0=None,
               n_param_rows=0):
    """What dex_eval* execute for tree t of Population.folded(): the scalar segments of its
    constant subtrees first (prepass kernel, csrc/dex_eval.cu fold_tree: every check applies,
    whatever early_exit says), their results stored into the constant slots of the folded
    tape, then the folded tape over the samples."""
    off = img["offsets"]
    main = img["tape"][off[t]:off[t + 1]].copy()
    ok = True
    empty = np.zeros((0, 1), dtype=dtype)
    for k in range(img["seg_offsets"][t], img["seg_offsets"][t + 1]):
        b, e, target = (int(v) for v in img["segs"][k])
        assert off[t] <= target < off[t + 1], "segment target outside its tree"
        val, sok = run_tape(img["scalar_tape"][b:e], empty, 32, opcode_info, dtype, early_exit=True)
        ok = ok and sok
        w = main[target - off[t]]
        assert ((int(w[0]) >> 16) & 3) == SRC_CONST or ((int(w[0]) >> 18) & 3) == SRC_CONST
        raw = np.asarray(val[:1], dtype=dtype).view(np.uint32)
        w[2] = raw[0]
        w[3] = raw[1] if dtype == np.float64 else 0
    y, mok = run_tape(main, X, max_stack, opcode_info, dtype, early_exit=early_exit, params=params,
                      classes0=classes0, n_param_rows=n_param_rows)
    return y, (ok and mok)
