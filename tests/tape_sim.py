"""numpy interpreter for the device evaluation tape (csrc/dex_tape.h) — TEST ONLY.

It executes exactly what csrc/dex_eval.cu executes (same operand sources, PUSH rows,
check flags, GUARD), with numpy ufuncs as the arithmetic, so that the host-side
flattener (csrc/dex_flatten.cpp) can be checked against the CPU oracle in this
GPU-less container.  It is not part of the product and is never imported by it.
"""
import numpy as np

from oracle.oracle import _np_ops

SRC_ACC, SRC_ROW, SRC_CONST, SRC_PARAM = 0, 1, 2, 3
F_PUSH, F_CHK_OUT, F_CHK_A, F_CHK_B, F_ALWAYS, F_GUARD = (1 << 12, 1 << 13, 1 << 14, 1 << 15,
                                                           1 << 16, 1 << 17)


def run_tape(ins, X, max_stack, opcode_info, dtype, early_exit=True, params=None, classes0=None):
    """ins: uint32[n, 4] of one tree; X: (F, N).  Returns (out[N], ok)."""
    u, b, t = _np_ops()
    F, N = X.shape
    rows = np.zeros((max_stack + F, N), dtype=dtype)
    rows[max_stack:] = X
    acc = np.zeros(N, dtype=dtype)
    ok = True

    def const_of(w):
        if dtype == np.float32:
            return np.array([w[2]], dtype=np.uint32).view(np.float32)[0]
        return np.array([w[2], w[3]], dtype=np.uint32).view(np.float64)[0]

    def fetch(src, row, c):
        if src == SRC_ROW:
            return rows[row].copy()
        if src == SRC_CONST:
            return np.full(N, c, dtype=dtype)
        if src == SRC_PARAM:
            return np.asarray(params, dtype=dtype)[row, classes0]
        return acc.copy()

    for w in ins:
        w0 = int(w[0])
        if w0 & F_PUSH:
            rows[w0 >> 24] = acc
        c = const_of(w)
        va = fetch((w0 >> 8) & 3, int(w[1]) & 0xFFFF, c)
        vb = fetch((w0 >> 10) & 3, int(w[1]) >> 16, c)
        chk = early_exit or bool(w0 & F_ALWAYS)
        if chk and (w0 & F_CHK_A) and not np.isfinite(va).all():
            ok = False
        if chk and (w0 & F_CHK_B) and not np.isfinite(vb).all():
            ok = False
        code = w0 & 0xFF
        sym, deg, _ = opcode_info[code]
        with np.errstate(all="ignore"):
            if deg == 1:
                r = u[sym](va)
            elif deg == 2:
                r = b[sym](va, vb)
            else:
                r = t[sym](va, vb, acc)
        r = np.asarray(r, dtype=dtype)
        if w0 & F_GUARD:
            r = np.where(np.isfinite(va), r, np.inf).astype(dtype)
        acc = r
        if chk and (w0 & F_CHK_OUT) and not np.isfinite(r).all():
            ok = False
    return acc, ok
