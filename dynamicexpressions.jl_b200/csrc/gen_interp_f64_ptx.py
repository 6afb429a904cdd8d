#!/usr/bin/env python
"""Generates dex_interp_f64.inc: the Float64 inner interpreter loop of dex_eval.cu as ONE inline-PTX
block (early_exit = true launches, 4 samples per thread as two 16-byte chunks).

Same tape, same contract and the same structure as the Float32 loop of gen_interp_ptx.py (jump-table
dispatch on w0 & 127, PUSH variants as table entries, one shared tail, out-of-line checked tail with
the warp vote that implements early exit); the arithmetic is scalar `*.rn.f64`.  Native handlers:
loads, KEEP, + - * / max min in all operand forms, neg abs square cube inv sqrt safe_sqrt relu.  The
transcendental handlers (exp, log, sin, cos, tanh, ...) and the generic handler return to the C++ code
for that one instruction (the library's double-precision sequences) and the loop is re-entered.

Operands: 0 pc | 1..4 acc (f64) | nf0 nf1 (f64) | 4 instruction words | ip n my_s tile_b cs_b
"""
import os
import struct

from gen_interp_ptx import F_CHK_A, F_CHK_B, F_CHK_OUT, handler_names

HERE = os.path.dirname(os.path.abspath(__file__))
K = 4
L = []


def emit(s=""):
    L.append(s)


def op(name):
    return "%" + str({"pc": 0, "nf": 1 + K, "ins": 3 + K, "ip": 7 + K, "n": 8 + K, "my": 9 + K, "tile": 10 + K,
                      "cs": 11 + K}[name])


def dhex(x):
    return "0d%016X" % struct.unpack("<Q", struct.pack("<d", x))[0]


A = [f"A{i}" for i in range(K)]
X = [f"X{i}" for i in range(K)]
Y = [f"Y{i}" for i in range(K)]
NAN = "0d7FF8000000000000"


def load_row(regs, addr):
    if addr == "ra":
        emit(f"and.b32 ra, w1, 65535; mad.lo.s32 ra, ra, {op('tile')}, {op('my')};")
    else:
        emit(f"shr.u32 rb, w1, 16; mad.lo.s32 rb, rb, {op('tile')}, {op('my')};")
    emit(f"ld.shared.v2.f64 {{{regs[0]}, {regs[1]}}}, [{addr}];")
    emit(f"add.s32 t, {addr}, {op('cs')};")
    emit(f"ld.shared.v2.f64 {{{regs[2]}, {regs[3]}}}, [t];")


def chk_vec(regs, flag, lab):
    emit(f"and.b32 t, w0, {flag}; setp.eq.b32 p, t, 0; @p bra.uni {lab};")
    for i, r in enumerate(regs):
        nf = "NF" if i % 2 == 0 else "NG"
        emit(f"fma.rn.f64 {nf}, {r}, ZZ, {nf};")
    emit(f"{lab}:")


def chk_const(lab, flag):
    emit(f"and.b32 t, w0, {flag}; setp.eq.b32 p, t, 0; @p bra.uni {lab};")
    emit("fma.rn.f64 NF, CC, ZZ, NF;")
    emit(f"{lab}:")


def jmax(d, a, b, is_max):
    """dex::j_max / j_min (Julia semantics): NaN if either is NaN, -0 < +0 (PTX max/min order the zeros
    the same way but return the OTHER operand for a NaN)."""
    emit(f"{'max' if is_max else 'min'}.f64 T0, {a}, {b}; setp.nan.f64 p, {a}, {b}; selp.f64 {d}, {NAN}, T0, p;")


def binary(name, sym):
    pat = name.rsplit("_", 1)[1]
    lab = f"H_{name}"
    emit(f"{lab}:")
    srcs = []
    for pos, ch in enumerate(pat):
        if ch == "A":
            srcs.append(A)
        elif ch == "R":
            regs = X if pos == 0 else Y
            load_row(regs, "ra" if pos == 0 else "rb")
            if (sym == "DIV" and pos == 1) or sym in ("MAX", "MIN"):
                chk_vec(regs, F_CHK_A if pos == 0 else F_CHK_B, f"{lab}_c{pos}")
            srcs.append(regs)
        else:
            emit("mov.b64 CC, {n2s, n3s};")
            chk_const(f"{lab}_cc", F_CHK_A if pos == 0 else F_CHK_B)
            srcs.append(["CC"] * K)
    a, b = srcs
    for k in range(K):
        if sym in ("ADD", "SUB", "MUL", "DIV"):
            emit(f"{ {'ADD': 'add', 'SUB': 'sub', 'MUL': 'mul', 'DIV': 'div'}[sym] }.rn.f64 {A[k]}, {a[k]}, {b[k]};")
        else:
            jmax(A[k], a[k], b[k], sym == "MAX")
    emit("bra.uni TAIL;")


def unary(name, sym):
    kind = name.rsplit("_", 1)[1]
    lab = f"H_{name}"
    emit(f"{lab}:")
    if kind == "R":
        load_row(X, "ra")
        chk_vec(X, F_CHK_A, f"{lab}_ca")
        src = X
    else:
        src = A
    for k in range(K):
        s, d = src[k], A[k]
        if sym == "NEG":
            emit(f"neg.f64 {d}, {s};")
        elif sym == "ABS":
            emit(f"abs.f64 {d}, {s};")
        elif sym == "SQUARE":
            emit(f"mul.rn.f64 {d}, {s}, {s};")
        elif sym == "CUBE":
            emit(f"mul.rn.f64 T0, {s}, {s}; mul.rn.f64 {d}, T0, {s};")
        elif sym == "INV":
            emit(f"rcp.rn.f64 {d}, {s};")
        elif sym in ("SQRT", "SAFE_SQRT"):       # sqrt of a negative number is NaN either way
            emit(f"sqrt.rn.f64 {d}, {s};")
        elif sym == "RELU":
            emit(f"setp.lt.f64 p, {s}, 0d0000000000000000; selp.f64 {d}, 0d0000000000000000, {s}, p;")
        else:
            raise KeyError(sym)
    emit("bra.uni TAIL;")


NATIVE_UNARY = {"NEG", "ABS", "SQUARE", "CUBE", "INV", "SQRT", "SAFE_SQRT", "RELU"}
NATIVE_BINARY = {"ADD", "SUB", "MUL", "DIV", "MAX", "MIN"}


def generate():
    names = handler_names()
    targets = []
    for nm in names:
        sym = nm.rsplit("_", 1)[0]
        native = nm in ("LOAD_R", "LOAD_C", "KEEP") or sym in NATIVE_UNARY or sym in NATIVE_BINARY
        targets.append(f"H_{nm}" if native else "EXIT")
    assert len(names) < 64
    targets += ["EXIT"] * (63 - len(names))
    targets.append("CHK_TAIL")       # never produced by the flattener; keeps the block out of line
    targets += [("P_" + t[2:]) if t.startswith("H_") else "EXIT" for t in targets[:64]]
    pc, nf, ins = op("pc"), int(op("nf")[1:]), int(op("ins")[1:])

    emit("{")
    emit(".reg .pred p, q;")
    emit(".reg .b32 w0, w1, n0, n1, n2, n3, n2s, n3s, h, t, ra, rb, rp;")
    emit(".reg .f64 " + ", ".join(A + X + Y) + ", CC, ZZ, NF, NG, T0;")
    emit(".reg .b64 ad;")
    emit(" ".join(f"mov.f64 {A[k]}, %{1 + k};" for k in range(K)))
    emit(f"mov.f64 NF, %{nf}; mov.f64 NG, %{nf + 1};")
    emit("mov.f64 ZZ, 0d0000000000000000;")
    emit(f"mov.b32 n0, %{ins}; mov.b32 n1, %{ins + 1}; mov.b32 n2, %{ins + 2}; mov.b32 n3, %{ins + 3};")
    emit("TBL: .branchtargets " + ", ".join(targets) + ";")
    emit("LOOP:")
    emit("and.b32 h, n0, 127;")
    emit("mov.b32 w0, n0; mov.b32 w1, n1; mov.b32 n2s, n2; mov.b32 n3s, n3;")
    emit(f"add.s32 {pc}, {pc}, 1; setp.ne.s32 q, {pc}, {op('n')};")
    emit(f"mul.wide.s32 ad, {pc}, 16; add.s64 ad, ad, {op('ip')};")
    emit("ld.global.nc.v4.u32 {n0, n1, n2, n3}, [ad];")
    emit("brx.idx.uni h, TBL;")

    for nm, tg in zip(names, targets[:len(names)]):
        if tg == "EXIT":
            continue
        emit(f"P_{nm}:")
        emit(f"shr.u32 rp, w0, 27; mad.lo.s32 rp, rp, {op('tile')}, {op('my')};")
        emit(f"st.shared.v2.f64 [rp], {{{A[0]}, {A[1]}}};")
        emit(f"add.s32 rp, rp, {op('cs')}; st.shared.v2.f64 [rp], {{{A[2]}, {A[3]}}};")
        emit(f"bra.uni H_{nm};")

    emit("H_LOAD_R:")
    load_row(A, "ra")
    chk_vec(A, F_CHK_A, "H_LOAD_R_ca")
    emit("bra.uni TAIL;")
    emit("H_LOAD_C:")
    emit("mov.b64 CC, {n2s, n3s};")
    chk_const("H_LOAD_C_cc", F_CHK_A)
    emit(" ".join(f"mov.f64 {r}, CC;" for r in A))
    emit("bra.uni TAIL;")
    emit("H_KEEP:")
    emit("bra.uni TAIL;")
    for nm in names[3:]:
        if nm == "KEEP":
            continue
        sym, pat = nm.rsplit("_", 1)
        if len(pat) == 1:
            if sym in NATIVE_UNARY:
                unary(nm, sym)
        elif sym in NATIVE_BINARY:
            binary(nm, sym)

    emit("TAIL:")
    emit(f"and.b32 t, w0, {F_CHK_OUT}; setp.ne.b32 p, t, 0; @p bra.uni CHK_TAIL;")
    emit("NEXT:")
    emit("@q bra.uni LOOP;")
    emit("bra.uni OUT;")
    # checked result + early exit: see gen_interp_ptx.py
    emit("CHK_TAIL:")
    emit(f"fma.rn.f64 T0, {A[0]}, ZZ, ZZ;")
    for r in A[1:]:
        emit(f"fma.rn.f64 T0, {r}, ZZ, T0;")
    emit("setp.nan.f64 p, T0, T0; vote.sync.any.pred p, p, 0xffffffff; @p bra.uni BAIL;")
    emit("bra.uni NEXT;")
    emit("BAIL:")
    emit("add.rn.f64 NF, NF, T0;")
    emit(f"mov.s32 {pc}, {op('n')};")
    emit(f"mul.wide.s32 ad, {pc}, 16; add.s64 ad, ad, {op('ip')};")
    emit("ld.global.nc.v4.u32 {n0, n1, n2, n3}, [ad];")
    emit("bra.uni OUT;")
    emit("EXIT:")
    emit(f"sub.s32 {pc}, {pc}, 1;")
    emit("OUT:")
    emit(" ".join(f"mov.f64 %{1 + k}, {A[k]};" for k in range(K)))
    emit("add.rn.f64 NF, NF, NG;")
    emit(f"mov.f64 %{nf}, NF; mov.f64 %{nf + 1}, ZZ;")
    emit(f"mov.b32 %{ins}, n0; mov.b32 %{ins + 1}, n1; mov.b32 %{ins + 2}, n2; mov.b32 %{ins + 3}, n3;")
    emit("}")

    out = os.path.join(HERE, "dex_interp_f64.inc")
    with open(out, "w") as f:
        f.write("// GENERATED by gen_interp_f64_ptx.py — do not edit.  Float64 interpreter loop as inline PTX (4 samples per thread).\n")
        for line in L:
            esc = line.replace("\\", "\\\\").replace('"', '\\"')
            f.write(f'"{esc}\\n\\t"\n')
    print(f"wrote {out}: {len(L)} PTX lines, {sum(t.startswith('H_') for t in targets)} native handlers of {len(names)}")


if __name__ == "__main__":
    generate()
