"""Replays one round of tests/soak.py and prints the disagreeing trees (debug aid)."""
import sys
import numpy as np
import dexb200
from dexb200 import treegen, device as D
from oracle import oracle
from tests import soak

def main(seed, target):
    rng = np.random.default_rng(seed)
    for r in range(target + 1):
        spec, nu, nb, dtype, depth, P, N, F, pol, tseed, X = soak.draw(rng)
    ops = dexb200.OperatorEnum(spec)
    nodes, offsets = treegen.gen_population(P, depth, nu, nb, F, seed=tseed, dtype=dtype)
    print(spec, dtype.__name__, depth, P, N, F, pol)
    oracle.lib()
    of = (oracle.EARLY_EXIT if pol.get("early_exit", True) else 0) | (oracle.USE_FUSED if pol.get("use_fused", True) else 0) | (oracle.BUMPER if pol.get("bumper") else 0)
    ref, rok = oracle.eval_population(nodes, offsets, ops.opcodes, X, of)
    _, rok_e = oracle.eval_population(nodes, offsets, ops.opcodes, X, of | oracle.ELEMENTWISE)
    pop0 = D.Population(None, ops, dtype, wire=(nodes, offsets), bumper=pol.get("bumper", False), use_fused=pol.get("use_fused", True))
    o0, k0 = pop0.eval(X, early_exit=pol.get("early_exit", True))
    k0 = k0.cpu().numpy().astype(bool); o0 = o0.cpu().numpy()
    for t in np.nonzero(k0 != rok_e)[0]:
        tree = dexb200.from_wire(nodes[offsets[t]:offsets[t + 1]])
        print("FLAG DIFF tree", t, "device", k0[t], "oracle sum-rule", rok[t], "elementwise", rok_e[t], dexb200.string_tree(tree, ops))
        r1, _ = oracle.eval_population(nodes[offsets[t]:offsets[t + 1]], np.array([0, offsets[t + 1] - offsets[t]]), ops.opcodes, X, of & ~oracle.EARLY_EXIT)
        bad = ~np.isfinite(r1[0])
        print("   oracle (no early exit) non-finite samples:", np.nonzero(bad)[0][:5], "X there", X[:, bad][:, :3].T, "device row there", o0[t][bad][:3])
    for name, kw, early in (("as drawn", dict(bumper=pol.get("bumper", False), use_fused=pol.get("use_fused", True)), pol.get("early_exit", True)),
                            ("early_exit on", dict(bumper=pol.get("bumper", False), use_fused=pol.get("use_fused", True)), True),
                            ("no bumper", dict(), pol.get("early_exit", True))):
        pop = D.Population(None, ops, dtype, wire=(nodes, offsets), **kw)
        out, ok = pop.eval(X, early_exit=early)
        out = out.cpu().numpy()
        bad = []
        for t in np.nonzero(rok)[0]:
            fin = np.isfinite(ref[t]) & np.isfinite(out[t])
            if fin.any():
                e = np.linalg.norm(out[t][fin].astype(np.float64) - ref[t][fin]) / max(np.linalg.norm(ref[t][fin].astype(np.float64)), 1e-300)
                if e > 1e-3: bad.append((int(t), float(e)))
        print(name, "bad trees:", bad[:8])
    for t, _ in bad[:2] if bad else []:
        pass
    # print the first bad tree of the drawn policy
    pop = D.Population(None, ops, dtype, wire=(nodes, offsets), bumper=pol.get("bumper", False), use_fused=pol.get("use_fused", True))
    out, ok = pop.eval(X, early_exit=pol.get("early_exit", True)); out = out.cpu().numpy()
    for t in np.nonzero(rok)[0]:
        fin = np.isfinite(ref[t]) & np.isfinite(out[t])
        if fin.any() and np.linalg.norm(out[t][fin].astype(np.float64) - ref[t][fin]) > 1e-3 * np.linalg.norm(ref[t][fin].astype(np.float64)):
            tree = dexb200.from_wire(nodes[offsets[t]:offsets[t + 1]])
            print("tree", t, dexb200.string_tree(tree, ops))
            sub = (nodes[offsets[t]:offsets[t + 1]], np.array([0, offsets[t + 1] - offsets[t]]))
            r64, _ = oracle.eval_population(sub[0], sub[1], ops.opcodes, X.astype(np.float64), of)
            d = np.where(fin, np.abs(out[t].astype(np.float64) - ref[t]), 0)
            for j in np.argsort(-d)[:6]:
                print("sample", j, "X", X[:, j], "device", out[t][j], "oracle", ref[t][j], "oracle f64", r64[0][j])
            ins, off = pop.tape()
            print("tape words of the tree:")
            for k in range(off[t], off[t + 1]):
                print("   ", [hex(int(v)) for v in np.asarray(ins[k]).view(np.uint32).ravel()[:4]])
            break

if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]))
