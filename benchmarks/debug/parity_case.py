"""Diagnose one tree of a parity test on the GPU box: prints the tree and its worst elements."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import dexb200
from dexb200 import device as D, treegen
from oracle import oracle
from tests.parity_util import part_verdict, relerr, RTOL

def show(name, w, ops, out, ref, yards):
    print("==", name, dexb200.string_tree(dexb200.from_wire(w), ops)[:600])
    v = part_verdict(np.float32, out, ref, yards)
    print("verdict", v)
    out = out.astype(np.float64).ravel(); ref = ref.astype(np.float64).ravel()
    ys = [np.asarray(y, dtype=np.float64).ravel() for y in yards]
    d = np.abs(out - ref)
    sens = np.zeros_like(ref)
    for y in ys:
        sens = np.fmax(sens, np.where(np.isfinite(y), np.abs(y - ref), np.inf))
    well = 30 * sens <= 1e-4 * np.abs(ref)
    nw = lambda v: float(np.linalg.norm(v))
    print("well frac", well.mean(), "err_well", nw(d[well]) / max(nw(ref[well]), 1e-300), "ill: d", nw(d[~well]), "30s", 30 * nw(np.where(np.isfinite(sens[~well]), sens[~well], 0)))
    bad = np.argsort(-(d * well))[:8]
    for j in bad:
        print("  well-elem", j, "out", out[j], "ref", ref[j], "d/|ref|", d[j] / abs(ref[j]), "sens/|ref|", sens[j] / abs(ref[j]), [y[j] for y in ys])
    bad = np.argsort(-(d * ~well))[:5]
    for j in bad:
        print("  ill-elem", j, "out", out[j], "ref", ref[j], "d", d[j], "sens", sens[j])

ops = dexb200.OperatorEnum(treegen.OPSET_A)
which = sys.argv[1]
if which == "C5":
    n_params, n_classes, F, N = 3, 10, 5, 1 << 18
    nodes, offsets = treegen.gen_population(1000, 8, 2, 4, F, seed=0, n_params=n_params)
    t = 93 * 4
    w = nodes[offsets[t]:offsets[t + 1]]
    rng = np.random.default_rng(5)
    X = rng.standard_normal((F, N)).astype(np.float32)
    params = rng.standard_normal((250, n_params, n_classes)).astype(np.float32)[93:94]
    cls0 = rng.integers(0, n_classes, N)
    so = np.array([0, len(w)])
    pop = D.Population(None, ops, np.float32, wire=(w, so))
    out, ok = pop.eval_parametric(X, params, cls0)
    out = out.cpu().numpy()
    ref, rok = oracle.eval_parametric_population(w, so, ops.opcodes, X, params, cls0)
    ref64, _ = oracle.eval_parametric_population(w, so, ops.opcodes, X.astype(np.float64), params.astype(np.float64), cls0)
    ref_p, _ = oracle.eval_parametric_population(w, so, ops.opcodes, np.nextafter(X, np.float32(np.inf)), params, cls0)
    show("C5/93", w, ops, out[0], ref[0], (ref64[0], ref_p[0]))
else:
    nodes, offsets = treegen.gen_population(1000, 8, 2, 4, 5, seed=0)
    t = 96 * 4
    w = nodes[offsets[t]:offsets[t + 1]]
    so = np.array([0, len(w)])
    N = 1 << 16
    X = np.random.default_rng(0).standard_normal((5, N)).astype(np.float32)
    pop = D.Population(None, ops, np.float32, wire=(w, so))
    out, grad, off, ok = pop.eval_grad(X, D.GRAD_FEATURES)
    out, grad = out.cpu().numpy(), grad.cpu().numpy()
    mode = oracle.GRAD_FEATURES
    ref, rgrads, rok = oracle.eval_grad_population(w, so, ops.opcodes, X, mode)
    ref_p, rgrads_p, _ = oracle.eval_grad_population(w, so, ops.opcodes, np.nextafter(X, np.float32(np.inf)), mode)
    ref64, rgrads64, _ = oracle.eval_grad_population(w, so, ops.opcodes, X.astype(np.float64), mode)
    g = grad.reshape(N, 5).T
    show("C3/96 value", w, ops, out[0], ref[0], (ref_p[0], ref64[0]))
    show("C3/96 grad", w, ops, g, rgrads[0], (rgrads_p[0], rgrads64[0]))
