"""Replays the gradient leg of one round of tests/soak.py and prints the disagreeing tree (debug aid)."""
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
    mode = [D.GRAD_FEATURES, D.GRAD_CONSTANTS, D.GRAD_BOTH][target % 3]
    omode = [oracle.GRAD_FEATURES, oracle.GRAD_CONSTANTS, oracle.GRAD_BOTH][target % 3]
    oracle.lib()
    pop = D.Population(None, ops, dtype, wire=(nodes, offsets))
    out, grad, off, gok = pop.eval_grad(X, mode)
    out, grad, gok = out.cpu().numpy(), grad.cpu().numpy(), gok.cpu().numpy().astype(bool)
    ref, rgrads, rok = oracle.eval_grad_population(nodes, offsets, ops.opcodes, X, omode)
    print(spec, dtype.__name__, "mode", target % 3, "N", N, "F", F)
    shown = 0
    for t in np.nonzero(rok)[0]:
        G = rgrads[t].shape[0]
        if not G:
            continue
        g = grad[off[t]:off[t + 1]].reshape(N, G).T
        with np.errstate(all="ignore"):
            d = np.abs(g.astype(np.float64) - rgrads[t])
            rel = d / np.maximum(np.abs(rgrads[t]), 1e-300)
        if np.nanmax(np.where(d > 0, rel, 0)) > 1e-3:
            tree = dexb200.from_wire(nodes[offsets[t]:offsets[t + 1]])
            print("tree", t, dexb200.string_tree(tree, ops))
            i, j = np.unravel_index(np.nanargmax(np.where(d > 0, rel, 0)), d.shape)
            print("  direction", i, "sample", j, "X", X[:, j], "device", g[i, j], "oracle", rgrads[t][i, j], "value dev/oracle", out[t][j], ref[t][j])
            shown += 1
            if shown >= 3:
                break


if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]))
