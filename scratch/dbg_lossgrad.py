import sys; sys.path.insert(0, ".")
import numpy as np, torch, dexb200
from dexb200 import device as D, treegen
ops = dexb200.OperatorEnum({1: ("cos", "exp", "sin"), 2: ("+", "-", "*", "/")})
for dtype in (np.float32, np.float64):
    nodes, offsets = treegen.gen_population(90, 6, 3, 4, 3, seed=77, dtype=dtype)
    pop = D.Population(None, ops, dtype, wire=(nodes, offsets))
    for N in (100, 512, 513, 1024, 1500, 4096):
        rng = np.random.default_rng(21)
        X = rng.standard_normal((3, N)).astype(dtype); y = rng.standard_normal(N).astype(dtype)
        loss, grad, off, ok = pop.eval_loss_grad(X, y, D.GRAD_CONSTANTS)
        out, g2, off2, ok2 = pop.eval_grad(X, D.GRAD_CONSTANTS)
        o = out.cpu().numpy().astype(np.float64)
        want = ((o - y[None].astype(np.float64))**2).mean(axis=1)
        good = ok2.cpu().numpy().astype(bool)
        err = np.abs(loss.cpu().numpy()[good] - want[good]) / np.maximum(want[good], 1e-30)
        bad = np.nonzero(err > 1e-6)[0]
        print(dtype.__name__, N, "max rel err", err.max(), "n bad", len(bad), "of", good.sum(), "bad trees", np.nonzero(good)[0][bad][:8])
