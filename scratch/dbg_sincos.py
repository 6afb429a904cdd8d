import sys; sys.path.insert(0, ".")
import numpy as np, torch, dexb200
from dexb200 import device as D, treegen
ops = dexb200.OperatorEnum(treegen.OPSET_B)
nodes, offsets = treegen.gen_population(40, 6, 4, 4, 3, seed=9)
N = 1024
X = np.random.default_rng(N).standard_normal((3, N)).astype(np.float32)
pop = D.Population(None, ops, np.float32, wire=(nodes, offsets))
out, ok = pop.eval(X)
out = out.cpu().numpy()
t = 28
w = nodes[offsets[t]:offsets[t + 1]]
tree = dexb200.from_wire(w)
print(dexb200.string_tree(tree, ops))
y64, ok64 = None, None
import math
# evaluate in float64 with numpy by hand through the python mirror? use eval on float64 population
pop64 = D.Population(None, ops, np.float64, wire=(nodes, offsets))
o64, _ = pop64.eval(X.astype(np.float64))
o64 = o64.cpu().numpy()
err = np.abs(out[t] - o64[t]) / np.maximum(np.abs(o64[t]), 1e-30)
idx = np.argsort(-err)[:8]
print("worst samples", idx, err[idx], out[t][idx], o64[t][idx])
print("norm relerr", np.linalg.norm(out[t] - o64[t]) / np.linalg.norm(o64[t]))
print(X[:, idx])
# direct check of sin/cos handlers on a grid
ops2 = dexb200.OperatorEnum({1: ("sin", "cos"), 2: ("*",)})
N_ = dexb200.Node
xs = np.concatenate([np.linspace(-50, 50, 200001), np.random.default_rng(0).standard_normal(100000) * 1000]).astype(np.float32)
for nm, i, f in (("sin", 1, np.sin), ("cos", 2, np.cos)):
    for form, tree2 in (("R", N_(i, N_(feature=1))), ("A", N_(i, N_(1, N_(feature=1), N_(feature=2))))):
        y, okk = dexb200.eval_tree_array(tree2, np.stack([xs, np.ones_like(xs)]), ops2)
        ref = f(xs.astype(np.float64))
        ulp = np.spacing(np.abs(ref).astype(np.float32)).astype(np.float64)
        e = np.abs(y.astype(np.float64) - ref) / ulp
        print(nm, form, "max ulp", e.max(), "at", xs[e.argmax()], "n>3ulp", int((e > 3).sum()))
