import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import dexb200
from dexb200 import device as D, treegen
from oracle import oracle
from tests.test_gpu_parity import _subset
ops = dexb200.OperatorEnum(treegen.OPSET_A)
n_params, n_classes, F, N = 3, 10, 5, 1 << 18
nodes, offsets = treegen.gen_population(1000, 8, 2, 4, F, seed=0, n_params=n_params)
sn, so = _subset(nodes, offsets, np.arange(0, 1000, 4))
rng = np.random.default_rng(5)
X = rng.standard_normal((F, N)).astype(np.float32)
params = rng.standard_normal((250, n_params, n_classes)).astype(np.float32)
cls0 = rng.integers(0, n_classes, N)
pop = D.Population(None, ops, np.float32, wire=(sn, so))
out, ok = pop.eval_parametric(X, params, cls0)
out, ok = out.cpu().numpy(), ok.cpu().numpy()
ref, rok = oracle.eval_parametric_population(sn, so, ops.opcodes, X, params, cls0)
t = 93
w = sn[so[t]:so[t+1]]
print("population: ok", ok[t], rok[t], "isfinite", np.isfinite(out[t]).mean())
d = np.abs(out[t].astype(np.float64) - ref[t])
j = np.argsort(-d)[:6]
print("worst", [(int(i), float(out[t][i]), float(ref[t][i])) for i in j])
yn = oracle.numpy_eval(w, ops.opcodes, X.astype(np.float64), dexb200.OPCODE_INFO, parameters=params[t].astype(np.float64), classes0=cls0)
print("numpy64 at worst", [float(yn[i]) for i in j])
# single tree
p1 = D.Population(None, ops, np.float32, wire=(w, np.array([0, len(w)])))
o1, k1 = p1.eval_parametric(X, params[t:t+1], cls0)
o1 = o1.cpu().numpy()[0]
print("single: ok", k1.cpu().numpy(), "isfinite", np.isfinite(o1).mean(), "same as population", np.array_equal(o1, out[t]))
o2, k2 = p1.eval_parametric(X, np.ascontiguousarray(params[t:t+1]), cls0, early_exit=False)
o2 = o2.cpu().numpy()[0]
print("single no early exit: isfinite", np.isfinite(o2).mean(), "same as population", np.array_equal(o2, out[t]), np.abs(o2 - out[t]).max())
