import sys; sys.path.insert(0, ".")
import time, numpy as np, torch, dexb200
from dexb200 import device as D, treegen
ops = dexb200.OperatorEnum(treegen.OPSET_A)
torch.cuda.init()
ctx = D.Context.get(0)
for P in (100, 1000, 10000):
    nodes, offsets = treegen.gen_population(P, 8, 2, 4, 5, seed=0)
    ts = []
    for _ in range(8):
        t0 = time.perf_counter(); pop = D.Population(None, ops, np.float32, wire=(nodes, offsets), ctx=ctx); ts.append(time.perf_counter() - t0); del pop
    print(P, "trees: pack+upload min %.2f ms median %.2f ms" % (min(ts) * 1e3, np.median(ts) * 1e3))
