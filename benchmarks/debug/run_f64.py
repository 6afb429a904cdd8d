import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import dexb200
from dexb200 import device as D, treegen
ops = dexb200.OperatorEnum(treegen.OPSET_A)
nodes, offsets = treegen.gen_population(1000, 8, 2, 4, 5, seed=0, dtype=np.float64)
pop = D.Population(None, ops, np.float64, wire=(nodes, offsets))
X = torch.randn((1 << 16, 5), device="cuda", dtype=torch.float64)
out = torch.empty((1000, 1 << 16), device="cuda", dtype=torch.float64); ok = torch.empty(1000, dtype=torch.uint8, device="cuda")
for _ in range(4):
    pop.eval(X.T, out=out, ok=ok)
torch.cuda.synchronize()
