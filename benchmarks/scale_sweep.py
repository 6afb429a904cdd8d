import sys, os; sys.path.insert(0, '.')
import numpy as np, torch
import dexb200
from dexb200 import device as D, treegen
ops = dexb200.OperatorEnum(treegen.OPSET_A)
P, N = int(sys.argv[1]), int(sys.argv[2])
nodes, offsets = treegen.gen_population(P, 8, 2, 4, 5, seed=0)
pop = D.Population(None, ops, np.float32, wire=(nodes, offsets))
Xd = torch.randn((N, 5), device="cuda", dtype=torch.float32)
out = torch.empty((P, N), device="cuda", dtype=torch.float32)
ok = torch.empty(P, device="cuda", dtype=torch.uint8)
f = lambda: pop.eval(Xd.T, out=out, ok=ok)
f(); torch.cuda.synchronize()
ts = []
for _ in range(3):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
ms = min(ts)
print(f"chunk_instr={os.environ.get('DEXB200_CHUNK_INSTR')} P={P} N={N}: {ms:.2f} ms node-ops/s {pop.info['n_nodes'] * N / ms * 1e3:.3e} frac {P*N*24/6547.8e9/(ms*1e-3):.3f}")
