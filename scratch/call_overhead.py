import sys; sys.path.insert(0, ".")
import time, numpy as np, torch, dexb200, ctypes as C
from dexb200 import device as D, treegen
ops = dexb200.OperatorEnum(treegen.OPSET_A)
nodes, offsets = treegen.gen_population(100, 6, 2, 4, 5, seed=0)
pop = D.Population(None, ops, np.float32, wire=(nodes, offsets))
N = 1000
X = torch.randn((N, 5), device="cuda")
out = torch.empty((100, N), device="cuda"); ok = torch.empty(100, dtype=torch.uint8, device="cuda")
for _ in range(10): pop.eval(X.T, out=out, ok=ok)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(2000): pop.eval(X.T, out=out, ok=ok)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("python eval call: %.2f us issue, %.2f us incl. drain" % ((t1 - t0) / 2000 * 1e6, (t2 - t0) / 2000 * 1e6))
lib = D.lib(); ctx = pop.ctx
Xp, outp, okp = X.data_ptr(), out.data_ptr(), ok.data_ptr()
t0 = time.perf_counter()
for _ in range(2000): lib.dex_eval(ctx.h, pop.h, C.c_void_p(Xp), 5, N, 5, C.c_void_p(outp), N, C.c_void_p(okp), 1)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("raw dex_eval call:  %.2f us issue, %.2f us incl. drain" % ((t1 - t0) / 2000 * 1e6, (t2 - t0) / 2000 * 1e6))
