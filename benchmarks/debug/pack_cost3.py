import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
import torch
import dexb200
from dexb200 import device as D, treegen
ops = dexb200.OperatorEnum(treegen.OPSET_A)
nodes, offsets = treegen.gen_population(1000, 8, 2, 4, 5, seed=0)
ctx = D.Context.get(0)
X = torch.randn((1 << 16, 5)).pin_memory()
oh = torch.empty((1000, 1 << 16)).pin_memory(); kh = torch.empty(1000, dtype=torch.uint8).pin_memory()
def T(f, n=10):
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(n): r=f()
    torch.cuda.synchronize(); return (time.perf_counter()-t)/n*1e3
def step():
    p2 = D.Population(None, ops, np.float32, wire=(nodes, offsets), ctx=ctx)
    p2.eval_host(X, oh, kh)
    del p2
print("before nvml: create+destroy", T(lambda: D.Population(None, ops, np.float32, wire=(nodes, offsets), ctx=ctx)), "step", T(step))
with bench.ClockSampler(0) as clk:
    time.sleep(0.05)
print("after sampler: create+destroy", T(lambda: D.Population(None, ops, np.float32, wire=(nodes, offsets), ctx=ctx)), "step", T(step))
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
pop = D.Population(None, ops, np.float32, wire=(nodes, offsets), ctx=ctx)
Xd = X.cuda(); out = torch.empty((1000, 1<<16), device="cuda"); ok = torch.empty(1000, dtype=torch.uint8, device="cuda")
for _ in range(20):
    flush.fill_(1); pop.eval(Xd.T, out=out, ok=ok)
print("after device loop: create+destroy", T(lambda: D.Population(None, ops, np.float32, wire=(nodes, offsets), ctx=ctx)), "step", T(step))
import pynvml; pynvml.nvmlShutdown()
print("after nvml shutdown: create+destroy", T(lambda: D.Population(None, ops, np.float32, wire=(nodes, offsets), ctx=ctx)), "step", T(step))
