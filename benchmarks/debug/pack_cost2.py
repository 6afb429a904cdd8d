import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
if len(sys.argv) > 1 and sys.argv[1] == "bind":
    print("bound", bench.bind_to_gpu_numa(0), sorted(os.sched_getaffinity(0)))
import torch
import dexb200
from dexb200 import device as D, treegen
ops = dexb200.OperatorEnum(treegen.OPSET_A)
nodes, offsets = treegen.gen_population(1000, 8, 2, 4, 5, seed=0)
ctx = D.Context.get(0)
if len(sys.argv) > 2:
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda"); big = torch.empty((1000, 1<<16), device="cuda")
def T(f, n=10):
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(n): r=f()
    torch.cuda.synchronize(); return (time.perf_counter()-t)/n*1e3
for i in range(3):
    print("create+destroy", T(lambda: D.Population(None, ops, np.float32, wire=(nodes, offsets), ctx=ctx)))
