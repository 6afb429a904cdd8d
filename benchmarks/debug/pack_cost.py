import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import dexb200
from dexb200 import device as D, treegen
ops = dexb200.OperatorEnum(treegen.OPSET_A)
nodes, offsets = treegen.gen_population(1000, 8, 2, 4, 5, seed=0)
ctx = D.Context.get(0)
X = torch.randn((1 << 16, 5)).pin_memory()
oh = torch.empty((1000, 1 << 16)).pin_memory(); kh = torch.empty(1000, dtype=torch.uint8).pin_memory()
pop = D.Population(None, ops, np.float32, wire=(nodes, offsets), ctx=ctx)
for _ in range(3): pop.eval_host(X, oh, kh)
def T(f, n=10):
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(n): r=f()
    torch.cuda.synchronize(); return (time.perf_counter()-t)/n*1e3
print("warm eval_host ms", T(lambda: pop.eval_host(X, oh, kh)))
ts = {"create":0,"first_eval":0,"second_eval":0,"destroy":0}
for _ in range(10):
    torch.cuda.synchronize(); t0=time.perf_counter()
    p2 = D.Population(None, ops, np.float32, wire=(nodes, offsets), ctx=ctx)
    t1=time.perf_counter(); p2.eval_host(X, oh, kh); t2=time.perf_counter(); p2.eval_host(X, oh, kh); t3=time.perf_counter()
    del p2; torch.cuda.synchronize(); t4=time.perf_counter()
    for k,v in zip(ts, (t1-t0,t2-t1,t3-t2,t4-t3)): ts[k]+=v*100
print(ts)
os.environ["DEXB200_PACK_THREADS"]="1"
print("create+destroy 1 thread", T(lambda: D.Population(None, ops, np.float32, wire=(nodes, offsets), ctx=ctx)))
os.environ["DEXB200_PACK_THREADS"]="4"
print("create+destroy 4 threads", T(lambda: D.Population(None, ops, np.float32, wire=(nodes, offsets), ctx=ctx)))
hc = D.host_context()
print("host-only create (no CUDA) 4 thr", T(lambda: D.Population(None, ops, np.float32, wire=(nodes, offsets), ctx=hc)))
os.environ["DEXB200_PACK_THREADS"]="16"
print("host-only create (no CUDA) 16 thr", T(lambda: D.Population(None, ops, np.float32, wire=(nodes, offsets), ctx=hc)))
