import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
import dexb200
from dexb200 import device as D, treegen
ops = dexb200.OperatorEnum(treegen.OPSET_A)
pops = {}
def pop_for(P):
    if P not in pops:
        nodes, offsets = treegen.gen_population(P, 8, 2, 4, 5, seed=0)
        pops[P] = D.Population(None, ops, np.float32, wire=(nodes, offsets))
    return pops[P]
def run(P, N, loss=False, reps=3):
    pop = pop_for(P)
    Xd = torch.randn((N, 5), device="cuda", dtype=torch.float32)
    if loss:
        y = torch.randn(N, device="cuda")
        f = lambda: pop.eval_loss(Xd.T, y)
    else:
        out = torch.empty((P, N), device="cuda", dtype=torch.float32)
        ok = torch.empty(P, device="cuda", dtype=torch.uint8)
        f = lambda: pop.eval(Xd.T, out=out, ok=ok)
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ms = min(ts)
    print(f"P={P:6d} N={N:8d} loss={loss}: {ms:8.2f} ms  node-ops/s {pop.info['n_nodes'] * N / ms * 1e3:.3e}")
run(1000, 1 << 16); run(10000, 1 << 16); run(1000, 1 << 20); run(10000, 1 << 20)
run(1000, 1 << 16, True); run(10000, 1 << 20, True)
import subprocess; print(subprocess.run(["nvidia-smi","--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active","--format=csv,noheader"],capture_output=True,text=True).stdout)
