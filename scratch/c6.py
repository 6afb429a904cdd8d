# scaled-up configurations: C6 (10k depth-8 trees x 2^20 samples, F=5) and C4-shape (depth 12, F=10)
import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
import dexb200
from dexb200 import device as D, treegen
ops = dexb200.OperatorEnum(treegen.OPSET_A)
def run(name, P, depth, F, N, reps=3):
    t0 = time.time()
    nodes, offsets = treegen.gen_population(P, depth, 2, 4, F, seed=0, max_nodes=(2**depth - 1))
    pop = D.Population(None, ops, np.float32, wire=(nodes, offsets))
    print(name, "gen+pack s", round(time.time() - t0, 2), pop.info)
    Xd = torch.randn((N, F), device="cuda", dtype=torch.float32)
    out = torch.empty((P, N), device="cuda", dtype=torch.float32)
    ok = torch.empty(P, device="cuda", dtype=torch.uint8)
    pop.eval(Xd.T, out=out, ok=ok); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); pop.eval(Xd.T, out=out, ok=ok); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = min(ts)
    nodeops = pop.info["n_nodes"] * N
    alg = P * N * (F * 4 + 4)
    print(f"{name}: {ms:.2f} ms  node-ops/s {nodeops / ms * 1e3:.3e}  HBM-roofline frac {alg / 6547.8e9 / (ms * 1e-3):.3f}  ok frac {ok.float().mean().item():.3f}")
run("C2", 1000, 8, 5, 1 << 16)
run("C6", 10000, 8, 5, 1 << 20)
run("C4-shard(1/8)", 10000, 12, 10, 1 << 17)
