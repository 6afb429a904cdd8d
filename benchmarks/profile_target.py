"""Launches one configuration a few times so that ncu can capture its kernel (profiles/run_profile.sh):
    python benchmarks/profile_target.py C3 | C4 | C6"""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dexb200
from dexb200 import device as D, treegen
which = sys.argv[1] if len(sys.argv) > 1 else "C3"
ops = dexb200.OperatorEnum(treegen.OPSET_A)
if which == "C3":
    nodes, offsets = treegen.gen_population(1000, 8, 2, 4, 5, seed=0)
    pop = D.Population(None, ops, np.float32, wire=(nodes, offsets))
    X = torch.randn((1 << 16, 5), device="cuda")
    for _ in range(4):
        pop.eval_grad(X.T, D.GRAD_FEATURES)
elif which == "C4":      # one of the 8 sample shards of configs[3]
    nodes, offsets = treegen.gen_population(10000, 12, 2, 4, 10, seed=0)
    pop = D.Population(None, ops, np.float32, wire=(nodes, offsets))
    X = torch.randn((1 << 17, 10), device="cuda")
    out = torch.empty((10000, 1 << 17), device="cuda"); ok = torch.empty(10000, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        pop.eval(X.T, out=out, ok=ok)
elif which == "C6":
    nodes, offsets = treegen.gen_population(10000, 8, 2, 4, 5, seed=0)
    pop = D.Population(None, ops, np.float32, wire=(nodes, offsets))
    X = torch.randn((1 << 20, 5), device="cuda")
    out = torch.empty((10000, 1 << 20), device="cuda"); ok = torch.empty(10000, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        pop.eval(X.T, out=out, ok=ok)
torch.cuda.synchronize()
