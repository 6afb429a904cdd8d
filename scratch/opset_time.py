import sys; sys.path.insert(0, ".")
import numpy as np, torch, dexb200
from dexb200 import device as D, treegen
N = 1 << 16
X = torch.randn((N, 5), device="cuda")
def timeit(f, reps=10):
    f(); torch.cuda.synchronize(); ts=[]
    for _ in range(reps):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
for name, spec, nu in (("A: cos exp", {1: ("cos", "exp"), 2: ("+", "-", "/", "*")}, 2),
                 ("cos exp log", {1: ("cos", "exp", "log"), 2: ("+", "-", "/", "*")}, 3),
                 ("cos exp safe_log tanh", {1: ("cos", "exp", "safe_log", "tanh"), 2: ("+", "-", "/", "*")}, 4),
                 ("sqrt abs square sin", {1: ("sqrt", "abs", "square", "sin"), 2: ("+", "-", "/", "*")}, 4)):
    ops = dexb200.OperatorEnum(spec)
    nodes, offsets = treegen.gen_population(1000, 8, nu, 4, 5, seed=0)
    pop = D.Population(None, ops, np.float32, wire=(nodes, offsets))
    out = torch.empty((1000, N), device="cuda"); ok = torch.empty(1000, dtype=torch.uint8, device="cuda")
    t = timeit(lambda: pop.eval(X.T, out=out, ok=ok))
    tg = timeit(lambda: pop.eval_grad(X.T, D.GRAD_FEATURES))
    print(f"{name:28s} eval {t:.3f} ms  {pop.info['n_nodes']*N/t*1e-9:.0f} Gnode-ops/s | grad {tg:.3f} ms | instrs {pop.info['n_folded_instructions']} ok {float(ok.float().mean()):.2f}")
for name, spec, nu, nb in (("pow: + - * / ^ cos exp", {1: ("cos", "exp"), 2: ("+", "-", "/", "*", "^")}, 2, 5),
                           ("ternary fma", {1: ("cos", "exp"), 2: ("+", "-", "/", "*"), 3: ("fma",)}, 2, 4)):
    ops = dexb200.OperatorEnum(spec)
    nodes, offsets = treegen.gen_population(1000, 8, nu, nb, 5, seed=0)
    pop = D.Population(None, ops, np.float32, wire=(nodes, offsets))
    out = torch.empty((1000, N), device="cuda"); ok = torch.empty(1000, dtype=torch.uint8, device="cuda")
    t = timeit(lambda: pop.eval(X.T, out=out, ok=ok))
    tg = timeit(lambda: pop.eval_grad(X.T, D.GRAD_FEATURES))
    print(f"{name:28s} eval {t:.3f} ms | grad {tg:.3f} ms | instrs {pop.info['n_folded_instructions']} generic {pop.info['n_generic']} ok {float(ok.float().mean()):.2f}")
