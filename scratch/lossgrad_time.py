import sys; sys.path.insert(0, ".")
import numpy as np, torch, dexb200
from dexb200 import device as D, treegen
ops = dexb200.OperatorEnum(treegen.OPSET_A)
nodes, offsets = treegen.gen_population(1000, 8, 2, 4, 5, seed=0)
pop = D.Population(None, ops, np.float32, wire=(nodes, offsets))
N = 1 << 16
X = torch.randn((N, 5), device="cuda"); y = torch.randn(N, device="cuda")
def timeit(f, reps=10):
    f(); torch.cuda.synchronize(); ts=[]
    for _ in range(reps):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
for mode,name in ((D.GRAD_FEATURES,"features"),(D.GRAD_CONSTANTS,"constants"),(D.GRAD_BOTH,"both")):
    t1 = timeit(lambda: pop.eval_grad(X.T, mode)); t2 = timeit(lambda: pop.eval_loss_grad(X.T, y, mode))
    print(name, "eval_grad ms", round(t1,3), "fused loss+grad ms", round(t2,3))
