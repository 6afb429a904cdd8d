import sys; sys.path.insert(0, '.')
import numpy as np, torch
import dexb200
from dexb200 import device as D, treegen
from oracle import oracle
spec = treegen.OPSET_B
ops = dexb200.OperatorEnum(spec)
dtype = np.float32
nodes, offsets = treegen.gen_population(300, 8, len(spec[1]), len(spec[2]), 5, seed=11, dtype=dtype)
X = np.random.default_rng(0).standard_normal((5, 3000)).astype(dtype)
pop = D.Population(None, ops, dtype, wire=(nodes, offsets))
out, ok = pop.eval(X)
out = out.cpu().numpy(); ok = ok.cpu().numpy().astype(bool)
ref, rok = oracle.eval_population(nodes, offsets, ops.opcodes, X)
bad = np.nonzero(ok != rok)[0]
print("mismatch trees", bad)
ins, off = pop.tape()
for t in bad[:3]:
    tree = dexb200.from_wire(nodes[offsets[t]:offsets[t+1]])
    print(t, "gpu ok", ok[t], "oracle ok", rok[t], dexb200.string_tree(tree, ops))
    print(" gpu finite all:", np.isfinite(out[t]).all(), " oracle finite all:", np.isfinite(ref[t]).all())
    j = np.nonzero(~np.isfinite(out[t]) | ~np.isfinite(ref[t]))[0][:5]
    print(" nonfinite idx", j, out[t][j], ref[t][j])
    for w in ins[off[t]:off[t+1]]:
        w0=int(w[0]); print("   ", D.lib().dex_handler_name(w0&0x3f).decode(), dexb200.OPCODE_INFO[(w0>>8)&0xff][0], "src", (w0>>16)&3, (w0>>18)&3, "flags", [n for n,b in [("PUSH",20),("OUT",21),("A",22),("B",23),("ALW",24),("GRD",25),("CC",26)] if w0&(1<<b)], "rows", int(w[1])&0xffff, int(w[1])>>16, w0>>27, np.array([w[2]],dtype=np.uint32).view(np.float32)[0])
    # elementwise compare
    d = np.abs(out[t]-ref[t]); print(" max abs diff", np.nanmax(d))
