#!/usr/bin/env python
"""Times the C2 population shape (1k depth-8 trees, 5 features, Float32, 2^16 samples) under
several operator sets: the tree SHAPES are those of bench.py (the generator draws operator
indices, not operators), only the meaning of each index changes.  Shows what an operator costs
when it leaves the natively implemented handlers of the Float32 loop (`^`, tanh, ...).

    python benchmarks/opset_sweep.py [--reps 5] [--grad]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SETS = {
    "A: cos exp | + - / *  (bench.py)": {1: ("cos", "exp"), 2: ("+", "-", "/", "*")},
    "cos exp | + - ^ *": {1: ("cos", "exp"), 2: ("+", "-", "^", "*")},
    "sin log | + - / *": {1: ("sin", "log"), 2: ("+", "-", "/", "*")},
    "safe_sqrt safe_log | + - / *": {1: ("safe_sqrt", "safe_log"), 2: ("+", "-", "/", "*")},
    "tanh abs | + - max *": {1: ("tanh", "abs"), 2: ("+", "-", "max", "*")},
    "square cube | + - / *": {1: ("square", "cube"), 2: ("+", "-", "/", "*")},
    "atan erf | + - / *": {1: ("atan", "erf"), 2: ("+", "-", "/", "*")},
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--grad", action="store_true")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    args = ap.parse_args()
    import torch
    import dexb200
    from dexb200 import device as D, treegen
    P, N, F = 1000, 1 << 16, 5
    npdt, tdt = (np.float32, torch.float32) if args.dtype == "f32" else (np.float64, torch.float64)
    nodes, offsets = treegen.gen_population(P, 8, 2, 4, F, seed=0)
    X = torch.randn((N, F), device="cuda", dtype=tdt).T      # (F, N) view, column-major memory
    out = torch.empty((P, N), device="cuda", dtype=tdt)
    ok = torch.empty((P,), device="cuda", dtype=torch.uint8)
    for name, spec in SETS.items():
        ops = dexb200.OperatorEnum(spec)
        pop = D.Population(None, ops, npdt, wire=(nodes, offsets))
        ts = []
        for _ in range(args.reps + 1):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            pop.eval(X, out=out, ok=ok)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        line = {"opset": name, "dtype": args.dtype, "eval_ms": min(ts[1:]), "complete": float(ok.float().mean()),
                "generic_instructions": pop.info["n_generic"], "instructions": pop.info["n_instructions"]}
        if args.grad:
            ts = []
            for _ in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                pop.eval_grad(X)                 # d/dX; its buffers come from torch's caching allocator
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            line["grad_ms"] = min(ts[1:])
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
