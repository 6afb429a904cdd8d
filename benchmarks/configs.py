#!/usr/bin/env python
"""Times every configuration of BASELINE.json (plus C6, the north_star target sentence) on
one GPU, device-resident, and prints one JSON line per configuration.

    python benchmarks/configs.py [--only C2,C3] [--reps 5]

bench.py is the contract benchmark (configs[1]); this script is the companion that shows the
other configurations: C1 README single tree (Float64), C3 gradients of C2, C4 one of the 8
sample shards of 10k depth-12 trees x 2^20 samples, C5 parametric, C6 10k trees x 2^20 samples.
Roofline fraction = algorithmic bytes / measured HBM copy bandwidth / time.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def timeit(f, reps):
    import torch
    f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        f()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    only = set(args.only.split(",")) if args.only else None
    import torch
    import dexb200
    from dexb200 import device as D, treegen
    peak = hbm_peak()
    ops = dexb200.OperatorEnum(treegen.OPSET_A)

    def emit(name, desc, ms, ms_med, nodeops, alg_bytes, extra=None):
        line = {"config": name, "desc": desc, "ms": ms, "ms_median": ms_med,
                "node_ops_per_s": nodeops / (ms * 1e-3), "algorithmic_bytes": alg_bytes,
                "hbm_roofline_frac": alg_bytes / (peak * 1e9) / (ms * 1e-3), "hbm_peak_gbs": peak}
        line.update(extra or {})
        print(json.dumps(line), flush=True)

    def want(n):
        return only is None or n in only

    if want("C1"):
        o1 = dexb200.OperatorEnum({1: ("cos",), 2: ("+", "-", "*")})
        dexb200.extend_operators(o1)
        tree = dexb200.Node(feature=1, T=np.float64) * dexb200.call("cos", dexb200.Node(feature=2, T=np.float64) - 3.2)
        X = torch.randn((100, 2), device="cuda", dtype=torch.float64)
        pop = D.Population([tree], o1, np.float64)
        ms, med = timeit(lambda: pop.eval(X.T), 50)
        emit("C1", "README tree x1*cos(x2-3.2), Float64, X 2x100 (launch-latency bound; reference CPU: 607 ns)",
             ms, med, 6 * 100, 100 * (2 * 8 + 8))

    pops = {}

    def population(P, depth, F, n_params=0):
        key = (P, depth, F, n_params)
        if key not in pops:
            nodes, offsets = treegen.gen_population(P, depth, 2, 4, F, seed=0, n_params=n_params)
            pops[key] = D.Population(None, ops, np.float32, wire=(nodes, offsets))
        return pops[key]

    def eval_cfg(name, desc, P, depth, F, N):
        pop = population(P, depth, F)
        X = torch.randn((N, F), device="cuda", dtype=torch.float32)
        out = torch.empty((P, N), device="cuda", dtype=torch.float32)
        ok = torch.empty(P, device="cuda", dtype=torch.uint8)
        ms, med = timeit(lambda: pop.eval(X.T, out=out, ok=ok), args.reps)
        emit(name, desc, ms, med, pop.info["n_nodes"] * N, P * N * (F * 4 + 4),
             {"n_nodes": pop.info["n_nodes"], "tape_instructions": pop.info["n_instructions"],
              "stack_rows": pop.info["max_stack"], "complete_fraction": float(ok.float().mean())})

    if want("C2"):
        eval_cfg("C2", "1k depth-8 trees, 5 features, Float32, 2^16 samples", 1000, 8, 5, 1 << 16)
    if want("C2-f64"):
        P, F, N = 1000, 5, 1 << 16
        nodes, offsets = treegen.gen_population(P, 8, 2, 4, F, seed=0, dtype=np.float64)
        pop64 = D.Population(None, ops, np.float64, wire=(nodes, offsets))
        X = torch.randn((N, F), device="cuda", dtype=torch.float64)
        out = torch.empty((P, N), device="cuda", dtype=torch.float64)
        ok = torch.empty(P, device="cuda", dtype=torch.uint8)
        ms, med = timeit(lambda: pop64.eval(X.T, out=out, ok=ok), args.reps)
        emit("C2-f64", "the C2 population in Float64 (C++ handlers, no PTX loop)", ms, med, pop64.info["n_nodes"] * N,
             P * N * (F * 8 + 8))
        ms, med = timeit(lambda: pop64.eval_grad(X.T, D.GRAD_FEATURES), args.reps)
        emit("C3-f64", "d/dX of the C2 population in Float64", ms, med, pop64.info["n_nodes"] * N,
             P * N * (F * 8 + (1 + F) * 8))
    if want("C3"):
        P, F, N = 1000, 5, 1 << 16
        pop = population(P, 8, F)
        X = torch.randn((N, F), device="cuda", dtype=torch.float32)
        ms, med = timeit(lambda: pop.eval_grad(X.T, D.GRAD_FEATURES), args.reps)
        emit("C3", "eval_grad_tree_array d/dX (G=5) of the C2 population", ms, med, pop.info["n_nodes"] * N,
             P * N * (F * 4 + (1 + F) * 4))
    if want("C4"):
        eval_cfg("C4/8", "one of 8 sample shards: 10k depth-12 trees, 10 features, Float32, 2^17 of 2^20 samples",
                 10000, 12, 10, 1 << 17)
    if want("C5"):
        P, F, N, npar, ncls = 1000, 5, 1 << 18, 3, 10
        pop = population(P, 8, F, n_params=npar)
        X = torch.randn((N, F), device="cuda", dtype=torch.float32)
        params = torch.randn((P, npar, ncls), dtype=torch.float32)
        cls = torch.randint(0, ncls, (N,))
        out = torch.empty((P, N), device="cuda", dtype=torch.float32)
        ok = torch.empty(P, device="cuda", dtype=torch.uint8)
        ms, med = timeit(lambda: pop.eval_parametric(X.T, params, cls, out=out, ok=ok), args.reps)
        emit("C5", "ParametricExpression: 1k trees, 3 params x 10 classes, Float32, 2^18 samples", ms, med,
             pop.info["n_nodes"] * N, P * N * (F * 4 + 4 + npar * 4 + 4),
             {"n_generic_instructions": pop.info["n_generic"], "tape_instructions": pop.info["n_instructions"]})
    if want("C6"):
        eval_cfg("C6", "north_star target: 10k depth-8 trees, 5 features, Float32, 2^20 samples, 1 GPU",
                 10000, 8, 5, 1 << 20)
    if want("LOSS"):
        P, F, N = 10000, 5, 1 << 20
        pop = population(P, 8, F)
        X = torch.randn((N, F), device="cuda", dtype=torch.float32)
        y = torch.randn(N, device="cuda")
        ms, med = timeit(lambda: pop.eval_loss(X.T, y), args.reps)
        emit("C6-loss", "fused MSE per tree on the C6 workload (no result matrix written)", ms, med,
             pop.info["n_nodes"] * N, P * N * (F * 4))


if __name__ == "__main__":
    main()
