#!/usr/bin/env python
"""Per-call cost of SMALL evaluations (symbolic-regression datasets are often a few hundred to a
few thousand rows): wall time of one synchronous call through the C ABI, device-resident
(`dex_eval` + stream synchronise) and host-to-host (`dex_eval_host`).  One JSON line per shape.
(For the CPU side of the comparison see `bench.py --impl reference`: the C port of the reference
algorithm sustains about 1.9 x 10^10 node-ops/s on 16 threads, 1.2 x 10^9 on one.)

    python benchmarks/small_calls.py [--reps 200]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=200)
    args = ap.parse_args()
    import torch
    import dexb200
    from dexb200 import device as D, treegen
    ops = dexb200.OperatorEnum(treegen.OPSET_A)
    F = 5
    for P, N in ((1, 100), (1, 10_000), (100, 100), (100, 1000), (1000, 100), (1000, 1000), (1000, 10_000)):
        nodes, offsets = treegen.gen_population(P, 8, 2, 4, F, seed=0)
        pop = D.Population(None, ops, np.float32, wire=(nodes, offsets))
        Xh = torch.from_numpy(np.random.default_rng(0).standard_normal((N, F)).astype(np.float32)).pin_memory()
        Xd = Xh.cuda()
        out = torch.empty((P, N), device="cuda")
        ok = torch.empty(P, dtype=torch.uint8, device="cuda")
        out_h = torch.empty((P, N)).pin_memory()
        ok_h = torch.empty(P, dtype=torch.uint8).pin_memory()

        def timed(f):
            for _ in range(5):
                f()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.reps):
                f()
            return (time.perf_counter() - t0) / args.reps * 1e6

        def dev_call():
            pop.eval(Xd.T, out=out, ok=ok)
            pop.ctx.synchronize()

        def dev_async():
            pop.eval(Xd.T, out=out, ok=ok)

        line = {"n_trees": P, "nsamples": N, "nodes": int(offsets[-1]),
                "dex_eval_plus_sync_us": timed(dev_call),
                "dex_eval_host_us": timed(lambda: pop.eval_host(Xh, out_h, ok_h))}
        t = timed(dev_async)
        torch.cuda.synchronize()
        line["dex_eval_enqueue_only_us"] = t
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
