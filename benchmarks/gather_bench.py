#!/usr/bin/env python
"""Multi-GPU exchange step of the path, timed separately from the kernel:

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 benchmarks/gather_bench.py

Per rank: the bench.py workload (1 000 depth-8 trees, 2^16 samples per GPU).  Three lines:
  kernel        local evaluation only (what bench.py times)
  nccl_gather   local evaluation, then NCCL all-gather of the result rows (sharded.gather_results)
  fused_gather  the interpreter stores straight into the root GPU's (P, N_total) matrix through
                NVLink peer memory (sharded.FusedGather): no collective on the data path
Times: CUDA events per rank, max over ranks, best of `reps`."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dexb200  # noqa: E402
from dexb200 import device as D, sharded, treegen  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    P, NL, F, reps = 1000, 1 << 16, 5, 10
    ops = dexb200.OperatorEnum(treegen.OPSET_A)
    nodes, offsets = treegen.gen_population(P, 8, 2, 4, F, seed=0)
    ctx = D.Context.get(local)
    pop = D.Population(None, ops, np.float32, wire=(nodes, offsets), ctx=ctx)
    N = NL * world
    Xl = torch.randn((NL, F), device=f"cuda:{local}").T
    out = torch.empty((P, NL), device=f"cuda:{local}")
    ok = torch.empty(P, dtype=torch.uint8, device=f"cuda:{local}")
    fg = sharded.FusedGather(ctx, P, N, torch.float32, root=0) if world > 1 else None

    def timed(f):
        best = 1e30
        for _ in range(reps + 2):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            f()
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b)], device=f"cuda:{local}", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t.item()))
        return best

    t_kernel = timed(lambda: pop.eval(Xl, out=out, ok=ok))
    lines = {"n_gpus": world, "trees": P, "samples_per_gpu": NL, "result_bytes_per_gpu": P * NL * 4,
             "kernel_ms": t_kernel}
    if world > 1:
        def nccl():
            o, k = pop.eval(Xl, out=out, ok=ok)
            sharded.gather_results(o, k, N)
        lines["nccl_allgather_ms"] = timed(nccl)
        lines["fused_peer_gather_ms"] = timed(lambda: fg.eval(pop, Xl))
        fg.close()
    if rank == 0:
        print(json.dumps(lines), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
