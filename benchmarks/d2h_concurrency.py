#!/usr/bin/env python
"""Why does the host-result path (dex_eval_host, `e2e` of bench.py) not scale across GPUs?
Plain cudaMemcpyAsync D2H of the bench's result size (262 MB per rank, pinned destination) with
1, 2, 4, ... ranks copying at the same time — no kernels of this repository involved:

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 benchmarks/d2h_concurrency.py

Prints the aggregate host-ingest rate per number of concurrent ranks: the platform ceiling that `e2e`
runs into (the fused-loss entry point, which returns P doubles instead of P x N floats, avoids it)."""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    nbytes = 1000 * 65536 * 4
    src = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dst = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    res = {}
    k = 1
    while k <= world:
        for _ in range(2):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if rank < k:
                for _ in range(5):
                    dst.copy_(src, non_blocking=True)
                torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            t = torch.tensor([dt if rank < k else 0.0], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[k] = {"ms_per_copy": float(t.item()) / 5 * 1e3, "aggregate_GBps": k * nbytes * 5 / float(t.item()) / 1e9}
        k *= 2
    if rank == 0:
        print(json.dumps({"bytes_per_rank": nbytes, "concurrent_ranks": res}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
