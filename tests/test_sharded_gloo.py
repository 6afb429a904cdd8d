"""Host-side logic of the multi-GPU path on CPU: column blocks, result gather and flag
reduction with the gloo backend at world_size 2 and 3 (uneven blocks).  The per-rank
"evaluation" is the CPU oracle standing in for the device kernel: what is under test here
is the sharding plumbing of dexb200/sharded.py, not arithmetic."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dexb200
from dexb200 import sharded, treegen


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, N, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle
        ops = dexb200.OperatorEnum(treegen.OPSET_A)
        nodes, offsets = treegen.gen_population(12, 5, 2, 4, 3, seed=3)
        X = np.random.default_rng(0).standard_normal((3, N)).astype(np.float32)
        s, e = sharded.column_block(N, rank, world)
        out_l, ok_l = oracle.eval_population(nodes, offsets, ops.opcodes, np.ascontiguousarray(X[:, s:e]),
                                             oracle.DEFAULT_FLAGS | oracle.ELEMENTWISE, nthreads=1)
        out, ok = sharded.gather_results(torch.from_numpy(out_l), torch.from_numpy(ok_l.astype(np.uint8)), N)
        full, ok_full = oracle.eval_population(nodes, offsets, ops.opcodes, X,
                                               oracle.DEFAULT_FLAGS | oracle.ELEMENTWISE, nthreads=1)
        a, b = out.numpy(), full
        same = bool(((a == b) | (np.isnan(a) & np.isnan(b))).all())
        q.put((rank, same, bool((ok.numpy().astype(bool) == ok_full).all()), (s, e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,N", [(2, 1000), (3, 1001), (2, 7)])
def test_gather_and_flag_reduce(world, N):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    blocks = sorted(r[3] for r in res)
    assert blocks[0][0] == 0 and blocks[-1][1] == N
    assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))      # contiguous cover
    assert max(e - s for s, e in blocks) - min(e - s for s, e in blocks) <= 1     # balanced
    for rank, same, ok_same, _ in res:
        assert same, f"rank {rank}: gathered rows differ from the unsharded evaluation"
        assert ok_same, f"rank {rank}: reduced flags differ"


def test_column_block_edge_cases():
    assert sharded.column_block(0, 0, 4) == (0, 0)
    assert [sharded.column_block(3, r, 8) for r in range(8)] == \
        [(0, 1), (1, 2), (2, 3), (3, 3), (3, 3), (3, 3), (3, 3), (3, 3)]
    assert sharded.column_block(1 << 20, 7, 8) == (7 << 17, 8 << 17)
