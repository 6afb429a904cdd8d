"""Multi-GPU path on real devices (needs >= 2 GPUs; skipped otherwise): sample-sharded
evaluation, the NCCL all-gather of result rows and the fused evaluate + gather-to-root over
NVLink peer memory (dexb200/sharded.py) must all reproduce the single-GPU evaluation bit for bit
(the same kernel evaluates the same columns; only the tiling differs)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, N, q):
    import torch.distributed as dist
    import dexb200
    from dexb200 import device as D, sharded, treegen
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        ops = dexb200.OperatorEnum(treegen.OPSET_A)
        nodes, offsets = treegen.gen_population(64, 6, 2, 4, 5, seed=11)
        X = np.random.default_rng(5).standard_normal((5, N)).astype(np.float32)
        ctx = D.Context.get(rank)
        pop = D.Population(None, ops, np.float32, wire=(nodes, offsets), ctx=ctx)
        s, e = sharded.column_block(N, rank, world)
        Xl = np.ascontiguousarray(X[:, s:e])
        # (1) local block + NCCL all-gather of the rows
        out_l, ok_l = pop.eval(Xl)
        full_nccl, ok_nccl = sharded.gather_results(out_l, ok_l, N)
        # (2) fused evaluate + gather-to-root through peer memory
        fg = sharded.FusedGather(ctx, pop.n_trees, N, torch.float32, root=0)
        full_fused, ok_fused = fg.eval(pop, Xl)
        full_fused2, _ = fg.eval(pop, Xl)          # reusable
        # (3) the unsharded evaluation on this GPU
        ref, ok_ref = pop.eval(X)
        torch.cuda.synchronize()

        def same(a, b):
            return bool(torch.equal(a.cpu(), b.cpu()))

        good = ok_ref.bool()    # rows of incomplete trees are unspecified (early exit), as in the reference
        assert int(good.sum()) > 10
        res = {"rank": rank, "nccl": same(full_nccl[good], ref[good]), "ok_nccl": same(ok_nccl, ok_ref),
               "ok_fused": same(ok_fused, ok_ref)}
        if rank == 0:
            res["fused"] = same(full_fused[good], ref[good]) and same(full_fused2[good], ref[good])
        fg.close()
        q.put(res)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N", [4096 + 37, 100_000])
def test_sharded_gather_and_fused_peer_gather_match_single_gpu(N):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for r in res:
        assert r["nccl"], f"rank {r['rank']}: NCCL-gathered rows differ from the unsharded evaluation"
        assert r["ok_nccl"] and r["ok_fused"], f"rank {r['rank']}: reduced flags differ"
        if r["rank"] == 0:
            assert r["fused"], "fused peer-memory gather differs from the unsharded evaluation"


def test_single_process_shard_eval_over_two_devices():
    """dex_shard_eval_host: one host thread, one context + packed population per device; the rows
    of every column block land in place and equal the one-device evaluation bit for bit."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    import dexb200
    from dexb200 import device as D, treegen
    ops = dexb200.OperatorEnum(treegen.OPSET_A)
    nodes, offsets = treegen.gen_population(300, 7, 2, 4, 5, seed=4)
    N = 50_000 + 13
    Xh = np.random.default_rng(3).standard_normal((N, 5)).astype(np.float32)
    pops = [D.Population(None, ops, np.float32, wire=(nodes, offsets), ctx=D.Context.get(d)) for d in (0, 1)]
    out = np.empty((300, N), np.float32)
    ok = np.empty(300, np.uint8)
    D.shard_eval_host(pops, Xh, out, ok)
    ref = np.empty((300, N), np.float32)
    rok = np.empty(300, np.uint8)
    pops[0].eval_host(Xh, ref, rok)
    good = rok.astype(bool)     # rows of incomplete trees are unspecified (early exit)
    assert good.sum() > 50 and (ok == rok).all() and np.array_equal(out[good], ref[good])


def test_single_process_shard_eval_gathers_on_the_root_device():
    """dex_shard_eval: every device's kernel stores its column block straight into the root device's
    matrix through peer memory; flags are AND-reduced on the root; asynchronous in the root stream."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    import dexb200
    from dexb200 import device as D, treegen
    ops = dexb200.OperatorEnum(treegen.OPSET_A)
    nodes, offsets = treegen.gen_population(300, 7, 2, 4, 5, seed=4)
    N = 70_000 + 5
    Xh = np.random.default_rng(3).standard_normal((N, 5)).astype(np.float32)
    pops = [D.Population(None, ops, np.float32, wire=(nodes, offsets), ctx=D.Context.get(d)) for d in (0, 1)]
    for root in (0, 1):
        blocks = []
        for d in (0, 1):
            s, e = N * d // 2, N * (d + 1) // 2
            blocks.append(torch.from_numpy(Xh[s:e]).to(f"cuda:{d}").T)       # (F, n_d), column-major memory
        out = torch.full((300, N), -3.0, device=f"cuda:{root}")
        ok = torch.full((300,), 7, dtype=torch.uint8, device=f"cuda:{root}")
        D.shard_eval(pops, blocks, out, ok, root=root)
        pops[root].ctx.synchronize()
        ref, rok = pops[root].eval(torch.from_numpy(Xh).to(f"cuda:{root}").T)
        torch.cuda.synchronize(root)
        good = rok.bool()
        assert int(good.sum()) > 50 and torch.equal(ok, rok)
        assert torch.equal(out[good], ref[good])
