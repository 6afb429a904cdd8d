"""Sample-sharded evaluation across GPUs (one process per GPU, torch.distributed).

The path shards embarrassingly along the sample axis: columns of the column-major ``X`` are
contiguous, so rank r evaluates the contiguous block ``X[:, start_r:stop_r]`` against the
replicated population and produces ``out_r[P, stop_r - start_r]``.  There is NO data-path
collective.  Two optional exchange steps exist for callers that want the reference's
single-array view:

* :func:`gather_results` — all-gather of the result rows (NCCL over NVLink on GPUs, gloo in
  the CPU tests) into ``out[P, N]`` on every rank;
* the ``min`` all-reduce of the per-tree ``complete`` flags (a tree is complete iff it is
  complete on every shard).

Nothing here touches arithmetic; it is plumbing around :class:`dexb200.device.Population`.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def column_block(N: int, rank: int, world: int):
    """Contiguous, balanced column block of rank ``rank``: the first ``N % world`` ranks get
    one extra column.  Returns (start, stop)."""
    base, extra = divmod(int(N), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def allreduce_ok(ok_local: torch.Tensor, group=None) -> torch.Tensor:
    """complete[t] = min over shards (uint8 0/1)."""
    ok = ok_local.to(torch.int32)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    return ok.to(torch.uint8)


def gather_results(out_local: torch.Tensor, ok_local: torch.Tensor, N: int, group=None):
    """All-gather the per-rank result blocks into ``out[P, N]`` (every rank gets the whole
    matrix) and reduce the flags.  ``out_local`` is ``[P, stop_r - start_r]`` for this rank's
    :func:`column_block`.  Blocks may differ by one column; they are padded for the
    collective and trimmed afterwards."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return out_local, ok_local
    world = dist.get_world_size(group)
    P = out_local.shape[0]
    width = -(-int(N) // world)  # widest block
    padded = out_local.new_zeros((P, width))
    padded[:, : out_local.shape[1]] = out_local
    flat = out_local.new_empty((world * P, width))  # rank-major concatenation along dim 0
    dist.all_gather_into_tensor(flat, padded.contiguous(), group=group)
    gathered = flat.view(world, P, width)
    out = out_local.new_empty((P, int(N)))
    for r in range(world):
        s, e = column_block(N, r, world)
        out[:, s:e] = gathered[r, :, : e - s]
    return out, allreduce_ok(ok_local, group)


def eval_population_sharded(pop, X_local, *, early_exit=True, gather=False, N_total=None, group=None):
    """Evaluate this rank's column block; optionally gather.  ``X_local`` has shape (F, n_local)
    (the reference's layout).  Returns (out, ok) — local block, or the gathered (P, N_total)."""
    out, ok = pop.eval(X_local, early_exit=early_exit)
    if not gather:
        return out, allreduce_ok(ok, group)
    if N_total is None:
        n = torch.tensor([out.shape[1]], device=out.device, dtype=torch.int64)
        if dist.is_initialized():
            dist.all_reduce(n, group=group)
        N_total = int(n.item())
    return gather_results(out, ok, N_total, group)
