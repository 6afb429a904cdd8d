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

* :class:`FusedGather` — gather-to-root WITHOUT a collective: every rank's interpreter kernel
  stores its result rows straight into the root GPU's ``(P, N)`` matrix through an NVLink peer
  mapping (CUDA IPC), so the transfer overlaps the arithmetic tile by tile and nothing is
  staged or copied afterwards.

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


class FusedGather:
    """Fused evaluate + gather-to-root over NVLink peer memory (one process per GPU).

    ``root`` owns a ``(P, N_total)`` result matrix allocated with :class:`dexb200.device.PeerBuffer`;
    its CUDA IPC handle is exchanged once (``all_gather_object``) and every other rank maps it.
    ``eval(pop, X_local)`` then launches the ordinary interpreter kernel with
    ``out = root_matrix + first_column_of_this_rank`` and row stride ``N_total``: the result stores
    of rank r travel through NVLink and land in place on the root — no staging buffer, no
    collective on the data path.  Completion: each rank synchronises its stream, then a barrier;
    the flags take the tiny ``min`` all-reduce.  Reusable across calls (same P, N_total, dtype)."""

    def __init__(self, ctx, n_trees, N_total, dtype=torch.float32, root=0, group=None):
        from .device import PeerBuffer
        self.ctx, self.P, self.N, self.dtype, self.root, self.group = ctx, int(n_trees), int(N_total), dtype, root, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.es = torch.empty((), dtype=dtype).element_size()
        self.own = None
        handle = None
        if self.rank == root:
            self.own = PeerBuffer(ctx, max(self.P * self.N * self.es, 1))
            handle = self.own.handle()
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, handle, group=group)
            handle = handles[root]
        self.peer = self.own if self.rank == root else PeerBuffer.open(ctx, handle)
        self.start, self.stop = column_block(self.N, self.rank, self.world)

    def result(self):
        """The gathered ``(P, N_total)`` matrix (root only)."""
        assert self.rank == self.root
        return self.own.as_tensor((self.P, self.N), self.dtype)

    def eval(self, pop, X_local, *, early_exit=True):
        """Evaluate this rank's column block into the root's matrix; returns the reduced flags
        (and, on the root, the gathered matrix) once every rank's stores have landed."""
        assert X_local.shape[1] == self.stop - self.start
        ok = pop.eval_into(X_local, self.peer.ptr + self.start * self.es, self.N, early_exit=early_exit)
        torch.cuda.current_stream().synchronize()
        okr = allreduce_ok(ok, self.group)      # also orders every rank's completed stores before the read
        if self.world > 1:
            dist.barrier(group=self.group)
        return (self.result() if self.rank == self.root else None), okr

    def close(self):
        if self.world > 1:
            dist.barrier(group=self.group)      # nobody unmaps / frees while a peer may still write
        if self.peer is not self.own and self.peer is not None:
            self.peer.close()
        self.peer = None
        if self.world > 1:
            dist.barrier(group=self.group)
        if self.own is not None:
            self.own.close()
            self.own = None
