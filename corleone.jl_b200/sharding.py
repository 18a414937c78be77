"""Multi-GPU sharding of the (interval x candidate) index space -- one process per GPU.

Every (interval, candidate) unit is independent given the node values (that is the point of multiple
shooting, src/multiple_shooting.jl:118-130).  The *candidate* axis is partitioned contiguously across
ranks so that shooting defects x_end(i) - x0(i+1) (:159), quadrature running sums (:171-177) and
per-candidate objectives stay rank-local.  The reference's counterpart is `pmap` over intervals
(src/Corleone.jl:24), which ships every interval's arguments over TCP; here nothing but the
cross-candidate partial sums crosses NVLink:

    allreduce(SUM)   objective_sum (1), grad_sum (nvars), fim_sum (n*n)   -- one flat buffer, one call
    allgather        per-candidate rows (defects, objectives) only when a caller wants them assembled

When there are fewer candidates than GPUs (B < world size: the reference's own B = 1 use on a long horizon), the
*interval* axis is partitioned instead (`IntervalShardedEvaluator`): rank r integrates the contiguous interval range
`partition(I, world, r)` of every candidate (cb200_set_interval_range).  A matching defect belongs to the interval that
ends at the node and needs only that interval's end state and the next interval's variables, which every rank holds,
so nothing is exchanged before the epilogue; the per-rank outputs are full-size and zero outside the rank's range and
assemble with

    allreduce(SUM)   [defects | objective | gradient | constraint-Jacobian values | end states ...] -- one flat buffer

Works on any torch.distributed backend: NCCL on the GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Shard:
    start: int
    stop: int

    @property
    def size(self) -> int:
        return self.stop - self.start


def partition(n_candidates: int, world: int, rank: int) -> Shard:
    """Contiguous split of range(n_candidates); the first n % world ranks hold one extra candidate."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} of {world}")
    q, r = divmod(int(n_candidates), world)
    start = rank * q + min(rank, r)
    return Shard(start, start + q + (1 if rank < r else 0))


def counts(n_candidates: int, world: int) -> list[int]:
    return [partition(n_candidates, world, r).size for r in range(world)]


def _dist():
    import torch.distributed as dist
    return dist


def combine_sums(parts: dict, device=None, group=None) -> dict:
    """All-reduce (SUM) of the cross-candidate partial sums in ONE collective: the arrays in `parts`
    (name -> 1-D float64 array / tensor, same shapes on every rank) are packed into one flat buffer."""
    import torch
    dist = _dist()
    names = sorted(parts)
    flats = [torch.as_tensor(np.asarray(parts[k], dtype=np.float64) if not torch.is_tensor(parts[k]) else parts[k],
                             dtype=torch.float64).reshape(-1) for k in names]
    if device is not None:
        flats = [f.to(device) for f in flats]
    buf = torch.cat(flats) if flats else torch.zeros(0, dtype=torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    out, off = {}, 0
    for k, f in zip(names, flats):
        out[k] = buf[off:off + f.numel()].reshape(np.shape(parts[k]) if not torch.is_tensor(parts[k]) else parts[k].shape)
        off += f.numel()
    return out


def gather_candidates(local, n_candidates: int, group=None, device=None):
    """All-gather per-candidate rows: local is (B_local, n) on this rank -> (n_candidates, n) on every
    rank, candidates in global order (ranks hold contiguous shards, possibly of unequal size).  `device`: where the
    collective runs (the rank's GPU for NCCL; None = where `local` lives, e.g. the CPU for gloo)."""
    import torch
    dist = _dist()
    local = torch.as_tensor(local)
    if device is not None:
        local = local.to(device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    cnt = counts(n_candidates, world)
    assert local.shape[0] == cnt[dist.get_rank(group)], (local.shape, cnt)
    mx = max(cnt)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:c] for b, c in zip(bufs, cnt)], dim=0)


class ShardedEvaluator:
    """Batch-sharded evaluation of one layer.

    `evaluate_local(ps_rows) -> dict` is the rank-local hot path (NativeLayer.eval_host on this rank's GPU);
    per-candidate results stay local, the keys listed in `sum_keys` are combined with one all-reduce.
    """

    def __init__(self, evaluate_local, sum_keys=("objective_sum", "grad_sum", "fim_sum"), group=None, device=None):
        self.evaluate_local = evaluate_local
        self.sum_keys = tuple(sum_keys)
        self.group = group
        self.device = device

    def rank_world(self):
        dist = _dist()
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(self.group), dist.get_world_size(self.group)
        return 0, 1

    def __call__(self, ps_all: np.ndarray, gather: tuple = ()):
        """ps_all: (B, nvars) known on every rank (or only this rank's rows matter).  Returns the local
        result dict with the summed keys replaced by their global values and `gather` keys assembled."""
        rank, world = self.rank_world()
        sh = partition(ps_all.shape[0], world, rank)
        res = dict(self.evaluate_local(ps_all[sh.start:sh.stop]))
        parts = {k: res[k] for k in self.sum_keys if k in res}
        if parts:
            res.update({k: v.cpu().numpy() for k, v in combine_sums(parts, self.device, self.group).items()})
        for k in gather:
            res[k] = gather_candidates(res[k], ps_all.shape[0], self.group, self.device).cpu().numpy()
        res["_shard"] = sh
        return res


class IntervalShardedEvaluator:
    """Interval-sharded evaluation of one layer for B < world size.

    `evaluate_range(ps_all, first, last) -> dict` is the rank-local hot path restricted to the intervals first:last
    (1-based, inclusive; NativeLayer.set_interval_range + eval_host): every array has its full shape and is zero where
    other intervals would have written.  All arrays are assembled by ONE all-reduce(SUM) of their concatenation
    (`status` flags are summed too: non-zero = some interval of the candidate failed)."""

    def __init__(self, evaluate_range, n_intervals: int, group=None, device=None):
        self.evaluate_range = evaluate_range
        self.n_intervals = int(n_intervals)
        self.group = group
        self.device = device

    def rank_world(self):
        dist = _dist()
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(self.group), dist.get_world_size(self.group)
        return 0, 1

    def __call__(self, ps_all: np.ndarray):
        rank, world = self.rank_world()
        sh = partition(self.n_intervals, world, rank)
        res = dict(self.evaluate_range(ps_all, sh.start + 1, sh.stop))     # an empty shard is first:first-1
        keys = [k for k in sorted(res) if isinstance(res[k], np.ndarray)]
        parts = {k: res[k].astype(np.float64).ravel() for k in keys}
        summed = combine_sums(parts, self.device, self.group)
        for k in keys:
            res[k] = summed[k].cpu().numpy().reshape(res[k].shape).astype(res[k].dtype)
        res["_shard"] = sh
        return res
