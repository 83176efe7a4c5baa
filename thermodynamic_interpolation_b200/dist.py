"""Multi-GPU plumbing: trajectories are independent, so the molecule axis is sharded in contiguous
chunks with replicated weights and NO collective on the data path.  Collectives
(torch.distributed; NCCL over NVLink on GPUs, gloo in the CPU tests) appear only in
  * `gather_samples`   - final all-gather of the per-rank samples,
  * `allreduce_stats`  - one fp64 all-reduce(SUM) of the reweighting partial sums,
  * `norm_allreduce`   - optional 2-scalar all-reduce per dopri5 step attempt so that shards share
                         torchdiffeq's batch-global RMS error norm (one step sequence for the job).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from .batch import MolBatch


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced split: the first (n_items % world) ranks get one extra item."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch, rank: int, world: int) -> MolBatch:
    """The molecules [lo,hi) of `batch` as a self-contained batch (node/edge arrays sliced, edge_index
    and batch vector re-based)."""
    ptr = batch.ptr
    n_mol = int(ptr.numel() - 1)
    lo, hi = shard_range(n_mol, rank, world)
    n_lo, n_hi = int(ptr[lo]), int(ptr[hi])
    counts = (ptr[1:] - ptr[:-1])
    ecount = counts * (counts - 1)
    eptr = torch.zeros_like(ptr)
    eptr[1:] = torch.cumsum(ecount, 0)
    e_lo, e_hi = int(eptr[lo]), int(eptr[hi])
    N, E = int(ptr[-1]), int(eptr[-1])
    out = {}
    for k in batch.keys():
        v = batch[k]
        if not torch.is_tensor(v):
            out[k] = v
        elif k == "ptr":
            out[k] = ptr[lo:hi + 1] - n_lo
        elif k == "batch":
            out[k] = v[n_lo:n_hi] - lo
        elif k == "edge_index":
            out[k] = v[:, e_lo:e_hi] - n_lo
        elif v.dim() >= 1 and v.shape[0] == N:
            out[k] = v[n_lo:n_hi]
        elif v.dim() >= 1 and v.shape[0] == E:
            out[k] = v[e_lo:e_hi]
        elif v.dim() >= 1 and v.shape[0] == n_mol:
            out[k] = v[lo:hi]
        else:
            out[k] = v
    return MolBatch(**out)


def _world(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def gather_samples(x_local: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather of per-rank samples along dim 0 (ragged across ranks: padded to the longest shard)."""
    world = _world(group)
    if world == 1:
        return x_local
    n = torch.tensor([x_local.shape[0]], dtype=torch.long, device=x_local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s) for s in sizes]
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    buf[: x_local.shape[0]] = x_local
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)
    return torch.cat([o[:s] for o, s in zip(outs, sizes)], dim=0)


def allreduce_stats(partials: torch.Tensor, group=None) -> torch.Tensor:
    """fp64 SUM all-reduce of the partial sums of stats.reweight_partials (in place, returned)."""
    if _world(group) > 1:
        dist.all_reduce(partials, op=dist.ReduceOp.SUM, group=group)
    return partials


def norm_allreduce(device, group=None):
    """Callback for MoleculeIntegrator(norm_allreduce=...): sums {sum_sq, count} over ranks."""
    def cb(ptr, _user):
        t = torch.tensor([ptr[0], ptr[1]], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        vals = t.tolist()
        ptr[0], ptr[1] = vals[0], vals[1]
    return cb if _world(group) > 1 else None
