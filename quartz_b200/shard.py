"""Multi-GPU plumbing for the one place the path shards: independent voices.

One process per GPU; rank r owns a contiguous voice range and its own context/bank; there is NO data-path
collective — outputs are disjoint row ranges that the host gathers.  torch.distributed (NCCL on the GPU box, gloo in
the CPU tests) is used only for barriers, the max-over-ranks time and, optionally, gathering outputs to rank 0."""
import numpy as np


def voice_range(n_voices, world, rank, group=1):
    """Contiguous, group-aligned partition of [0, n_voices): rank r gets [lo, hi)."""
    n_groups = n_voices // group
    base, rem = divmod(n_groups, world)
    lo_g = rank * base + min(rank, rem)
    hi_g = lo_g + base + (1 if rank < rem else 0)
    return lo_g * group, hi_g * group


def shard_workload(make, world, rank, strong=False, **kw):
    """weak scaling: every rank renders a full-size bank over its own voice ids [rank*V, (rank+1)*V);
    strong scaling: the global bank of V voices is split across ranks."""
    full = make(**kw)
    if isinstance(full, list):          # multi-bank workload (one bank per voice archetype)
        total = sum(w.V for w in full)
        return make(v0=rank * total, **kw) if not strong else make(V=total // world, v0=rank * (total // world), **kw)
    if full.V == 1:
        return full
    if not strong:
        return make(v0=rank * full.V, **kw)
    lo, hi = voice_range(full.V, world, rank, full.group)
    kw2 = dict(kw)
    kw2.update(V=hi - lo, v0=lo)
    return make(**kw2)


def max_over_ranks(value, dist=None, device="cpu"):
    """device-side max of a scalar over ranks (time-like quantities are reported as the slowest rank's)"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_rows(local_rows, dist=None):
    """Gather voice-major output rows [rows_r, T] from every rank to rank 0 (host side, no GPU collective)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local_rows
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, np.ascontiguousarray(local_rows))
    return np.concatenate(parts, axis=0)
