"""Multi-GPU plumbing: the receive chain shards by channel and nothing else.

Every channel owns its FIR history, biquad state, mode and tables (no cross-channel term anywhere in
Minimal-SDR.ino:518-775 or filter_biquad.cpp:33-82), so rank r of W simply owns a contiguous channel range and runs its own
chain object; there is NO collective on the data path.  torch.distributed is only used to agree on timing (max over ranks)
and, if asked, to gather results or checksums.
"""
from dataclasses import dataclass


@dataclass(frozen=True)
class Shard:
    rank: int
    world: int
    ch0: int      # first global channel owned by this rank
    n: int        # number of channels owned


def plan(total_channels, world, rank=None, granule=32):
    """Contiguous ranges, sizes differing by at most one `granule` (32 = the kernel's channel group), every channel owned
    exactly once.  Returns the list of all shards, or this rank's shard when `rank` is given."""
    if total_channels < 0 or world < 1:
        raise ValueError("bad arguments")
    groups = (total_channels + granule - 1) // granule
    base, extra = divmod(groups, world)
    shards, g0 = [], 0
    for r in range(world):
        g = base + (1 if r < extra else 0)
        ch0 = min(g0 * granule, total_channels)
        ch1 = min((g0 + g) * granule, total_channels)
        shards.append(Shard(r, world, ch0, ch1 - ch0))
        g0 += g
    return shards if rank is None else shards[rank]


def weak_scaling_shard(channels_per_rank, world, rank):
    """Benchmark layout: per-GPU work fixed, rank r owns global channels [r*C, (r+1)*C)."""
    return Shard(rank, world, rank * channels_per_rank, channels_per_rank)


def max_over_ranks(value, device=None):
    """Timing rule: a multi-GPU number is the max over ranks.  Works on any backend (gloo on CPU, nccl on GPU)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_rows(local_rows, shard, total_channels, device=None):
    """Optional, untimed: assemble [total_channels, n] on every rank from per-rank [shard.n, n] int16 tensors."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_rows
    shards = plan(total_channels, shard.world)
    nmax = max(s.n for s in shards)
    pad = torch.zeros((nmax, local_rows.shape[1]), dtype=local_rows.dtype, device=local_rows.device)
    pad[:shard.n] = local_rows
    raw = pad.view(torch.uint8)  # gloo has no int16 collectives; bytes travel on every backend
    parts = [torch.empty_like(raw) for _ in shards]
    dist.all_gather(parts, raw)
    return torch.cat([p.view(local_rows.dtype)[:s.n] for p, s in zip(parts, shards)], dim=0)
