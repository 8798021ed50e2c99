"""Time sharding of the LEC hot path across the GPUs of one box.

Time steps are independent except that dT/dt needs T at the neighbouring slots, so rank r gets a
contiguous block of steps plus a one-slot halo on each side (taken from the input, no
communication).  The only collective is one all-gather of the per-step results
(16 scalars + 19 x nlev per-level values, fp64); it is latency-bound (a few MB for a month of
hourly ERA5), so it runs over ``torch.distributed`` (NCCL over NVLink on GPUs, gloo in CPU tests).
"""

from __future__ import annotations

import numpy as np


def is_rank0() -> bool:
    """True on a single process and on rank 0 of a ``torch.distributed`` job (the rank that writes files)."""
    try:
        import torch.distributed as dist
        return not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0
    except ImportError:
        return True


def time_shards(nsteps: int, world: int):
    """Contiguous blocks of ceil(nsteps / world) steps; trailing ranks may be short or empty."""
    per = -(-nsteps // world)
    return [(min(r * per, nsteps), min((r + 1) * per, nsteps)) for r in range(world)]


def shard_slots(start: int, stop: int, nslots: int):
    """Slot window [lo, hi) a rank must hold to evaluate steps [start, stop): +-1 halo, clipped."""
    if stop <= start:
        return 0, 0
    return max(start - 1, 0), min(stop + 1, nslots)


def shard_steps(global_steps: np.ndarray, start: int, stop: int, lo: int):
    """The rank's steps with slot indices rebased to its slot window (coefficients untouched:
    they are the np.gradient coefficients of the GLOBAL time axis, so edges stay one-sided only at
    the true ends of the series)."""
    local = global_steps[start:stop].copy()
    for k in ("slot", "slot_m", "slot_p"):
        local[k] -= lo
    return local


def gather_results(local: "torch.Tensor", shards, group=None):
    """All-gather of per-step result rows ``[n_local, ...]`` into ``[nsteps, ...]`` (same on every
    rank).  Shards may be uneven: rows are padded to the largest shard for the collective."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    assert len(shards) == world
    per = max(b - a for a, b in shards)
    tail = tuple(local.shape[1:])
    buf = torch.zeros((per,) + tail, dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = torch.empty((world * per,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    out = out.view((world, per) + tail)
    return torch.cat([out[r, : b - a] for r, (a, b) in enumerate(shards)], dim=0)
