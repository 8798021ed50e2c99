"""world_size-2 gloo test of the multi-GPU host logic (time shards, halo slots, rebased steps,
all-gather of uneven shards).  The per-step "engine" here is a deterministic stand-in that
uses exactly the information a rank has (its slot window and rebased steps), so a wrong halo or
a wrong rebase changes the gathered result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lorenzcycletoolkit_b200 import engine as E
from lorenzcycletoolkit_b200 import sharding as S


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_engine(T_window, steps):
    """dT/dt-like stencil on a [slots] series + the slot id, per step."""
    out = np.empty((len(steps), 3))
    for n, st in enumerate(steps):
        out[n, 0] = st["ct_m"] * T_window[st["slot_m"]] + st["ct_0"] * T_window[st["slot"]] + st["ct_p"] * T_window[st["slot_p"]]
        out[n, 1] = T_window[st["slot"]]
        out[n, 2] = st["i1"]
    return out


def _worker(rank, world, port, nsteps, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tsec = np.cumsum(np.r_[0.0, np.random.default_rng(0).uniform(3000, 4000, nsteps - 1)])   # non-uniform axis
    T = np.sin(tsec / 7000.0) * 10 + 250
    gsteps = E.time_stencil(tsec, E.make_steps(nsteps))
    gsteps["i1"] = np.arange(nsteps)
    shards = S.time_shards(nsteps, world)
    a, b = shards[rank]
    lo, hi = S.shard_slots(a, b, nsteps)
    local = S.shard_steps(gsteps, a, b, lo)
    res = _fake_engine(T[lo:hi], local)
    full = S.gather_results(torch.from_numpy(res), shards).numpy()
    want = _fake_engine(T, gsteps)
    q.put((rank, bool(np.array_equal(full, want)), full.shape))
    dist.destroy_process_group()


@pytest.mark.parametrize("nsteps", [7, 8, 2])
def test_time_sharded_gather_equals_single_rank(nsteps):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nsteps, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, shape in got:
        assert ok and shape == (nsteps, 3), (rank, ok, shape)


def test_shard_geometry():
    assert S.time_shards(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert S.time_shards(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert S.shard_slots(0, 3, 10) == (0, 4) and S.shard_slots(3, 6, 10) == (2, 7) and S.shard_slots(9, 10, 10) == (8, 10)
    assert S.shard_slots(2, 2, 10) == (0, 0)
