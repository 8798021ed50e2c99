"""CPU timing legs of ``bench.py`` (the ``cpu_baseline`` object and ``--impl reference``).

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The reference (pure Python on xarray/MetPy/pint)
cannot be imported in this image, so the CPU arm times the numpy
oracle -- a lower bound on the real reference's cost (no pint/xarray dispatch, no per-level
CSV writes).  One unit of work = all LEC terms of ONE time step on a latitude band of the
C4 workload (synthetic ERA5 0.25 deg x 37 levels), evaluated the way the moving framework
evaluates one step (``BoxData`` on [level][lat][lon] with dT/dt from the t-1/t+1 slots).
"""
from __future__ import annotations

import time

import numpy as np

from . import lec_oracle as O

NAMES = ["Air Temperature", "Eastward Wind Component", "Northward Wind Component",
         "Omega Velocity", "Geopotential"]
UNITS = ["K", "m/s", "m/s", "Pa/s", "m**2/s**2"]


def make_prepared(grid, fields, j_lo, j_hi):
    """Oracle dataset over rows [j_lo, j_hi) of ``grid`` from five [3][L][rows][nlon] arrays."""
    P = O.Prepared()
    P.fields = dict(zip(NAMES, fields))
    P.units = dict(zip(NAMES, UNITS))
    P.time = np.datetime64("2020-01-01T00") + np.arange(3) * np.timedelta64(1, "h")
    P.level = grid["level"]
    P.lat, P.rlats, P.coslats = grid["lat"][j_lo:j_hi], grid["rlats"][j_lo:j_hi], grid["coslats"][j_lo:j_hi]
    P.lon, P.rlons = grid["lon"], grid["rlons"]
    return P


def one_step(P):
    """All terms of the middle slot of a 3-slot dataset; returns (terms dict, seconds)."""
    t0 = time.perf_counter()
    tsec = 3600.0 * np.arange(3)
    dTdt = O.differentiate(P.fields["Air Temperature"], tsec, 0)
    b = O.BoxState(P, float(P.lon[0]), float(P.lon[-1]), float(P.lat[0]), float(P.lat[-1]),
                   fixed=False, dTdt=dTdt[1], tsel=1)
    lv, terms = {}, {}
    terms.update(O.energy_contents(b, lv))
    terms.update(O.conversion_terms(b, lv))
    terms.update(O.boundary_terms(b))
    terms.update(O.generation_terms(b, lv))
    return {k: float(v) for k, v in terms.items()}, time.perf_counter() - t0


def _worker(args):
    """Generate the worker's own 3-slot band (untimed), then time ``reps`` evaluations."""
    nlon, nlat, rows, t, reps, seed = args
    import torch
    torch.set_num_threads(1)
    from lorenzcycletoolkit_b200 import synthetic as S
    grid = S.era5_grid(nlon=nlon, nlat=nlat)
    j_lo = max(1, (nlat - rows) // 2)
    sub = dict(grid)
    for k in ("lat", "rlats", "coslats"):
        sub[k] = grid[k][j_lo:j_lo + rows]
    fields = [x.numpy() for x in S.synth_fields(sub, 3, np.float32, "cpu", seed=seed, t0=t)]
    P = make_prepared(sub, fields, 0, rows)
    times = []
    for _ in range(reps):
        _, dt = one_step(P)
        times.append(dt)
    return times


def _loop(idx, nlon, nlat, rows, nsteps, seed, bar):
    import torch
    torch.set_num_threads(1)
    from lorenzcycletoolkit_b200 import synthetic as S
    grid = S.era5_grid(nlon=nlon, nlat=nlat)
    j_lo = max(1, (nlat - rows) // 2)
    sub = dict(grid)
    for k in ("lat", "rlats", "coslats"):
        sub[k] = grid[k][j_lo:j_lo + rows]
    fields = [x.numpy() for x in S.synth_fields(sub, 3, np.float32, "cpu", seed=seed, t0=3 * idx)]
    P = make_prepared(sub, fields, 0, rows)
    bar.wait()                       # data ready
    for _ in range(nsteps):
        bar.wait()                   # start of a timed step
        one_step(P)
        bar.wait()                   # end of the step


def run_parallel(nproc, nlon, nlat, rows, nsteps, seed=1234):
    """``nproc`` processes each evaluate one band-step per bench step, in lockstep.
    Returns the wall time of every step (seconds)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    bar = ctx.Barrier(nproc + 1)
    procs = [ctx.Process(target=_loop, args=(i, nlon, nlat, rows, nsteps, seed, bar), daemon=True)
             for i in range(nproc)]
    for p in procs:
        p.start()
    bar.wait(timeout=1800)
    out = []
    for _ in range(nsteps):
        bar.wait(timeout=1800)
        t0 = time.perf_counter()
        bar.wait(timeout=1800)
        out.append(time.perf_counter() - t0)
    for p in procs:
        p.join(timeout=60)
    return out
