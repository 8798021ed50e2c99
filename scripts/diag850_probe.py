"""Timing of the 850-hPa track diagnostics on a C5-shaped case: 0.1 deg planes
of the track extent (600 x 400), one 151 x 151 box per step; lec_diag850_host (planes cross PCIe inside the
call) against the numpy restatement on the host."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lorenzcycletoolkit_b200 import engine as E
from oracle import lec_oracle as O

nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 240
nlon, nlat, half = 600, 400, 75
rng = np.random.default_rng(5)
lon = (-80.0 + 0.1 * np.arange(nlon)).astype(np.float32)
lat = (-50.0 + 0.1 * np.arange(nlat)).astype(np.float32)
u, v, z = (rng.normal(size=(nsteps, nlat, nlon)).astype(np.float32) for _ in range(3))
ci = np.linspace(half + 5, nlon - half - 6, nsteps).astype(int)
cj = np.linspace(half + 5, nlat - half - 6, nsteps).astype(int)
steps = np.zeros(nsteps, dtype=E.DIAG_STEP_DTYPE)
steps["slot"] = np.arange(nsteps)
steps["i0"], steps["i1"], steps["j0"], steps["j1"] = ci - half, ci + half, cj - half, cj + half
steps["ic"], steps["jc"] = ci, cj
E.diag850_host(u[:2], v[:2], z[:2], lon, lat, steps[:2])          # context + module load
best = 1e9
for _ in range(5):
    t0 = time.perf_counter()
    vals, idx = E.diag850_host(u, v, z, lon, lat, steps)
    best = min(best, time.perf_counter() - t0)
t0 = time.perf_counter()
ovals, oidx = O.diag850(u, v, z, lon, lat, [tuple(int(s[k]) for k in ("slot", "i0", "i1", "j0", "j1")) for s in steps],
                        centres=[(int(a), int(b)) for a, b in zip(ci, cj)])
t_np = time.perf_counter() - t0
assert np.array_equal(vals, ovals) and np.array_equal(idx, oidx)
print(f"diag850: {nsteps} steps, GPU call (pageable host planes, {3 * u.nbytes / 1e6:.0f} MB H2D inside) {best * 1e3:.2f} ms "
      f"= {nsteps / best:.0f} steps/s; numpy restatement {t_np * 1e3:.1f} ms = {nsteps / t_np:.0f} steps/s; bit-identical")
