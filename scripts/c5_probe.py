"""Throughput probe for BASELINE.json configs[4]-shaped work: 0.1 deg grid, 55 levels, Semi-Lagrangian
151 x 151 boxes along a track, one box per time step, all steps in ONE engine call.  Only the track
extent +- margin is resident (what slice_domain hands over), not the 7 GB/step global field."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lorenzcycletoolkit_b200 import engine as E, synthetic as S

nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 120
half = int(sys.argv[2]) if len(sys.argv) > 2 else 75          # box = (2*half+1)^2 points; 75 = the C5 box
bands = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]   # latitude-band heights to try (0 = auto)
side = 2 * half + 1
nlon, nlat, nlev = max(600, (side + 449) // 4 * 4), max(400, side + 249), 55
ALIGN = int(os.environ.get("C5_ALIGN", "0"))      # traffic experiment: row pitch and i0 multiples of ALIGN elements
if ALIGN:
    nlon = (nlon + ALIGN - 1) // ALIGN * ALIGN
lon = (-80.0 + 0.1 * np.arange(nlon)).astype(np.float32)
lat = (-50.0 + 0.1 * np.arange(nlat)).astype(np.float32)
lev = np.linspace(1000.0, 100000.0, nlev)
grid = dict(lon=lon, lat=lat, level=lev, rlons=np.deg2rad(lon), rlats=np.deg2rad(lat), coslats=np.cos(np.deg2rad(lat)))
if os.environ.get("C5_L2GRAN"):      # traffic experiment: cudaLimitMaxL2FetchGranularity (0x05)
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    torch.zeros(1, device="cuda:0")
    r = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(os.environ["C5_L2GRAN"])))
    v = ctypes.c_size_t(0); rt.cudaDeviceGetLimit(ctypes.byref(v), 5)
    print("L2 fetch granularity set rc", r, "now", v.value, flush=True)
fields = S.synth_fields(grid, nsteps, np.float32, "cuda:0")
f64 = lambda a: np.asarray(a, dtype=np.float64)
steps = E.time_stencil(3600.0 * np.arange(nsteps), E.make_steps(nsteps))
ci = np.linspace(half + 5, nlon - half - 6, nsteps).astype(int)
cj = np.linspace(half + 5, nlat - half - 6, nsteps).astype(int)
if ALIGN:
    ci = (ci - half) // ALIGN * ALIGN + half
steps["i0"], steps["i1"], steps["j0"], steps["j1"] = ci - half, ci + half, cj - half, cj + half
if os.environ.get("C5_NONEIGH") == "1":      # traffic experiment: no time neighbours (T(t+-1) alias T(t))
    steps["slot_m"] = steps["slot_p"] = steps["slot"]
B = 5 * nlev * side * side * 4
for band in bands:
    eng = E.LecEngine(f64(lon), f64(lat), f64(grid["rlons"]), f64(grid["rlats"]), f64(grid["coslats"]), lev, np.float32,
                      max_steps=int(os.environ.get("C5_MAXSTEPS", nsteps)), max_box_rows=side, band_rows=band)
    best = 1e9
    for it in range(4):
        terms, levels, flags = eng.run_torch(fields, steps)
        torch.cuda.synchronize()
        a, b, c = eng.last_timing()
        best = min(best, a)
    print(f"C5-shape ({side}x{side}) band_rows={band}: {nsteps} steps, rows {best:.3f} ms fin {b:.3f} ms call {c:.3f} ms -> "
          f"{nsteps / c * 1e3:.0f} steps/s, {B * nsteps / best / 1e6:.0f} GB/s algorithmic in the row kernel "
          f"({B * nsteps / best / 1e6 / 6535.7:.3f})  flags {int(flags.max().item())} Az[0] {float(terms[0, 0]):.6e}", flush=True)
    eng.close()
