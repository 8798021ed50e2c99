"""First-light performance probe of the row-moment kernel on C4-shaped data."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lorenzcycletoolkit_b200 import engine as E, synthetic as S

nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
band = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dt = np.float64 if (len(sys.argv) > 3 and sys.argv[3] == 'f64') else np.float32
math = int(sys.argv[4]) if len(sys.argv) > 4 else 0
g = S.era5_grid()
f64 = lambda a: np.asarray(a, dtype=np.float64)
t0 = time.time()
fields = S.synth_fields(g, nsteps + 2, dt, "cuda:0")
torch.cuda.synchronize()
print("synth s", time.time() - t0, flush=True)
eng = E.LecEngine(f64(g["lon"]), f64(g["lat"]), f64(g["rlons"]), f64(g["rlats"]), f64(g["coslats"]), g["level"],
                  dt, max_steps=nsteps, max_box_rows=719, band_rows=band, math=math)
tsec = 3600.0 * np.arange(nsteps + 2)
steps = E.time_stencil(tsec, E.make_steps(nsteps + 2))[1:-1]
steps["i0"], steps["i1"], steps["j0"], steps["j1"] = 0, 1439, 1, 719
B = 5 * 37 * 719 * 1440 * np.dtype(dt).itemsize
for it in range(4):
    terms, levels, flags = eng.run_torch(fields, steps)
    torch.cuda.synchronize()
    a, b, c = eng.last_timing()
    print(f"rows {a:.3f} ms  fin {b:.3f} ms  call {c:.3f} ms  -> {nsteps / c * 1e3:.1f} steps/s, "
          f"rows-kernel {B * nsteps / a / 1e6:.0f} GB/s", flush=True)
print('flags', int(flags.max().item()))
