"""The whole Semi-Lagrangian drop-in on an ERA5-convention file (0.25 deg, latitude north->south, 37 levels
surface-first in hPa, int16-packed; written here by tests/helpers.write_era5_like): CLI -> raw-backed dataset ->
lec_diag850_host + lec_run_host_raw -> CSVs.  Run under `ncu --metrics gpu__time_duration.sum -k regex:lec_` to list
every kernel of the product path (ingest, narrow row kernel, finalize x3, 850-hPa diagnostics)."""
import os, shutil, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
from lorenzcycletoolkit_b200 import cli

tmp = tempfile.mkdtemp()
os.makedirs(os.path.join(tmp, "inputs"))
inp = os.path.join(H.GOLDEN, "inputs")
shutil.copy(os.path.join(inp, "namelist_ERA5"), os.path.join(tmp, "inputs", "namelist"))
shutil.copy(os.path.join(inp, "track_testdata_ERA5"), os.path.join(tmp, "inputs", "track"))
nc = os.path.join(tmp, "testdata_ERA5.nc")
H.write_era5_like(nc, True, 160)
os.chdir(tmp)
t0 = time.perf_counter()
cli.main([nc, "-r", "-t"])
print(f"CLI track run on {os.path.getsize(nc) / 1e6:.0f} MB packed file: {time.perf_counter() - t0:.2f} s wall")
out = os.path.join(tmp, "LEC_Results", "testdata_ERA5_track")
print(open(os.path.join(out, "testdata_ERA5_track_trackfile")).read())
