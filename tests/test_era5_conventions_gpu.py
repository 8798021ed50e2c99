"""BASELINE.json configs[1] (ERA5, Semi-Lagrangian track): the bundled samples/testdata_ERA5.nc is not
in the reference checkout (.MISSING_LARGE_BLOBS), so a synthetic file with ERA5's on-disk conventions is
written here -- 0.25 deg grid, latitude stored north-to-south, 37 levels in hPa stored surface-first,
plain float32 or packed int16 (scale_factor / add_offset, which xarray decodes to float64) -- and run
through the CLI with the reference's own inputs/namelist_ERA5 and inputs/track_testdata_ERA5."""
import os
import shutil

import numpy as np
import pandas as pd
import pytest

from oracle import lec_oracle as O
import helpers as H
from lorenzcycletoolkit_b200 import cli

pytestmark = pytest.mark.gpu
INP = os.path.join(H.GOLDEN, "inputs")


_write_era5_like = H.write_era5_like


@pytest.mark.parametrize("packed,nlon", [(False, 160), (False, 161), (True, 160)])
def test_era5_track_cli_matches_oracle(tmp_path, monkeypatch, packed, nlon):
    os.makedirs(tmp_path / "inputs")
    shutil.copy(os.path.join(INP, "namelist_ERA5"), tmp_path / "inputs" / "namelist")
    shutil.copy(os.path.join(INP, "track_testdata_ERA5"), tmp_path / "inputs" / "track")
    nc = str(tmp_path / "testdata_ERA5.nc")
    _write_era5_like(nc, packed, nlon)
    monkeypatch.chdir(tmp_path)
    cli.main([nc, "-r", "-t"])
    df = pd.read_csv(tmp_path / "LEC_Results" / "testdata_ERA5_track" / "testdata_ERA5_track_results.csv", index_col=0)

    raw = O.read_netcdf3(nc)
    nl = O.read_namelist(os.path.join(INP, "namelist_ERA5"))
    tr = O.read_track(os.path.join(INP, "track_testdata_ERA5"))
    P = O.slice_domain_track(O.process_data(raw, nl, tr), tr)
    assert P.fields["Air Temperature"].dtype == (np.float64 if packed else np.float32)
    assert len(P.level) == 32 and P.level[0] == 1000.0 and np.all(np.diff(P.lat) > 0)   # <10 hPa dropped, lat sorted
    odf, _, boxes = O.lec_moving(P, tr, mode="fp64")
    assert all(b[1][1] - b[1][0] == 60 and b[1][3] - b[1][2] == 60 for b in boxes)       # 15 x 15 deg = 61 x 61 points
    assert list(df.columns) == list(odf.columns) and len(df) == 5
    tol = 1e-9 if packed else 1e-5
    for c in odf.columns:
        assert H.series_err(df[c].values, odf[c].values) <= tol, (c, H.series_err(df[c].values, odf[c].values))
    lv = pd.read_csv(tmp_path / "LEC_Results" / "testdata_ERA5_track" / "results_vertical_levels" / "Ck_level.csv", index_col=0)
    assert lv.shape == (5, 32) and float(lv.columns[-1]) == 100000.0
