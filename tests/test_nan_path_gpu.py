"""The reference's NaN path (``_handle_nans``: energy_contents.py:190-208, conversion_terms.py:247,
boundary_terms.py:420-438, generation_and_dissipation_terms.py:190) THROUGH THE DROP-IN: an NCEP-style file
with missing values (``_FillValue``) on pressure levels, ``lec_fixed`` and ``lec_moving`` against the oracle,
which applies ``handle_nans`` inside every term exactly where the reference does.

Cases planted (``_FillValue`` entries in copies of the bundled NetCDF-3 files):
  * an interior level partly / wholly missing      -> interpolated along p
  * an edge level (1000 hPa) missing at ONE time    -> cannot be interpolated: the fixed framework drops the level
    for all times (per term), the moving framework only at that step."""
import argparse
import logging
import os

import numpy as np
import pandas as pd
import pytest

from oracle import lec_oracle as O
import helpers as H
from lorenzcycletoolkit_b200.utils import preprocessing as PP

pytestmark = pytest.mark.gpu
INP = os.path.join(H.GOLDEN, "inputs")
SAM = os.path.join(H.GOLDEN, "samples")
FILL = np.float32(1e20)


def _plant_catarina(name, data):
    # raw order: levels 10 ... 1000 hPa (index 11 = 500, 16 = 1000), 36 times, 7 x 8 points
    if name == "TMP_2_ISBL":
        data[:, 11, 2:5, 1:6] = FILL             # interior level, part of the box, every time
    if name == "V_VEL_2_ISBL":
        data[10, 16, :, :] = FILL                # the 1000-hPa level at one time
    if name == "U_GRD_2_ISBL":
        data[20, 3, :, :] = FILL                 # a whole interior level (50 hPa) at one time
    return data


def _plant_testdata(name, data):
    # raw order: levels 600 700 850 925 1000 hPa, 5 times, 33 x 41 points (lat 0 ... -80, lon -100 ... 0)
    if name == "TMP_2_ISBL":
        data[:, 2, 10:20, 12:30] = FILL          # 850 hPa over the track box, every time
    if name == "V_VEL_2_ISBL":
        data[3, 4, :, :] = FILL                  # the 1000-hPa level at one time
    return data


@pytest.fixture(scope="module")
def nan_files(tmp_path_factory):
    d = tmp_path_factory.mktemp("nan")
    cat, tst = str(d / "Catarina_NaN.nc"), str(d / "testdata_NaN.nc")
    H.write_netcdf3_copy(os.path.join(SAM, "Catarina_NCEP-R2.nc"), cat, _plant_catarina)
    H.write_netcdf3_copy(os.path.join(SAM, "testdata_NCEP-R2.nc"), tst, _plant_testdata)
    return cat, tst


def _oracle_prepared(path, track=None):
    raw = O.read_netcdf3(path)
    nl = O.read_namelist(os.path.join(INP, "namelist_NCEP-R2"))
    tr = O.read_track(os.path.join(INP, track)) if track else None
    return O.process_data(raw, nl, tr), tr


def _check_frame(df, odf, tol):
    assert list(df.columns) == list(odf.columns)
    for c in odf.columns:
        a, b = df[c].values, odf[c].values
        assert np.array_equal(np.isnan(a), np.isnan(b)), c
        ok = ~np.isnan(b)
        assert ok.any(), c                                      # the NaN path must leave numbers, not NaN columns
        assert H.series_err(a[ok], b[ok]) <= tol, (c, H.series_err(a[ok], b[ok]))


@pytest.mark.parametrize("lazy", ["1", "0"])
def test_lec_fixed_with_missing_levels(nan_files, tmp_path, monkeypatch, lazy):
    monkeypatch.setenv("LEC_DEVICE_INGEST", lazy)              # raw records decoded on the GPU / host-prepared
    from lorenzcycletoolkit_b200.frameworks import lec_fixed
    nan_file = nan_files[0]
    nl = PP.read_namelist(os.path.join(INP, "namelist_NCEP-R2"))
    boxf = tmp_path / "box"
    boxf.write_text("min_lon;-55\nmax_lon;-36\nmin_lat;-35\nmax_lat;-20\n")
    args = argparse.Namespace(infile=nan_file, fixed=True, track=False, choose=False, residuals=True,
                              box_limits=str(boxf), outname=None, plots=False, cdsapi=False, mpas=False)
    data = PP.prepare_data(args, os.path.join(INP, "namelist_NCEP-R2"), box_limits_file=args.box_limits)
    lv_dir = tmp_path / "lv"
    os.makedirs(lv_dir)
    df = lec_fixed(data, nl, str(tmp_path), str(lv_dir), logging.getLogger("t"), args)

    P, _ = _oracle_prepared(nan_file)
    box = (-55, -36, -35, -20)
    P = O.slice_domain_fixed(P, *box)
    F = P.fields
    assert all(np.isnan(F[n]).any() for n in ("Air Temperature", "Omega Velocity", "Eastward Wind Component"))
    odf, olv, extra = O.lec_fixed(P, *box, mode="fp64")
    _check_frame(df, odf, 1e-5)
    # per-level files: the cleaned integrand (interpolated, levels dropped) under the full header
    nlev = len(P.level)
    for name in ("Az", "Ke", "Ce", "Ck", "Gz", "Ca_2"):
        want = np.asarray(olv[name], dtype=np.float64)
        got = pd.read_csv(lv_dir / f"{name}_lv_ISBL3.csv", header=None, skiprows=1, index_col=0).values
        assert got.shape == want.shape, (name, got.shape, want.shape)
        assert np.array_equal(np.isnan(got), np.isnan(want)), name
        ok = ~np.isnan(want)
        assert H.series_err(got[ok], want[ok]) <= 1e-5, name
    # omega is missing on an edge level at one time: every omega term loses that level for ALL times,
    # the T-only / wind-only terms keep every level (interior gaps are interpolated)
    assert np.asarray(olv["Ce"]).shape[1] == nlev - 1
    assert np.asarray(olv["Az"]).shape[1] == nlev and np.asarray(olv["Ke"]).shape[1] == nlev


def test_lec_moving_with_missing_levels(nan_files, tmp_path):
    from lorenzcycletoolkit_b200.frameworks import lec_moving
    nan_file = nan_files[1]
    nl = PP.read_namelist(os.path.join(INP, "namelist_NCEP-R2"))
    trk = os.path.join(INP, "track_testdata_NCEP-R2")
    args = argparse.Namespace(infile=nan_file, fixed=False, track=True, choose=False, residuals=True,
                              trackfile=trk, cdsapi=False, mpas=False, zeta=False, plots=False)
    data = PP.prepare_data(args, os.path.join(INP, "namelist_NCEP-R2"))
    lv_dir = tmp_path / "lv"
    os.makedirs(lv_dir)
    df = lec_moving(data, nl, None, str(tmp_path), str(tmp_path), str(lv_dir), logging.getLogger("t"), args)
    P, tr = _oracle_prepared(nan_file, "track_testdata_NCEP-R2")
    P = O.slice_domain_track(P, tr)
    odf, olv, _ = O.lec_moving(P, tr, mode="fp64")
    _check_frame(df, odf, 1e-5)
    # per step: at the step with the missing 1000-hPa omega, the omega terms integrate over one level less
    assert isinstance(olv["Ce"], list) and olv["Ce"][3].shape[-1] == olv["Ce"][0].shape[-1] - 1
