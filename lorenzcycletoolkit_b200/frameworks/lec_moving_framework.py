"""Semi-Lagrangian moving-box driver (reference: ``src/frameworks/lec_moving_framework.py:35-841``).

The reference loops over time steps in Python and builds one ``BoxData`` (about 150 xarray
expressions) per step.  Here the per-step boxes are collected first (``get_limits``) and the
whole track is evaluated by ONE engine call (``BoxBatch``); the per-step term objects, the
per-level CSV rows, the results CSV and the ``*_trackfile`` are then written exactly as before."""

from __future__ import annotations

import logging
import os
from pathlib import Path

import numpy as np
import pandas as pd

from ..sharding import is_rank0 as _is_rank0
from ..analysis import BoundaryTerms, ConversionTerms, EnergyContents, GenerationDissipationTerms
from ..utils.box_data import BoxBatch, unit_factor, G
from ..utils.calc_budget_and_residual import calc_budget_diff, calc_residuals
from ..utils.preprocessing import read_track, _label_slice
from .lec_fixed_framework import create_level_files

RE = 6371008.7714


def create_terms_dict(args):
    """lec_moving_framework.py:35-55."""
    terms = ["Az", "Ae", "Kz", "Ke", "Cz", "Ca", "Ck", "Ce", "BAz", "BAe", "BKz", "BKe", "BΦZ", "BΦE"]
    terms += ["Gz", "Ge"] if args.residuals else ["Gz", "Ge", "Dz", "De"]
    return {t: [] for t in terms}


def handle_track_file(data, times, LonIndexer, LatIndexer, TimeIndexer, args, app_logger):
    """Read the track and check that it lies inside the data (lec_moving_framework.py:58-160)."""
    trackfile = args.trackfile
    try:
        track = read_track(trackfile)
    except FileNotFoundError:
        app_logger.error(f"❌ Track file {trackfile} not found.")
        raise
    lon, lat = np.asarray(data.lon), np.asarray(data.lat)
    if track["Lon"].min() < lon.min() or track["Lon"].max() > lon.max():
        raise ValueError("Track longitude limits are outside the data domain")
    if track["Lat"].min() < lat.min() or track["Lat"].max() > lat.max():
        raise ValueError("Track latitude limits are outside the data domain")
    t = pd.to_datetime(times)
    if track.index.min() > t.max() or track.index.max() < t.min():
        raise ValueError("Track time limits do not overlap the data time limits")
    return track


def get_limits(args, t, data850, track=None):
    """Box of one time step (lec_moving_framework.py:199-266): the track row nearest in time,
    centre +- width/2 (default 15 x 15 degrees)."""
    if not getattr(args, "track", False):
        raise NotImplementedError("the interactive --choose framework needs a display (out of scope)")
    closest = int(np.argmin(np.abs(track.index - t)))
    row = track.iloc[closest]
    central_lat, central_lon = row["Lat"], row["Lon"]
    width, length = row.get("width", 15), row.get("length", 15)
    return {"datestr": pd.to_datetime(t).strftime("%Y-%m-%d-%H%M"),
            "central_lat": central_lat, "central_lon": central_lon, "length": length, "width": width,
            "min_lon": central_lon - width / 2, "max_lon": central_lon + width / 2,
            "min_lat": central_lat - length / 2, "max_lat": central_lat + length / 2}


def diagnostics_850(data, variable_list_df, it=None):
    """850-hPa wind speed, relative vorticity and geopotential height (lec_moving_framework.py:650-663)
    for ALL time steps in one vectorised numpy pass (the reference recomputes them per step inside its
    time loop); ``it`` selects one step.  Vorticity in spherical form
    zeta = dv/dx - du/dy + (u/a) tan(lat) -- MetPy's WGS-84 geodesic spacing is not available here
    (SURVEY.md B.7), so these trackfile diagnostics are unpinned."""
    k = int(np.argmin(np.abs(np.asarray(data.level, dtype=np.float64) - 85000.0)))
    if float(data.level[k]) != 85000.0:
        raise KeyError("85000 Pa level not found (the moving framework needs 850 hPa)")
    sel = slice(None) if it is None else slice(it, it + 1)

    def field(row):
        var = variable_list_df.loc[row]["Variable"]
        return np.asarray(data[var][sel, k], dtype=np.float64) * unit_factor(variable_list_df.loc[row]["Units"], row)
    u, v = field("Eastward Wind Component"), field("Northward Wind Component")
    if "Geopotential" in variable_list_df.index:
        hgt = field("Geopotential") / G
    else:
        hgt = field("Geopotential Height")
    rlat = np.deg2rad(np.asarray(data.lat, dtype=np.float64))
    rlon = np.deg2rad(np.asarray(data.lon, dtype=np.float64))
    dvdx = np.gradient(v, rlon, axis=2) / (RE * np.cos(rlat)[None, :, None])
    dudy = np.gradient(u, rlat, axis=1) / RE
    zeta = dvdx - dudy + u * np.tan(rlat)[None, :, None] / RE
    out = {"izeta_850": zeta, "ihgt_850": hgt, "iwspd_850": np.sqrt(u * u + v * v), "iu_850": u, "iv_850": v}
    if it is not None:
        out = {k2: a[0] for k2, a in out.items()}
    out["lat"], out["lon"] = np.asarray(data.lat), np.asarray(data.lon)
    return out


def get_position(track, limits, d850, args):
    """Extrema inside the (unsnapped, label-sliced) box (lec_moving_framework.py:269-417)."""
    lat, lon = d850["lat"], d850["lon"]
    js, is_ = _label_slice(lat, limits["min_lat"], limits["max_lat"]), _label_slice(lon, limits["min_lon"], limits["max_lon"])
    zeta, hgt, wspd = d850["izeta_850"][js, is_], d850["ihgt_850"][js, is_], d850["iwspd_850"][js, is_]
    row = track.loc[pd.to_datetime(limits["datestr"], format="%Y-%m-%d-%H%M")] if track is not None and \
        pd.to_datetime(limits["datestr"], format="%Y-%m-%d-%H%M") in track.index else None
    south = limits["min_lat"] < 0

    def from_track(col):
        return row is not None and col in track.columns and not pd.isna(row[col])

    min_max_zeta = float(row["min_max_zeta_850"]) if from_track("min_max_zeta_850") else \
        float(np.nanmin(zeta) if south else np.nanmax(zeta))
    min_hgt = float(row["min_hgt_850"]) if from_track("min_hgt_850") else float(hgt.min())
    max_wind = float(row["max_wind_850"]) if from_track("max_wind_850") else float(wspd.max())

    def where(a, which):
        idx = np.unravel_index(a.argmin() if which == "min" else a.argmax(), a.shape)
        return lat[js][idx[0]], lon[is_][idx[1]]

    zlat, zlon = where(zeta, "min" if lat[js].min() < 0 else "max")
    hlat, hlon = where(hgt, "min")
    wlat, wlon = where(wspd, "max")
    return {"min_max_zeta_850_lat": zlat, "min_max_zeta_850_lon": zlon, "min_max_zeta_850": min_max_zeta,
            "min_hgt_850_lat": hlat, "min_hgt_850_lon": hlon, "min_hgt_850": min_hgt,
            "max_wind_850_lat": wlat, "max_wind_850_lon": wlon, "max_wind_850": max_wind}


def compute_and_store_terms(box_obj, terms_dict, app_logger):
    """One step's 16 scalars appended to the lists (lec_moving_framework.py:430-495)."""
    try:
        ec = EnergyContents(box_obj, "moving", app_logger)
        for n, f in (("Az", ec.calc_az), ("Ae", ec.calc_ae), ("Kz", ec.calc_kz), ("Ke", ec.calc_ke)):
            terms_dict[n].append(f())
        ct = ConversionTerms(box_obj, "moving", app_logger)
        for n, f in (("Cz", ct.calc_cz), ("Ca", ct.calc_ca), ("Ck", ct.calc_ck), ("Ce", ct.calc_ce)):
            terms_dict[n].append(f())
        bt = BoundaryTerms(box_obj, "moving", app_logger)
        for n, f in (("BAz", bt.calc_baz), ("BAe", bt.calc_bae), ("BKz", bt.calc_bkz), ("BKe", bt.calc_bke),
                     ("BΦZ", bt.calc_boz), ("BΦE", bt.calc_boe)):
            terms_dict[n].append(f())
        gd = GenerationDissipationTerms(box_obj, "moving", app_logger)
        terms_dict["Gz"].append(gd.calc_gz())
        terms_dict["Ge"].append(gd.calc_ge())
        if "Dz" in terms_dict:
            terms_dict["Dz"].append(gd.calc_dz())
            terms_dict["De"].append(gd.calc_de())
    except Exception:
        app_logger.exception("❌ An exception occurred while computing the LEC terms of a step")
        raise
    return terms_dict


def finalize_results(times, terms_dict, args, results_subdirectory, out_track, app_logger, write=True):
    """Results CSV + trackfile (lec_moving_framework.py:498-543)."""
    df = pd.DataFrame(terms_dict, index=pd.to_datetime(times), dtype=float)
    app_logger.info("📈 Estimating budget terms...")
    df = calc_budget_diff(df, np.asarray(times), app_logger)
    if args.residuals:
        app_logger.info("🧮 Computing residuals...")
        df = calc_residuals(df, app_logger)
    method = "track" if args.track else "choose"
    infile_name = os.path.basename(args.infile).split(".nc")[0]
    results_file = os.path.join(results_subdirectory, f"{infile_name}_{method}_results.csv")
    if not write:
        return results_file, df
    df.to_csv(results_file)
    app_logger.info(f"💾 Results saved to {results_file}")
    out_track = out_track.rename(columns={"datestr": "time", "central_lat": "Lat", "central_lon": "Lon"})
    output_trackfile = os.path.join(results_subdirectory, f"{infile_name}_{method}_trackfile")
    out_track.to_csv(output_trackfile, index=False, sep=";")
    app_logger.info(f"📍 System track saved to {output_trackfile}")
    return results_file, df


def lec_moving(data, variable_list_df, dTdt, results_subdirectory, figures_directory,
               results_subdirectory_vertical_levels, app_logger, args, engine_options=None):
    """``dTdt`` is accepted for signature compatibility and ignored: the row kernel evaluates the
    same centred time difference of T over the track-selected times (lorenzcycletoolkit.py:184-186)
    from the adjacent time slots of the resident field."""
    app_logger = app_logger or logging.getLogger("lorenzcycletoolkit")
    app_logger.info("📊 Computing energetics using moving framework")
    LonIndexer = variable_list_df.loc["Longitude"]["Variable"]
    LatIndexer = variable_list_df.loc["Latitude"]["Variable"]
    TimeName = variable_list_df.loc["Time"]["Variable"]
    VerticalCoordIndexer = variable_list_df.loc["Vertical Level"]["Variable"]
    write = _is_rank0()          # under torchrun every rank computes its time shard; rank 0 writes the files
    if write:
        create_level_files(results_subdirectory_vertical_levels, TimeName, VerticalCoordIndexer, data.level)
    else:
        results_subdirectory_vertical_levels = None
    times = pd.to_datetime(np.asarray(data.time))
    if len(times) == 0:
        raise ValueError("Mismatch between trackfile and data! Check that the track times exist in the file.")
    track = handle_track_file(data, times, LonIndexer, LatIndexer, TimeName, args, app_logger)

    limits_list, rows = [], []
    d850_all = diagnostics_850(data, variable_list_df)
    for it, t in enumerate(times):
        d850 = {k2: (a[it] if k2 not in ("lat", "lon") else a) for k2, a in d850_all.items()}
        limits = get_limits(args, t, d850, track)
        position = get_position(track, limits, d850, args)
        app_logger.info(f"🗺️ {t}: box center=({limits['central_lat']:.2f}, {limits['central_lon']:.2f}), "
                        f"size={limits['length']}°x{limits['width']}°")
        limits_list.append(limits)
        rows.append({**limits, **position})
    out_track = pd.DataFrame(rows)

    try:
        batch = BoxBatch(data, variable_list_df, limits_list, args, results_subdirectory,
                         results_subdirectory_vertical_levels, engine_options=engine_options)
    except Exception:
        app_logger.exception("❌ An exception occurred while creating BoxData object")
        raise
    _, _, call_ms = batch.timing_ms
    app_logger.info(f"🚀 B200 engine: {len(times)} steps in {call_ms:.2f} ms")
    terms_dict = create_terms_dict(args)
    for it in range(len(times)):
        terms_dict = compute_and_store_terms(batch.step(it), terms_dict, app_logger)
    results_file, df = finalize_results(times, terms_dict, args, results_subdirectory, out_track, app_logger,
                                        write=write)
    return df
