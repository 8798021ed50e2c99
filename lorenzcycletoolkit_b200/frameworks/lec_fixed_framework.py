"""Eulerian fixed-box driver (reference: ``src/frameworks/lec_fixed_framework.py:30-348``).

Same signature, same files on disk (21 per-level CSVs + ``<stem>_fixed_results.csv``) and the
same column order; the arithmetic of all time steps is ONE batched engine call made by
``BoxData`` instead of ~150 full-array xarray expressions."""

from __future__ import annotations

import logging
import os
from pathlib import Path

import numpy as np
import pandas as pd

from ..sharding import is_rank0 as _is_rank0
from ..analysis import BoundaryTerms, ConversionTerms, EnergyContents, GenerationDissipationTerms
from ..utils.box_data import BoxData
from ..utils.calc_budget_and_residual import calc_budget_diff, calc_residuals
from ..utils.preprocessing import read_box_limits

LEVEL_FILE_TERMS = ["Az", "Ae", "Kz", "Ke", "Ge", "Gz", "Cz", "Cz_1", "Cz_2", "Ca", "Ca_1", "Ca_2",
                    "Ce", "Ce_1", "Ce_2", "Ck", "Ck_1", "Ck_2", "Ck_3", "Ck_4", "Ck_5"]


def create_level_files(directory, TimeName, VerticalCoordIndexer, pressure):
    """Header-only per-level CSVs (lec_fixed_framework.py:172-197)."""
    columns = [TimeName] + [float(i) for i in pressure]
    for term in LEVEL_FILE_TERMS:
        pd.DataFrame(columns=columns).to_csv(Path(directory, f"{term}_{VerticalCoordIndexer}.csv"), index=None)


def lec_fixed(data, variable_list_df, results_subdirectory, results_subdirectory_vertical_levels,
              app_logger, args, engine_options=None):
    app_logger = app_logger or logging.getLogger("lorenzcycletoolkit")
    app_logger.info("📊 Computing energetics using fixed framework...")
    try:
        lim = read_box_limits(args.box_limits)
    except FileNotFoundError:
        app_logger.error("❌ Box limits file not found!")
        raise FileNotFoundError(f"Box limits file not found: {os.path.abspath(args.box_limits)}. "
                                "Create one or use --box_limits to specify path.")
    min_lon, max_lon, min_lat, max_lat = lim["min_lon"], lim["max_lon"], lim["min_lat"], lim["max_lat"]
    data = data.compute()
    TimeName = variable_list_df.loc["Time"]["Variable"]
    VerticalCoordIndexer = variable_list_df.loc["Vertical Level"]["Variable"]
    app_logger.info(f"🗺️ Bounding box: lon=[{min_lon}, {max_lon}], lat=[{min_lat}, {max_lat}]")
    write = _is_rank0()          # under torchrun every rank computes its time shard; rank 0 writes the files
    if write:
        create_level_files(results_subdirectory_vertical_levels, TimeName, VerticalCoordIndexer, data.level)
    else:
        results_subdirectory_vertical_levels = None

    try:
        box_obj = BoxData(data, variable_list_df, min_lon, max_lon, min_lat, max_lat, args,
                          results_subdirectory, results_subdirectory_vertical_levels,
                          engine_options=engine_options)
    except Exception:
        app_logger.exception("❌ An exception occurred while creating BoxData object")
        raise
    rows_ms, fin_ms, call_ms = box_obj.timing_ms
    app_logger.info(f"🚀 B200 engine: {len(box_obj.times)} steps in {call_ms:.2f} ms "
                    f"({1e3 * len(box_obj.times) / max(call_ms, 1e-9):.1f} timesteps/s incl. host<->device copies)")

    try:
        ec_obj = EnergyContents(box_obj, "fixed", app_logger)
        energy_list = [ec_obj.calc_az(), ec_obj.calc_ae(), ec_obj.calc_kz(), ec_obj.calc_ke()]
    except Exception:
        app_logger.exception("❌ An exception occurred while computing EnergyContents")
        raise
    app_logger.info("⚡ Computed energy contents (Az, Ae, Kz, Ke)")
    try:
        ct_obj = ConversionTerms(box_obj, "fixed", app_logger)
        conversion_list = [ct_obj.calc_cz(), ct_obj.calc_ca(), ct_obj.calc_ck(), ct_obj.calc_ce()]
    except Exception:
        app_logger.exception("❌ An exception occurred while computing ConversionTerms")
        raise
    app_logger.info("🔄 Computed conversion terms (Cz, Ca, Ck, Ce)")
    try:
        bt_obj = BoundaryTerms(box_obj, "fixed", app_logger)
        boundary_list = [bt_obj.calc_baz(), bt_obj.calc_bae(), bt_obj.calc_bkz(), bt_obj.calc_bke(),
                         bt_obj.calc_boz(), bt_obj.calc_boe()]     # B-Phi computed, then dropped (:287-290)
    except Exception:
        app_logger.exception("❌ An exception occurred while computing BoundaryTerms")
        raise
    app_logger.info("🏁 Computed boundary terms (BAz, BAe, BKz, BKe, BΦZ, BΦE)")
    try:
        gdt_obj = GenerationDissipationTerms(box_obj, "fixed", app_logger)
        gen_diss_list = [gdt_obj.calc_gz(), gdt_obj.calc_ge()] if args.residuals else \
            [gdt_obj.calc_gz(), gdt_obj.calc_ge(), gdt_obj.calc_dz(), gdt_obj.calc_de()]
    except Exception:
        app_logger.exception("❌ An exception occurred while computing GenerationDissipationTerms")
        raise
    app_logger.info("🔥 Computed generation/dissipation terms (Gz, Ge, Dz, De)")

    dates = np.asarray(data.time)
    df = pd.DataFrame(index=dates.astype("datetime64[ns]"))
    for i, col in enumerate(["Az", "Ae", "Kz", "Ke"]):
        df[col] = energy_list[i]
    for i, col in enumerate(["Cz", "Ca", "Ck", "Ce"]):
        df[col] = conversion_list[i]
    for i, col in enumerate(["BAz", "BAe", "BKz", "BKe", "Gz", "Ge", "Dz", "De"][: len(gen_diss_list) + 4]):
        df[col] = boundary_list[i] if i < 4 else gen_diss_list[i - 4]
    df = calc_budget_diff(df, dates, app_logger)
    df = calc_residuals(df, app_logger)
    app_logger.info("📈 Computed budget and residuals")

    if getattr(args, "outname", None):
        results_filename = args.outname
    else:
        infile_name = os.path.basename(args.infile).split(".nc")[0]
        results_filename = f"{infile_name}_fixed_results"
    results_file = Path(results_subdirectory, f"{results_filename}.csv")
    if write:
        df.to_csv(results_file)
        app_logger.info(f"💾 Results saved to {results_file}")
    if getattr(args, "plots", False):
        app_logger.warning("⚠️ plots are produced by the reference's src/plots from these CSVs; "
                           "matplotlib/cartopy are not part of the B200 engine")
    return df
