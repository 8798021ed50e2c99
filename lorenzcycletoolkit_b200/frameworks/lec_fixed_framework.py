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


# (class, [(output column, method)]) in the order the reference evaluates them (lec_fixed_framework.py:215-279)
TERM_GROUPS = [
    (EnergyContents, [("Az", "calc_az"), ("Ae", "calc_ae"), ("Kz", "calc_kz"), ("Ke", "calc_ke")]),
    (ConversionTerms, [("Cz", "calc_cz"), ("Ca", "calc_ca"), ("Ck", "calc_ck"), ("Ce", "calc_ce")]),
    (BoundaryTerms, [("BAz", "calc_baz"), ("BAe", "calc_bae"), ("BKz", "calc_bkz"), ("BKe", "calc_bke"),
                     ("BΦZ", "calc_boz"), ("BΦE", "calc_boe")]),
    (GenerationDissipationTerms, [("Gz", "calc_gz"), ("Ge", "calc_ge"), ("Dz", "calc_dz"), ("De", "calc_de")]),
]
# columns of <stem>_fixed_results.csv: the B-Phi terms are evaluated and then dropped (:281-290)
RESULT_COLUMNS = ["Az", "Ae", "Kz", "Ke", "Cz", "Ca", "Ck", "Ce", "BAz", "BAe", "BKz", "BKe", "Gz", "Ge", "Dz", "De"]


def lec_fixed(data, variable_list_df, results_subdirectory, results_subdirectory_vertical_levels,
              app_logger, args, engine_options=None):
    log = app_logger or logging.getLogger("lorenzcycletoolkit")
    log.info("📊 Fixed (Eulerian) framework on the B200 engine")
    try:
        lim = read_box_limits(args.box_limits)
    except FileNotFoundError:
        log.error(f"❌ no box limits file at {os.path.abspath(args.box_limits)}")
        raise FileNotFoundError(f"Box limits file not found: {os.path.abspath(args.box_limits)}. "
                                "Create one or use --box_limits to specify path.")
    data = data.compute()
    time_name = variable_list_df.loc["Time"]["Variable"]
    level_name = variable_list_df.loc["Vertical Level"]["Variable"]
    log.info("🗺️ box lon=[%s, %s] lat=[%s, %s]", lim["min_lon"], lim["max_lon"], lim["min_lat"], lim["max_lat"])
    write = _is_rank0()          # under torchrun every rank computes its time shard; rank 0 writes the files
    if write:
        create_level_files(results_subdirectory_vertical_levels, time_name, level_name, data.level)
    else:
        results_subdirectory_vertical_levels = None

    try:
        box_obj = BoxData(data, variable_list_df, lim["min_lon"], lim["max_lon"], lim["min_lat"], lim["max_lat"],
                          args, results_subdirectory, results_subdirectory_vertical_levels,
                          engine_options=engine_options)
    except Exception:
        log.exception("❌ BoxData could not be built")
        raise
    nsteps, call_ms = len(box_obj.times), box_obj.timing_ms[2]
    log.info("🚀 %d time steps in %.2f ms on the GPU (%.1f timesteps/s incl. host<->device copies)",
             nsteps, call_ms, 1e3 * nsteps / max(call_ms, 1e-9))

    columns = {}
    for cls, calls in TERM_GROUPS:
        try:
            obj = cls(box_obj, "fixed", log)
            for name, method in calls:
                if name in ("Dz", "De") and args.residuals:
                    continue                                  # -r: dissipation comes out of the residuals
                columns[name] = getattr(obj, method)()
        except Exception:
            log.exception("❌ %s failed", cls.__name__)
            raise
        log.info("✅ %s: %s", cls.__name__, ", ".join(n for n, _ in calls))

    dates = np.asarray(data.time)
    df = pd.DataFrame({c: columns[c] for c in RESULT_COLUMNS if c in columns}, index=dates.astype("datetime64[ns]"))
    df = calc_residuals(calc_budget_diff(df, dates, log), log)     # always, with or without -r (:292-293)

    df.attrs["engine_ms"] = {"row_kernels": box_obj.timing_ms[0], "finalize_kernels": box_obj.timing_ms[1],
                             "call_incl_copies": box_obj.timing_ms[2]}       # device time of the engine call
    stem = args.outname if getattr(args, "outname", None) else \
        os.path.basename(args.infile).split(".nc")[0] + "_fixed_results"
    results_file = Path(results_subdirectory, f"{stem}.csv")
    if write:
        df.to_csv(results_file)
        log.info("💾 %s", results_file)
    if getattr(args, "plots", False):
        log.warning("⚠️ figures are made by the reference's src/plots from these CSVs; not part of the engine")
    return df
