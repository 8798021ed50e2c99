"""Shared plumbing of the four term classes: per-level CSV side effects
(``_save_vertical_levels``) and the reference's NaN path (``_handle_nans``) applied to the
per-level integrands the engine returns."""

from __future__ import annotations

import numpy as np
import pandas as pd

G = 9.80665
RD = 8.314462618 / 28.96546e-3


def trapz_levels(f, p):
    """``DataArray.integrate(level)``: trapezoid along the last axis."""
    return np.sum((p[1:] - p[:-1]) * 0.5 * (f[..., 1:] + f[..., :-1]), axis=-1)


def handle_nans(f, p):
    """``_handle_nans`` (energy_contents.py:190-208 and its copies): linear interpolation
    along the level coordinate (interior gaps only), then drop every level that still holds a
    NaN at any time.  Returns the cleaned integrand and its level coordinate."""
    f = np.array(f, dtype=np.float64, copy=True)
    if not np.isnan(f).any():
        return f, p
    for row in f.reshape(-1, f.shape[-1]):
        bad = np.isnan(row)
        if bad.any() and not bad.all():
            row[:] = np.interp(p, p[~bad], row[~bad], left=np.nan, right=np.nan)
    if np.isnan(f).any():
        keep = ~np.isnan(f.reshape(-1, f.shape[-1])).any(axis=0)
        f, p = f[..., keep], p[keep]
    return f, p


class TermBase:
    def __init__(self, box_obj, method, app_logger):
        if method not in ("fixed", "moving"):
            raise ValueError("method must be 'fixed' or 'moving'")
        self.box_obj = box_obj
        self.method = method
        self.app_logger = app_logger
        self.results_subdirectory = box_obj.results_subdirectory
        self.results_subdirectory_vertical_levels = box_obj.results_subdirectory_vertical_levels
        self.TimeName = box_obj.TimeName
        self.VerticalCoordIndexer = box_obj.VerticalCoordIndexer
        self.PressureData = box_obj.PressureData

    # -- engine results ------------------------------------------------------------------ #
    def _levels(self, name):
        return self.box_obj.level_term(name)

    def _result(self, values):
        """time-dimensioned array in fixed mode, scalar in moving mode (as the reference)."""
        values = np.asarray(values, dtype=np.float64)
        if self.method == "fixed" or getattr(self.box_obj, "batched", False):
            return values                 # (batched: all steps of a moving run, one CSV write per file)
        return float(values.reshape(-1)[0])

    def _volume_term(self, name, factor=1.0):
        """Integrated term ``name``: the device value, or -- if the engine flagged a non-finite
        integrand -- the reference's NaN path on the per-level integrand, re-integrated here."""
        f = self._levels(name)
        if self.box_obj.has_nonfinite and np.isnan(f).any():
            f, p = handle_nans(f, self.PressureData)
            self._save_vertical_levels(f, name, p)
            return self._result(trapz_levels(f, p) * factor)
        self._save_vertical_levels(f, name)
        return self._result(self.box_obj.term(name))

    # -- per-level CSV side effect --------------------------------------------------------- #
    def _save_vertical_levels(self, function, variable_name, levels=None):
        """Append the per-level rows to ``<term>_<level name>.csv`` (header written by the
        framework): one row per time; moving mode labels rows ``%Y-%m-%d %H:%M:%S``."""
        if self.results_subdirectory_vertical_levels is None:
            return function
        levels = self.PressureData if levels is None else levels
        path = f"{self.results_subdirectory_vertical_levels}/{variable_name}_{self.VerticalCoordIndexer}.csv"
        function = np.asarray(function)
        if function.ndim == 1:                       # time-less integrand (Cz_1, Ce_1)
            if self.method == "fixed":
                df = pd.DataFrame({self.VerticalCoordIndexer: levels, variable_name: function}).T
                df.to_csv(path, mode="a", header=None)
                return function
            function = np.broadcast_to(function, (len(self.box_obj.times), function.size))   # one row per step
        index = pd.DatetimeIndex(self.box_obj.times)
        if self.method != "fixed":
            index = index.strftime("%Y-%m-%d %H:%M:%S")
        df = pd.DataFrame(function, index=index, columns=[float(x) for x in levels])
        df.to_csv(path, mode="a", header=None)
        return function
