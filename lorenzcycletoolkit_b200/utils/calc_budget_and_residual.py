"""Budget derivatives and residuals (reference: ``src/utils/calc_budget_and_residual.py:32-56,
131-154``): O(T) host epilogue on the gathered per-step scalars."""
import numpy as np


def calc_budget_diff(df, time, app_logger=None):
    """``np.gradient`` of Az, Ae, Kz, Ke in time, dt = times[1] - times[0] in seconds."""
    time = np.asarray(time)
    if len(time) < 2:
        raise ValueError("the budget needs at least two time steps")
    dt = float((time[1] - time[0]) / np.timedelta64(1, "s"))
    for term in ("Az", "Ae", "Kz", "Ke"):
        df[f"∂{term}/∂t (finite diff.)"] = np.gradient(df[term], dt)
    return df


def calc_residuals(df, app_logger=None):
    df["RGz"] = df["∂Az/∂t (finite diff.)"] + df["Cz"] + df["Ca"] - df["BAz"]
    df["RKz"] = df["∂Kz/∂t (finite diff.)"] - df["Cz"] - df["Ck"] - df["BKz"]
    df["RGe"] = df["∂Ae/∂t (finite diff.)"] - df["Ca"] + df["Ce"] - df["BAe"]
    df["RKe"] = df["∂Ke/∂t (finite diff.)"] - df["Ce"] + df["Ck"] - df["BKe"]
    return df
