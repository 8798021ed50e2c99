"""GPU parity: the CUDA engine (through the C ABI) against the numpy oracle on the
reference's bundled datasets.  Tolerances (BASELINE.json north_star):
  fp64 arithmetic : 1e-9  (series-scaled: max_t|a-b| / max_t|b|)
  fp32-input mode : 1e-5
"""
import numpy as np
import pytest

from oracle import lec_oracle as O
from lorenzcycletoolkit_b200 import engine as E
import helpers as H

pytestmark = pytest.mark.gpu

TOL64 = 1e-9
TOL32 = 1e-5

CASES = {
    "catarina": dict(nc="Catarina_NCEP-R2.nc", box=(-55, -36, -35, -20)),
    "testdata_reg1": dict(nc="testdata_NCEP-R2.nc", box=(-60, -30, -42.5, -17.5)),
    "testdata_testcase": dict(nc="testdata_NCEP-R2.nc", box=(-53, -44, -31, -24)),
}


def _fixed_case(name):
    c = CASES[name]
    P, _ = H.load_prepared(c["nc"])
    W, Ea, S, N = c["box"]
    P = O.slice_domain_fixed(P, W, Ea, S, N)
    df, lv, extra = O.lec_fixed(P, W, Ea, S, N, mode="fp64")
    return P, c["box"], df, lv, extra


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("variant", ["f64", "f32_math64", "f32"])
def test_fixed_box_matches_oracle(name, variant):
    P, (W, Ea, S, N), df, lv, extra = _fixed_case(name)
    dtype = np.float64 if variant == "f64" else np.float32
    math = E.LEC_MATH_F64 if variant == "f32_math64" else E.LEC_MATH_AUTO
    tol = TOL32 if variant == "f32" else TOL64
    fields, scale = H.engine_inputs(P, dtype)
    with H.make_engine(P, dtype, scale, math=math) as eng:
        steps = H.fixed_steps(P, W, Ea, S, N)
        terms, levels, flags = eng.run_host(fields, steps)
    assert not (flags & E.FLAG_NONFINITE).any()
    errs = H.compare_terms(terms, df, extra=extra)
    assert set(errs) == set(E.TERM_NAMES)
    bad = {k: v for k, v in errs.items() if not v <= tol}
    assert not bad, f"{name}/{variant}: {bad}"
    lerrs = H.compare_levels(levels, lv)
    bad = {k: v for k, v in lerrs.items() if not v <= tol}
    assert not bad, f"{name}/{variant} levels: {bad}"


@pytest.mark.parametrize("variant", ["f64", "f32"])
def test_moving_box_matches_oracle(variant):
    P, tr = H.load_prepared("testdata_NCEP-R2.nc", track="track_testdata_NCEP-R2")
    P = O.slice_domain_track(P, tr)
    df, lv, boxes = O.lec_moving(P, tr, mode="fp64")
    dtype = np.float64 if variant == "f64" else np.float32
    tol = TOL64 if variant == "f64" else TOL32
    fields, scale = H.engine_inputs(P, dtype)
    steps = H.moving_steps(P, tr)
    # box/track index selection is bit-exact against the pandas model
    for it, (_, idx) in enumerate(boxes):
        assert (steps["i0"][it], steps["i1"][it], steps["j0"][it], steps["j1"][it]) == idx
    with H.make_engine(P, dtype, scale) as eng:
        terms, levels, flags = eng.run_host(fields, steps)
    errs = H.compare_terms(terms, df)
    assert set(errs) == set(E.TERM_NAMES)
    bad = {k: v for k, v in errs.items() if not v <= tol}
    assert not bad, f"moving/{variant}: {bad}"
    lerrs = H.compare_levels(levels, lv)
    bad = {k: v for k, v in lerrs.items() if not v <= tol}
    assert not bad, f"moving/{variant} levels: {bad}"


@pytest.mark.parametrize("variant", ["f64", "f32"])
def test_tiled_row_kernel_matches_oracle(variant, monkeypatch):
    """LEC_ROW_KERNEL=tile (persistent CTAs, one producer thread feeding a shared-memory ring with TMA tensor
    loads, one consumer warp per box row) writes the same row records: same gates as the direct-load kernel.
    (LEC_NARROW=0: the 8-column Catarina box would otherwise take the sub-warp kernel.)"""
    monkeypatch.setenv("LEC_ROW_KERNEL", "tile")
    monkeypatch.setenv("LEC_NARROW", "0")
    P, (W, Ea, S, N), df, lv, extra = _fixed_case("catarina")       # nlon = 8: rows are 16-byte aligned
    dtype = np.float64 if variant == "f64" else np.float32
    tol = TOL64 if variant == "f64" else TOL32
    fields, scale = H.engine_inputs(P, dtype)
    with H.make_engine(P, dtype, scale) as eng:
        terms, levels, flags = eng.run_host(fields, H.fixed_steps(P, W, Ea, S, N))
    errs = H.compare_terms(terms, df, extra=extra)
    bad = {k: v for k, v in errs.items() if not v <= tol}
    assert not bad, bad
    lerrs = H.compare_levels(levels, lv)
    bad = {k: v for k, v in lerrs.items() if not v <= tol}
    assert not bad, bad
