"""Properties at the benchmark size (BASELINE.json configs[3]: 1440 x 721 x 37, fp32, box = all
longitudes x 719 rows).  The oracle needs ~10 s per full step, so the full box is checked through
size-independent properties and a latitude band of it directly against the oracle."""
import numpy as np
import pytest

from oracle import lec_oracle as O
from lorenzcycletoolkit_b200 import engine as E, synthetic as S
import helpers as H

pytestmark = pytest.mark.gpu
NT = 4


@pytest.fixture(scope="module")
def c4():
    import torch
    grid = S.era5_grid()
    dev = S.synth_fields(grid, NT, np.float32, "cuda:0")
    f64 = lambda a: np.asarray(a, dtype=np.float64)
    eng = E.LecEngine(f64(grid["lon"]), f64(grid["lat"]), f64(grid["rlons"]), f64(grid["rlats"]),
                      f64(grid["coslats"]), grid["level"], np.float32, max_steps=NT, max_box_rows=719)
    steps = E.time_stencil(3600.0 * np.arange(NT), E.make_steps(NT))
    steps["i0"], steps["i1"], steps["j0"], steps["j1"] = 0, 1439, 1, 719
    yield grid, dev, eng, steps
    eng.close()
    del dev
    torch.cuda.empty_cache()


def _run(eng, dev, steps):
    import torch
    t, l, f = eng.run_torch(dev, steps)
    torch.cuda.synchronize()
    return t.cpu().numpy(), l.cpu().numpy(), f.cpu().numpy()


def test_full_box_is_finite_reproducible_and_physical(c4):
    grid, dev, eng, steps = c4
    t1, l1, f1 = _run(eng, dev, steps)
    t2, l2, _ = _run(eng, dev, steps)
    assert np.array_equal(t1, t2) and np.array_equal(l1, l2)            # bit-reproducible
    assert not f1.any() and np.isfinite(t1).all() and np.isfinite(l1).all()
    az, ae, kz, ke = (t1[:, i] for i in range(4))
    assert (az > 0).all() and (ae > 0).all() and (kz > 0).all() and (ke > 0).all()
    # per-level energy integrands are non-negative (sums of squares over sigma > 0)
    for n in ("Az", "Ae", "Kz", "Ke"):
        assert (l1[:, E.LEVEL_TERM_NAMES.index(n)] >= 0).all()
    # the scalar terms are the trapezoid of their per-level integrands (checksum of checksums)
    p = grid["level"]
    trap = lambda y: np.sum((p[1:] - p[:-1]) * 0.5 * (y[:, 1:] + y[:, :-1]), axis=1)
    assert np.allclose(trap(l1[:, E.LEVEL_TERM_NAMES.index("Az")]), az, rtol=1e-12)
    assert np.allclose(trap(l1[:, E.LEVEL_TERM_NAMES.index("Kz")]) / (2 * O.g), kz, rtol=1e-12)
    assert np.allclose(trap(l1[:, E.LEVEL_TERM_NAMES.index("Ck")]) / O.g, t1[:, 6], rtol=1e-12)
    assert np.allclose(trap(l1[:, E.LEVEL_TERM_NAMES.index("Ge")]), t1[:, 15], rtol=1e-12)


def test_scaling_laws(c4):
    """Doubling the winds quadruples Kz/Ke and leaves Az/Ae unchanged (to fp32 rounding)."""
    grid, dev, eng, steps = c4
    t1, _, _ = _run(eng, dev, steps)
    dev2 = [dev[0], dev[1] * 2, dev[2] * 2, dev[3], dev[4]]
    t2, _, _ = _run(eng, dev2, steps)
    assert np.allclose(t2[:, 2], 4 * t1[:, 2], rtol=2e-6) and np.allclose(t2[:, 3], 4 * t1[:, 3], rtol=2e-6)
    assert np.allclose(t2[:, 0], t1[:, 0], rtol=1e-12) and np.allclose(t2[:, 1], t1[:, 1], rtol=1e-12)
    assert np.allclose(t2[:, 7], t1[:, 7], rtol=1e-12)                  # Ce = -(R/pg) AA(w'T') does not see u, v


def test_latitude_band_against_oracle(c4):
    """A 96-row band (24 degrees of latitude) of the full-width grid against the oracle (fp64 arithmetic
    on the same fp32 values): the 1440-wide rows, 12 sweep iterations per row, all 37 levels.
    (On a much narrower band the area eddies [X]_j - [[X]] become ~1e-3 of the zonal eddies and the
    fp32 row sums limit sub-terms such as Cz_2 to ~2e-5: DESIGN.md section 5.)"""
    grid, dev, eng, steps = c4
    j0, j1 = 380, 475
    st = steps.copy()
    st["j0"], st["j1"] = j0, j1
    terms, levels, flags = _run(eng, dev, st)
    assert not flags.any()
    host = [d[:, :, j0:j1 + 1, :].cpu().numpy() for d in dev]
    P = H.prepared_from_arrays(host, grid["lon"], grid["lat"][j0:j1 + 1], grid["level"],
                               np.datetime64("2020-01-01T00") + np.arange(NT) * np.timedelta64(1, "h"))
    df, lv, extra = O.lec_fixed(P, float(grid["lon"][0]), float(grid["lon"][-1]), float(P.lat[0]), float(P.lat[-1]), mode="fp64")
    errs = H.compare_terms(terms, df, extra=extra)
    bad = {k: v for k, v in errs.items() if not v <= 1e-5}
    assert not bad, bad
    lerrs = H.compare_levels(levels, lv)
    bad = {k: v for k, v in lerrs.items() if not v <= 1e-5}
    assert not bad, bad


def test_narrow_latitude_band_against_oracle(c4):
    """A 24-row band (6 degrees) of the full-width grid: the area eddies [X]_j - [[X]] are ~1e-3 of the zonal
    eddies, so every term built from them amplifies the rounding of the zonal means.  Long rows + few rows is
    the shape that switches the fp32 row kernels to compensated linear sums (lec_lin_add); round 1 measured
    Cz_2 / Ca_2 at 1.7e-5 here without them."""
    grid, dev, eng, steps = c4
    j0, j1 = 400, 423
    st = steps.copy()
    st["j0"], st["j1"] = j0, j1
    terms, levels, flags = _run(eng, dev, st)
    assert not flags.any()
    host = [d[:, :, j0:j1 + 1, :].cpu().numpy() for d in dev]
    P = H.prepared_from_arrays(host, grid["lon"], grid["lat"][j0:j1 + 1], grid["level"],
                               np.datetime64("2020-01-01T00") + np.arange(NT) * np.timedelta64(1, "h"))
    df, lv, extra = O.lec_fixed(P, float(grid["lon"][0]), float(grid["lon"][-1]), float(P.lat[0]), float(P.lat[-1]), mode="fp64")
    errs = H.compare_terms(terms, df, extra=extra)
    errs.update({"lv:" + k: v for k, v in H.compare_levels(levels, lv).items()})
    bad = {k: v for k, v in errs.items() if not v <= 1e-5}
    assert not bad, bad


# --------------------------------------------------------------------------------------------------------- #
# The benchmark's own configuration against the oracle: the FULL C4 box (1440 x 719 rows x 37 levels), an
# edge time step (one-sided dT/dt) and an interior one (centred), all 16 terms + 19 per-level families, in
# the three arithmetic variants, at the tolerances BASELINE.json's north_star states (1e-5 for fp32-input
# mode, 1e-9 in fp64).  The oracle evaluates each step as the moving framework does (BoxData on
# [level][lat][lon] with the dT/dt of np.gradient over the 3-slot time axis) in fp64 on the same fp32 values.
@pytest.fixture(scope="module")
def c4_full_oracle():
    import torch
    grid = S.era5_grid()
    dev = S.synth_fields(grid, 3, np.float32, "cuda:0")
    host = [d[:, :, 1:720, :].cpu().numpy().astype(np.float64) for d in dev]
    P = H.prepared_from_arrays(host, grid["lon"], grid["lat"][1:720], grid["level"],
                               np.datetime64("2020-01-01T00") + np.arange(3) * np.timedelta64(1, "h"))
    for n in ("lat", "lon", "rlats", "coslats", "rlons"):      # fp64 semantics: the stored float32 coordinate values,
        setattr(P, n, getattr(P, n).astype(np.float64))         # upcast (oracle.to_mode), never recomputed in double
    tsec = 3600.0 * np.arange(3)
    dTdt = O.differentiate(P.fields["Air Temperature"], tsec, 0)
    want = []
    for it in (0, 1):
        b = O.BoxState(P, float(P.lon[0]), float(P.lon[-1]), float(P.lat[0]), float(P.lat[-1]),
                       fixed=False, dTdt=dTdt[it], tsel=it)
        lv, terms = {}, {}
        terms.update(O.energy_contents(b, lv))
        terms.update(O.conversion_terms(b, lv))
        terms.update(O.boundary_terms(b))
        terms.update(O.generation_terms(b, lv))
        want.append(({k: float(v) for k, v in terms.items()}, {k: np.asarray(v, dtype=np.float64) for k, v in lv.items()}))
        del b
    del P, host, dTdt
    yield grid, dev, want
    del dev
    torch.cuda.empty_cache()


@pytest.mark.parametrize("variant,tol", [("f32", 1e-5), ("math64", 1e-9), ("f64", 1e-9)])
def test_full_c4_box_against_oracle(c4_full_oracle, variant, tol):
    import torch
    grid, dev, want = c4_full_oracle
    f64 = lambda a: np.asarray(a, dtype=np.float64)
    dtype = np.float64 if variant == "f64" else np.float32
    fields = [d.double() for d in dev] if variant == "f64" else dev
    eng = E.LecEngine(f64(grid["lon"]), f64(grid["lat"]), f64(grid["rlons"]), f64(grid["rlats"]),
                      f64(grid["coslats"]), grid["level"], dtype, max_steps=3, max_box_rows=719,
                      math=E.LEC_MATH_F64 if variant == "math64" else E.LEC_MATH_AUTO)
    steps = E.time_stencil(3600.0 * np.arange(3), E.make_steps(3))
    steps["i0"], steps["i1"], steps["j0"], steps["j1"] = 0, 1439, 1, 719
    t, l, f = eng.run_torch(fields, steps)
    torch.cuda.synchronize()
    terms, levels, flags = t.cpu().numpy(), l.cpu().numpy(), f.cpu().numpy()
    eng.close()
    del fields
    assert not flags.any()
    # series-scaled error over the two compared steps (max_t |a - b| / max_t |b|)
    bad = {}
    for i, name in enumerate(E.TERM_NAMES):
        ref = np.array([want[it][0][name] for it in (0, 1)])
        e = H.series_err(terms[:2, i], ref)
        if not e <= tol:
            bad[name] = e
    worst = {}
    for i, name in enumerate(E.LEVEL_TERM_NAMES):
        ref = np.stack([want[it][1][name] for it in (0, 1)])
        e = H.series_err(levels[:2, i, :], ref)
        worst[name] = e
        if not e <= tol:
            bad["lv:" + name] = e
    print(f"full C4 box, {variant}: worst per-level error {max(worst.values()):.2e} "
          f"({max(worst, key=worst.get)}), tolerance {tol:g}")
    assert not bad, bad
