"""850-hPa track diagnostics kernel (lec_diag850_host) against the numpy restatement oracle.diag850 (MetPy
1.6.2 ``vorticity`` on the lat / lon grid: geodesic grid deltas, map factors, 3-point ``first_derivative``):
fp64 values with the oracle's bits, arg-reduction indices exact (numpy argmin / argmax semantics)."""
import numpy as np
import pytest

from lorenzcycletoolkit_b200 import engine as E
from oracle import lec_oracle as O

pytestmark = pytest.mark.gpu


def _planes(rng, nt, nlat, nlon, dtype):
    lon = np.linspace(0, 2 * np.pi, nlon)[None, None, :]
    lat = np.linspace(-1, 1, nlat)[None, :, None]
    t = np.arange(nt)[:, None, None]
    u = 20 * np.cos(2 * lat) + 8 * np.sin(3 * lon + 0.3 * t) + rng.normal(0, 1, (nt, nlat, nlon))
    v = 6 * np.sin(2 * lon - 0.2 * t) * np.cos(lat) + rng.normal(0, 1, (nt, nlat, nlon))
    z = 1500 + 60 * np.cos(lon + 0.1 * t) * np.sin(2 * lat) + rng.normal(0, 2, (nt, nlat, nlon))
    return [np.ascontiguousarray(a, dtype=dtype) for a in (u, v, z)]


def _steps(boxes, centres=None):
    return E.diag_steps(boxes, centres)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("uniform", [True, False])
def test_matches_oracle_bitwise(dtype, uniform):
    rng = np.random.default_rng(7)
    nt, nlat, nlon = 6, 41, 73
    if uniform:
        lon = (-80 + 0.5 * np.arange(nlon)).astype(np.float32)     # the degree axis is exact, radians are not
        lat = (-50 + 0.5 * np.arange(nlat)).astype(np.float32)
    else:
        lon = np.cumsum(rng.uniform(0.3, 0.9, nlon)) - 80
        lat = np.cumsum(rng.uniform(0.3, 0.9, nlat)) - 40
    u, v, z = _planes(rng, nt, nlat, nlon, dtype)
    boxes = [(0, 0, nlon - 1, 0, nlat - 1),            # whole domain: one-sided stencils at the domain edges
             (1, 5, 35, 3, 33), (2, 0, 10, 30, 40), (3, 60, 72, 0, 5), (4, 11, 11, 7, 7), (5, 20, 50, 10, 11),
             (5, 1, 71, 1, 39)]
    # grid point nearest to a track centre per step (the -z branch): corners, edges and interior points
    centres = [(0, 0), (20, 18), (nlon - 1, nlat - 1), (66, 2), (11, 7), (0, 11), (36, 20)]
    for scale, z_div in (((1.0, 1.0, 1.0), 1.0), ((0.514444, 0.514444, 9.80665), O.g)):
        vals, idx = E.diag850_host(u, v, z, lon, lat, _steps(boxes, centres), scale=scale, z_div=z_div)
        ovals, oidx = O.diag850(u, v, z, lon, lat, boxes, scale=scale, z_div=z_div, centres=centres)
        assert vals.shape == (len(boxes), 5) and np.isfinite(vals).all()
        assert np.array_equal(idx, oidx)
        assert np.array_equal(vals, ovals), np.abs(vals / ovals - 1).max()
    # without centres the fifth value is NaN
    vals, _ = E.diag850_host(u, v, z, lon, lat, _steps(boxes))
    assert np.isnan(vals[:, 4]).all()


def test_vorticity_of_solid_body_rotation():
    """Known answer: u = U0 cos(lat), v = 0 has zeta = 2 U0 sin(lat) / a on the sphere; MetPy's ellipsoidal
    metric (as restated) stays within 1 % of it -- a gross check that the map-factor / curvature term is there."""
    nlat, nlon = 81, 60
    lon = (-60 + 0.5 * np.arange(nlon)).astype(np.float32)
    lat = (-50 + 0.5 * np.arange(nlat)).astype(np.float32)
    U0 = 30.0
    u = np.ascontiguousarray(np.broadcast_to(U0 * np.cos(np.deg2rad(lat.astype(np.float64)))[None, :, None], (1, nlat, nlon)))
    v = np.zeros_like(u); z = np.zeros_like(u)
    jc = 30
    vals, _ = E.diag850_host(u, v, z, lon, lat, _steps([(0, 0, nlon - 1, 0, nlat - 1)], [(10, jc)]))
    want = 2 * U0 * np.sin(np.deg2rad(float(lat[jc]))) / 6371008.7714
    assert abs(vals[0, 4] / want - 1) < 0.01


def test_nan_and_tie_semantics():
    rng = np.random.default_rng(8)
    nt, nlat, nlon = 4, 20, 24
    lon = -60 + 2.5 * np.arange(nlon); lat = -45 + 2.5 * np.arange(nlat)
    u, v, z = _planes(rng, nt, nlat, nlon, np.float32)
    z[0, 6, 9] = np.nan; z[0, 8, 3] = np.nan          # argmin -> the first NaN of the box, value skips NaNs
    u[1, 5, 5] = np.nan                                # zeta NaN at the point and its lat neighbours, wind NaN at the point
    z[2] = 1234.5                                      # ties -> first occurrence
    u[3, 4:12, 2:10] = np.nan; v[3, 4:12, 2:10] = np.nan; z[3, 4:12, 2:10] = np.nan     # all-NaN box
    boxes = [(0, 2, 15, 4, 12), (1, 2, 15, 2, 12), (2, 3, 20, 1, 18), (3, 4, 7, 6, 9)]
    vals, idx = E.diag850_host(u, v, z, lon, lat, _steps(boxes))
    ovals, oidx = O.diag850(u, v, z, lon, lat, boxes)
    assert np.array_equal(idx, oidx)
    assert np.array_equal(vals, ovals, equal_nan=True)
    assert idx[0, 2] == (6 - 4) * 14 + (9 - 2) and idx[2, 2] == 0 and np.isnan(vals[3]).all()


def test_errors():
    rng = np.random.default_rng(9)
    u, v, z = _planes(rng, 2, 10, 12, np.float32)
    lon = np.arange(12.0); lat = np.arange(10.0)
    with pytest.raises(IndexError):
        E.diag850_host(u, v, z, lon, lat, _steps([(0, 0, 12, 0, 9)]))
    with pytest.raises(IndexError):
        E.diag850_host(u, v, z, lon, lat, _steps([(2, 0, 11, 0, 9)]))
    with pytest.raises(ValueError):
        E.diag850_host(u, v, z[:, :5], lon, lat, _steps([(0, 0, 11, 0, 9)]))
    vals, idx = E.diag850_host(u, v, z, lon, lat, _steps([]))
    assert vals.shape == (0, 5)
    with pytest.raises(ValueError):                      # first_derivative needs three points per axis
        E.diag850_host(u[:, :2], v[:, :2], z[:, :2], lon, lat[:2], _steps([(0, 0, 11, 0, 1)]))
