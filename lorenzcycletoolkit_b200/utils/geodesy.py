"""Grid metrics of a 1-D latitude / longitude grid as MetPy 1.6.2 hands them to ``vorticity``
(reference: ``src/frameworks/lec_moving_framework.py:660-663`` -> ``metpy.calc.vorticity(u_850, v_850)``;
MetPy: ``metpy/xarray.py`` ``grid_deltas``, ``metpy/calc/tools.py`` ``nominal_lat_lon_grid_deltas`` and
``parse_grid_arguments``).  Host-side tables for ``lec_diag850_*``: O(nlon + nlat) numbers per run.

* ``dx = a * diff(lon [rad])``: the "nominal" zonal spacing, taken on the equator;
* ``dy``: geodesic distance between consecutive latitudes along a meridian (``Geod.inv``), i.e. the meridian arc of
  the ellipsoid, negative where latitude decreases;
* ``parallel_scale``, ``meridional_scale``: PROJ's map factors of the geographic "projection" x = lambda, y = phi
  (``Proj.get_factors``): k = sqrt(1 - e^2 sin^2 phi) / cos phi, h = (1 - e^2 sin^2 phi)^(3/2) / (1 - e^2).

Ellipsoid: the one ``CRS('+proj=latlon')`` carries (PROJ's default, GRS80)."""

from __future__ import annotations

import numpy as np

GRS80_A = 6378137.0
GRS80_F = 1.0 / 298.257222101
GRS80_E2 = GRS80_F * (2.0 - GRS80_F)

# 8-point Gauss-Legendre rule on [-1, 1]
_X8 = np.array([-0.9602898564975363, -0.7966664774136267, -0.5255324099163290, -0.1834346424956498,
                0.1834346424956498, 0.5255324099163290, 0.7966664774136267, 0.9602898564975363])
_W8 = np.array([0.1012285362903763, 0.2223810344533745, 0.3137066458778873, 0.3626837833783620,
                0.3626837833783620, 0.3137066458778873, 0.2223810344533745, 0.1012285362903763])


def meridian_arcs(lat_deg):
    """Signed meridian arc length [m] between consecutive latitudes: the integral of the meridional radius of
    curvature a (1 - e^2) / (1 - e^2 sin^2 phi)^(3/2), one 8-point Gauss-Legendre rule per interval (exact to
    rounding for spacings of a few degrees)."""
    phi = np.asarray(lat_deg, dtype=np.float64) * (np.pi / 180.0)
    half = 0.5 * (phi[1:] - phi[:-1])
    mid = 0.5 * (phi[1:] + phi[:-1])
    x = mid[:, None] + half[:, None] * _X8[None, :]
    radius = GRS80_A * (1.0 - GRS80_E2) / (1.0 - GRS80_E2 * np.sin(x) ** 2) ** 1.5
    return half * np.sum(radius * _W8[None, :], axis=1)


def latlon_grid_metrics(lon_deg, lat_deg):
    """``(dx[nlon-1], dy[nlat-1], parallel_scale[nlat], meridional_scale[nlat])`` in float64 from the stored
    coordinate values (upcast, never recomputed)."""
    lon = np.asarray(lon_deg, dtype=np.float64) * (np.pi / 180.0)
    phi = np.asarray(lat_deg, dtype=np.float64) * (np.pi / 180.0)
    dx = GRS80_A * np.diff(lon)
    dy = meridian_arcs(lat_deg)
    t = 1.0 - GRS80_E2 * np.sin(phi) ** 2
    return dx, dy, np.sqrt(t) / np.cos(phi), t * np.sqrt(t) / (1.0 - GRS80_E2)
