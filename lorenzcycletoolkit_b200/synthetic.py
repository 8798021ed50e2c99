"""Seeded synthetic ERA5-shaped fields for benchmarks and full-size property tests
(SURVEY.md section 8(d)): realistic stratification (sigma stays far above the 0.03
floor), a jet, planetary waves with random phases and white noise.  Generated on the
device with torch; no file or network access."""

from __future__ import annotations

import math

import numpy as np

ERA5_LEVELS_HPA = [1, 2, 3, 5, 7, 10, 20, 30, 50, 70, 100, 125, 150, 175, 200, 225, 250, 300, 350,
                   400, 450, 500, 550, 600, 650, 700, 750, 775, 800, 825, 850, 875, 900, 925, 950,
                   975, 1000]


def era5_grid(nlon=1440, nlat=721, levels_hpa=None, coord_dtype=np.float32):
    """Coordinates exactly as ``process_data`` hands them over (preprocessing.py:275-365):
    lon in [-180, 180), lat ascending, level ascending in Pa, radians/cos in the
    coordinate dtype."""
    lev = np.asarray(ERA5_LEVELS_HPA if levels_hpa is None else levels_hpa, dtype=np.float64) * 100.0
    lon = (-180.0 + 360.0 / nlon * np.arange(nlon)).astype(coord_dtype)
    lat = np.linspace(-90.0, 90.0, nlat).astype(coord_dtype)
    return dict(lon=lon, lat=lat, level=lev, rlons=np.deg2rad(lon), rlats=np.deg2rad(lat),
                coslats=np.cos(np.deg2rad(lat)))


def synth_fields(grid, nslots, dtype, device, seed=1234, t0=0, dt_hours=1.0, out=None):
    """Five tensors ``[slot][level][lat][lon]`` (T, u, v, omega, Phi) on ``device``.
    Slot ``s`` depends only on (seed, t0 + s), so shards generate consistent halos."""
    import torch

    tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    dev = torch.device(device)
    lon = torch.as_tensor(np.asarray(grid["rlons"], dtype=np.float64), device=dev)
    lat = torch.as_tensor(np.asarray(grid["rlats"], dtype=np.float64), device=dev)
    p = torch.as_tensor(np.asarray(grid["level"], dtype=np.float64), device=dev)
    L, ny, nx = p.numel(), lat.numel(), lon.numel()
    if out is None:
        out = [torch.empty((nslots, L, ny, nx), dtype=tdt, device=dev) for _ in range(5)]
    x = (p / 1.0e5)[:, None, None]
    coslat = torch.cos(lat)[None, :, None]
    sinlat = torch.sin(lat)[None, :, None]
    Tbar = 288.0 - 60.0 * (1.0 - x ** 0.19) + 15.0 * (coslat ** 2 - 0.5)
    jet = 20.0 * torch.exp(-((lat[None, :, None].abs() - 0.78) / 0.26) ** 2) * torch.sin(math.pi * x.clamp(max=1.0)) ** 0.5
    # hydrostatic Phi of the basic state: Phi(p) = Rd * int_p^ps T dlnp (trapezoid on the level axis)
    Rd = 287.04749097718457
    lnp = torch.log(p)
    Tcol = Tbar.expand(L, ny, 1)
    seg = 0.5 * (Tcol[1:] + Tcol[:-1]) * (lnp[1:] - lnp[:-1])[:, None, None]
    Phibar = Rd * torch.flip(torch.cumsum(torch.flip(seg, [0]), 0), [0])
    Phibar = torch.cat([Phibar, torch.zeros_like(Phibar[:1])], 0)
    gen = torch.Generator(device=dev)
    for s in range(nslots):
        t = t0 + s
        rng = np.random.default_rng([seed, t])
        hours = t * dt_hours
        # zonal-mean meridional circulation (Hadley/Ferrel-like cells): [v] ~ 1 m/s, [omega] ~ 0.05 Pa/s
        cell = torch.sin(math.pi * x.clamp(max=1.0))
        vbar = 1.0 * torch.sin(2 * lat)[None, :, None] * torch.cos(math.pi * x.clamp(max=1.0))
        wbar = 0.05 * torch.cos(3 * lat)[None, :, None] * cell
        fields = [Tbar.expand(L, ny, nx).clone(), jet.expand(L, ny, nx).clone(),
                  vbar.expand(L, ny, nx).clone(), wbar.expand(L, ny, nx).clone(),
                  Phibar.expand(L, ny, nx).clone()]
        amp = [3.0, 8.0, 6.0, 0.2, 300.0]
        for m in range(1, 9):
            ph0 = np.random.default_rng([seed, 7, m]).uniform(0, 2 * math.pi, size=5)
            speed = 2 * math.pi / (24.0 * (2 + m))          # phase speed, rad/hour
            tilt = torch.as_tensor(rng.uniform(0.5, 1.5), device=dev)
            for f in range(5):
                phase = m * lon[None, None, :] + ph0[f] - speed * hours + 0.6 * m * (1.0 - x) * tilt
                env = (coslat ** 2) * (1.0 + 0.3 * sinlat * (f % 2))
                fields[f] += (amp[f] / m) * env * torch.cos(phase)
        gen.manual_seed(int(seed) * 1_000_003 + t)
        noise = [0.5, 1.0, 1.0, 0.02, 20.0]
        for f in range(5):
            fields[f] += noise[f] * torch.randn((L, ny, nx), dtype=torch.float32, device=dev, generator=gen)
            out[f][s].copy_(fields[f])
    return out
