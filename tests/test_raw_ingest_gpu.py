"""lec_run_host_raw: fields still in FILE layout (latitude north-to-south, longitudes 0..360, levels in any
order with extra ones above 10 hPa, int16-packed or float, a time selection) against lec_run_host on the
arrays the host pipeline would have prepared -- bit-identical results, fewer PCIe bytes."""
import numpy as np
import pytest

import helpers as H
from lorenzcycletoolkit_b200 import engine as E
from test_engine_synthetic_gpu import _dataset

pytestmark = pytest.mark.gpu


def _raw_layout(rng, fields, extra_levels=2, roll=7):
    """Scramble engine-layout fields into a file layout; returns raw arrays and the engine->raw maps."""
    nt, nlev, nlat, nlon = fields[0].shape
    lev_perm = rng.permutation(nlev + extra_levels)            # raw level order
    lev_map = np.array([int(np.where(lev_perm == k)[0][0]) for k in range(nlev)], dtype=np.int32)
    lat_map = np.arange(nlat - 1, -1, -1, dtype=np.int32)      # stored north -> south
    lon_map = ((np.arange(nlon) + roll) % nlon).astype(np.int32)
    raws = []
    for f in fields:
        raw = (rng.normal(size=(nt, nlev + extra_levels, nlat, nlon)) * 1e3).astype(f.dtype)    # junk in unused levels
        raw[:, lev_map[:, None, None], lat_map[None, :, None], lon_map[None, None, :]] = f
        raws.append(np.ascontiguousarray(raw))
    return raws, lon_map, lat_map, lev_map


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_float_records_any_layout(dtype):
    rng = np.random.default_rng(11)
    P, fields = _dataset(48, 19, 6, 7, dtype)
    steps = H.fixed_steps(P, P.lon[2], P.lon[44], P.lat[1], P.lat[17])
    raws, lon_map, lat_map, lev_map = _raw_layout(rng, fields)
    raws[0][2, lev_map[3], lat_map[5], lon_map[9]] = -9999.0       # a fill value inside the box
    ref_fields = [f.copy() for f in fields]
    ref_fields[0][2, 3, 5, 9] = np.nan
    with H.make_engine(P, dtype, [1.0] * 5) as eng:
        a = eng.run_host(ref_fields, steps)
        h2d_a = eng.last_transfer()[0]
        b = eng.run_host_raw(raws, lon_map, lat_map, lev_map, np.arange(7), steps,
                             decode=[dict(fills=[-9999.0])] + [{}] * 4)
        h2d_b = eng.last_transfer()[0]
    assert (a[2][1:4] & E.FLAG_NONFINITE).all()
    for x, y in zip(a, b):
        assert np.array_equal(x, y, equal_nan=True)
    nk = int(lev_map.max() - lev_map.min() + 1)      # the raw level RANGE the maps touch crosses PCIe
    assert h2d_b * 6 == h2d_a * nk


def test_time_selection_and_crop():
    """Engine grid = a crop of the raw grid (rows and columns), slots = every other record."""
    rng = np.random.default_rng(12)
    P, fields = _dataset(40, 15, 5, 6, np.float32)
    nt, nlev, nlat, nlon = fields[0].shape
    big = [rng.normal(size=(2 * nt, nlev, nlat + 9, nlon + 12)).astype(np.float32) for _ in range(5)]
    j0, i0 = 4, 5
    for b, f in zip(big, fields):
        b[::2, :, j0:j0 + nlat, i0:i0 + nlon] = f
    steps = H.fixed_steps(P, P.lon[1], P.lon[37], P.lat[1], P.lat[13])
    with H.make_engine(P, np.float32, [1.0] * 5) as eng:
        a = eng.run_host(fields, steps)
        b = eng.run_host_raw(big, i0 + np.arange(nlon), j0 + np.arange(nlat), np.arange(nlev), 2 * np.arange(nt), steps)
        # only the touched rows cross PCIe (whole rows of them)
        assert eng.last_transfer()[0] == 5 * nt * nlev * nlat * (nlon + 12) * 4
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("with_offset", [True, False])
def test_packed_int16(with_offset):
    """ERA5-style packing: decoded like xarray (float64 with an add_offset, float32 without), a quarter /
    half of the PCIe bytes of the decoded arrays."""
    rng = np.random.default_rng(13)
    P, fields = _dataset(36, 17, 6, 5, np.float64)
    steps = H.fixed_steps(P, P.lon[1], P.lon[33], P.lat[1], P.lat[15])
    packed, decode, decoded = [], [], []
    for f in fields:
        lo, hi = f.min(), f.max()
        scale = np.float64((hi - lo) / 65000.0)
        offset = np.float64(0.5 * (hi + lo)) if with_offset else None
        q = np.round((f - (offset if with_offset else 0.0)) / scale)
        if not with_offset:
            scale = np.float64(np.abs(f).max() / 32000.0)
            q = np.round(f / scale)
        q = q.astype(np.int16)
        packed.append(q)
        if with_offset:
            d = q.astype(np.float64); d *= scale; d += offset
        else:
            d = q.astype(np.float32); d *= scale
        decoded.append(d)
        decode.append(dict(scale=scale, offset=offset, float32=not with_offset))
    dt = np.float64 if with_offset else np.float32
    raws, lon_map, lat_map, lev_map = _raw_layout(rng, packed, extra_levels=1, roll=18)
    raws[3][1, lev_map[2], lat_map[6], lon_map[10]] = -32767
    decoded[3][1, 2, 6, 10] = np.nan
    decode[3]["fills"] = [-32767]
    with H.make_engine(P, dt, [1.0] * 5) as eng:
        a = eng.run_host(decoded, steps)
        h2d_a = eng.last_transfer()[0]
        b = eng.run_host_raw(raws, lon_map, lat_map, lev_map, np.arange(5), steps, decode=decode)
        h2d_b = eng.last_transfer()[0]
    for x, y in zip(a, b):
        assert np.array_equal(x, y, equal_nan=True)
    nk = int(lev_map.max() - lev_map.min() + 1)
    assert h2d_b * np.dtype(dt).itemsize * 6 == h2d_a * 2 * nk      # int16 records, raw level range nk for 6 engine levels


@pytest.mark.parametrize("dtype", [np.int16, np.float32, np.float64])
def test_big_endian_interleaved_records(dtype):
    """NetCDF-3 classic layout: big-endian values, the records of the five variables interleaved in one
    buffer (record r of T, of u, ... then record r + 1) -- consumed in place."""
    rng = np.random.default_rng(14)
    fdt = np.float32 if dtype == np.int16 else dtype
    P, fields = _dataset(32, 13, 5, 6, fdt)
    steps = H.fixed_steps(P, P.lon[1], P.lon[30], P.lat[1], P.lat[11])
    nt, nlev, nlat, nlon = fields[0].shape
    if dtype == np.int16:
        scales = [np.float64(np.abs(f).max() / 32000.0) for f in fields]
        native = [np.round(f / s).astype(np.int16) for f, s in zip(fields, scales)]
        ref = []
        for q, s in zip(native, scales):
            d = q.astype(np.float32); d *= s; ref.append(d)
        decode = [dict(scale=s, float32=True) for s in scales]
    else:
        native, ref, decode = fields, fields, None
    be = np.dtype(dtype).newbyteorder(">")
    file_buf = np.empty((nt, 5, nlev, nlat, nlon), dtype=be)           # [record][variable][level][lat][lon]
    for f in range(5):
        file_buf[:, f] = native[f]
    raws = [file_buf[:, f] for f in range(5)]
    assert not raws[0].flags.c_contiguous
    ident = (np.arange(nlon), np.arange(nlat), np.arange(nlev))
    with H.make_engine(P, fdt, [1.0] * 5) as eng:
        a = eng.run_host(ref, steps)
        b = eng.run_host_raw(raws, *ident, np.arange(nt), steps, decode=decode)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_raw_records_in_staged_chunks():
    """Several staged chunks (small windows): the two raw staging buffers and the ingest stream are reused
    across chunks and calls; T halo slots come device-to-device from the previous chunk."""
    rng = np.random.default_rng(15)
    P, fields = _dataset(40, 17, 6, 13, np.float32)
    steps = H.fixed_steps(P, P.lon[1], P.lon[37], P.lat[1], P.lat[15])
    raws, lon_map, lat_map, lev_map = _raw_layout(rng, fields, extra_levels=0)
    slot_bytes = fields[0][0].nbytes
    with H.make_engine(P, np.float32, [1.0] * 5) as eng:
        a = eng.run_host(fields, steps)
    with H.make_engine(P, np.float32, [1.0] * 5, max_steps=3) as eng:
        for _ in range(2):
            b = eng.run_host_raw(raws, lon_map, lat_map, lev_map, np.arange(13), steps)
            assert eng.last_transfer()[0] == 5 * 13 * slot_bytes
            for x, y in zip(a, b):
                assert np.array_equal(x, y)
        c = eng.run_host(fields, steps)                    # the two host paths share the staging buffers
        assert np.array_equal(a[0], c[0])
    with H.make_engine(P, np.float32, [1.0] * 5, host_stage_bytes=2 * 5 * 4 * slot_bytes) as eng:
        d = eng.run_host_raw(raws, lon_map, lat_map, lev_map, np.arange(13), steps[::-1].copy())   # steps in reverse order
        assert np.array_equal(d[0], a[0][::-1]) and np.array_equal(d[1], a[1][::-1])


def test_raw_errors():
    P, fields = _dataset(24, 11, 4, 3, np.float32)
    steps = H.fixed_steps(P, P.lon[1], P.lon[20], P.lat[1], P.lat[9])
    ident = (np.arange(24), np.arange(11), np.arange(4))
    with H.make_engine(P, np.float32, [1.0] * 5) as eng:
        with pytest.raises(IndexError):
            eng.run_host_raw(fields, np.arange(24) + 1, ident[1], ident[2], np.arange(3), steps)
        with pytest.raises(IndexError):
            eng.run_host_raw(fields, *ident, [0, 1, 3], steps)
        with pytest.raises(ValueError):
            eng.run_host_raw([f.astype(np.float64) for f in fields], *ident, np.arange(3), steps)
        with pytest.raises(ValueError):
            eng.run_host_raw(fields, ident[0][:-1], ident[1], ident[2], np.arange(3), steps)
        t, _, _ = eng.run_host_raw(fields, *ident, np.arange(3), steps)
        assert np.array_equal(t, eng.run_host(fields, steps)[0])
