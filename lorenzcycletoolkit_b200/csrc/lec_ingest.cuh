// File layout -> engine layout on the device (reference: get_data / process_data,
// src/utils/preprocessing.py:35-147 and :149-371).  The reference decodes packed variables and then
// re-sorts the whole dataset on the host, one full copy per transform (longitude wrap + sort, latitude
// sort, level sort, drop of the levels above 10 hPa, track-time selection, domain crop).  Here the RAW
// records cross PCIe as stored (int16 packed, float32 or float64) and one pass of this kernel writes the
// engine's [level ascending][lat ascending][lon ascending] slot, decoding on the way:
//   packed:  x = double(raw) * scale_factor  (+ add_offset), rounded to float32 first when the decoded
//            variable is float32 (scale only) -- the arithmetic of xarray's CFScaleOffsetCoder / numpy
//            in-place ops with float64 attribute scalars, with explicit roundings (no FMA)
//   _FillValue / missing_value -> NaN
// The index maps (canonical -> raw) come from the host, which derives them from the coordinates alone.
#pragma once
#include "lec_common.cuh"

namespace lec {

struct IngestParams {
  const void* src;            // raw sub-volume [nk_raw][nj_raw][rlon] (levels kr_lo.., rows jr_lo.. of one record)
  void* dst;                  // engine slot [nlev][nlat][nlon]
  const int* lon_map; const int* lat_map; const int* lev_map;     // device copies, canonical -> raw
  int nlon, nlat, nlev;       // engine grid
  int pitch;                  // row length of the engine slot (>= nlon; the pad columns are never written)
  int rlon, nj_raw, jr_lo, kr_lo;
  double scale, offset, fill0, fill1;
  int use_scale, use_offset, round32, nfill, big_endian;
};

constexpr int kIngestThreads = 256;

// raw element <-> its bit pattern, byte-swapped when the file is big-endian
template <typename RT> struct IngestBits;
template <> struct IngestBits<short> {
  using U = unsigned short;
  static __device__ __forceinline__ short get(U u, int swap) { return swap ? short(U((u >> 8) | (u << 8))) : short(u); }
};
template <> struct IngestBits<float> {
  using U = unsigned;
  static __device__ __forceinline__ float get(U u, int swap) { return __uint_as_float(swap ? __byte_perm(u, 0, 0x0123) : u); }
};
template <> struct IngestBits<double> {
  using U = unsigned long long;
  static __device__ __forceinline__ double get(U u, int swap) {
    if (!swap) return __longlong_as_double((long long)u);
    const unsigned lo = __byte_perm(unsigned(u >> 32), 0, 0x0123), hi = __byte_perm(unsigned(u), 0, 0x0123);
    return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
  }
};

// Four engine columns per thread: one 8/16/32-byte load when the four raw columns are consecutive and
// aligned (the usual case: a crop, a wrap or the identity), element loads otherwise; 16-byte stores.
template <typename RT, typename FT>
__global__ void __launch_bounds__(kIngestThreads) lec_ingest_kernel(const IngestParams p) {
  using U = typename IngestBits<RT>::U;
  const int i4 = (blockIdx.x * kIngestThreads + threadIdx.x) * 4;
  if (i4 >= p.nlon) return;
  const int j = blockIdx.y, k = blockIdx.z;
  const int rj = __ldg(p.lat_map + j) - p.jr_lo, rk = __ldg(p.lev_map + k) - p.kr_lo;
  const U* __restrict__ srow = static_cast<const U*>(p.src) + ((long long)rk * p.nj_raw + rj) * p.rlon;
  FT* __restrict__ drow = static_cast<FT*>(p.dst) + ((long long)k * p.nlat + j) * p.pitch + i4;
  const int n = min(4, p.nlon - i4);
  int r[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) r[e] = __ldg(p.lon_map + min(i4 + e, p.nlon - 1));
  U u[4];
  const bool run = n == 4 && r[1] == r[0] + 1 && r[2] == r[0] + 2 && r[3] == r[0] + 3 &&
                   (reinterpret_cast<unsigned long long>(srow + r[0]) % (4 * sizeof(U)) == 0);
  if (run) {
    if constexpr (sizeof(U) == 2) {
      const uint2 t = *reinterpret_cast<const uint2*>(srow + r[0]);
      u[0] = U(t.x & 0xffffu); u[1] = U(t.x >> 16); u[2] = U(t.y & 0xffffu); u[3] = U(t.y >> 16);
    } else if constexpr (sizeof(U) == 4) {
      const uint4 t = *reinterpret_cast<const uint4*>(srow + r[0]);
      u[0] = t.x; u[1] = t.y; u[2] = t.z; u[3] = t.w;
    } else {
      const ulonglong2 t0 = *reinterpret_cast<const ulonglong2*>(srow + r[0]);
      const ulonglong2 t1 = *reinterpret_cast<const ulonglong2*>(srow + r[0] + 2);
      u[0] = t0.x; u[1] = t0.y; u[2] = t1.x; u[3] = t1.y;
    }
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) u[e] = srow[r[e]];
  }
  FT out[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const RT raw = IngestBits<RT>::get(u[e], p.big_endian);
    double x = double(raw);
    if (p.use_scale) x = __dmul_rn(x, p.scale);
    if (p.use_offset) x = __dadd_rn(x, p.offset);
    if (p.round32) x = double(float(x));
    if ((p.nfill > 0 && raw == RT(p.fill0)) || (p.nfill > 1 && raw == RT(p.fill1)))
      x = __longlong_as_double(0x7ff8000000000000LL);
    out[e] = FT(x);
  }
  if (n == 4) {        // rows are padded to whole 16-byte chunks and i4 is a multiple of 4
    if constexpr (sizeof(FT) == 4) {
      *reinterpret_cast<float4*>(drow) = make_float4(out[0], out[1], out[2], out[3]);
    } else {
      *reinterpret_cast<double2*>(drow) = make_double2(out[0], out[1]);
      *reinterpret_cast<double2*>(drow + 2) = make_double2(out[2], out[3]);
    }
  } else {
    for (int e = 0; e < n; ++e) drow[e] = out[e];
  }
}

}  // namespace lec
