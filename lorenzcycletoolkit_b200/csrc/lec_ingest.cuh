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

// one raw element, byte-swapped when the file is big-endian
template <typename RT> __device__ __forceinline__ RT ingest_load(const RT* q, int swap);
template <> __device__ __forceinline__ short ingest_load<short>(const short* q, int swap) {
  const unsigned short u = *reinterpret_cast<const unsigned short*>(q);
  return swap ? short((u >> 8) | (u << 8)) : short(u);
}
template <> __device__ __forceinline__ float ingest_load<float>(const float* q, int swap) {
  const unsigned u = *reinterpret_cast<const unsigned*>(q);
  return __uint_as_float(swap ? __byte_perm(u, 0, 0x0123) : u);
}
template <> __device__ __forceinline__ double ingest_load<double>(const double* q, int swap) {
  const unsigned long long u = *reinterpret_cast<const unsigned long long*>(q);
  if (!swap) return __longlong_as_double((long long)u);
  const unsigned lo = __byte_perm(unsigned(u >> 32), 0, 0x0123), hi = __byte_perm(unsigned(u), 0, 0x0123);
  return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}

template <typename RT, typename FT>
__global__ void __launch_bounds__(kIngestThreads) lec_ingest_kernel(const IngestParams p) {
  const int i = blockIdx.x * kIngestThreads + threadIdx.x;
  if (i >= p.nlon) return;
  const int j = blockIdx.y, k = blockIdx.z;
  const int ri = __ldg(p.lon_map + i), rj = __ldg(p.lat_map + j) - p.jr_lo, rk = __ldg(p.lev_map + k) - p.kr_lo;
  const RT raw = ingest_load<RT>(static_cast<const RT*>(p.src) + ((long long)rk * p.nj_raw + rj) * p.rlon + ri, p.big_endian);
  double x = double(raw);
  if (p.use_scale) x = __dmul_rn(x, p.scale);
  if (p.use_offset) x = __dadd_rn(x, p.offset);
  if (p.round32) x = double(float(x));
  if ((p.nfill > 0 && raw == RT(p.fill0)) || (p.nfill > 1 && raw == RT(p.fill1)))
    x = __longlong_as_double(0x7ff8000000000000LL);
  static_cast<FT*>(p.dst)[((long long)k * p.nlat + j) * p.pitch + i] = FT(x);
}

}  // namespace lec
