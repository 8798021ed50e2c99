// 850-hPa track diagnostics of the moving framework (reference: src/frameworks/lec_moving_framework.py
// :650-663 wind speed + relative vorticity over the pre-sliced domain, :269-417 get_position, and
// src/utils/tools.py:95-128 find_extremum_coordinates): per time step, over the label-sliced box, the
// extrema of zeta, geopotential height and wind speed and WHERE they are.
//
// Everything is fp64 with explicitly rounded operations in numpy's evaluation order (no FMA contraction),
// so zeta has the bits of the numpy restatement and the arg-reductions select the same grid point:
//   zeta = dv/dx - du/dy + u tan(lat) / a      (spherical form; MetPy's geodesic spacing is not restated)
//   d/dx, d/dy = np.gradient over the DOMAIN axes (one-sided at the domain edges, not at the box edges).
// One CTA per time step; the box of one level is at most a few 10^4 points.
#pragma once
#include "lec_common.cuh"

namespace lec {

struct DiagAxis {          // np.gradient(f, x) along one axis
  const double* a;         // non-uniform interior coefficients (nullptr on a uniform axis)
  const double* b;
  const double* c;
  double two_dx;           // uniform axis: 2 * dx
  double dx_first, dx_last;
  int n;
};

struct DiagStepDev { int slot, i0, i1, j0, j1; };

struct DiagParams {
  const void* u; const void* v; const void* z;      // [slot][lat][lon] planes of the 850-hPa level
  DiagAxis ax, ay;
  const double* coslat; const double* tanlat;
  double su, sv, sz, zdiv;                            // unit factors; hgt = (z * sz) / zdiv
  const DiagStepDev* steps;
  double* out_val;                                     // [nsteps][4]: zeta min, zeta max, hgt min, wind max (NaNs skipped)
  int* out_idx;                                        // [nsteps][4]: numpy argmin/argmax of the box, row-major flat index
  int nlon, nlat;
};

__device__ __forceinline__ double diag_grad(const DiagAxis& A, int i, double fm, double f0, double fp) {
  if (i == 0) return __ddiv_rn(__dsub_rn(fp, f0), A.dx_first);
  if (i == A.n - 1) return __ddiv_rn(__dsub_rn(f0, fm), A.dx_last);
  if (A.a == nullptr) return __ddiv_rn(__dsub_rn(fp, fm), A.two_dx);
  return __dadd_rn(__dadd_rn(__dmul_rn(A.a[i], fm), __dmul_rn(A.b[i], f0)), __dmul_rn(A.c[i], fp));
}

// running extremum with numpy semantics: first occurrence wins ties, the first NaN wins argmin/argmax,
// the VALUE skips NaNs (nanmin / xarray's skipna)
struct DiagExt {
  double val; int idx; int nan_idx;
  __device__ __forceinline__ void init() { val = 0.0; idx = 0x7fffffff; nan_idx = 0x7fffffff; }
  template <bool MIN>
  __device__ __forceinline__ void take(double x, int i) {
    if (x != x) { nan_idx = min(nan_idx, i); return; }
    const bool better = (idx == 0x7fffffff) || (MIN ? x < val : x > val) || (x == val && i < idx);
    if (better) { val = x; idx = i; }
  }
  template <bool MIN>
  __device__ __forceinline__ void merge(double x, int i, int ni) {
    nan_idx = min(nan_idx, ni);
    if (i == 0x7fffffff) return;
    const bool better = (idx == 0x7fffffff) || (MIN ? x < val : x > val) || (x == val && i < idx);
    if (better) { val = x; idx = i; }
  }
};

template <bool MIN>
__device__ __forceinline__ void diag_warp_reduce(DiagExt& e) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const double x = __shfl_xor_sync(0xffffffffu, e.val, o);
    const int i = __shfl_xor_sync(0xffffffffu, e.idx, o);
    const int ni = __shfl_xor_sync(0xffffffffu, e.nan_idx, o);
    e.merge<MIN>(x, i, ni);
  }
}

constexpr int kDiagThreads = 256;

template <typename FT>
__global__ void __launch_bounds__(kDiagThreads) lec_diag850_kernel(const DiagParams p) {
  const DiagStepDev st = p.steps[blockIdx.x];
  const int nx = st.i1 - st.i0 + 1, ny = st.j1 - st.j0 + 1;
  const long long plane = (long long)p.nlat * p.nlon;
  const FT* __restrict__ U = static_cast<const FT*>(p.u) + st.slot * plane;
  const FT* __restrict__ V = static_cast<const FT*>(p.v) + st.slot * plane;
  const FT* __restrict__ Z = static_cast<const FT*>(p.z) + st.slot * plane;

  DiagExt zmin, zmax, hmin, wmax;
  zmin.init(); zmax.init(); hmin.init(); wmax.init();
  for (int q = threadIdx.x; q < nx * ny; q += kDiagThreads) {
    const int jr = q / nx, ir = q - jr * nx;
    const int j = st.j0 + jr, i = st.i0 + ir;
    const long long o = (long long)j * p.nlon + i;
    const int im = max(i - 1, 0), ip = min(i + 1, p.nlon - 1), jm = max(j - 1, 0), jp = min(j + 1, p.nlat - 1);
    const double u0 = __dmul_rn(double(U[o]), p.su), v0 = __dmul_rn(double(V[o]), p.sv);
    const double vW = __dmul_rn(double(V[(long long)j * p.nlon + im]), p.sv);
    const double vE = __dmul_rn(double(V[(long long)j * p.nlon + ip]), p.sv);
    const double uS = __dmul_rn(double(U[(long long)jm * p.nlon + i]), p.su);
    const double uN = __dmul_rn(double(U[(long long)jp * p.nlon + i]), p.su);
    const double dvdx = __ddiv_rn(diag_grad(p.ax, i, vW, v0, vE), __dmul_rn(kRe, p.coslat[j]));
    const double dudy = __ddiv_rn(diag_grad(p.ay, j, uS, u0, uN), kRe);
    const double zeta = __dadd_rn(__dsub_rn(dvdx, dudy), __ddiv_rn(__dmul_rn(u0, p.tanlat[j]), kRe));
    const double wspd = __dsqrt_rn(__dadd_rn(__dmul_rn(u0, u0), __dmul_rn(v0, v0)));
    const double hgt = __ddiv_rn(__dmul_rn(double(Z[o]), p.sz), p.zdiv);
    zmin.take<true>(zeta, q); zmax.take<false>(zeta, q); hmin.take<true>(hgt, q); wmax.take<false>(wspd, q);
  }
  diag_warp_reduce<true>(zmin); diag_warp_reduce<false>(zmax); diag_warp_reduce<true>(hmin); diag_warp_reduce<false>(wmax);

  __shared__ double s_val[kDiagThreads / 32][4];
  __shared__ int s_idx[kDiagThreads / 32][4], s_nan[kDiagThreads / 32][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    s_val[warp][0] = zmin.val; s_idx[warp][0] = zmin.idx; s_nan[warp][0] = zmin.nan_idx;
    s_val[warp][1] = zmax.val; s_idx[warp][1] = zmax.idx; s_nan[warp][1] = zmax.nan_idx;
    s_val[warp][2] = hmin.val; s_idx[warp][2] = hmin.idx; s_nan[warp][2] = hmin.nan_idx;
    s_val[warp][3] = wmax.val; s_idx[warp][3] = wmax.idx; s_nan[warp][3] = wmax.nan_idx;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    const int m = threadIdx.x;
    DiagExt e; e.init();
    for (int w = 0; w < kDiagThreads / 32; ++w) {
      if (m == 0 || m == 2) e.merge<true>(s_val[w][m], s_idx[w][m], s_nan[w][m]);
      else e.merge<false>(s_val[w][m], s_idx[w][m], s_nan[w][m]);
    }
    p.out_val[blockIdx.x * 4 + m] = (e.idx == 0x7fffffff) ? __longlong_as_double(0x7ff8000000000000LL) : e.val;
    p.out_idx[blockIdx.x * 4 + m] = (e.nan_idx != 0x7fffffff) ? e.nan_idx : e.idx;
  }
}

}  // namespace lec
