// Kernel A of the LEC engine: the only HBM-heavy kernel.
//
// One warp owns one box row (step, level, lat) and sweeps it along longitude in
// 128-bit chunks.  In a SINGLE pass over T, u, v, omega, Phi it evaluates
//   * the diabatic-heating residual Q pointwise (thermodynamics.py:76-124): centred
//     differences in lon/lat/p/t on T written in difference form a(T[-1]-T)+c(T[+1]-T)
//     so fp32 arithmetic does not cancel,
//   * the 22 shifted zonal trapezoid moments of lec_common.cuh (the zonal means and every
//     eddy product of box_data.py:157-231 / src/analysis/*.py follow from them exactly),
//   * the raw west/east edge values needed by boundary_terms.py.
// Shifted single-pass moments (shift = first box value of the row) replace the
// reference's "mean first, anomalies second" two passes: the central moments are
// recovered in fp64 by the finalize kernel, so no rounded mean ever biases the anomalies.
// Lanes reduce with a 23-shuffle fp64 halving butterfly; the sum order is fixed, so
// results are bit-reproducible across launches, shards and GPUs.
#pragma once
#include "lec_common.cuh"

namespace lec {

constexpr int kRowsPerCta = 8;          // warps per CTA, one row each (adjacent rows share L1 lines)
constexpr int kRowThreads = kRowsPerCta * 32;

struct RowParams {
  const void* field[5];
  GridDev g;
  const StepDev* steps;
  double* rec;
  int nsteps;
  int max_ny;            // record rows reserved per (step, level)
  int tiles_per_band;    // CTA row-tiles per latitude band
  int nbands;
  long long slot_stride; // elements per slot = nlev*nlat*nlon
};

template <typename FT, int VEC> struct VecLoad;
template <> struct VecLoad<float, 4> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  static __device__ __forceinline__ void ld_stream(const float* p, float (&v)[4]) {
    float4 t = __ldcs(reinterpret_cast<const float4*>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
};
template <> struct VecLoad<double, 2> {
  static __device__ __forceinline__ void ld(const double* p, double (&v)[2]) {
    double2 t = __ldg(reinterpret_cast<const double2*>(p)); v[0] = t.x; v[1] = t.y; }
  static __device__ __forceinline__ void ld_stream(const double* p, double (&v)[2]) {
    double2 t = __ldcs(reinterpret_cast<const double2*>(p)); v[0] = t.x; v[1] = t.y; }
};
template <typename FT> struct VecLoad<FT, 1> {
  static __device__ __forceinline__ void ld(const FT* p, FT (&v)[1]) { v[0] = __ldg(p); }
  static __device__ __forceinline__ void ld_stream(const FT* p, FT (&v)[1]) { v[0] = __ldcs(p); }
};

__device__ __forceinline__ double shfl_xor_f64(double v, int m) {
  return __shfl_xor_sync(0xffffffffu, v, m);
}

// Halving butterfly: N values per lane -> lane L ends with the full-warp sum of value
// bitrev5(L) (zero for indices >= N).  23 shuffles for N = 22 instead of 110.
template <int N>
__device__ __forceinline__ double butterfly_reduce(double (&v)[N], int lane) {
  static_assert(N <= 32, "at most one value per lane");
  double a[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) a[i] = (i < N) ? v[i] : 0.0;
  int n = 32;
#pragma unroll
  for (int bit = 16; bit >= 1; bit >>= 1) {
    n >>= 1;
    const bool hi = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      // statically skip pairs that are known to be all-zero padding
      const int live = (N + (32 / (2 * n)) - 1) / (32 / (2 * n));   // live values before this step
      if (2 * i < live) {
        const double keep = hi ? a[2 * i + 1] : a[2 * i];
        const double send = hi ? a[2 * i] : a[2 * i + 1];
        a[i] = keep + shfl_xor_f64(send, bit);
      } else {
        a[i] = 0.0;
      }
    }
  }
  return a[0];
}

__device__ __forceinline__ int bitrev5(int x) {
  return ((x & 1) << 4) | ((x & 2) << 2) | (x & 4) | ((x & 8) >> 2) | ((x & 16) >> 4);
}

template <typename FT, typename CT, int VEC, bool LON_TABLE>
__global__ void __launch_bounds__(kRowThreads)
lec_row_moments_kernel(const RowParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CTA order: band-major, then step, level, row-tile.  Sweeping time inside a latitude
  // band keeps T(t+1) (first touched as the time neighbour of step t) in L2 until it is
  // the centre of step t+1 and the t-1 neighbour of step t+2.
  long long id = blockIdx.x;
  const int jt = int(id % p.tiles_per_band); id /= p.tiles_per_band;
  const int k = int(id % p.g.nlev); id /= p.g.nlev;
  const int s = int(id % p.nsteps);
  const int band = int(id / p.nsteps);

  const StepDev* __restrict__ st = p.steps + s;
  const int i0 = st->i0, i1 = st->i1, j0 = st->j0, j1 = st->j1;
  const int jrel = (band * p.tiles_per_band + jt) * kRowsPerCta + warp;
  if (jrel > j1 - j0) return;                 // warp-uniform; no block barriers below
  const int j = j0 + jrel;
  const int nlon = p.g.nlon, nlat = p.g.nlat, nlev = p.g.nlev;

  const long long plane = (long long)nlat * nlon;
  const long long row_c = ((long long)st->slot * nlev + k) * plane + (long long)j * nlon;
  const FT* __restrict__ Tg = static_cast<const FT*>(p.field[0]);
  const FT* Tc_row = Tg + row_c;
  const FT* Tm_row = Tg + row_c + (long long)(st->slot_m - st->slot) * p.slot_stride;
  const FT* Tp_row = Tg + row_c + (long long)(st->slot_p - st->slot) * p.slot_stride;
  const FT* Tkm_row = (k > 0) ? Tc_row - plane : Tc_row;
  const FT* Tkp_row = (k < nlev - 1) ? Tc_row + plane : Tc_row;
  const FT* Tjm_row = (j > j0) ? Tc_row - nlon : Tc_row;
  const FT* Tjp_row = (j < j1) ? Tc_row + nlon : Tc_row;
  const FT* U_row = static_cast<const FT*>(p.field[1]) + row_c;
  const FT* V_row = static_cast<const FT*>(p.field[2]) + row_c;
  const FT* W_row = static_cast<const FT*>(p.field[3]) + row_c;
  const FT* F_row = static_cast<const FT*>(p.field[4]) + row_c;

  // row-level scalars (all folded on the host in fp64, rounded once to CT here)
  const double sT = p.g.scale[0];
  const CT scT = CT(sT), scU = CT(p.g.scale[1]), scV = CT(p.g.scale[2]),
           scW = CT(p.g.scale[3]), scF = CT(p.g.scale[4]);
  const CT ct_m = CT(st->ct_m * sT), ct_p = CT(st->ct_p * sT), ct_s = CT(st->ct_s * sT);
  const CT cy_m = CT(((j == j0) ? 0.0 : (j == j1) ? -st->cyN : p.g.cya[j]) * sT);
  const CT cy_p = CT(((j == j1) ? 0.0 : (j == j0) ? st->cyS : p.g.cyc[j]) * sT);
  const CT s_m = CT(p.g.sm[k] * sT), s_p = CT(p.g.sp[k] * sT), s_s = CT(p.g.ss[k] * sT);
  const CT inv_cos = CT(1.0 / p.g.coslat[j]);
  const CT cxW = CT(st->cxW * sT), cxE = CT(st->cxE * sT);
  const CT wW = CT(st->wW), wE = CT(st->wE);
  const CT wl_u = CT(p.g.wl_u), cxa_u = CT(p.g.cxa_u * sT), cxc_u = CT(p.g.cxc_u * sT);
  const CT cp = CT(kCp);

  // shifts: raw first-in-box values of the row (broadcast loads)
  const FT shT = __ldg(Tc_row + i0), shU = __ldg(U_row + i0), shV = __ldg(V_row + i0),
           shW = __ldg(W_row + i0), shF = __ldg(F_row + i0);

  CT S[R_NSUM];
#pragma unroll
  for (int n = 0; n < R_NSUM; ++n) S[n] = CT(0);

  double* __restrict__ rec = p.rec + (((long long)s * nlev + k) * p.max_ny + jrel) * LEC_NREC;

  const int c0 = i0 / VEC, c1 = i1 / VEC;
  const int niter = (c1 - c0 + 32) / 32;
  for (int it = 0; it < niter; ++it) {
    const int c_raw = c0 + it * 32 + lane;
    const bool lane_on = c_raw <= c1;
    const int c = lane_on ? c_raw : c1;     // clamp so every load stays inside the row
    const int col = c * VEC;

    FT Tc[VEC], Tm[VEC], Tp[VEC], Tkm[VEC], Tkp[VEC], Tjm[VEC], Tjp[VEC], U[VEC], V[VEC], W[VEC], F[VEC];
    VecLoad<FT, VEC>::ld(Tc_row + col, Tc);
    VecLoad<FT, VEC>::ld(Tm_row + col, Tm);
    VecLoad<FT, VEC>::ld(Tp_row + col, Tp);
    VecLoad<FT, VEC>::ld(Tkm_row + col, Tkm);
    VecLoad<FT, VEC>::ld(Tkp_row + col, Tkp);
    VecLoad<FT, VEC>::ld(Tjm_row + col, Tjm);
    VecLoad<FT, VEC>::ld(Tjp_row + col, Tjp);
    VecLoad<FT, VEC>::ld_stream(U_row + col, U);
    VecLoad<FT, VEC>::ld_stream(V_row + col, V);
    VecLoad<FT, VEC>::ld_stream(W_row + col, W);
    VecLoad<FT, VEC>::ld_stream(F_row + col, F);

    // lon neighbours of the chunk ends: adjacent lanes, or a scalar load at the warp ends
    FT Tl = __shfl_up_sync(0xffffffffu, Tc[VEC - 1], 1);
    FT Tr = __shfl_down_sync(0xffffffffu, Tc[0], 1);
    if (lane == 0) Tl = (col - 1 >= i0) ? __ldg(Tc_row + col - 1) : Tc[0];
    if (lane == 31 || c_raw >= c1) Tr = (col + VEC <= i1) ? __ldg(Tc_row + col + VEC) : Tc[VEC - 1];

    CT wl_t[VEC], cxa_t[VEC], cxc_t[VEC];
    if (LON_TABLE) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        wl_t[e] = CT(__ldg(p.g.wl + col + e));
        cxa_t[e] = CT(__ldg(p.g.cxa + col + e) * sT);
        cxc_t[e] = CT(__ldg(p.g.cxc + col + e) * sT);
      }
    }

#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const int i = col + e;
      const bool in = lane_on && i >= i0 && i <= i1;
      CT wgt = LON_TABLE ? wl_t[e] : wl_u;
      CT ca = LON_TABLE ? cxa_t[e] : cxa_u;
      CT cc = LON_TABLE ? cxc_t[e] : cxc_u;
      if (i == i0) { wgt = wW; ca = CT(0); cc = cxW; }
      if (i == i1) { wgt = wE; ca = -cxE; cc = CT(0); }
      const FT tl = (e > 0) ? Tc[e > 0 ? e - 1 : 0] : Tl;
      const FT tr = (e < VEC - 1) ? Tc[e < VEC - 1 ? e + 1 : 0] : Tr;
      const CT tc = CT(Tc[e]);
      // stencils in difference form: the differences of neighbouring temperatures are
      // exact in fp32 (Sterbenz), so the fp32 path does not cancel against T ~ 250 K
      const CT dtdt = ct_m * (CT(Tm[e]) - tc) + ct_p * (CT(Tp[e]) - tc) + ct_s * tc;
      const CT dTx = ca * (CT(tl) - tc) + cc * (CT(tr) - tc);
      const CT dTy = cy_m * (CT(Tjm[e]) - tc) + cy_p * (CT(Tjp[e]) - tc);
      const CT Ss = s_m * (CT(Tkm[e]) - tc) + s_p * (CT(Tkp[e]) - tc) + s_s * tc;
      const CT u = CT(U[e]) * scU, v = CT(V[e]) * scV, om = CT(W[e]) * scW;
      CT q = cp * (dtdt + u * dTx * inv_cos + v * dTy - Ss * om);
      CT a = (tc - CT(shT)) * scT, b = (CT(U[e]) - CT(shU)) * scU, cv = (CT(V[e]) - CT(shV)) * scV,
         w = (CT(W[e]) - CT(shW)) * scW, f = (CT(F[e]) - CT(shF)) * scF;
      if (!in) { wgt = CT(0); a = b = cv = w = f = q = CT(0); }

      const CT Wa = wgt * a, Wb = wgt * b, Wc = wgt * cv, Ww = wgt * w;
      S[R_A] += Wa; S[R_B] += Wb; S[R_C] += Wc; S[R_W] += Ww;
      S[R_F] += wgt * f; S[R_Q] += wgt * q;
      const CT pbb = Wb * b, pcc = Wc * cv, pca = Wc * a, pwa = Ww * a;
      S[R_AA] += Wa * a; S[R_BB] += pbb; S[R_CC] += pcc; S[R_BC] += Wb * cv;
      S[R_CA] += pca; S[R_WA] += pwa; S[R_WB] += Ww * b; S[R_WC] += Ww * cv;
      S[R_WF] += Ww * f; S[R_QA] += Wa * q;
      S[R_CAA] += pca * a; S[R_WAA] += pwa * a; S[R_BBC] += pbb * cv; S[R_CCC] += pcc * cv;
      S[R_BBW] += pbb * w; S[R_CCW] += pcc * w;

      if (in && i == i0) {
        rec[R_UW] = double(U[e]) * p.g.scale[1]; rec[R_VW] = double(V[e]) * p.g.scale[2];
        rec[R_TW] = double(Tc[e]) * sT;
      }
      if (in && i == i1) {
        rec[R_UE] = double(U[e]) * p.g.scale[1]; rec[R_VE] = double(V[e]) * p.g.scale[2];
        rec[R_TE] = double(Tc[e]) * sT;
      }
    }
  }

  double Sd[R_NSUM];
#pragma unroll
  for (int n = 0; n < R_NSUM; ++n) Sd[n] = double(S[n]);
  const double tot = butterfly_reduce<R_NSUM>(Sd, lane);
  const int idx = bitrev5(lane);
  if (idx < R_NSUM) rec[idx] = tot;
  if (lane == 1) {   // bitrev5(1) = 16 < 22 as well, any lane will do; spread the stores
    rec[R_SH_T] = double(shT) * sT;
    rec[R_SH_U] = double(shU) * p.g.scale[1];
    rec[R_SH_V] = double(shV) * p.g.scale[2];
    rec[R_SH_W] = double(shW) * p.g.scale[3];
    rec[R_SH_F] = double(shF) * p.g.scale[4];
  }
}

}  // namespace lec
