// Two-wide arithmetic for the row kernels.
//
// Blackwell (sm_100) issues packed fp32 FFMA2 / FADD2 / FMUL2 (PTX add/sub/mul/fma .f32x2): one
// warp instruction works on two fp32 values per lane, with a scalar operand broadcast for free.
// The row kernels are bound by instruction issue, not by the FMA pipe or DRAM, so every
// per-point operation is done on a PAIR of adjacent longitudes.  Pair<double> is the same
// interface with two scalar fp64 operations (fp64 fields / LEC_MATH_F64).
#pragma once
#include "lec_common.cuh"

namespace lec {

template <typename CT> struct Pair;

#ifndef LEC_SCALAR_PAIR
template <> struct Pair<float> {
  unsigned long long v;
  __device__ __forceinline__ static Pair make(float lo, float hi) {
    Pair p; asm("mov.b64 %0, {%1, %2};" : "=l"(p.v) : "f"(lo), "f"(hi)); return p;
  }
  __device__ __forceinline__ static Pair bcast(float s) { return make(s, s); }
  __device__ __forceinline__ float lo() const { return __uint_as_float((unsigned)(v & 0xffffffffull)); }
  __device__ __forceinline__ float hi() const { return __uint_as_float((unsigned)(v >> 32)); }
};
__device__ __forceinline__ Pair<float> operator+(Pair<float> a, Pair<float> b) {
  Pair<float> d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v)); return d;
}
__device__ __forceinline__ Pair<float> operator-(Pair<float> a, Pair<float> b) {
  Pair<float> d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v)); return d;
}
__device__ __forceinline__ Pair<float> operator*(Pair<float> a, Pair<float> b) {
  Pair<float> d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v)); return d;
}
__device__ __forceinline__ Pair<float> pfma(Pair<float> a, Pair<float> b, Pair<float> c) {
  Pair<float> d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return d;
}

#else   // A/B switch: plain scalar fp32 pairs (FFMA instead of FFMA2)
template <> struct Pair<float> {
  float l, h;
  __device__ __forceinline__ static Pair make(float lo, float hi) { Pair p; p.l = lo; p.h = hi; return p; }
  __device__ __forceinline__ static Pair bcast(float s) { return make(s, s); }
  __device__ __forceinline__ float lo() const { return l; }
  __device__ __forceinline__ float hi() const { return h; }
};
__device__ __forceinline__ Pair<float> operator+(Pair<float> a, Pair<float> b) { return Pair<float>::make(a.l + b.l, a.h + b.h); }
__device__ __forceinline__ Pair<float> operator-(Pair<float> a, Pair<float> b) { return Pair<float>::make(a.l - b.l, a.h - b.h); }
__device__ __forceinline__ Pair<float> operator*(Pair<float> a, Pair<float> b) { return Pair<float>::make(a.l * b.l, a.h * b.h); }
__device__ __forceinline__ Pair<float> pfma(Pair<float> a, Pair<float> b, Pair<float> c) {
  return Pair<float>::make(fmaf(a.l, b.l, c.l), fmaf(a.h, b.h, c.h));
}
#endif

template <> struct Pair<double> {
  double l, h;
  __device__ __forceinline__ static Pair make(double lo, double hi) { Pair p; p.l = lo; p.h = hi; return p; }
  __device__ __forceinline__ static Pair bcast(double s) { return make(s, s); }
  __device__ __forceinline__ double lo() const { return l; }
  __device__ __forceinline__ double hi() const { return h; }
};
__device__ __forceinline__ Pair<double> operator+(Pair<double> a, Pair<double> b) { return Pair<double>::make(a.l + b.l, a.h + b.h); }
__device__ __forceinline__ Pair<double> operator-(Pair<double> a, Pair<double> b) { return Pair<double>::make(a.l - b.l, a.h - b.h); }
__device__ __forceinline__ Pair<double> operator*(Pair<double> a, Pair<double> b) { return Pair<double>::make(a.l * b.l, a.h * b.h); }
__device__ __forceinline__ Pair<double> pfma(Pair<double> a, Pair<double> b, Pair<double> c) {
  return Pair<double>::make(fma(a.l, b.l, c.l), fma(a.h, b.h, c.h));
}

// Per-row constants of the pointwise Q, every factor folded (rounded once to CT).
template <typename CT>
struct RowCoef {
  CT ct_m, ct_p, ct_s;     // cp * sT * time stencil
  CT cy_m, cy_p;           // cp * sT * sV * lat stencil / dy
  CT s_m, s_p, s_s;        // -cp * sT * sW * static-stability stencil
  CT fx;                   // cp * sT * sU / cos(lat): multiplies the lon stencil
  CT shT, shU, shV, shW, shF;   // shifts (raw first-in-box values of the row)
};

// The 22 running moments of one row, two lanes wide.
template <typename CT>
struct RowAcc {
  using P = Pair<CT>;
  P S[R_NSUM];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int n = 0; n < R_NSUM; ++n) S[n] = P::bcast(CT(0));
  }
  // W* = weight x value (weight already applied), plain = value
  __device__ __forceinline__ void add(P Wa, P Wb, P Wc, P Ww, P Wf, P Wq, P a, P b, P c, P w, P f, P q) {
    S[R_A] = S[R_A] + Wa; S[R_B] = S[R_B] + Wb; S[R_C] = S[R_C] + Wc; S[R_W] = S[R_W] + Ww;
    S[R_F] = S[R_F] + Wf; S[R_Q] = S[R_Q] + Wq;
    const P pbb = Wb * b, pcc = Wc * c, pca = Wc * a, pwa = Ww * a;
    S[R_AA] = pfma(Wa, a, S[R_AA]); S[R_BB] = S[R_BB] + pbb; S[R_CC] = S[R_CC] + pcc;
    S[R_BC] = pfma(Wb, c, S[R_BC]); S[R_CA] = S[R_CA] + pca; S[R_WA] = S[R_WA] + pwa;
    S[R_WB] = pfma(Ww, b, S[R_WB]); S[R_WC] = pfma(Ww, c, S[R_WC]); S[R_WF] = pfma(Ww, f, S[R_WF]);
    S[R_QA] = pfma(Wa, q, S[R_QA]);
    S[R_CAA] = pfma(pca, a, S[R_CAA]); S[R_WAA] = pfma(pwa, a, S[R_WAA]); S[R_BBC] = pfma(pbb, c, S[R_BBC]);
    S[R_CCC] = pfma(pcc, c, S[R_CCC]); S[R_BBW] = pfma(pbb, w, S[R_BBW]); S[R_CCW] = pfma(pcc, w, S[R_CCW]);
  }
  __device__ __forceinline__ void to_double(double (&Sd)[R_NSUM]) const {
#pragma unroll
    for (int n = 0; n < R_NSUM; ++n) Sd[n] = double(S[n].lo()) + double(S[n].hi());
  }
};

// Interior pair of adjacent longitudes: every value is a Pair (lo = column i, hi = column i+1).
// tl/tr = temperature of the columns to the west / east of each lane of the pair.
template <typename CT, bool WEIGHTED>
__device__ __forceinline__ void interior_pair(RowAcc<CT>& acc, const RowCoef<CT>& rc, Pair<CT> tc, Pair<CT> tl,
                                              Pair<CT> tr, Pair<CT> tm, Pair<CT> tp, Pair<CT> tkm, Pair<CT> tkp,
                                              Pair<CT> tjm, Pair<CT> tjp, Pair<CT> u, Pair<CT> v, Pair<CT> om,
                                              Pair<CT> ph, Pair<CT> ca, Pair<CT> cc, Pair<CT> wg) {
  using P = Pair<CT>;
  const P dtdt = pfma(P::bcast(rc.ct_m), tm - tc, pfma(P::bcast(rc.ct_p), tp - tc, P::bcast(rc.ct_s) * tc));
  const P dTx = pfma(ca, tl - tc, cc * (tr - tc));
  const P dTy = pfma(P::bcast(rc.cy_m), tjm - tc, P::bcast(rc.cy_p) * (tjp - tc));
  const P Ss = pfma(P::bcast(rc.s_m), tkm - tc, pfma(P::bcast(rc.s_p), tkp - tc, P::bcast(rc.s_s) * tc));
  const P q = pfma(u, dTx, pfma(v, dTy, pfma(om, Ss, dtdt)));
  const P a = tc - P::bcast(rc.shT), b = u - P::bcast(rc.shU), c = v - P::bcast(rc.shV),
          w = om - P::bcast(rc.shW), f = ph - P::bcast(rc.shF);
  if (WEIGHTED) acc.add(wg * a, wg * b, wg * c, wg * w, wg * f, wg * q, a, b, c, w, f, q);
  else acc.add(a, b, c, w, f, q, a, b, c, w, f, q);
}

// One column with explicit (edge-aware) stencil coefficients and weight; returns the values
// the caller packs into pairs.  `in` = column inside the box (masked by select, so NaNs outside
// the box cannot leak).
template <typename CT>
struct PointVals { CT wg, a, b, c, w, f, q; };

template <typename CT>
__device__ __forceinline__ PointVals<CT> edge_point(const RowCoef<CT>& rc, bool in, CT wg, CT ca, CT cc, CT tc, CT tl,
                                                    CT tr, CT tm, CT tp, CT tkm, CT tkp, CT tjm, CT tjp, CT u, CT v,
                                                    CT om, CT ph) {
  const CT dtdt = rc.ct_m * (tm - tc) + rc.ct_p * (tp - tc) + rc.ct_s * tc;
  const CT dTx = ca * (tl - tc) + cc * (tr - tc);
  const CT dTy = rc.cy_m * (tjm - tc) + rc.cy_p * (tjp - tc);
  const CT Ss = rc.s_m * (tkm - tc) + rc.s_p * (tkp - tc) + rc.s_s * tc;
  PointVals<CT> r;
  r.q = dtdt + u * dTx + v * dTy + om * Ss;
  r.a = tc - rc.shT; r.b = u - rc.shU; r.c = v - rc.shV; r.w = om - rc.shW; r.f = ph - rc.shF;
  r.wg = wg;
  if (!in) { r.wg = CT(0); r.a = r.b = r.c = r.w = r.f = r.q = CT(0); }
  return r;
}

template <typename CT>
__device__ __forceinline__ void add_points(RowAcc<CT>& acc, const PointVals<CT>& x, const PointVals<CT>& y) {
  using P = Pair<CT>;
  const P wg = P::make(x.wg, y.wg), a = P::make(x.a, y.a), b = P::make(x.b, y.b), c = P::make(x.c, y.c),
          w = P::make(x.w, y.w), f = P::make(x.f, y.f), q = P::make(x.q, y.q);
  acc.add(wg * a, wg * b, wg * c, wg * w, wg * f, wg * q, a, b, c, w, f, q);
}

template <typename CT>
__device__ __forceinline__ PointVals<CT> zero_point() {
  PointVals<CT> r; r.wg = r.a = r.b = r.c = r.w = r.f = r.q = CT(0); return r;
}

}  // namespace lec
