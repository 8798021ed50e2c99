// Kernel A, TMA-pipelined variant (the default when the rows are 16-byte aligned).
//
// Same arithmetic and the same row records as lec_row_moments_kernel, different data movement
// and work split:
//  * a persistent CTA per SM walks the (band, step, level, row-tile) list; for every 512-byte
//    longitude chunk of a tile of R rows ONE thread issues 9 TMA tensor loads
//        T centre tile with a +-1 row / +-1 vector halo  ((R+2) x 136 floats)
//        T(t-1), T(t+1), T(k-1), T(k+1), u, v, omega, Phi tiles (R x 128 floats each)
//    into a shared-memory ring guarded by full/empty mbarriers.  Bytes in flight no longer depend
//    on registers or occupancy, the j+-1 halo is shared by the rows of the tile, and TMA takes
//    arbitrary start columns, so boxes need no alignment handling (chunks start at i0).
//  * the kernel is bound by instruction issue / dependency latency, so it wants as many resident
//    warps as possible without register spills.  Each row is therefore served by TWO warps with
//    disjoint accumulator sets reading the same shared-memory tiles:
//        role A ("thermo"): the pointwise Q (all seven T stencil points, u, v, omega) and the
//                           sums S_a S_q S_aa S_qa                               (4 accumulators)
//        role B ("wind")  : u, v, omega, Phi, T and the other 18 sums            (18 accumulators)
//    16 compute warps per SM at <= 128 registers, two-wide packed fp32 math (lec_packed.cuh).
#pragma once
#include <cuda.h>

#include "lec_common.cuh"
#include "lec_packed.cuh"
#include "lec_row_moments.cuh"

namespace lec {

#ifndef LEC_TMA_ROWS
#define LEC_TMA_ROWS 8
#endif
#ifndef LEC_TMA_STAGES
#define LEC_TMA_STAGES 5
#endif
constexpr int kTmaRows = LEC_TMA_ROWS;            // rows per tile
constexpr int kTmaStages = LEC_TMA_STAGES;
constexpr int kTmaThreads = 2 * kTmaRows * 32;    // warps [0,R): role A, [R,2R): role B; thread 0 drives TMA
constexpr int kTmaCtasPerSm = 1;

struct alignas(64) TmaMaps {
  CUtensorMap t_halo;    // T, box (C + 2 VEC) x (R + 2)
  CUtensorMap t_plain;   // T, box C x R
  CUtensorMap u, v, w, f;
};

template <typename FT>
struct TmaGeom {
  static constexpr int VEC = 16 / sizeof(FT);
  static constexpr int C = 32 * VEC;                         // columns per chunk (512 B per row)
  static constexpr int HP = C + 2 * VEC;                     // halo tile pitch
  static constexpr int halo_bytes = (kTmaRows + 2) * HP * sizeof(FT);
  static constexpr int halo_bytes_pad = (halo_bytes + 127) / 128 * 128;
  static constexpr int tile_bytes = kTmaRows * C * sizeof(FT);
  static constexpr int stage_bytes = halo_bytes_pad + 8 * tile_bytes;
  static constexpr int tx_bytes = halo_bytes + 8 * tile_bytes;           // what the 9 loads deliver
  static constexpr int smem_bytes = kTmaStages * stage_bytes + 1024;     // + barriers
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void tma_load_4d(unsigned dst, const CUtensorMap* map, unsigned bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

template <typename FT, int VEC>
__device__ __forceinline__ void lds_vec(const FT* p, FT (&v)[VEC]) {
  if constexpr (sizeof(FT) == 4) { float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  else { double2 t = *reinterpret_cast<const double2*>(p); v[0] = t.x; v[1] = t.y; }
}

struct TileId { int band, s, k, jt; };
__device__ __forceinline__ TileId decode_tile(unsigned id, const RowParams& p) {
  TileId t;                                        // the host guarantees grid < 2^31: 32-bit divides
  const unsigned a = id / (unsigned)p.tiles_per_band;
  t.jt = int(id - a * (unsigned)p.tiles_per_band);
  const unsigned b = a / (unsigned)p.g.nlev;
  t.k = int(a - b * (unsigned)p.g.nlev);
  t.band = int(b / (unsigned)p.nsteps);
  t.s = int(b - (unsigned)t.band * (unsigned)p.nsteps);
  return t;
}

// per-column trapezoid weights and folded lon-stencil coefficients of one chunk (LONW == 1)
template <typename CT, int VEC>
__device__ __forceinline__ void load_lon_tables(const GridDev& g, int i0, int col, CT fx, double fxd, bool want_stencil,
                                                CT (&wl_t)[VEC], CT (&cxa_t)[VEC], CT (&cxc_t)[VEC]) {
  if (sizeof(CT) == 4 && (i0 % VEC) == 0 && col + VEC <= g.nlon) {
    if constexpr (sizeof(CT) == 4) {
      float w4[VEC];
      VecLoad<float, VEC>::ld(g.wl32 + col, w4);
#pragma unroll
      for (int e = 0; e < VEC; ++e) wl_t[e] = CT(w4[e]);
      if (want_stencil) {
        float a4[VEC], c4[VEC];
        VecLoad<float, VEC>::ld(g.cxa32 + col, a4);
        VecLoad<float, VEC>::ld(g.cxc32 + col, c4);
#pragma unroll
        for (int e = 0; e < VEC; ++e) { cxa_t[e] = fx * CT(a4[e]); cxc_t[e] = fx * CT(c4[e]); }
      }
    }
  } else {
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const int i = min(col + e, g.nlon - 1);
      if constexpr (sizeof(CT) == 4) {
        wl_t[e] = CT(__ldg(g.wl32 + i));
        if (want_stencil) { cxa_t[e] = fx * CT(__ldg(g.cxa32 + i)); cxc_t[e] = fx * CT(__ldg(g.cxc32 + i)); }
      } else {
        wl_t[e] = CT(__ldg(g.wl + i));
        if (want_stencil) { cxa_t[e] = CT(fxd * __ldg(g.cxa + i)); cxc_t[e] = CT(fxd * __ldg(g.cxc + i)); }
      }
    }
  }
}

template <typename FT, typename CT, int LONW>
__global__ void __launch_bounds__(kTmaThreads, 1)
lec_row_moments_tma_kernel(const __grid_constant__ TmaMaps maps, const RowParams p) {
  using G = TmaGeom<FT>;
  using P = Pair<CT>;
  constexpr int VEC = G::VEC, C = G::C, HP = G::HP, R = kTmaRows;
  extern __shared__ __align__(1024) unsigned char smem[];   // plain shared pointer: keeps LDS (not generic LD)
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + kTmaStages * G::stage_bytes);
  const unsigned full0 = smem_u32(bars), empty0 = smem_u32(bars + kTmaStages);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool roleA = warp < R;
  const int row = roleA ? warp : warp - R;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kTmaStages; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 2 * R); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int nlev = p.g.nlev;

  // ---- producer state (meaningful in the producer thread only): the (tile, chunk) sequence,
  // kTmaStages-1 chunks ahead of the consumers.  The producer is lane 0 of the first role-B warp
  // (role B has less arithmetic than role A, so the extra work does not sit on the critical path).
  // Tile coordinates are decoded once per tile; a chunk costs one wait, one expect_tx, 9 TMA issues.
  const bool producer = threadIdx.x == R * 32;
  unsigned p_id = blockIdx.x, p_n = 0;
  int p_ch = 0, p_nch = 0, p_i0 = 0, p_jt0 = 0, p_k = 0, p_km = 0, p_kp = 0, p_slot = 0, p_slot_m = 0, p_slot_p = 0;
  // L2 prefetch of the DRAM-sourced rows (u, v, omega, Phi and T of the next time slot) of a whole
  // tile, one tile ahead of the TMA loads: the loads then hit L2, so kTmaStages-1 chunks in flight
  // cover the (much shorter) L2 latency.
  auto p_prefetch = [&](unsigned id) {
    if (!(p.prefetch_mode & 1) || id >= (unsigned)p.grid) return;
    const TileId t = decode_tile(id, p);
    const StepDev* __restrict__ st = p.steps + t.s;
    const int jr = (t.band * p.tiles_per_band + t.jt) * R;
    const int nrow = min(R, st->j1 - st->j0 + 1 - jr);
    const long long plane = (long long)p.g.nlat * p.g.nlon;
    for (int f = 0; f < 5; ++f) {
      const int slot = (f == 0) ? st->slot_p : st->slot;
      const FT* base = static_cast<const FT*>(p.field[f]) + ((long long)slot * nlev + t.k) * plane;
      for (int r = 0; r < nrow; ++r)
        prefetch_l2_range(base + (long long)(st->j0 + jr + r) * p.g.nlon, (long long)st->i0 * sizeof(FT),
                          (long long)(st->i1 + 1) * sizeof(FT));
    }
  };
  auto p_seek = [&]() {      // move to the next tile that has rows inside its step's box and cache its coordinates
    while (p_id < (unsigned)p.grid) {
      const TileId t = decode_tile(p_id, p);
      const StepDev* __restrict__ st = p.steps + t.s;
      const int jr = (t.band * p.tiles_per_band + t.jt) * R;
      if (jr <= st->j1 - st->j0) {
        p_nch = (st->i1 - st->i0 + C) / C; p_i0 = st->i0; p_jt0 = st->j0 + jr;
        p_k = t.k; p_km = t.k > 0 ? t.k - 1 : t.k; p_kp = t.k < nlev - 1 ? t.k + 1 : t.k;
        p_slot = st->slot; p_slot_m = st->slot_m; p_slot_p = st->slot_p;
        p_prefetch(p_id + gridDim.x);
        return;
      }
      p_id += gridDim.x;
    }
    p_nch = 0;
  };
  auto p_issue = [&]() {     // issue the 9 loads of chunk (p_id, p_ch) into stage p_n % S, then advance
    const int stg = p_n % kTmaStages;
    mbar_wait(empty0 + 8 * stg, ((p_n / kTmaStages) & 1) ^ 1);
    const unsigned fb = full0 + 8 * stg;
    mbar_expect_tx(fb, G::tx_bytes);
    const unsigned base = smem_u32(smem + (size_t)stg * G::stage_bytes);
    const int col = p_i0 + p_ch * C;
    tma_load_4d(base, &maps.t_halo, fb, col - VEC, p_jt0 - 1, p_k, p_slot);
    unsigned d = base + G::halo_bytes_pad;
    tma_load_4d(d, &maps.t_plain, fb, col, p_jt0, p_k, p_slot_m); d += G::tile_bytes;
    tma_load_4d(d, &maps.t_plain, fb, col, p_jt0, p_k, p_slot_p); d += G::tile_bytes;
    tma_load_4d(d, &maps.t_plain, fb, col, p_jt0, p_km, p_slot); d += G::tile_bytes;
    tma_load_4d(d, &maps.t_plain, fb, col, p_jt0, p_kp, p_slot); d += G::tile_bytes;
    tma_load_4d(d, &maps.u, fb, col, p_jt0, p_k, p_slot); d += G::tile_bytes;
    tma_load_4d(d, &maps.v, fb, col, p_jt0, p_k, p_slot); d += G::tile_bytes;
    tma_load_4d(d, &maps.w, fb, col, p_jt0, p_k, p_slot); d += G::tile_bytes;
    tma_load_4d(d, &maps.f, fb, col, p_jt0, p_k, p_slot);
    ++p_n;
    if (++p_ch == p_nch) { p_ch = 0; p_id += gridDim.x; p_seek(); }
  };
  if (producer) {
    p_prefetch(p_id);
    p_seek();
    for (int i = 0; i < kTmaStages - 1 && p_nch > 0; ++i) p_issue();
  }

  unsigned n = 0;
  for (unsigned id = blockIdx.x; id < (unsigned)p.grid; id += gridDim.x) {
    const TileId t = decode_tile(id, p);
    const StepDev* __restrict__ st = p.steps + t.s;
    const int i0 = st->i0, i1 = st->i1, j0 = st->j0, j1 = st->j1;
    const int jrel0 = (t.band * p.tiles_per_band + t.jt) * R;
    if (jrel0 > j1 - j0) continue;
    const int jrel = jrel0 + row;
    const bool row_on = jrel <= j1 - j0;
    const int j = row_on ? j0 + jrel : j1;
    const int k = t.k;
    const int nch = (i1 - i0 + C) / C;
    double* __restrict__ rec = p.rec + (((long long)t.s * nlev + k) * p.max_ny + jrel) * LEC_NREC;
    const double wnorm = (LONW == 0) ? 1.0 / p.g.wl_u : 1.0;
    const CT wW = CT(st->wW * wnorm), wE = CT(st->wE * wnorm);
    const int lo = lane * VEC;

    if (roleA) {
      // =============================== role A: Q and the T/Q sums ===============================
      RowCoef<CT> rc;
      rc.ct_m = CT(st->ct_m); rc.ct_p = CT(st->ct_p); rc.ct_s = CT(st->ct_s);
      rc.cy_m = CT((j == j0) ? 0.0 : (j == j1) ? -st->cyN : p.g.cya[j]);
      rc.cy_p = CT((j == j1) ? 0.0 : (j == j0) ? st->cyS : p.g.cyc[j]);
      rc.s_m = CT(p.g.sm[k]); rc.s_p = CT(p.g.sp[k]); rc.s_s = CT(p.g.ss[k]);
      const double fxd = p.g.fxj[j];
      rc.fx = CT(fxd);
      rc.shT = rc.shU = rc.shV = rc.shW = rc.shF = CT(0);
      const CT cxa_u = CT(fxd * p.g.cxa_u), cxc_u = CT(fxd * p.g.cxc_u);
      const CT cxW = CT(fxd * st->cxW), cxE = CT(fxd * st->cxE);
      // halo-tile rows: own row is row+1; the j-1 / j+1 rows fall back to the own row at the box
      // edges (their stencil coefficient is zero; never touch values outside the box)
      const int r_c = row + 1, r_m = (j > j0) ? row : row + 1, r_p = (j < j1) ? row + 2 : row + 1;
      P Sa = P::bcast(CT(0)), Sq = Sa, Saa = Sa, Sqa = Sa;

      for (int ch = 0; ch < nch; ++ch, ++n) {
        const int stg = n % kTmaStages;
        mbar_wait(full0 + 8 * stg, (n / kTmaStages) & 1);
        if (row_on) {
          const FT* halo = reinterpret_cast<const FT*>(smem + (size_t)stg * G::stage_bytes);
          const FT* tl0 = reinterpret_cast<const FT*>(smem + (size_t)stg * G::stage_bytes + G::halo_bytes_pad) + row * C + lo;
          FT Tc[VEC], Tm[VEC], Tp[VEC], Tkm[VEC], Tkp[VEC], Tjm[VEC], Tjp[VEC], U[VEC], V[VEC], W[VEC];
          lds_vec<FT, VEC>(halo + r_c * HP + VEC + lo, Tc);
          lds_vec<FT, VEC>(halo + r_m * HP + VEC + lo, Tjm);
          lds_vec<FT, VEC>(halo + r_p * HP + VEC + lo, Tjp);
          lds_vec<FT, VEC>(tl0, Tm);
          lds_vec<FT, VEC>(tl0 + 1 * R * C, Tp);
          lds_vec<FT, VEC>(tl0 + 2 * R * C, Tkm);
          lds_vec<FT, VEC>(tl0 + 3 * R * C, Tkp);
          lds_vec<FT, VEC>(tl0 + 4 * R * C, U);
          lds_vec<FT, VEC>(tl0 + 5 * R * C, V);
          lds_vec<FT, VEC>(tl0 + 6 * R * C, W);
          FT Tl = __shfl_up_sync(0xffffffffu, Tc[VEC - 1], 1);
          FT Tr = __shfl_down_sync(0xffffffffu, Tc[0], 1);
          if (lane == 0) Tl = halo[r_c * HP + VEC - 1];
          if (lane == 31) Tr = halo[r_c * HP + VEC + C];
          const int col = i0 + ch * C + lo;
          if (ch == 0) {   // shift = first box value of the row (lane 0, element 0 of the first chunk)
            const FT s0 = __shfl_sync(0xffffffffu, Tc[0], 0);
            rc.shT = CT(s0);
            if (lane == 0) rec[R_SH_T] = double(s0);
          }
          CT wl_t[VEC], cxa_t[VEC], cxc_t[VEC];
          if constexpr (LONW == 1) load_lon_tables<CT, VEC>(p.g, i0, col, rc.fx, fxd, true, wl_t, cxa_t, cxc_t);
          else {
#pragma unroll
            for (int e = 0; e < VEC; ++e) { wl_t[e] = CT(1); cxa_t[e] = cxa_u; cxc_t[e] = cxc_u; }
          }
          CT tcv[VEC], tlv[VEC], trv[VEC];
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            tcv[e] = CT(Tc[e]);
            tlv[e] = CT((e > 0) ? Tc[e > 0 ? e - 1 : 0] : Tl);
            trv[e] = CT((e < VEC - 1) ? Tc[e < VEC - 1 ? e + 1 : 0] : Tr);
          }
          const bool edge_iter = (ch == 0) || (ch == nch - 1);
          if (edge_iter) {
            // box edges: half weights, one-sided lon stencil; columns east of the box are replaced by
            // finite dummies (weight 0), so the packed body below needs no branches
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
              const int i = col + e;
              if (i == i0) { wl_t[e] = wW; cxa_t[e] = CT(0); cxc_t[e] = cxW; tlv[e] = tcv[e]; rec[R_TW] = double(Tc[e]); }
              if (i == i1) { wl_t[e] = wE; cxa_t[e] = -cxE; cxc_t[e] = CT(0); trv[e] = tcv[e]; rec[R_TE] = double(Tc[e]); }
              if (i > i1) {
                wl_t[e] = CT(0); cxa_t[e] = cxc_t[e] = CT(0);
                tcv[e] = tlv[e] = trv[e] = rc.shT;
                Tm[e] = Tp[e] = Tkm[e] = Tkp[e] = Tjm[e] = Tjp[e] = FT(rc.shT);
                U[e] = V[e] = W[e] = FT(0);
              }
            }
          }
          const bool weighted = (LONW == 1) || edge_iter;
#pragma unroll
          for (int e = 0; e < VEC; e += 2) {
            const P tc = P::make(tcv[e], tcv[e + 1]);
            const P dtdt = pfma(P::bcast(rc.ct_m), P::make(CT(Tm[e]), CT(Tm[e + 1])) - tc,
                                pfma(P::bcast(rc.ct_p), P::make(CT(Tp[e]), CT(Tp[e + 1])) - tc, P::bcast(rc.ct_s) * tc));
            const P dTx = pfma(P::make(cxa_t[e], cxa_t[e + 1]), P::make(tlv[e], tlv[e + 1]) - tc,
                               P::make(cxc_t[e], cxc_t[e + 1]) * (P::make(trv[e], trv[e + 1]) - tc));
            const P dTy = pfma(P::bcast(rc.cy_m), P::make(CT(Tjm[e]), CT(Tjm[e + 1])) - tc,
                               P::bcast(rc.cy_p) * (P::make(CT(Tjp[e]), CT(Tjp[e + 1])) - tc));
            const P Ss = pfma(P::bcast(rc.s_m), P::make(CT(Tkm[e]), CT(Tkm[e + 1])) - tc,
                              pfma(P::bcast(rc.s_p), P::make(CT(Tkp[e]), CT(Tkp[e + 1])) - tc, P::bcast(rc.s_s) * tc));
            const P q = pfma(P::make(CT(U[e]), CT(U[e + 1])), dTx,
                             pfma(P::make(CT(V[e]), CT(V[e + 1])), dTy, pfma(P::make(CT(W[e]), CT(W[e + 1])), Ss, dtdt)));
            const P a = tc - P::bcast(rc.shT);
            if (weighted) {
              const P wg = P::make(wl_t[e], wl_t[e + 1]);
              const P Wa = wg * a;
              Sa = Sa + Wa; Sq = pfma(wg, q, Sq); Saa = pfma(Wa, a, Saa); Sqa = pfma(Wa, q, Sqa);
            } else {
              Sa = Sa + a; Sq = Sq + q; Saa = pfma(a, a, Saa); Sqa = pfma(a, q, Sqa);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * stg);
      }
      if (row_on) {
        double Sd[4] = {double(Sa.lo()) + double(Sa.hi()), double(Sq.lo()) + double(Sq.hi()),
                        double(Saa.lo()) + double(Saa.hi()), double(Sqa.lo()) + double(Sqa.hi())};
        double tot = butterfly_reduce<4>(Sd, lane);
        if (LONW == 0) tot *= p.g.wl_u;
        const int idx = bitrev5(lane);
        if (idx == 0) rec[R_A] = tot;
        if (idx == 1) rec[R_Q] = tot;
        if (idx == 2) rec[R_AA] = tot;
        if (idx == 3) rec[R_QA] = tot;
      }
    } else {
      // =============================== role B: wind / omega / Phi sums ==========================
      CT shT = CT(0), shU = CT(0), shV = CT(0), shW = CT(0), shF = CT(0);
      P S[18];
#pragma unroll
      for (int q = 0; q < 18; ++q) S[q] = P::bcast(CT(0));
      enum { B_B, B_C, B_W, B_F, B_BB, B_CC, B_BC, B_CA, B_WA, B_WB, B_WC, B_WF, B_CAA, B_WAA, B_BBC, B_CCC, B_BBW, B_CCW };

      for (int ch = 0; ch < nch; ++ch, ++n) {
        const int stg = n % kTmaStages;
        if (producer && p_nch > 0) p_issue();                  // keep kTmaStages-1 chunks in flight
        mbar_wait(full0 + 8 * stg, (n / kTmaStages) & 1);
        if (row_on) {
          const FT* halo = reinterpret_cast<const FT*>(smem + (size_t)stg * G::stage_bytes);
          const FT* tl0 = reinterpret_cast<const FT*>(smem + (size_t)stg * G::stage_bytes + G::halo_bytes_pad) + row * C + lo;
          FT Tc[VEC], U[VEC], V[VEC], W[VEC], F[VEC];
          lds_vec<FT, VEC>(halo + (row + 1) * HP + VEC + lo, Tc);
          lds_vec<FT, VEC>(tl0 + 4 * R * C, U);
          lds_vec<FT, VEC>(tl0 + 5 * R * C, V);
          lds_vec<FT, VEC>(tl0 + 6 * R * C, W);
          lds_vec<FT, VEC>(tl0 + 7 * R * C, F);
          const int col = i0 + ch * C + lo;
          if (ch == 0) {
            const FT sT = __shfl_sync(0xffffffffu, Tc[0], 0), sU = __shfl_sync(0xffffffffu, U[0], 0),
                     sV = __shfl_sync(0xffffffffu, V[0], 0), sW = __shfl_sync(0xffffffffu, W[0], 0),
                     sF = __shfl_sync(0xffffffffu, F[0], 0);
            shT = CT(sT); shU = CT(sU); shV = CT(sV); shW = CT(sW); shF = CT(sF);
            if (lane == 0) { rec[R_SH_U] = double(sU); rec[R_SH_V] = double(sV); rec[R_SH_W] = double(sW); rec[R_SH_F] = double(sF); }
          }
          CT wl_t[VEC], dummy_a[VEC], dummy_c[VEC];
          if constexpr (LONW == 1) load_lon_tables<CT, VEC>(p.g, i0, col, CT(0), 0.0, false, wl_t, dummy_a, dummy_c);
          else {
#pragma unroll
            for (int e = 0; e < VEC; ++e) wl_t[e] = CT(1);
          }
          const bool edge_iter = (ch == 0) || (ch == nch - 1);
          if (edge_iter) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
              const int i = col + e;
              if (i == i0) { wl_t[e] = wW; rec[R_UW] = double(U[e]); rec[R_VW] = double(V[e]); }
              if (i == i1) { wl_t[e] = wE; rec[R_UE] = double(U[e]); rec[R_VE] = double(V[e]); }
              if (i > i1) { wl_t[e] = CT(0); Tc[e] = FT(shT); U[e] = FT(shU); V[e] = FT(shV); W[e] = FT(shW); F[e] = FT(shF); }
            }
          }
          const bool weighted = (LONW == 1) || edge_iter;
#pragma unroll
          for (int e = 0; e < VEC; e += 2) {
            const P a = P::make(CT(Tc[e]), CT(Tc[e + 1])) - P::bcast(shT);
            const P b = P::make(CT(U[e]), CT(U[e + 1])) - P::bcast(shU);
            const P c = P::make(CT(V[e]), CT(V[e + 1])) - P::bcast(shV);
            const P w = P::make(CT(W[e]), CT(W[e + 1])) - P::bcast(shW);
            const P f = P::make(CT(F[e]), CT(F[e + 1])) - P::bcast(shF);
            P Wb = b, Wc = c, Ww = w, Wf = f;
            if (weighted) { const P wg = P::make(wl_t[e], wl_t[e + 1]); Wb = wg * b; Wc = wg * c; Ww = wg * w; Wf = wg * f; }
            S[B_B] = S[B_B] + Wb; S[B_C] = S[B_C] + Wc; S[B_W] = S[B_W] + Ww; S[B_F] = S[B_F] + Wf;
            const P pbb = Wb * b, pcc = Wc * c, pca = Wc * a, pwa = Ww * a;
            S[B_BB] = S[B_BB] + pbb; S[B_CC] = S[B_CC] + pcc; S[B_BC] = pfma(Wb, c, S[B_BC]);
            S[B_CA] = S[B_CA] + pca; S[B_WA] = S[B_WA] + pwa;
            S[B_WB] = pfma(Ww, b, S[B_WB]); S[B_WC] = pfma(Ww, c, S[B_WC]); S[B_WF] = pfma(Ww, f, S[B_WF]);
            S[B_CAA] = pfma(pca, a, S[B_CAA]); S[B_WAA] = pfma(pwa, a, S[B_WAA]);
            S[B_BBC] = pfma(pbb, c, S[B_BBC]); S[B_CCC] = pfma(pcc, c, S[B_CCC]);
            S[B_BBW] = pfma(pbb, w, S[B_BBW]); S[B_CCW] = pfma(pcc, w, S[B_CCW]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * stg);
      }
      if (row_on) {
        double Sd[18];
#pragma unroll
        for (int q = 0; q < 18; ++q) Sd[q] = double(S[q].lo()) + double(S[q].hi());
        double tot = butterfly_reduce<18>(Sd, lane);
        if (LONW == 0) tot *= p.g.wl_u;
        const int idx = bitrev5(lane);
        // role-B slot -> record slot
        const int map[18] = {R_B, R_C, R_W, R_F, R_BB, R_CC, R_BC, R_CA, R_WA, R_WB, R_WC, R_WF,
                             R_CAA, R_WAA, R_BBC, R_CCC, R_BBW, R_CCW};
        if (idx < 18) rec[map[idx]] = tot;
      }
    }
  }
}

}  // namespace lec
