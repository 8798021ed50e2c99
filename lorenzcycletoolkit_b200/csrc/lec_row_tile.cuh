// Kernel A, TMA-tiled variant: the same sweep as lec_row_moments_kernel (same lane <-> column mapping,
// same per-lane accumulation order, same butterfly -- so the row records have the SAME BITS), fed from
// shared memory by the Tensor Memory Accelerator instead of per-lane global loads.
//
//  * One persistent CTA per SM walks the tile list (band, step, level, row-tile) -- the band-major order
//    of the direct kernel, so T(t+1) is still reused out of L2.  A tile is R consecutive box rows of one
//    (step, level); it is swept in chunks of 32 x 128 bit per row.
//  * ONE producer thread (its own warp, no arithmetic) issues, per chunk, 9 TMA tensor loads
//        T centre tile with the j-1 / j+1 rows and one vector of lon halo  ((R+2) x (C+2 VEC))
//        T(t-1), T(t+1), T(k-1), T(k+1), u, v, omega, Phi                     (R x C each)
//    into an S-stage shared-memory ring (full / empty mbarriers, expect_tx byte counting).  The j+-1 rows
//    are fetched once per tile instead of once per row (L2 -> SM traffic 9 + 2/R row streams instead of
//    11 + 2 scalar neighbours), and the bytes in flight ((S-1) stages, 110-170 KB per SM) no longer
//    depend on registers or occupancy.
//  * R consumer warps, one box row each: wait(full) -> 11 LDS.128 + 2 LDS.32 -> the shared iteration
//    body (lec_row_body.inc) -> arrive(empty).  No global addressing, no long-scoreboard stalls in the
//    arithmetic warps; rows outside the box or the grid are zero-filled by TMA and masked as before.
#pragma once
#include <cuda.h>

#include "lec_common.cuh"
#include "lec_packed.cuh"
#include "lec_row_moments.cuh"

namespace lec {

struct alignas(64) TmaMaps {
  CUtensorMap t_halo;    // T, box (C + 2 VEC) x (R + 2)
  CUtensorMap t_plain;   // T, box C x R
  CUtensorMap u, v, w, f;
};

template <typename FT, int R, int S>
struct TileGeom {
  static constexpr int VEC = 16 / sizeof(FT);
  static constexpr int C = 32 * VEC;                         // columns per chunk (512 B per row)
  static constexpr int HP = C + 2 * VEC;                     // halo tile pitch (elements)
  static constexpr int halo_bytes = (R + 2) * HP * sizeof(FT);
  static constexpr int halo_bytes_pad = (halo_bytes + 127) / 128 * 128;
  static constexpr int tile_bytes = R * C * sizeof(FT);
  static constexpr int stage_bytes = halo_bytes_pad + 8 * tile_bytes;
  static constexpr int tx_bytes = halo_bytes + 8 * tile_bytes;           // what the 9 loads deliver
  static constexpr int smem_bytes = S * stage_bytes + 128;               // + barriers
  static constexpr int threads = (R + 1) * 32;                           // R consumer warps + the producer warp
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

template <typename FT, typename CT, int LONW, int R, int S>
__global__ void __launch_bounds__((R + 1) * 32, 1)
lec_row_moments_tile_kernel(const __grid_constant__ TmaMaps maps, const RowParams p) {
  using G = TileGeom<FT, R, S>;
  constexpr int VEC = G::VEC, C = G::C, HP = G::HP;
  extern __shared__ __align__(1024) unsigned char smem[];   // plain shared pointer: keeps LDS (not generic LD)
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + S * G::stage_bytes);
  const unsigned full0 = smem_u32(bars), empty0 = smem_u32(bars + S);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, R); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int nlev = p.g.nlev;

  if (warp == R) {
    // ===================================== producer ==========================================
    // One thread walks the same (tile, chunk) sequence as the consumers and keeps the ring full; a chunk
    // costs one wait on its stage's empty barrier, one expect_tx and 9 tensor loads.
    if (lane == 0) {
      int stg = 0;
      unsigned phase = 0;
      for (unsigned id = blockIdx.x; id < (unsigned)p.grid; id += gridDim.x) {
        const TileId t = decode_tile(id, p);
        const StepDev* __restrict__ st = p.steps + t.s;
        const int jr = (t.band * p.tiles_per_band + t.jt) * R;
        if (jr > st->j1 - st->j0) continue;
        const int c0 = st->i0 / VEC, c1 = st->i1 / VEC;
        const int nch = (c1 - c0 + 32) / 32;
        const int jt0 = st->j0 + jr;
        const int k = t.k, km = k > 0 ? k - 1 : k, kp = k < nlev - 1 ? k + 1 : k;
        const int slot = st->slot, slot_m = st->slot_m, slot_p = st->slot_p;
        int col = c0 * VEC;
        for (int ch = 0; ch < nch; ++ch, col += C) {
          mbar_wait(empty0 + 8 * stg, phase ^ 1);
          const unsigned fb = full0 + 8 * stg;
          mbar_expect_tx(fb, G::tx_bytes);
          const unsigned base = smem_u32(smem) + (unsigned)stg * G::stage_bytes;
          tma_load_4d(base, &maps.t_halo, fb, col - VEC, jt0 - 1, k, slot);
          unsigned d = base + G::halo_bytes_pad;
          tma_load_4d(d, &maps.u, fb, col, jt0, k, slot); d += G::tile_bytes;
          tma_load_4d(d, &maps.v, fb, col, jt0, k, slot); d += G::tile_bytes;
          tma_load_4d(d, &maps.w, fb, col, jt0, k, slot); d += G::tile_bytes;
          tma_load_4d(d, &maps.f, fb, col, jt0, k, slot); d += G::tile_bytes;
          tma_load_4d(d, &maps.t_plain, fb, col, jt0, k, slot_p); d += G::tile_bytes;
          tma_load_4d(d, &maps.t_plain, fb, col, jt0, k, slot_m); d += G::tile_bytes;
          tma_load_4d(d, &maps.t_plain, fb, col, jt0, km, slot); d += G::tile_bytes;
          tma_load_4d(d, &maps.t_plain, fb, col, jt0, kp, slot);
          if (++stg == S) { stg = 0; phase ^= 1; }
        }
      }
    }
    return;
  }

  // ======================================= consumers ===========================================
  int stg = 0;
  unsigned phase = 0;
  for (unsigned id = blockIdx.x; id < (unsigned)p.grid; id += gridDim.x) {
    const TileId t = decode_tile(id, p);
    const StepDev* __restrict__ st = p.steps + t.s;
    const int i0 = st->i0, i1 = st->i1, j0 = st->j0, j1 = st->j1;
    const int jrel0 = (t.band * p.tiles_per_band + t.jt) * R;
    if (jrel0 > j1 - j0) continue;                 // same test as the producer: no chunk was issued
    const int jrel = jrel0 + warp;
    const bool row_on = jrel <= j1 - j0;
    const int j = row_on ? j0 + jrel : j1;
    const int k = t.k;
    const int c0 = i0 / VEC, c1 = i1 / VEC;
    const int niter = (c1 - c0 + 32) / 32;

    // row-level coefficients: every constant factor was folded on the host (lec_engine.cu)
    RowCoefS<CT> rc;
    rc.ct_m = CT(st->ct_m); rc.ct_p = CT(st->ct_p); rc.ct_s = CT(st->ct_s);
    rc.cy_m = CT((j == j0) ? 0.0 : (j == j1) ? -st->cyN : p.g.cya[j]);
    rc.cy_p = CT((j == j1) ? 0.0 : (j == j0) ? st->cyS : p.g.cyc[j]);
    rc.s_m = CT(p.g.sm[k]); rc.s_p = CT(p.g.sp[k]); rc.s_s = CT(p.g.ss[k]);
    const double fxd = p.g.fxj[j];
    rc.fx = CT(fxd);
    const CT cxa_u = CT(fxd * p.g.cxa_u), cxc_u = CT(fxd * p.g.cxc_u);
    const CT cxW = CT(fxd * st->cxW), cxE = CT(fxd * st->cxE);
    const double wnorm = (LONW == 0) ? 1.0 / p.g.wl_u : 1.0;
    const CT wW = CT(st->wW * wnorm), wE = CT(st->wE * wnorm);
    // halo-tile rows: own row is warp + 1; at the box edges the j-1 / j+1 row falls back to the own row
    // (its stencil coefficient is zero; values outside the box are never touched)
    const int r_c = warp + 1, r_m = (j > j0) ? warp : warp + 1, r_p = (j < j1) ? warp + 2 : warp + 1;

    FT shT = FT(0), shU = FT(0), shV = FT(0), shW = FT(0), shF = FT(0);
    CT cshT = CT(0), cshU = CT(0), cshV = CT(0), cshW = CT(0), cshF = CT(0);
    CT Sacc[R_NSUM], Cc[R_NLIN];
#pragma unroll
    for (int n = 0; n < R_NSUM; ++n) Sacc[n] = CT(0);
#pragma unroll
    for (int n = 0; n < R_NLIN; ++n) Cc[n] = CT(0);
    double* __restrict__ rec = p.rec + (((long long)t.s * nlev + k) * p.max_ny + jrel) * LEC_NREC;

    for (int it = 0; it < niter; ++it) {
      mbar_wait(full0 + 8 * stg, phase);
      if (row_on) {
        const unsigned char* sb = smem + (size_t)stg * G::stage_bytes;
        const FT* halo = reinterpret_cast<const FT*>(sb);
        const FT* tl0 = reinterpret_cast<const FT*>(sb + G::halo_bytes_pad) + warp * C + lane * VEC;
        if (it == 0) {      // shifts: raw first-in-box values of the row (broadcast reads)
          const int e0 = i0 - c0 * VEC;
          shT = halo[r_c * HP + VEC + e0];
          shU = reinterpret_cast<const FT*>(sb + G::halo_bytes_pad)[warp * C + e0];
          shV = reinterpret_cast<const FT*>(sb + G::halo_bytes_pad)[(1 * R + warp) * C + e0];
          shW = reinterpret_cast<const FT*>(sb + G::halo_bytes_pad)[(2 * R + warp) * C + e0];
          shF = reinterpret_cast<const FT*>(sb + G::halo_bytes_pad)[(3 * R + warp) * C + e0];
          cshT = CT(shT); cshU = CT(shU); cshV = CT(shV); cshW = CT(shW); cshF = CT(shF);
        }
        const int c_raw = c0 + it * 32 + lane;
        const bool lane_on = c_raw <= c1;
        const int col = (lane_on ? c_raw : c1) * VEC;      // table index / box mask; the tile slot is the lane's own
        FT Tc[VEC], Tm[VEC], Tp[VEC], Tkm[VEC], Tkp[VEC], Tjm[VEC], Tjp[VEC], U[VEC], V[VEC], W[VEC], F[VEC];
        const FT* hc = halo + r_c * HP + VEC + lane * VEC;
        lds_vec<FT, VEC>(hc, Tc);
        lds_vec<FT, VEC>(halo + r_m * HP + VEC + lane * VEC, Tjm);
        lds_vec<FT, VEC>(halo + r_p * HP + VEC + lane * VEC, Tjp);
        lds_vec<FT, VEC>(tl0, U);
        lds_vec<FT, VEC>(tl0 + 1 * R * C, V);
        lds_vec<FT, VEC>(tl0 + 2 * R * C, W);
        lds_vec<FT, VEC>(tl0 + 3 * R * C, F);
        lds_vec<FT, VEC>(tl0 + 4 * R * C, Tp);
        lds_vec<FT, VEC>(tl0 + 5 * R * C, Tm);
        lds_vec<FT, VEC>(tl0 + 6 * R * C, Tkm);
        lds_vec<FT, VEC>(tl0 + 7 * R * C, Tkp);
        const FT Tl = hc[-1];
        const FT Tr = hc[VEC];
        CT (&S_)[R_NSUM] = Sacc;
#define S S_
#define LEC_TAB_WL p.g.wl32
#define LEC_TAB_CXA p.g.cxa32
#define LEC_TAB_CXC p.g.cxc32
#define LEC_TAB_LOAD(ptr, dst) VecLoad<float, VEC>::ld(ptr, dst)
#include "lec_row_body.inc"
#undef LEC_TAB_WL
#undef LEC_TAB_CXA
#undef LEC_TAB_CXC
#undef LEC_TAB_LOAD
#undef S
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8 * stg);
      if (++stg == S) { stg = 0; phase ^= 1; }
    }

    if (row_on) {
      double Sd[R_NSUM];
#pragma unroll
      for (int n = 0; n < R_NSUM; ++n) Sd[n] = double(Sacc[n]);
      if constexpr (sizeof(CT) == 4) {
#pragma unroll
        for (int n = 0; n < R_NLIN; ++n) Sd[n] += double(Cc[n]);
      }
      double tot = butterfly_reduce<R_NSUM>(Sd, lane);
      if (LONW == 0) tot *= p.g.wl_u;
      const int idx = bitrev5(lane);
      if (idx < R_NSUM) rec[idx] = tot;
      if (lane == 1) {   // raw (unscaled) shifts; the finalize kernel applies the unit scales
        rec[R_SH_T] = double(shT); rec[R_SH_U] = double(shU); rec[R_SH_V] = double(shV);
        rec[R_SH_W] = double(shW); rec[R_SH_F] = double(shF);
      }
    }
  }
}

}  // namespace lec
