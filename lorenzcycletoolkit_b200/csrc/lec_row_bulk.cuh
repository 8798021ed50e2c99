// Kernel A, bulk-TMA staged variant.
//
// Same arithmetic, row records and warp-per-row mapping as lec_row_moments_kernel (the iteration
// body is the shared lec_row_body.inc), but the 11 row chunks of a sweep iteration are not loaded
// into registers: every warp owns a PRIVATE double buffer in shared memory that lanes 0..10 fill with
// one 512-byte `cp.async.bulk` (TMA 1-D bulk copy, SASS UBLKCP) each, completion tracked by a
// per-buffer mbarrier.  The copies of iteration it+1 are in flight while iteration it is computed, so
//   * bytes in flight no longer cost registers (the direct kernel holds 44 of its 126 registers for
//     loads and exposes the full L2/DRAM latency once per iteration),
//   * warps stay independent (no CTA-wide in-order ring: the tiled TMA kernel of lec_row_tma.cuh lost
//     26 % of its time waiting on the slowest of ~300 sectors per stage),
//   * the +-1 longitude halo of T rides along in the same copy, and the per-column longitude tables
//     sit in shared memory, so no synchronous global load is left on the critical path.
// Persistent CTAs (2 per SM, 8 warps each): warp g processes rows g, g + G, g + 2G, ... of the same
// band-major row order as the direct kernel.
#pragma once
#include "lec_common.cuh"
#include "lec_packed.cuh"
#include "lec_row_moments.cuh"
#include "lec_row_tma.cuh"   // mbarrier helpers

namespace lec {

constexpr int kBulkWarps = 8;
constexpr int kBulkThreads = kBulkWarps * 32;
constexpr int kBulkTcBytes = 512 + 32;                       // T centre chunk with a 16-byte halo each side
constexpr int kBulkBufBytes = kBulkTcBytes + 10 * 512;       // + Tm Tp Tkm Tkp Tjm Tjp U V W F
constexpr int kBulkWarpBytes = 2 * kBulkBufBytes;            // double buffer

__device__ __forceinline__ void bulk_copy_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// smem: [3 x nlon_pad fp32 tables (TABS == 1)] [kBulkWarps x kBulkWarpBytes] [kBulkWarps x 2 mbarriers]
template <typename FT, typename CT, int LONW, int TABS>
__global__ void __launch_bounds__(kBulkThreads, 2)
lec_row_moments_bulk_kernel(const RowParams p) {
  constexpr int VEC = 16 / sizeof(FT);
  constexpr int CH = 32 * VEC;                               // columns per chunk
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nlon = p.g.nlon, nlat = p.g.nlat, nlev = p.g.nlev;
  const int nlon_pad = (nlon + 3) & ~3;
  const int tab_bytes = TABS ? 3 * nlon_pad * 4 : 0;
  const float* tab_wl = reinterpret_cast<const float*>(smem);
  const float* tab_cxa = tab_wl + nlon_pad;
  const float* tab_cxc = tab_cxa + nlon_pad;
  unsigned char* wbuf = smem + tab_bytes + warp * kBulkWarpBytes;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + tab_bytes + kBulkWarps * kBulkWarpBytes);
  const unsigned bar0 = smem_u32(bars + 2 * warp);

  if (TABS) {
    float* t = reinterpret_cast<float*>(smem);
    for (int i = threadIdx.x; i < nlon; i += kBulkThreads) {
      t[i] = p.g.wl32[i]; t[nlon_pad + i] = p.g.cxa32[i]; t[2 * nlon_pad + i] = p.g.cxc32[i];
    }
  }
  if (lane == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  const long long plane = (long long)nlat * nlon;
  const unsigned total_rows = (unsigned)p.grid;               // rows incl. the padding rows of the last band
  const unsigned gwarps = gridDim.x * kBulkWarps;
  unsigned fills = 0;                                          // buffers filled so far by this warp

  for (unsigned rid = blockIdx.x * kBulkWarps + warp; rid < total_rows; rid += gwarps) {
    // row order: band-major, then step, level, row (tiles_per_band = rows per band here)
    const unsigned q1 = rid / (unsigned)p.tiles_per_band;
    const int jr = int(rid - q1 * (unsigned)p.tiles_per_band);
    const unsigned q2 = q1 / (unsigned)nlev;
    const int k = int(q1 - q2 * (unsigned)nlev);
    const int band = int(q2 / (unsigned)p.nsteps);
    const int s = int(q2 - (unsigned)band * (unsigned)p.nsteps);
    const StepDev* __restrict__ st = p.steps + s;
    const int i0 = st->i0, i1 = st->i1, j0 = st->j0, j1 = st->j1;
    const int jrel = band * p.tiles_per_band + jr;
    if (jrel > j1 - j0) continue;
    const int j = j0 + jrel;

    const long long row_c = ((long long)st->slot * nlev + k) * plane + (long long)j * nlon;
    const FT* __restrict__ Tc_row = static_cast<const FT*>(p.field[0]) + row_c;
    const int d_m = int((long long)(st->slot_m - st->slot) * p.slot_stride);
    const int d_p = int((long long)(st->slot_p - st->slot) * p.slot_stride);
    const int d_km = (k > 0) ? -int(plane) : 0, d_kp = (k < nlev - 1) ? int(plane) : 0;
    const int d_jm = (j > j0) ? -nlon : 0, d_jp = (j < j1) ? nlon : 0;

    // lane a (0..10) drives the copies of array a: its row base pointer and its slot in the buffer
    const FT* my_src;
    {
      const int dl = (lane == 1) ? d_m : (lane == 2) ? d_p : (lane == 3) ? d_km : (lane == 4) ? d_kp
                   : (lane == 5) ? d_jm : (lane == 6) ? d_jp : 0;
      my_src = (lane < 7) ? Tc_row + dl : static_cast<const FT*>(p.field[lane < 11 ? lane - 6 : 1]) + row_c;
    }
    const unsigned my_dst = (lane == 0) ? 0u : unsigned(kBulkTcBytes + (lane - 1) * 512);

    if ((p.prefetch_mode & 1) && lane < 5) {
      const FT* base = (lane == 0) ? Tc_row + d_p : static_cast<const FT*>(p.field[lane]) + row_c;
      prefetch_l2_range(base, (long long)i0 * sizeof(FT), (long long)(i1 + 1) * sizeof(FT));
    }

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

    const int c0 = i0 / VEC, c1 = i1 / VEC;
    const int niter = (c1 - c0 + 32) / 32;

    // start the copies of sweep iteration it2 into buffer (fills & 1); called by the whole warp
    auto stage_issue = [&](int it2) {
      const int colw = (c0 + it2 * 32) * VEC;                 // first column of the warp's chunk
      const int end = min(colw + CH, nlon);
      const int hs = max(colw - VEC, 0), he = min(colw + CH + VEC, nlon);
      const unsigned bar = bar0 + 8 * (fills & 1);
      const unsigned base = smem_u32(wbuf + (fills & 1) * kBulkBufBytes);
      if (lane == 0)
        mbar_expect_tx(bar, unsigned((he - hs) + 10 * (end - colw)) * (unsigned)sizeof(FT));
      __syncwarp();
      if (lane == 0) {
        bulk_copy_g2s(base + unsigned(hs - (colw - VEC)) * (unsigned)sizeof(FT), my_src + hs,
                      unsigned(he - hs) * (unsigned)sizeof(FT), bar);
      } else if (lane < 11) {
        bulk_copy_g2s(base + my_dst, my_src + colw, unsigned(end - colw) * (unsigned)sizeof(FT), bar);
      }
      ++fills;
    };
    stage_issue(0);

    // shifts: raw first-in-box values of the row (broadcast loads; L2 hits once the prefetch landed)
    const FT* U_row = static_cast<const FT*>(p.field[1]) + row_c;
    const FT shT = __ldg(Tc_row + i0), shU = __ldg(U_row + i0),
             shV = __ldg(static_cast<const FT*>(p.field[2]) + row_c + i0),
             shW = __ldg(static_cast<const FT*>(p.field[3]) + row_c + i0),
             shF = __ldg(static_cast<const FT*>(p.field[4]) + row_c + i0);
    const CT cshT = CT(shT), cshU = CT(shU), cshV = CT(shV), cshW = CT(shW), cshF = CT(shF);

    CT S[R_NSUM];
#pragma unroll
    for (int n = 0; n < R_NSUM; ++n) S[n] = CT(0);
    double* __restrict__ rec = p.rec + (((long long)s * nlev + k) * p.max_ny + jrel) * LEC_NREC;

    for (int it = 0; it < niter; ++it) {
      const int c_raw = c0 + it * 32 + lane;
      const bool lane_on = c_raw <= c1;
      const int col = (lane_on ? c_raw : c1) * VEC;
      const unsigned cur = fills - 1;                          // buffer that holds iteration `it`
      if (it + 1 < niter) stage_issue(it + 1);
      mbar_wait(bar0 + 8 * (cur & 1), (cur >> 1) & 1);

      const unsigned char* b = wbuf + (cur & 1) * kBulkBufBytes;
      const FT* tcb = reinterpret_cast<const FT*>(b) + VEC;   // element `first column of the warp's chunk`
      const int lo = (lane_on ? lane : (c1 - c0 - it * 32)) * VEC;   // clamped lanes re-read the last valid chunk
      FT Tc[VEC], Tm[VEC], Tp[VEC], Tkm[VEC], Tkp[VEC], Tjm[VEC], Tjp[VEC], U[VEC], V[VEC], W[VEC], F[VEC];
      const FT* ab = reinterpret_cast<const FT*>(b + kBulkTcBytes) + lo;
      constexpr int E = 512 / sizeof(FT);
      lds_vec<FT, VEC>(tcb + lo, Tc);
      lds_vec<FT, VEC>(ab + 0 * E, Tm); lds_vec<FT, VEC>(ab + 1 * E, Tp); lds_vec<FT, VEC>(ab + 2 * E, Tkm);
      lds_vec<FT, VEC>(ab + 3 * E, Tkp); lds_vec<FT, VEC>(ab + 4 * E, Tjm); lds_vec<FT, VEC>(ab + 5 * E, Tjp);
      lds_vec<FT, VEC>(ab + 6 * E, U); lds_vec<FT, VEC>(ab + 7 * E, V); lds_vec<FT, VEC>(ab + 8 * E, W);
      lds_vec<FT, VEC>(ab + 9 * E, F);
      // lon neighbours of the chunk ends: adjacent lanes; the warp ends read the halo of the T chunk
      FT Tl = __shfl_up_sync(0xffffffffu, Tc[VEC - 1], 1);
      FT Tr = __shfl_down_sync(0xffffffffu, Tc[0], 1);
      if (lane == 0) Tl = tcb[-1];
      if (lane == 31) Tr = tcb[CH];

#define LEC_TAB_WL (TABS ? tab_wl : p.g.wl32)
#define LEC_TAB_CXA (TABS ? tab_cxa : p.g.cxa32)
#define LEC_TAB_CXC (TABS ? tab_cxc : p.g.cxc32)
#define LEC_TAB_LOAD(ptr, dst) \
  do { if (TABS) lds_vec<float, VEC>(ptr, dst); else VecLoad<float, VEC>::ld(ptr, dst); } while (0)
#include "lec_row_body.inc"
#undef LEC_TAB_WL
#undef LEC_TAB_CXA
#undef LEC_TAB_CXC
#undef LEC_TAB_LOAD
    }

    double Sd[R_NSUM];
#pragma unroll
    for (int n = 0; n < R_NSUM; ++n) Sd[n] = double(S[n]);
    double tot = butterfly_reduce<R_NSUM>(Sd, lane);
    if (LONW == 0) tot *= p.g.wl_u;
    const int idx = bitrev5(lane);
    if (idx < R_NSUM) rec[idx] = tot;
    if (lane == 1) {
      rec[R_SH_T] = double(shT); rec[R_SH_U] = double(shU); rec[R_SH_V] = double(shV);
      rec[R_SH_W] = double(shW); rec[R_SH_F] = double(shF);
    }
  }
}

}  // namespace lec
