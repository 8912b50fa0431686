// band_kernels.cu — the row-streaming step in lock step ("band kernel", backend BAND_TB, sm_100a).
//
// Same arithmetic as every other step kernel (flowUpdate.comp:12-63 + flowApply.comp:14-53, contract in
// cell_math.cuh / DESIGN.md section 2) and bit-identical results.  Like the ring kernel
// (stream_kernels.cu) it blocks K whole steps over one HBM round trip by SKEWING rows in time instead
// of recomputing halo rows: row y runs half-pass s once rows y-1 and y+1 have finished half-pass s-1.
// The ring kernel lets every row advance on its own (per-row mbarriers); its throughput is capped by
// the dependency cone — only NW - 2K of its NW row slots make progress per ring turn and every
// half-pass pays a wake-up latency (ncu: a third of the warp samples sit in the neighbour waits).  The
// band kernel runs the same schedule in lock step:
//
//   * a column strip (128 cells, one float4 group per lane) is marched in BANDS of R*NW rows; row
//     o = q*NW + w of a band (q < R) belongs to warp w, which keeps its R rows in registers
//     (24 per row) and updates them together — R independent dependency chains per thread;
//   * half-pass s of a band updates the R*NW consecutive rows (Yb - R*NW - s, Yb - s] (Yb = last row of
//     the band): the window slides up by one row per half-pass, which is exactly the skew the data
//     dependency needs, so no row is ever recomputed, and a window of R*NW consecutive rows holds
//     exactly R rows of every warp: all warps work in every half-pass and the only synchronisation is
//     one barrier per half-pass;
//   * rows q < R-1 and the rows q = R-1 of the warps w < NW-2K finish inside their band.  The last 2K
//     rows of a band cannot (the rows below them are not loaded yet): their warps ("carriers") take
//     them through NW-1-w half-passes, park them in shared memory, pick the row they parked one band
//     earlier up again and finish that one (half-passes NW-w .. 2K);
//   * what a vertical neighbour needs goes through one exchange slot per row (H | F+Y | F-Y, 12 B per
//     cell); the slot of a row must outlive the band when the row below it is carried, so the last
//     2K+1 rows of a band alternate between two slots by band parity; x neighbours are adjacent lanes;
//   * TMA lands a warp's R rows of the NEXT band (terrain: one 2-D box per row, the five state planes:
//     one 3-D box per row) in its private landing buffers while the current band is computed;
//     out-of-bounds zero fill is the reference's exterior rule, as everywhere;
//   * NGRP independent warp groups share the CTA, each with its own pieces of the (strip, row) list, its
//     own buffers and its own named barrier: one group's barrier waits overlap the other's arithmetic.
//
// Rows that have nothing useful to do in a half-pass (warm-up rows above a piece, feeder rows below it
// that are past their last needed half-pass, missing rows of the last band) are computed anyway: their
// results are never stored (store guards) and by the schedule never reach a kept value, and computing
// them keeps the half-pass free of per-row branches.
#include "cell_math.cuh"
#include "band_schedule.h"

#include <algorithm>
#include <climits>
#include <cstdlib>

namespace tws {

template <int K_, int NW_, int R_, int NGRP_>
struct BandCfg {
  static constexpr int K = K_, NW = NW_, R = R_, NGRP = NGRP_, NT = NW_ * NGRP_ * 32;
  static constexpr int SXW = 128;                        // one float4 group per lane
  static constexpr int HX = stream_hx(K);
  static constexpr int OX = SXW - 2 * HX;
  static constexpr int HP = 2 * K;                       // computing half-passes per row
  static constexpr int BR = R * NW;                      // rows per band
  static constexpr int NC = HP;                          // carrier warps: NW-HP .. NW-1 (their row q = R-1)
  static constexpr int NDB = HP + 1;                     // rows of a band with a double-buffered exchange slot
  static constexpr int NSLOT = BR + NDB;
  static constexpr int LAND = 6 * SXW;                   // floats per landing / parking buffer (h, d, F x4 of one row)
  static constexpr int XROW = 3 * SXW;                   // floats per exchange slot
  static constexpr int GROUP_FLOATS = BR * LAND + NSLOT * XROW + NC * LAND;
  static constexpr size_t SMEM = (size_t)NGRP * GROUP_FLOATS * sizeof(float);
  static_assert(NW >= HP + 1, "the rows of one warp-row of a band must be deeper than the dependency cone");
  static_assert(SMEM <= 227 * 1024, "band configuration does not fit in shared memory");
};

struct BandRow {       // per-row context (the lane's column offset and store mask are shared by its rows)
  int gy;              // global row
  bool row_in;         // the row exists in the global grid
  bool store;          // the row is an output row of this piece
  size_t go;           // element offset of the lane's group in the output planes
};

// Split-phase group barrier: one mbarrier per warp group, one arrival per warp.  arrive() right after a
// half-pass has published what the neighbour rows need, wait() right before the next half-pass reads
// what the neighbours published — the arithmetic that needs no neighbour data (own-row sums and
// products before the wait, scaling of the unpublished components and the HBM stores after the arrive)
// runs while the other warps catch up.  Every warp alternates arrive / wait strictly, so a single
// barrier and one phase bit per thread suffice (nobody can arrive for phase n+1 before it has seen
// phase n complete).
struct GroupSync {
  uint32_t bar;        // shared address of the group's mbarrier
  uint32_t phase;      // parity of the phase the next wait() waits for
  __device__ __forceinline__ void arrive() {
    __syncwarp();                                         // every lane's shared-memory stores are ordered before the release
    asm volatile(
        "{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\n@p mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];\n}" ::"r"(bar) : "memory");
  }
  __device__ __forceinline__ void wait() { wait_now(); }
  // Four polls per pass of the loop, and only a failed fourth one pays for the hang guard: a waiting warp shares its
  // scheduler with working ones, so every instruction of the wait loop is taken from a row that could make progress
  // (round-1 ncu: ~6 wake-ups per wait at 7 instructions each were 30 % of all issued instructions; now 2 per poll
  // + 4 per four polls).  try_wait itself suspends the warp for a hardware-defined time before it reports failure.
  __device__ __forceinline__ void wait_now() {
#define TWS_GS_POLL "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra TWS_GS_DONE;\n"
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        ".reg .u32 n;\n"
        "mov.u32 n, 0;\n"
        "TWS_GS_LOOP:\n"
        TWS_GS_POLL TWS_GS_POLL TWS_GS_POLL TWS_GS_POLL
        "add.u32 n, n, 1;\n"
        "setp.gt.u32 q, n, 4194304;\n"
        "@q trap;\n"                                      // a barrier that never completes is a bug; a trap beats a hung GPU
        "bra TWS_GS_LOOP;\n"
        "TWS_GS_DONE:\n"
        "}\n" ::"r"(bar), "r"(phase) : "memory");
#undef TWS_GS_POLL
    phase ^= 1u;
  }
};

// Strips: an output row that a neighbouring strip needs as halo is stored a second time, straight into that
// neighbour's halo row (peer store over NVLink).  y: own local row; go: the element offset used for the own store.
__device__ __forceinline__ void push_flux(const BandEdge& e, int y, size_t go, const float4& fx, const float4& fy, const float4& fz, const float4& fw) {
  if (y < e.up_end) { st4(e.up[1] + go, fx); st4(e.up[2] + go, fy); st4(e.up[3] + go, fz); st4(e.up[4] + go, fw); }
  if (y >= e.down_begin) { st4(e.down[1] + go, fx); st4(e.down[2] + go, fy); st4(e.down[3] + go, fz); st4(e.down[4] + go, fw); }
}
__device__ __forceinline__ void push_depth(const BandEdge& e, int y, size_t go, const float4& d) {
  if (y < e.up_end) st4(e.up[0] + go, d);
  if (y >= e.down_begin) st4(e.down[0] + go, d);
}

// flowUpdate.comp:34-62 for the lane's 4 cells of each of its R rows.  Reads the neighbour rows' water
// level, leaves the new outflow in registers, publishes its +-Y components; LAST also stores the flux
// planes to HBM.  up / dn / me: 32-bit shared addresses of this lane's group in plane 0 of the slots.
// Interior rows (!EDGE) are split around the barrier: the +-X outflow and the friction products need
// only the row itself and are computed before the wait; the x components are scaled and the HBM stores
// issued after the arrive.  Operation order per value is that of flux_raw4 / flux_raw (cell_math.cuh).
template <int R, int SXW, bool EDGE, bool LAST>
__device__ __forceinline__ void band_flux(GroupSync& sy, const uint32_t (&up)[R], const uint32_t (&dn)[R], const uint32_t (&me)[R],
                                          const float4 (&h)[R], const float4 (&d)[R], float4 (&fx)[R], float4 (&fy)[R], float4 (&fz)[R],
                                          float4 (&fw)[R], const BandRow (&rc)[R], const int gx, const bool st_col, const FusedOut& out,
                                          const Geom& g, const StepConsts& c, const BandEdge& edge) {
  float4 HC[R], HU[R], HD[R];
  float HL[R], HR[R];
  float total[R][4], scale[R][4];
  bool need = false;
#pragma unroll
  for (int q = 0; q < R; ++q) HC[q] = add4(d[q], h[q]);                              // a + r, flowUpdate.comp:34
  // x neighbours sit in the adjacent lanes; the two strip-edge cells get a wrapped value: they lie in the
  // x halo, whose results are never kept
#pragma unroll
  for (int q = 0; q < R; ++q) {
    HL[q] = __shfl_up_sync(0xffffffffu, HC[q].w, 1);
    HR[q] = __shfl_down_sync(0xffffffffu, HC[q].x, 1);
  }
  if (!EDGE && TWS_PACKED) {
    const f2 PHI = pk(c.friction, c.friction), ALPHA = pk(c.accel, c.accel), KAPPA = pk(c.area_inv, c.area_inv);
    float4 tz[R], tw[R];
    f2 sxy_lo[R], sxy_hi[R];
    // ---- before the wait: +-X outflow (:44-45, :53-54), friction products of +-Y, fx + fy ----
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const float Dm = __fsub_rn(HL[q], HC[q].x);
      const float D0 = __fsub_rn(HC[q].x, HC[q].y), D1 = __fsub_rn(HC[q].y, HC[q].z), D2 = __fsub_rn(HC[q].z, HC[q].w),
                  D3 = __fsub_rn(HC[q].w, HR[q]);
      const float Pm = __fmul_rn(Dm, c.accel);
      float P0, P1, P2, P3;
      upk(mul2(pk(D0, D1), ALPHA), P0, P1);
      upk(mul2(pk(D2, D3), ALPHA), P2, P3);
      const float4 tx = cat4(mul2(lo2(fx[q]), PHI), mul2(hi2(fx[q]), PHI));
      const float4 ty = cat4(mul2(lo2(fy[q]), PHI), mul2(hi2(fy[q]), PHI));
      tz[q] = cat4(mul2(lo2(fz[q]), PHI), mul2(hi2(fz[q]), PHI));
      tw[q] = cat4(mul2(lo2(fw[q]), PHI), mul2(hi2(fw[q]), PHI));
      fx[q].x = max0(__fadd_rn(tx.x, P0)); fx[q].y = max0(__fadd_rn(tx.y, P1)); fx[q].z = max0(__fadd_rn(tx.z, P2)); fx[q].w = max0(__fadd_rn(tx.w, P3));
      fy[q].x = max0(__fsub_rn(ty.x, Pm)); fy[q].y = max0(__fsub_rn(ty.y, P0)); fy[q].z = max0(__fsub_rn(ty.z, P1)); fy[q].w = max0(__fsub_rn(ty.w, P2));
      sxy_lo[q] = add2(lo2(fx[q]), lo2(fy[q]));
      sxy_hi[q] = add2(hi2(fx[q]), hi2(fy[q]));
    }
    sy.wait();                                   // the rows above and below have published their water level
#pragma unroll
    for (int q = 0; q < R; ++q) { HU[q] = lds4(up[q]); HD[q] = lds4(dn[q]); }
#pragma unroll
    for (int q = 0; q < R; ++q) {
      float Z0, Z1, Z2, Z3, W0, W1, W2, W3;
      upk(mul2(sub2(lo2(HC[q]), lo2(HD[q])), ALPHA), Z0, Z1);                        // +Y, :46
      upk(mul2(sub2(hi2(HC[q]), hi2(HD[q])), ALPHA), Z2, Z3);
      upk(mul2(sub2(lo2(HC[q]), lo2(HU[q])), ALPHA), W0, W1);                        // -Y, :47
      upk(mul2(sub2(hi2(HC[q]), hi2(HU[q])), ALPHA), W2, W3);
      fz[q].x = max0(__fadd_rn(tz[q].x, Z0)); fz[q].y = max0(__fadd_rn(tz[q].y, Z1)); fz[q].z = max0(__fadd_rn(tz[q].z, Z2)); fz[q].w = max0(__fadd_rn(tz[q].w, Z3));
      fw[q].x = max0(__fadd_rn(tw[q].x, W0)); fw[q].y = max0(__fadd_rn(tw[q].y, W1)); fw[q].z = max0(__fadd_rn(tw[q].z, W2)); fw[q].w = max0(__fadd_rn(tw[q].w, W3));
      const f2 slo = add2(add2(sxy_lo[q], lo2(fz[q])), lo2(fw[q]));                  // ((fx + fy) + fz) + fw, :57
      const f2 shi = add2(add2(sxy_hi[q], hi2(fz[q])), hi2(fw[q]));
      upk(mul2(slo, KAPPA), total[q][0], total[q][1]);
      upk(mul2(shi, KAPPA), total[q][2], total[q][3]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float dep = comp(d[q], i);
        const bool over = total[q][i] > dep;                                         // :58
        scale[q][i] = over ? 0.0f : 1.0f;        // a == 0 -> a/total == +0 ; total <= a -> no scaling (x*1 == x)
        need = need || (over && dep != 0.0f);
      }
    }
  } else {
    sy.wait();
#pragma unroll
    for (int q = 0; q < R; ++q) { HU[q] = lds4(up[q]); HD[q] = lds4(dn[q]); }
#pragma unroll
    for (int q = 0; q < R; ++q) {
      float* pfx = &fx[q].x; float* pfy = &fy[q].x; float* pfz = &fz[q].x; float* pfw = &fw[q].x;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float Hc = comp(HC[q], i);
        float hxp = (i < 3) ? comp(HC[q], i + 1) : HR[q];
        float hxm = (i > 0) ? comp(HC[q], i - 1) : HL[q];
        float hyp = comp(HD[q], i), hym = comp(HU[q], i);
        if (EDGE && c.closed) {
          const int x = gx + i;
          if (x + 1 >= g.W) hxp = Hc;
          if (x - 1 < 0) hxm = Hc;
          if (rc[q].gy + 1 >= g.Hg) hyp = Hc;
          if (rc[q].gy - 1 < 0) hym = Hc;
        }
        total[q][i] = flux_raw(Hc, hxp, hxm, hyp, hym, pfx[i], pfy[i], pfz[i], pfw[i], c);
        const float dep = comp(d[q], i);
        const bool over = total[q][i] > dep;                                         // :58
        scale[q][i] = over ? 0.0f : 1.0f;
        need = need || (over && dep != 0.0f);
      }
    }
  }
  if (need) {                                    // the rare IEEE divisions: a wet cell that would drain completely
#pragma unroll
    for (int q = 0; q < R; ++q)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float dep = comp(d[q], i);
        if (total[q][i] > dep && dep != 0.0f) scale[q][i] = __fdiv_rn(dep, total[q][i]);   // :59
      }
  }
  // ---- scale and publish +-Y first: that is all the neighbour rows wait for ----
  // (interior rows: packed multiplies; the products are only ever added to in a later half-pass, behind
  // a barrier — tests/test_abi.py checks that ptxas did not contract any of them into an FFMA2)
#pragma unroll
  for (int q = 0; q < R; ++q) {
    if (!EDGE && TWS_PACKED) {
      const f2 slo = pk(scale[q][0], scale[q][1]), shi = pk(scale[q][2], scale[q][3]);
      fz[q] = cat4(mul2(lo2(fz[q]), slo), mul2(hi2(fz[q]), shi));
      fw[q] = cat4(mul2(lo2(fw[q]), slo), mul2(hi2(fw[q]), shi));
    } else {
      float* pfz = &fz[q].x; float* pfw = &fw[q].x;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        pfz[i] = __fmul_rn(pfz[i], scale[q][i]); pfw[i] = __fmul_rn(pfw[i], scale[q][i]);
        if (EDGE && !(rc[q].row_in && (unsigned)(gx + i) < (unsigned)g.W)) { pfz[i] = 0.f; pfw[i] = 0.f; }
      }
    }
    sts4(me[q] + 4 * SXW, fz[q]);                // plane 1: +Y outflow, read by the row below
    sts4(me[q] + 8 * SXW, fw[q]);                // plane 2: -Y outflow, read by the row above
  }
  sy.arrive();
#pragma unroll
  for (int q = 0; q < R; ++q) {
    if (!EDGE && TWS_PACKED) {
      const f2 slo = pk(scale[q][0], scale[q][1]), shi = pk(scale[q][2], scale[q][3]);
      fx[q] = cat4(mul2(lo2(fx[q]), slo), mul2(hi2(fx[q]), shi));
      fy[q] = cat4(mul2(lo2(fy[q]), slo), mul2(hi2(fy[q]), shi));
    } else {
      float* pfx = &fx[q].x; float* pfy = &fy[q].x;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        pfx[i] = __fmul_rn(pfx[i], scale[q][i]); pfy[i] = __fmul_rn(pfy[i], scale[q][i]);
        if (EDGE && !(rc[q].row_in && (unsigned)(gx + i) < (unsigned)g.W)) { pfx[i] = 0.f; pfy[i] = 0.f; }
      }
    }
    if (EDGE && rc[q].store && st_col) ledger_add(c, g, gx, rc[q].gy, fx[q], fy[q], fz[q], fw[q]);   // every sub-step, owner lanes only
    if (LAST && rc[q].store && st_col) {
      st4(out.F[0] + rc[q].go, fx[q]); st4(out.F[1] + rc[q].go, fy[q]); st4(out.F[2] + rc[q].go, fz[q]); st4(out.F[3] + rc[q].go, fw[q]);
      if (edge.nlev_edge) push_flux(edge, rc[q].gy - g.row0, rc[q].go, fx[q], fy[q], fz[q], fw[q]);
    }
  }
}

// flowApply.comp:38-46 with the source/sink extension compiled in or out.
template <bool EXT>
__device__ __forceinline__ float band_apply_cell(float depth, float fx, float fy, float fz, float fw, float iX1, float iX0, float iY1,
                                                 float iY0, const StepConsts& c, float& vx, float& vy, float& ds) {
  return apply_cell_src(depth, fx, fy, fz, fw, iX1, iX0, iY1, iY0, c, EXT, vx, vy, ds);
}

// flowApply.comp:32-52 for the lane's 4 cells of each of its R rows.  Reads the neighbour rows' +-Y
// outflow; not LAST: the new depth stays in registers and the new water level is published; LAST: depth
// and the packed fp16 flow vector go to HBM.  Interior rows: the x inflow, the outflow sum and the x
// flow component are computed before the wait; operation order per value is that of apply4.
template <int R, int SXW, bool EDGE, bool LAST, bool EXT>
__device__ __forceinline__ void band_depth(GroupSync& sy, const uint32_t (&up)[R], const uint32_t (&dn)[R], const uint32_t (&me)[R],
                                           const float4 (&h)[R], float4 (&d)[R], const float4 (&fx)[R], const float4 (&fy)[R],
                                           const float4 (&fz)[R], const float4 (&fw)[R], const BandRow (&rc)[R], const int gx,
                                           const bool st_col, const FusedOut& out, const Geom& g, const StepConsts& c, const BandEdge& edge,
                                           double& src_acc) {
  float4 iy1[R], iy0[R];
  float l[R], r[R];
#pragma unroll
  for (int q = 0; q < R; ++q) {
    l[q] = __shfl_up_sync(0xffffffffu, fx[q].w, 1);       // F(x-1,y).x (:33): the left cell's +X outflow
    r[q] = __shfl_down_sync(0xffffffffu, fy[q].x, 1);     // F(x+1,y).y (:32): the right cell's -X outflow
  }
  if (!EDGE && TWS_PACKED) {
    const f2 KAPPA = pk(c.area_inv, c.area_inv);
    f2 ax_lo[R], ax_hi[R], out_lo[R], out_hi[R];
    float vx[R][4];
#pragma unroll
    for (int q = 0; q < R; ++q) {
      // iX1 + iX0 = F(x+1,y).y + F(x-1,y).x (:38): neighbours inside the group are misaligned pairs -> scalar
      const float a0 = __fadd_rn(fy[q].y, l[q]), a1 = __fadd_rn(fy[q].z, fx[q].x), a2 = __fadd_rn(fy[q].w, fx[q].y), a3 = __fadd_rn(r[q], fx[q].z);
      ax_lo[q] = pk(a0, a1); ax_hi[q] = pk(a2, a3);
      out_lo[q] = add2(add2(add2(lo2(fx[q]), lo2(fy[q])), lo2(fz[q])), lo2(fw[q]));  // :39
      out_hi[q] = add2(add2(add2(hi2(fx[q]), hi2(fy[q])), hi2(fz[q])), hi2(fw[q]));
      if (LAST) {                                                                     // vx = (iX1 - fx) - (iX0 - fy), :45
        const float b0 = __fsub_rn(fy[q].y, fx[q].x), b1 = __fsub_rn(fy[q].z, fx[q].y), b2 = __fsub_rn(fy[q].w, fx[q].z), b3 = __fsub_rn(r[q], fx[q].w);
        const float c0 = __fsub_rn(l[q], fy[q].x), c1 = __fsub_rn(fx[q].x, fy[q].y), c2 = __fsub_rn(fx[q].y, fy[q].z), c3 = __fsub_rn(fx[q].z, fy[q].w);
        upk(sub2(pk(b0, b1), pk(c0, c1)), vx[q][0], vx[q][1]);
        upk(sub2(pk(b2, b3), pk(c2, c3)), vx[q][2], vx[q][3]);
      }
    }
    sy.wait();                                   // the rows above and below have published their +-Y outflow
#pragma unroll
    for (int q = 0; q < R; ++q) {
      iy1[q] = lds4(dn[q] + 8 * SXW);            // F(x,y+1).w, flowApply.comp:34
      iy0[q] = lds4(up[q] + 4 * SXW);            // F(x,y-1).z, :35
    }
    float4 nd[R];
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const f2 inlo = add2(add2(ax_lo[q], lo2(iy1[q])), lo2(iy0[q]));                // :38
      const f2 inhi = add2(add2(ax_hi[q], hi2(iy1[q])), hi2(iy0[q]));
      float q0, q1, q2, q3;
      upk(mul2(sub2(inlo, out_lo[q]), KAPPA), q0, q1);
      upk(mul2(sub2(inhi, out_hi[q]), KAPPA), q2, q3);
      nd[q].x = max0(__fadd_rn(d[q].x, q0)); nd[q].y = max0(__fadd_rn(d[q].y, q1));  // :41
      nd[q].z = max0(__fadd_rn(d[q].z, q2)); nd[q].w = max0(__fadd_rn(d[q].w, q3));
      if (EXT) {                                                                      // EXT (same expression as apply_cell)
        const f2 RAIN = pk(c.rain_step, c.rain_step), EVAP = pk(c.evap_step, c.evap_step);
        float e0, e1, e2, e3;
        upk(sub2(add2(lo2(nd[q]), RAIN), EVAP), e0, e1);
        upk(sub2(add2(hi2(nd[q]), RAIN), EVAP), e2, e3);
        const float4 pre = nd[q];
        nd[q] = make_float4(max0(e0), max0(e1), max0(e2), max0(e3));
        if (rc[q].store && st_col)                       // ledger of the sources: what they changed, owner lanes, every sub-step
          src_acc += ((double)__fsub_rn(nd[q].x, pre.x) + (double)__fsub_rn(nd[q].y, pre.y)) +
                     ((double)__fsub_rn(nd[q].z, pre.z) + (double)__fsub_rn(nd[q].w, pre.w));
      }
      if (!LAST) {
        d[q] = nd[q];
        sts4(me[q], add4(d[q], h[q]));
      }
    }
    sy.arrive();
    if (LAST) {
#pragma unroll
      for (int q = 0; q < R; ++q) {
        float vy0, vy1, vy2, vy3;                                                     // vy = (iY1 - fz) - (iY0 - fw), :46
        upk(sub2(sub2(lo2(iy1[q]), lo2(fz[q])), sub2(lo2(iy0[q]), lo2(fw[q]))), vy0, vy1);
        upk(sub2(sub2(hi2(iy1[q]), hi2(fz[q])), sub2(hi2(iy0[q]), hi2(fw[q]))), vy2, vy3);
        if (rc[q].store && st_col) {
          st4(out.d + rc[q].go, nd[q]);
          if (edge.nlev_edge) push_depth(edge, rc[q].gy - g.row0, rc[q].go, nd[q]);
          *reinterpret_cast<uint4*>(out.v + rc[q].go) =
              make_uint4(pack_half2(vx[q][0], vy0), pack_half2(vx[q][1], vy1), pack_half2(vx[q][2], vy2), pack_half2(vx[q][3], vy3));
        }
      }
    }
    return;
  }
  sy.wait();
#pragma unroll
  for (int q = 0; q < R; ++q) {
    iy1[q] = lds4(dn[q] + 8 * SXW);
    iy0[q] = lds4(up[q] + 4 * SXW);
  }
  float nds[R][4]; uint32_t nvs[R][4];
#pragma unroll
  for (int q = 0; q < R; ++q) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float iX1 = (i < 3) ? comp(fy[q], i + 1) : r[q];
      const float iX0 = (i > 0) ? comp(fx[q], i - 1) : l[q];
      float vx, vy, ds;
      nds[q][i] = band_apply_cell<EXT>(comp(d[q], i), comp(fx[q], i), comp(fy[q], i), comp(fz[q], i), comp(fw[q], i), iX1, iX0,
                                       comp(iy1[q], i), comp(iy0[q], i), c, vx, vy, ds);
      nvs[q][i] = LAST ? pack_half2(vx, vy) : 0u;
      if (EDGE && !(rc[q].row_in && (unsigned)(gx + i) < (unsigned)g.W)) { nds[q][i] = 0.f; nvs[q][i] = 0u; ds = 0.f; }
      if (EXT && rc[q].store && st_col) src_acc += (double)ds;
    }
    if (!LAST) {
      d[q] = make_float4(nds[q][0], nds[q][1], nds[q][2], nds[q][3]);
      sts4(me[q], add4(d[q], h[q]));
    }
  }
  sy.arrive();
  if (LAST) {
#pragma unroll
    for (int q = 0; q < R; ++q)
      if (rc[q].store && st_col) {
        st4(out.d + rc[q].go, make_float4(nds[q][0], nds[q][1], nds[q][2], nds[q][3]));
        if (edge.nlev_edge) push_depth(edge, rc[q].gy - g.row0, rc[q].go, make_float4(nds[q][0], nds[q][1], nds[q][2], nds[q][3]));
        *reinterpret_cast<uint4*>(out.v + rc[q].go) = make_uint4(nvs[q][0], nvs[q][1], nvs[q][2], nvs[q][3]);
      }
  }
}

// Poll a flag in the own control block until the neighbour has published epoch >= value.  Bounded: after ~20 s
// the wait gives up and raises ctrl->error, so a dead peer cannot hang the GPU.
__device__ __noinline__ void band_wait_flag(const volatile uint32_t* flag, uint32_t value, uint32_t* error) {
  if ((int32_t)(*flag - value) >= 0) return;
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while ((int32_t)(*flag - value) < 0) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 20000000000ull) { *error = 1u; return; }
    __nanosleep(100);
  }
}

template <class C, bool EXT>
__global__ void __launch_bounds__(C::NT, 1) band_step_kernel(const __grid_constant__ CUtensorMap tm_h,
                                                             const __grid_constant__ CUtensorMap tm_s,
                                                             const __grid_constant__ BandSched sch,
                                                             const __grid_constant__ BandEdge edge,
                                                             FusedOut out, Geom g, StepConsts c, int lr0, int lr1, int nstrips,
                                                             int tma_y_bias, uint32_t* sched) {
  constexpr int K = C::K, NW = C::NW, R = C::R, BR = C::BR, SXW = C::SXW, HX = C::HX, OX = C::OX, HP = C::HP, LAND = C::LAND, XROW = C::XROW;
  extern __shared__ __align__(1024) float smem[];
  __shared__ uint64_t full_all[NW * C::NGRP];
  __shared__ uint64_t gbar_all[C::NGRP];
  __shared__ int next_piece[C::NGRP][2];
  // read through volatile asm: the thread index stays in a register instead of being re-read (S2R, ~20
  // cycles of latency) and re-derived in every half-pass
  uint32_t tid_u;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid_u));
  const int tid = (int)tid_u, lane = tid & 31, grp = (tid >> 5) / NW, warp = (tid >> 5) - grp * NW;   // warp: index inside the group
  float* gsm = smem + grp * C::GROUP_FLOATS;              // this group's landing buffers | exchange slots | parking
  uint64_t* full = full_all + grp * NW + warp;
  GroupSync sy;
  sy.bar = smem_u32(&gbar_all[grp]);
  sy.phase = 0;
  float* land = gsm + warp * (R * LAND);                  // R landing buffers of this warp
  const float* land_l = land + lane * 4;
  const uint32_t xch = smem_u32(gsm + BR * LAND + lane * 4);
  const int pw = (NW - 1 - warp) < HP ? (NW - 1 - warp) : HP;     // half-passes the last row gets inside its own band
  const bool carrier = pw < HP;
  double src_acc = 0.0;                                   // EXT ledger of the sources: per-thread partial sum, flushed at the end
  // the half-pass before which a carrier switches to the row it parked one band ago (0: never).  Pinned in a register:
  // left to the compiler it is re-derived from the warp index (8 instructions) in front of every half-pass
  const uint32_t swap_at = keep_u32(carrier ? (uint32_t)(pw + 1) : 0u);
  float* park = gsm + BR * LAND + C::NSLOT * XROW + (carrier ? warp - (NW - C::NC) : 0) * LAND + lane * 4;

  // Exchange slots.  Band row o = q*NW + warp uses slot o, except the last NDB rows of a band, which
  // alternate between two slots by band parity.  The row above o = 0 is the last row of the previous
  // band, the row below o = BR-1 the first row of the next one.
  auto slot_off = [&](int o, int& dbl) {                  // byte offset of the first slot of band row o; dbl: its parity stride
    const bool db = o >= BR - C::NDB;
    dbl = db ? XROW * 4 : 0;
    return (db ? (BR - C::NDB) + 2 * (o - (BR - C::NDB)) : o) * XROW * 4;
  };
  int sb_me[R], db_me[R], sb_up[R], db_up[R], ju[R], sb_dn[R], db_dn[R], jd[R];
#pragma unroll
  for (int q = 0; q < R; ++q) {
    const int o = q * NW + warp;
    sb_me[q] = slot_off(o, db_me[q]);
    ju[q] = o == 0 ? -1 : 0;
    sb_up[q] = slot_off(o == 0 ? BR - 1 : o - 1, db_up[q]);
    jd[q] = o == BR - 1 ? 1 : 0;
    sb_dn[q] = slot_off(o == BR - 1 ? 0 : o + 1, db_dn[q]);
  }

  if (tid == 0) {
#pragma unroll 1
    for (int i = 0; i < NW * C::NGRP; ++i) mbar_init(&full_all[i], 1);
#pragma unroll 1
    for (int i = 0; i < C::NGRP; ++i) mbar_init(&gbar_all[i], NW);      // one arrival per warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();                                        // the only CTA-wide barrier: the mbarriers exist
  sy.arrive();                                            // phase 0: "nobody is reading any exchange slot" (matched by the first wait)
  uint32_t parity = 0;

  // every group is a virtual CTA; its first piece is static, the following ones come from the counter
  const int vcta = (int)blockIdx.x * C::NGRP + grp, nvcta = (int)gridDim.x * C::NGRP;
  int piece = vcta, pp = 0;

#pragma unroll 1
  while (piece < sch.npieces) {
    // fetch the piece after this one now: the atomic's round trip hides behind the first TMA load of the piece (any
    // later and the fetching warp — and with it the group — waits for it: ptxas aggregates the atomic over the warp
    // and broadcasts the result with a shuffle right behind it, so the value cannot be left in flight); the group
    // reads the slot behind the piece's barriers (release / acquire), and the two slots alternate so that the next
    // write cannot overtake a reader.  Committing one piece ahead is harmless because pieces shrink along the list:
    // a group never holds more than a fraction of its fair share — which is also why a strip's tiny edge pieces come
    // LAST in the list (first, every group would fetch twice within microseconds and end up holding two long pieces).
    if (sched && warp == 0 && lane == 0) next_piece[grp][pp] = nvcta + (int)atomicAdd(sched, 1u);
    int lev, strip, ya_rel, yb_rel;
    band_decode(sch, piece, nstrips, lev, strip, ya_rel, yb_rel);
    const int ya = lr0 + ya_rel, yb = lr0 + yb_rel;
    const bool edge_piece = edge.nlev_edge != 0 && lev >= edge.lev_edge0;
    if (edge_piece) {
      // the halo rows this piece reads were stored by the neighbour's previous block: wait for its flag (it is
      // normally long there), then order this warp's loads — generic and TMA — behind the observation
      if (lane == 0) {
        if (edge.wait_up != nullptr && ya - HP < 0) band_wait_flag(edge.wait_up, edge.wait_value, edge.error);
        if (edge.wait_down != nullptr && yb + HP > g.rows) band_wait_flag(edge.wait_down, edge.wait_value, edge.error);
        __threadfence_system();
        asm volatile("fence.proxy.async;" ::: "memory");
      }
      __syncwarp();
    }
    const int sx0 = strip * OX - HX;
    const int ystart = ya - HP;                           // 2K warm-up rows above, 2K feeder rows below
    const int N = (yb - ya) + 2 * HP;
    // Bands of the piece.  The rows a band leaves unfinished (its last 2K, carried into the next band) are in
    // the last band always feeder rows or rows past the piece (N <= J*BR puts row N-2K at or before the first
    // carried row J*BR-2K), and a feeder at distance t below the last output row needs only 2K-t half-passes,
    // which its place in the window gives it — so no further band is needed to finish anything that is kept
    // (model-checked in tests/test_band_schedule.py).
    const int J = (N + BR - 1) / BR;
    const bool xedge = sx0 < 0 || sx0 + SXW > g.W;
    const int gx = sx0 + lane * 4;
    const bool st_col = lane * 4 >= HX && lane * 4 < HX + OX && gx < g.pitch;

    auto issue = [&](int i0) {                            // one lane: land the warp's R rows of the band starting at piece row i0
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(full, (uint32_t)(R * LAND * sizeof(float)));
#pragma unroll
      for (int q = 0; q < R; ++q) {
        const int ty = ystart + i0 + q * NW + warp + tma_y_bias;       // rows past the piece or the grid: fetched or zero-filled, never used
        tma_load_2d(land + q * LAND, &tm_h, sx0, ty, full);
        tma_load_3d(land + q * LAND + SXW, &tm_s, sx0, ty, 0, full);   // d, F+X, F-X, F+Y, F-Y in one operation
      }
    };
    if (lane == 0 && warp < N) issue(0);

    // the rows this warp is working on: registers, row context, where they exchange with their neighbours
    float4 h[R], d[R], fx[R], fy[R], fz[R], fw[R];
    BandRow rc[R];
    uint32_t x_me[R], x_up[R], x_dn[R];
    bool plain = false;                                   // all R rows are interior rows of the global grid (no masks needed)
    auto set_row = [&](int q, int jj, int i) {            // slot q <- piece row i of band jj
      const int y = ystart + i;
      rc[q].gy = g.row0 + y;
      rc[q].row_in = (unsigned)rc[q].gy < (unsigned)g.Hg;
      rc[q].store = y >= ya && y < yb;
      rc[q].go = (size_t)((long long)y * g.pitch) + gx;             // only dereferenced when rc.store (y >= 0)
      x_me[q] = keep_u32(xch + sb_me[q] + ((jj & 1) ? db_me[q] : 0));
      x_up[q] = keep_u32(xch + sb_up[q] + (((jj + ju[q]) & 1) ? db_up[q] : 0));
      x_dn[q] = keep_u32(xch + sb_dn[q] + (((jj + jd[q]) & 1) ? db_dn[q] : 0));
    };
    auto classify = [&]() {
      bool e = xedge;
#pragma unroll
      for (int q = 0; q < R; ++q) e = e || rc[q].gy <= 0 || rc[q].gy >= g.Hg - 1;
      plain = !e;
    };
    // park the new last row, resume the row parked one band ago (plane by plane through one temporary)
    auto swap_rows = [&](int jj) {
      constexpr int q = R - 1;
      float4 t;
      t = ld4(park);           st4(park, h[q]);            h[q] = t;
      t = ld4(park + SXW);     st4(park + SXW, d[q]);      d[q] = t;
      t = ld4(park + 2 * SXW); st4(park + 2 * SXW, fx[q]); fx[q] = t;
      t = ld4(park + 3 * SXW); st4(park + 3 * SXW, fy[q]); fy[q] = t;
      t = ld4(park + 4 * SXW); st4(park + 4 * SXW, fz[q]); fz[q] = t;
      t = ld4(park + 5 * SXW); st4(park + 5 * SXW, fw[q]); fw[q] = t;
      set_row(q, jj - 1, (jj - 1) * BR + q * NW + warp);            // band -1 (first swap of a piece): rows above the piece, never stored
      classify();
    };

#pragma unroll 1
    for (int j = 0; j < J; ++j) {
      const int i0 = j * BR;
#pragma unroll
      for (int q = 0; q < R; ++q) set_row(q, j, i0 + q * NW + warp);
      classify();
      // ---- half-pass 0: registers <- landing buffers; publish H; prefetch this warp's rows of the next band ----
      const bool loaded = i0 + warp < N;
      if (loaded) {
        mbar_wait(full, parity);
        parity ^= 1u;
#pragma unroll
        for (int q = 0; q < R; ++q) {
          const float* lq = land_l + q * LAND;
          h[q] = ld4(lq);            d[q] = ld4(lq + SXW);
          fx[q] = ld4(lq + 2 * SXW); fy[q] = ld4(lq + 3 * SXW);
          fz[q] = ld4(lq + 4 * SXW); fw[q] = ld4(lq + 5 * SXW);
        }
        __syncwarp();
        if (lane == 0 && i0 + BR + warp < N) issue(i0 + BR);
      }
      sy.wait();                                          // the last half-pass of the previous band has read the slots these rows reuse
      if (loaded) {
#pragma unroll
        for (int q = 0; q < R; ++q) sts4(x_me[q], add4(d[q], h[q]));
      }
      sy.arrive();

      // ---- the 2K computing half-passes.  Carriers switch their last row once, before half-pass pw + 1;
      // rows on the grid edge take the masked variants ----
#pragma unroll 1
      for (uint32_t lv = 1; lv < (uint32_t)K; ++lv) {
        if (2 * lv - 1 == swap_at) swap_rows(j);
        if (plain) band_flux<R, SXW, false, false>(sy, x_up, x_dn, x_me, h, d, fx, fy, fz, fw, rc, gx, st_col, out, g, c, edge);
        else band_flux<R, SXW, true, false>(sy, x_up, x_dn, x_me, h, d, fx, fy, fz, fw, rc, gx, st_col, out, g, c, edge);
        if (2 * lv == swap_at) swap_rows(j);
        if (plain) band_depth<R, SXW, false, false, EXT>(sy, x_up, x_dn, x_me, h, d, fx, fy, fz, fw, rc, gx, st_col, out, g, c, edge, src_acc);
        else band_depth<R, SXW, true, false, EXT>(sy, x_up, x_dn, x_me, h, d, fx, fy, fz, fw, rc, gx, st_col, out, g, c, edge, src_acc);
      }
      if (HP - 1 == swap_at) swap_rows(j);
      if (plain) band_flux<R, SXW, false, true>(sy, x_up, x_dn, x_me, h, d, fx, fy, fz, fw, rc, gx, st_col, out, g, c, edge);
      else band_flux<R, SXW, true, true>(sy, x_up, x_dn, x_me, h, d, fx, fy, fz, fw, rc, gx, st_col, out, g, c, edge);
      if (HP == swap_at) swap_rows(j);
      if (plain) band_depth<R, SXW, false, true, EXT>(sy, x_up, x_dn, x_me, h, d, fx, fy, fz, fw, rc, gx, st_col, out, g, c, edge, src_acc);
      else band_depth<R, SXW, true, true, EXT>(sy, x_up, x_dn, x_me, h, d, fx, fy, fz, fw, rc, gx, st_col, out, g, c, edge, src_acc);
    }
    if (edge_piece) {
      // every lane's stores into the neighbours' halos are visible system-wide before this warp counts itself done;
      // the warp that completes the last edge piece of the launch publishes the new epoch to the neighbours
      __threadfence_system();
      __syncwarp();
      if (lane == 0 && atomicAdd(sched + 2, 1u) == edge.n_edge_warps - 1u) {
        sched[2] = 0u;
        __threadfence_system();
        if (edge.post_up != nullptr) *edge.post_up = edge.post_value;
        if (edge.post_down != nullptr) *edge.post_down = edge.post_value;
        __threadfence_system();
      }
    }
    piece = sched ? *(volatile int*)&next_piece[grp][pp] : piece + nvcta;
    pp ^= 1;
  }
  if (EXT) ledger_src_flush(c.ledger_src, src_acc);
  // the last group to leave re-arms the counters for the next launch on this stream (every group that had a
  // piece has by then seen its final, failing fetch, so no atomic on the counter is still in flight)
  if (sched && vcta < sch.npieces && warp == 0 && lane == 0) {
    const unsigned active = (unsigned)(nvcta < sch.npieces ? nvcta : sch.npieces);
    __threadfence();
    if (atomicAdd(sched + 1, 1u) == active - 1u) { sched[0] = 0u; sched[1] = 0u; }
  }
}

// ---- host side ---------------------------------------------------------------------------
#ifndef TWS_BAND_R
#define TWS_BAND_R 1             // rows per warp and band
#endif
#ifndef TWS_BAND_NGRP
#define TWS_BAND_NGRP 2          // independent warp groups per CTA
#endif
#ifndef TWS_BAND_NW
#define TWS_BAND_NW 24           // warps per CTA (all groups)
#endif
// a group must be deeper than the dependency cone (2K + 1 warps): fewer groups for the larger K
constexpr int band_groups(int K) {
  return TWS_BAND_NW / TWS_BAND_NGRP >= 2 * K + 1 ? TWS_BAND_NGRP : (TWS_BAND_NW / 2 >= 2 * K + 1 ? 2 : 1);
}
template <int K> struct BandCfgFor { using type = BandCfg<K, TWS_BAND_NW / band_groups(K), TWS_BAND_R, band_groups(K)>; };

static int band_sm_count() {
  static int cached[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!cached[dev & 63]) cudaDeviceGetAttribute(&cached[dev & 63], cudaDevAttrMultiProcessorCount, dev);
  return cached[dev & 63] > 0 ? cached[dev & 63] : 148;
}

template <int K, bool EXT>
static cudaError_t launch_band_k(const Geom& g, const Planes& p, const TmaSet& tma, int src, const StepConsts& c, int lr0, int lr1,
                                 cudaStream_t st, uint32_t* sched, int cta_budget, const BandEdge* edge, int e_top, int e_bot) {
  using C = typename BandCfgFor<K>::type;
  if (C::SXW != stream_strip_width()) return cudaErrorInvalidValue;      // ring and band kernels share the row descriptors
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  auto kern = band_step_kernel<C, EXT>;
  if (!attr_set[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    if (e != cudaSuccess) return e;
    attr_set[dev & 63] = true;
  }
  if (lr1 <= lr0) return cudaSuccess;
  if (edge != nullptr && sched == nullptr) return cudaErrorInvalidValue;
  const int dst = 1 - src;
  const size_t row0_off = (size_t)TWS_HALO_ROWS * g.pitch;
  FusedOut out;
  out.d = p.d[dst] + row0_off;
  for (int i = 0; i < 4; ++i) out.F[i] = p.F[dst][i] + row0_off;
  out.v = p.v + row0_off;
  const int nstrips = (g.W + C::OX - 1) / C::OX;
  // one persistent CTA per SM (fewer when there are fewer pieces than warp groups); cta_budget > 0: at most that
  // many CTAs, < 0: leave that many SMs free
  const int all_sms = band_sm_count();
  const int sms = cta_budget > 0 ? std::min(cta_budget, all_sms) : std::max(1, all_sms + cta_budget);
  BandEdge ed{};
  if (edge != nullptr) ed = *edge;
  int lev_edge0 = 0, nlev_edge = 0, edge_segs = 0;
  // a strip's block (edge != nullptr): the interior, then the edge bands — tiny pieces that fill the tail of the
  // launch; the neighbours need them only for THEIR edge pieces at the end of their next block
  const BandSched sch = band_build_schedule(lr1 - lr0, nstrips, sms * C::NGRP, C::BR, C::HP, edge != nullptr ? e_top : 0,
                                            edge != nullptr ? e_bot : 0, &lev_edge0, &nlev_edge, &edge_segs);
  if (edge != nullptr) {
    ed.lev_edge0 = lev_edge0;
    ed.nlev_edge = nlev_edge;
    ed.n_edge_warps = (unsigned)edge_segs * (unsigned)nstrips * (unsigned)C::NW;
  }
  const int want = (sch.npieces + C::NGRP - 1) / C::NGRP;
  const int grid = want < 1 ? 1 : (want < sms ? want : sms);
  const int bias = g.has_up ? TWS_HALO_ROWS : 0;
  kern<<<grid, C::NT, C::SMEM, st>>>(tma.m[0], tma.m[1], sch, ed, out, g, c, lr0, lr1, nstrips, bias, sched);
  return cudaGetLastError();
}

cudaError_t launch_band(int K, const Geom& g, const Planes& p, const TmaSet& tma, int src, const StepConsts& c, int lr0, int lr1,
                        cudaStream_t st, uint32_t* sched, int cta_budget, const BandEdge* edge, int e_top, int e_bot) {
  const bool ext = c.ext_sources != 0;
  switch (K) {
    case 1: return ext ? launch_band_k<1, true>(g, p, tma, src, c, lr0, lr1, st, sched, cta_budget, edge, e_top, e_bot) : launch_band_k<1, false>(g, p, tma, src, c, lr0, lr1, st, sched, cta_budget, edge, e_top, e_bot);
    case 2: return ext ? launch_band_k<2, true>(g, p, tma, src, c, lr0, lr1, st, sched, cta_budget, edge, e_top, e_bot) : launch_band_k<2, false>(g, p, tma, src, c, lr0, lr1, st, sched, cta_budget, edge, e_top, e_bot);
    case 3: return ext ? launch_band_k<3, true>(g, p, tma, src, c, lr0, lr1, st, sched, cta_budget, edge, e_top, e_bot) : launch_band_k<3, false>(g, p, tma, src, c, lr0, lr1, st, sched, cta_budget, edge, e_top, e_bot);
    case 4: return ext ? launch_band_k<4, true>(g, p, tma, src, c, lr0, lr1, st, sched, cta_budget, edge, e_top, e_bot) : launch_band_k<4, false>(g, p, tma, src, c, lr0, lr1, st, sched, cta_budget, edge, e_top, e_bot);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace tws
