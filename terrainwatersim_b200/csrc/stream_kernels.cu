// stream_kernels.cu — the row-streaming, warp-autonomous form of the fused step (sm_100a).
//
// Same arithmetic as every other step kernel (flowUpdate.comp:12-63 + flowApply.comp:14-53,
// contract in cell_math.cuh / DESIGN.md section 2) and bit-identical results; what differs is how K
// whole steps are blocked over one HBM round trip:
//
//   * The grid is cut into column strips SXW = 128*G cells wide (OX = SXW - 2*HX of them are
//     outputs, HX = x halo of the K-step dependency cone) and every strip is marched top to
//     bottom.  A persistent CTA (one per SM) owns a contiguous piece of the (strip, row) work
//     list; all pieces have the same number of rows, so any grid size balances over 148 SMs.
//   * ONE WARP OWNS ONE ROW (lane L holds the float4 groups L, L+32, ...) for the row's whole life:
//     the TMA unit lands the row's h, d and four flux planes in the warp's private landing buffer
//     (`cp.async.bulk.tensor.2d`, mbarrier completion, out-of-bounds zero fill = the reference's
//     exterior rule), the warp pulls them into registers, and the 2K half-passes (flux_1, depth_1,
//     ... flux_K, depth_K) update them IN REGISTERS.  The moment a landing buffer has been read
//     the same warp issues the TMA load of ITS next row (row + NW), which lands while the
//     current row is being computed.
//   * Rows are skewed in time instead of recomputed: row y runs half-pass s as soon as rows y-1
//     and y+1 have finished half-pass s-1.  Only what a vertical neighbour needs goes through
//     shared memory (the water level H = d + h, the +Y and -Y outflow: 12 B per cell and slot);
//     x neighbours are adjacent lanes (warp shuffles).  A warp waits only for its two neighbour
//     slots, never for the CTA: there is no __syncthreads in the row loop, and warps in
//     different phases overlap their shared-memory, FP and store work.
//   * No y-halo recomputation: each piece pays 2K warm-up rows once, and HBM sees 48/K B per
//     cell-update plus the x-halo re-reads.
#include "cell_math.cuh"
#include "band_schedule.h"

#include <cstdio>

namespace tws {

template <int K_, int NW_, int G_>
struct StreamCfg {
  static constexpr int K = K_, NW = NW_, NT = NW_ * 32, G = G_;
  static constexpr int SXW = 128 * G;                   // staged strip width: G float4 groups per lane
  static constexpr int HX = stream_hx(K);               // x halo rounded to whole float4 groups
  static constexpr int OX = SXW - 2 * HX;               // output columns per strip
  static constexpr int NHP = 2 * K + 1;                 // half-passes per row, the load included
  static constexpr int LAND = 6 * SXW;                  // floats per landing buffer (h, d, F x4)
  static constexpr int XROW = 3 * SXW;                  // floats per exchange slot (H, F+Y, F-Y)
  static constexpr size_t SMEM = (size_t)NW * (LAND + XROW) * sizeof(float);
  static_assert(NW > 2 * K, "the row ring must be deeper than the dependency cone");
};

// ------------------------------------------------------------------------------------------
// Neighbour synchronisation: event mbarriers.  One mbarrier per (slot, half-pass, ring-turn parity), arrived on once per
// row, so a barrier completes once every second row of its slot and its phase parity is bit 1 of the slot's ring-turn
// count.  A waiting warp sleeps in hardware (mbarrier.try_wait with a suspend-time hint).  A parity wait is exact only
// while the awaited phase is the one in progress or the one just completed, i.e. while the neighbour slot is less than two
// ring turns away from the awaited row in either direction — which the dependency structure guarantees (with a single
// barrier per (slot, half-pass) it does not: in the first turn slot 0 can be a whole turn behind or ahead of the last
// slot).  Every slot arrives on all of its barriers in every ring turn of a piece, also for half-passes a row skips, so
// the phase counts are a pure function of the turn count and the barriers live across pieces.
// tests/test_stream_protocol.py model-checks this protocol.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kStreamSuspendNs = 20000;   // suspend-time hint of the neighbour waits

__device__ __forceinline__ void mbar_arrive(uint32_t addr) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}

// Per-row synchronisation cursor: advanced by one half-pass per wait() / signal().
struct RowSync {
  uint32_t up, dn, me;    // shared-memory addresses: neighbours' event barrier (or progress word), own
  uint32_t pu, pd;        // awaited phase parities
  bool up_on;             // false for the first row of a piece (nobody above)

  // rows idx-1 and idx+1 have finished the half-pass before the one about to run
  // One barrier, one tight PTX loop: try_wait (hardware suspend), loop back on failure.  A waiting
  // warp re-issues 5 instructions per wake-up instead of the 14 of the two-barrier C loop below —
  // waiting warps share their scheduler with working ones, so every spin instruction is stolen from a
  // row that could make progress.  The bounded spin count turns a protocol bug into a trap, not a hang.
  static __device__ __forceinline__ void wait_one(uint32_t bar, uint32_t par) {
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        ".reg .u32 n;\n"
        "mov.u32 n, 0;\n"
        "TWS_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra TWS_WAIT_DONE;\n"
        "add.u32 n, n, 1;\n"
        "setp.gt.u32 q, n, 4194304;\n"
        "@q trap;\n"
        "bra TWS_WAIT_LOOP;\n"
        "TWS_WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(par), "r"(kStreamSuspendNs) : "memory");
  }

  __device__ __forceinline__ void wait() {
    // the row below is the one that lags (it started later); the row above is almost always done
    wait_one(dn, pd);
    if (up_on) wait_one(up, pu);
    up += 16; dn += 16;                                  // evt[slot][s][turn & 1] -> evt[slot][s + 1][turn & 1]
    __syncwarp();
  }
  // this row has finished a half-pass: every lane's shared-memory stores are ordered before the release
  __device__ __forceinline__ void signal(int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(me);
    me += 16;
  }
};

template <int G>
struct RowCtx {
  int gy;          // global row
  int gx[G];       // global x of the lane's groups
  bool row_in;     // the row exists in the global grid
  bool store;      // the row is an output row of this piece
  bool st[G];      // the lane's groups are output columns
  size_t go[G];    // element offsets of the groups in the output planes
};

// x neighbours of the lane's groups; `w3` / `x0` are component 3 / 0 of the value in each group.
// Lane L owns groups L (and L+32 when G = 2), so left(a) = lane-1's a.w, left(b) = lane-1's b.w but
// lane 0 takes lane 31's a.w; right(a) = lane+1's a.x but lane 31 takes lane 0's b.x, right(b) =
// lane+1's b.x.  The two strip-edge cells get their own value back: they lie in the x halo, whose
// results are never kept.
template <int G>
__device__ __forceinline__ void x_neighbours(const float (&w3)[G], const float (&x0)[G], int lane, float (&l)[G], float (&r)[G]) {
  if (G == 1) {
    l[0] = __shfl_up_sync(0xffffffffu, w3[0], 1);
    r[0] = __shfl_down_sync(0xffffffffu, x0[0], 1);
  } else {
    l[0] = __shfl_up_sync(0xffffffffu, w3[0], 1);
    l[G - 1] = __shfl_sync(0xffffffffu, lane == 31 ? w3[0] : w3[G - 1], (lane + 31) & 31);
    r[0] = __shfl_sync(0xffffffffu, lane == 0 ? x0[G - 1] : x0[0], (lane + 1) & 31);
    r[G - 1] = __shfl_down_sync(0xffffffffu, x0[G - 1], 1);
  }
}

// flowApply.comp:38-46 with the source/sink extension compiled in or out.
template <bool EXT>
__device__ __forceinline__ float apply_cell_t(float depth, float fx, float fy, float fz, float fw, float iX1, float iX0, float iY1,
                                              float iY0, const StepConsts& c, float& vx, float& vy, float& ds) {
  return apply_cell_src(depth, fx, fy, fz, fw, iX1, iX0, iY1, iY0, c, EXT, vx, vy, ds);
}

// flowUpdate.comp:34-62 for the lane's 4*G cells.  Reads the neighbour rows' water level, leaves the
// new outflow in registers, publishes its +-Y components; LAST also stores the flux planes to HBM.
// The three slot addresses (32-bit shared window, bytes) are this lane's first group in plane 0 (H) of the slot.
template <int G, int SXW, bool EDGE, bool LAST>
__device__ __forceinline__ void stream_flux(const uint32_t up, const uint32_t dn, const uint32_t me, const int lane,
                                            const float4 (&h)[G], const float4 (&d)[G], float4 (&fx)[G], float4 (&fy)[G], float4 (&fz)[G],
                                            float4 (&fw)[G], const RowCtx<G>& rc, const FusedOut& out, const Geom& g, const StepConsts& c,
                                            double& out_acc) {
  float4 HC[G], HU[G], HD[G];
  float w3[G], x0[G], HL[G], HR[G];
#pragma unroll
  for (int q = 0; q < G; ++q) {
    HC[q] = add4(d[q], h[q]);                                                        // a + r, flowUpdate.comp:34
    HU[q] = lds4(up + 512 * q);
    HD[q] = lds4(dn + 512 * q);
    w3[q] = HC[q].w; x0[q] = HC[q].x;
  }
  x_neighbours<G>(w3, x0, lane, HL, HR);
  float total[G][4], scale[G][4];
  bool need = false;
#pragma unroll
  for (int q = 0; q < G; ++q) {
    if (!EDGE && TWS_PACKED) {
      flux_raw4(HC[q], HU[q], HD[q], HL[q], HR[q], fx[q], fy[q], fz[q], fw[q], c, total[q]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float dep = comp(d[q], i);
        const bool over = total[q][i] > dep;                                         // :58
        scale[q][i] = over ? 0.0f : 1.0f;
        need = need || (over && dep != 0.0f);
      }
      continue;
    }
    float* pfx = &fx[q].x; float* pfy = &fy[q].x; float* pfz = &fz[q].x; float* pfw = &fw[q].x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float Hc = comp(HC[q], i);
      float hxp = (i < 3) ? comp(HC[q], i + 1) : HR[q];
      float hxm = (i > 0) ? comp(HC[q], i - 1) : HL[q];
      float hyp = comp(HD[q], i), hym = comp(HU[q], i);
      if (EDGE && c.closed) {
        const int gx = rc.gx[q] + i;
        if (gx + 1 >= g.W) hxp = Hc;
        if (gx - 1 < 0) hxm = Hc;
        if (rc.gy + 1 >= g.Hg) hyp = Hc;
        if (rc.gy - 1 < 0) hym = Hc;
      }
      total[q][i] = flux_raw(Hc, hxp, hxm, hyp, hym, pfx[i], pfy[i], pfz[i], pfw[i], c);
      const float dep = comp(d[q], i);
      const bool over = total[q][i] > dep;                                           // :58
      scale[q][i] = over ? 0.0f : 1.0f;          // a == 0 -> a/total == +0 ; total <= a -> no scaling (x*1 == x)
      need = need || (over && dep != 0.0f);
    }
  }
  if (need) {                                    // the rare IEEE divisions: a wet cell that would drain completely
#pragma unroll
    for (int q = 0; q < G; ++q)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float dep = comp(d[q], i);
        if (total[q][i] > dep && dep != 0.0f) scale[q][i] = __fdiv_rn(dep, total[q][i]);   // :59
      }
  }
#pragma unroll
  for (int q = 0; q < G; ++q) {
    float* pfx = &fx[q].x; float* pfy = &fy[q].x; float* pfz = &fz[q].x; float* pfw = &fw[q].x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      pfx[i] = __fmul_rn(pfx[i], scale[q][i]); pfy[i] = __fmul_rn(pfy[i], scale[q][i]);
      pfz[i] = __fmul_rn(pfz[i], scale[q][i]); pfw[i] = __fmul_rn(pfw[i], scale[q][i]);
      if (EDGE && !(rc.row_in && (unsigned)(rc.gx[q] + i) < (unsigned)g.W)) { pfx[i] = 0.f; pfy[i] = 0.f; pfz[i] = 0.f; pfw[i] = 0.f; }
    }
    sts4(me + 4 * SXW + 512 * q, fz[q]);         // plane 1: +Y outflow, read by the row below
    sts4(me + 8 * SXW + 512 * q, fw[q]);         // plane 2: -Y outflow, read by the row above
    if (EDGE && rc.store && rc.st[q]) ledger_acc(out_acc, c, g, rc.gx[q], rc.gy, fx[q], fy[q], fz[q], fw[q]);   // every sub-step, owner lanes only
    if (LAST && rc.store && rc.st[q]) {
      st4(out.F[0] + rc.go[q], fx[q]); st4(out.F[1] + rc.go[q], fy[q]); st4(out.F[2] + rc.go[q], fz[q]); st4(out.F[3] + rc.go[q], fw[q]);
    }
  }
}

// flowApply.comp:32-52 for the lane's 4*G cells.  Reads the neighbour rows' +-Y outflow; not LAST:
// new depth stays in registers and the new water level is published; LAST: depth and the packed
// fp16 flow vector go to HBM.
template <int G, int SXW, bool EDGE, bool LAST, bool EXT>
__device__ __forceinline__ void stream_depth(const uint32_t up, const uint32_t dn, const uint32_t me, const int lane,
                                             const float4 (&h)[G], float4 (&d)[G], const float4 (&fx)[G], const float4 (&fy)[G],
                                             const float4 (&fz)[G], const float4 (&fw)[G], const RowCtx<G>& rc, const FusedOut& out,
                                             const Geom& g, const StepConsts& c, double& src_acc) {
  float4 iy1[G], iy0[G];
  float w3[G], x0[G], l[G], r[G];
#pragma unroll
  for (int q = 0; q < G; ++q) {
    iy1[q] = lds4(dn + 8 * SXW + 512 * q);     // F(x,y+1).w, flowApply.comp:34
    iy0[q] = lds4(up + 4 * SXW + 512 * q);     // F(x,y-1).z, :35
    w3[q] = fx[q].w; x0[q] = fy[q].x;
  }
  // F(x-1,y).x (:33) is the left cell's +X outflow, F(x+1,y).y (:32) the right cell's -X outflow
  x_neighbours<G>(w3, x0, lane, l, r);
#pragma unroll
  for (int q = 0; q < G; ++q) {
    float nd[4]; uint32_t nv[4];
    float4 ds4 = make_float4(0.f, 0.f, 0.f, 0.f);                                     // EXT ledger: what the sources changed
    if (!EDGE && TWS_PACKED) {
      float4 nd4; uint4 nv4 = make_uint4(0u, 0u, 0u, 0u);
      apply4<LAST>(d[q], fx[q], fy[q], fz[q], fw[q], l[q], r[q], iy1[q], iy0[q], c, EXT, nd4, nv4, EXT ? &ds4 : nullptr);
      nd[0] = nd4.x; nd[1] = nd4.y; nd[2] = nd4.z; nd[3] = nd4.w;
      nv[0] = nv4.x; nv[1] = nv4.y; nv[2] = nv4.z; nv[3] = nv4.w;
    } else {
      float* pds = &ds4.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float iX1 = (i < 3) ? comp(fy[q], i + 1) : r[q];
      const float iX0 = (i > 0) ? comp(fx[q], i - 1) : l[q];
      float vx, vy;
      nd[i] = apply_cell_t<EXT>(comp(d[q], i), comp(fx[q], i), comp(fy[q], i), comp(fz[q], i), comp(fw[q], i), iX1, iX0, comp(iy1[q], i),
                                comp(iy0[q], i), c, vx, vy, pds[i]);
      if (LAST) nv[i] = pack_half2(vx, vy);
      if (EDGE && !(rc.row_in && (unsigned)(rc.gx[q] + i) < (unsigned)g.W)) { nd[i] = 0.f; pds[i] = 0.f; if (LAST) nv[i] = 0u; }
    }
    }
    if (EXT && rc.store && rc.st[q]) src_acc += ((double)ds4.x + (double)ds4.y) + ((double)ds4.z + (double)ds4.w);   // every sub-step, owner lanes only
    if (!LAST) {
      d[q] = make_float4(nd[0], nd[1], nd[2], nd[3]);
      sts4(me + 512 * q, add4(d[q], h[q]));
    } else if (rc.store && rc.st[q]) {
      st4(out.d + rc.go[q], make_float4(nd[0], nd[1], nd[2], nd[3]));
      *reinterpret_cast<uint4*>(out.v + rc.go[q]) = make_uint4(nv[0], nv[1], nv[2], nv[3]);
    }
  }
}

template <class C, bool EXT>
__global__ void __launch_bounds__(C::NT, 1) stream_step_kernel(const __grid_constant__ CUtensorMap tm_h,
                                                               const __grid_constant__ CUtensorMap tm_s,
                                                               const __grid_constant__ BandSched sch,
                                                               FusedOut out, Geom g, StepConsts c, int lr0, int lr1, int nstrips,
                                                               int tma_y_bias, uint32_t* sched) {
  constexpr int K = C::K, NW = C::NW, G = C::G, SXW = C::SXW, HX = C::HX, OX = C::OX, NHP = C::NHP, LAND = C::LAND;
  extern __shared__ __align__(1024) float smem[];
  __shared__ uint64_t full[NW];
  __shared__ int next_piece[2];
  __shared__ __align__(16) uint64_t evt[NW][NHP][2];      // per slot, half-pass and ring-turn parity
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wup = (warp + NW - 1) % NW, wdn = (warp + 1) % NW;
  double src_acc = 0.0;                                 // EXT ledger of the sources: per-thread partial sum, flushed at the end
  double out_acc = 0.0;                                 // ... and of the boundary outflow (lanes that own an edge cell)

  // lane-resolved pointers: landing buffer; exchange slots [NW][H | F+Y | F-Y][SXW] of this row and its neighbours
  float* land = smem + warp * LAND;
  const float* land_l = land + lane * 4;
  const uint32_t xch = smem_u32(smem + NW * LAND + lane * 4);
  const uint32_t x_me = xch + warp * C::XROW * 4;
  const uint32_t x_up = xch + wup * C::XROW * 4;
  const uint32_t x_dn = xch + wdn * C::XROW * 4;

  if (tid == 0) {
#pragma unroll 1
    for (int i = 0; i < NW; ++i) mbar_init(&full[i], 1);
#pragma unroll 1
    for (int i = 0; i < NW * NHP * 2; ++i) mbar_init(&evt[0][0][0] + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t parity = 0;                                    // phase of this warp's landing barrier
  uint32_t turn_base = 0;                                 // ring turns completed by earlier pieces (event-barrier phases)

  // Pieces (row segment x column strip) come from the same guided work list as the band kernel's (band_schedule.h): the CTA
  // starts on the piece of its own index and then takes pieces from a device-wide counter, fetched one piece ahead and handed
  // to the other warps through shared memory behind the block barriers between pieces.  Round 1 split the list into equal
  // static shares: a launch then lasted as long as its slowest CTA — measured, the two edge column strips (masked scalar
  // arithmetic, about 13/8 of the cost of an interior row) and the lake / shore regions (IEEE divisions) set the pace.
  int piece = (int)blockIdx.x, pp = 0;
#pragma unroll 1
  while (piece < sch.npieces) {
    if (sched != nullptr && tid == 0) next_piece[pp] = (int)gridDim.x + (int)atomicAdd(sched, 1u);
    int lev, strip, ya_rel, yb_rel;
    band_decode(sch, piece, nstrips, lev, strip, ya_rel, yb_rel);
    const int ya = lr0 + ya_rel, yb = lr0 + yb_rel;
    const int sx0 = strip * OX - HX;
    const int ystart = ya - 2 * K;                        // 2K warm-up rows above, 2K feeder rows below
    const int nrows = (yb - ya) + 4 * K;
    const int turns = (nrows + NW - 1) / NW;
    const bool xedge = sx0 < 0 || sx0 + SXW > g.W;

    __syncthreads();                                      // every warp is done with the previous piece
    __syncthreads();

    auto issue = [&](int y) {                             // one lane: land row y in this warp's buffer
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&full[warp], (uint32_t)(LAND * sizeof(float)));
      const int ty = y + tma_y_bias;
      tma_load_2d(land, &tm_h, sx0, ty, &full[warp]);

      tma_load_3d(land + SXW, &tm_s, sx0, ty, 0, &full[warp]);     // d, F+X, F-X, F+Y, F-Y in one operation
    };
    if (lane == 0 && warp < nrows) issue(ystart + warp);

    RowCtx<G> rc;
#pragma unroll
    for (int q = 0; q < G; ++q) {
      const int o = (lane + 32 * q) * 4;
      rc.gx[q] = sx0 + o;
      rc.st[q] = o >= HX && o < HX + OX && rc.gx[q] < g.pitch;
    }

#pragma unroll 1
    for (int t = 0; t < turns; ++t) {
      const int idx = t * NW + warp;
      const uint32_t turn = turn_base + (uint32_t)t;
      RowSync sy;
      // neighbours' turns: the row above a slot-0 row lives in the previous turn, the row below a last-slot row in the next
      const uint32_t tu = warp == 0 ? turn - 1u : turn, td = warp == NW - 1 ? turn + 1u : turn;
      sy.up = smem_u32(&evt[wup][0][tu & 1u]); sy.pu = (tu >> 1) & 1u;
      sy.dn = smem_u32(&evt[wdn][0][td & 1u]); sy.pd = (td >> 1) & 1u;
      sy.me = smem_u32(&evt[warp][0][turn & 1u]);
      if (idx >= nrows) {                                 // no row for this slot in the last turn: keep the phase counts uniform
        if (lane == 0)
          for (int e = 0; e < NHP; ++e) mbar_arrive(sy.me + 16 * e);
        break;
      }
      sy.up_on = idx > 0;
      const int y = ystart + idx;                         // local row
      rc.gy = g.row0 + y;
      rc.row_in = (unsigned)rc.gy < (unsigned)g.Hg;
      rc.store = y >= ya && y < yb;
      const size_t rowoff = (size_t)((long long)y * g.pitch);      // only dereferenced when rc.store (y >= 0)
#pragma unroll
      for (int q = 0; q < G; ++q) rc.go[q] = rowoff + rc.gx[q];
      // rows below the piece only feed the rows above them: row yb-1+m stops after half-pass 2K-m
      const int smax = (y < yb) ? 2 * K : 2 * K - (y - yb + 1);
      const bool edge = xedge || rc.gy <= 0 || rc.gy >= g.Hg - 1;

      // ---- half-pass 0: registers <- landing buffer; publish H; prefetch this warp's next row ----
      float4 h[G], d[G], fx[G], fy[G], fz[G], fw[G];
      mbar_wait(&full[warp], parity);
      parity ^= 1u;
#pragma unroll
      for (int q = 0; q < G; ++q) {
        h[q] = ld4(land_l + 128 * q);            d[q] = ld4(land_l + SXW + 128 * q);
        fx[q] = ld4(land_l + 2 * SXW + 128 * q); fy[q] = ld4(land_l + 3 * SXW + 128 * q);
        fz[q] = ld4(land_l + 4 * SXW + 128 * q); fw[q] = ld4(land_l + 5 * SXW + 128 * q);
      }
      __syncwarp();
      if (lane == 0 && idx + NW < nrows) issue(y + NW);
#pragma unroll
      for (int q = 0; q < G; ++q) sts4(x_me + 512 * q, add4(d[q], h[q]));
      sy.signal(lane);

      if (!edge && smax == 2 * K) {
        // ---- the common row: interior, all 2K half-passes, no masks ----------------------------
#pragma unroll 1
        for (int lv = 1; lv < K; ++lv) {
          sy.wait();
          stream_flux<G, SXW, false, false>(x_up, x_dn, x_me, lane, h, d, fx, fy, fz, fw, rc, out, g, c, out_acc);
          sy.signal(lane);
          sy.wait();
          stream_depth<G, SXW, false, false, EXT>(x_up, x_dn, x_me, lane, h, d, fx, fy, fz, fw, rc, out, g, c, src_acc);
          sy.signal(lane);
        }
        sy.wait();
        stream_flux<G, SXW, false, true>(x_up, x_dn, x_me, lane, h, d, fx, fy, fz, fw, rc, out, g, c, out_acc);
        sy.signal(lane);
        sy.wait();
        stream_depth<G, SXW, false, true, EXT>(x_up, x_dn, x_me, lane, h, d, fx, fy, fz, fw, rc, out, g, c, src_acc);
        sy.signal(lane);
      } else {
        // ---- rows on the grid edge (exterior masks, boundary mode) and feeder rows that stop early ----
        int s = 1;
#pragma unroll 1
        for (; s <= smax; ++s) {
          sy.wait();
          const bool last = s >= 2 * K - 1;
          if (s & 1) {
            if (last) stream_flux<G, SXW, true, true>(x_up, x_dn, x_me, lane, h, d, fx, fy, fz, fw, rc, out, g, c, out_acc);
            else stream_flux<G, SXW, true, false>(x_up, x_dn, x_me, lane, h, d, fx, fy, fz, fw, rc, out, g, c, out_acc);
          } else {
            if (last) stream_depth<G, SXW, true, true, EXT>(x_up, x_dn, x_me, lane, h, d, fx, fy, fz, fw, rc, out, g, c, src_acc);
            else stream_depth<G, SXW, true, false, EXT>(x_up, x_dn, x_me, lane, h, d, fx, fy, fz, fw, rc, out, g, c, src_acc);
          }
          sy.signal(lane);
        }
        if (lane == 0)                                    // the half-passes this feeder row skipped
          for (; s <= 2 * K; ++s) { mbar_arrive(sy.me); sy.me += 16; }
      }
    }
    turn_base += (uint32_t)turns;
    __syncthreads();                                      // next_piece[pp] was written before the barriers at the top of this piece
    piece = sched != nullptr ? next_piece[pp] : piece + (int)gridDim.x;
    pp ^= 1;
  }
  // the last CTA to leave re-arms the counters for the next launch on this stream (every CTA that had a piece has by then seen
  // its final, failing fetch)
  if (sched != nullptr && (int)blockIdx.x < sch.npieces && tid == 0) {
    const unsigned active = (unsigned)((int)gridDim.x < sch.npieces ? (int)gridDim.x : sch.npieces);
    __threadfence();
    if (atomicAdd(sched + 1, 1u) == active - 1u) { sched[0] = 0u; sched[1] = 0u; }
  }
  if (EXT) ledger_src_flush(c.ledger_src, src_acc);
  ledger_src_flush(c.ledger, out_acc);
}

// ---- host side ---------------------------------------------------------------------------
#ifndef TWS_STREAM_G
#define TWS_STREAM_G 1           // float4 groups per lane: rows of 128 * G cells
#endif
#ifndef TWS_STREAM_NW
#define TWS_STREAM_NW (24 / TWS_STREAM_G)     // row slots (warps) per SM: 24 x 80 registers keeps the row state spill-free
#endif
template <int K> struct StreamCfgFor { using type = StreamCfg<K, TWS_STREAM_NW, TWS_STREAM_G>; };

int stream_strip_width() { return StreamCfgFor<1>::type::SXW; }

static int stream_sm_count() {
  static int cached[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!cached[dev & 63]) cudaDeviceGetAttribute(&cached[dev & 63], cudaDevAttrMultiProcessorCount, dev);
  return cached[dev & 63] > 0 ? cached[dev & 63] : 148;
}

template <int K, bool EXT>
static cudaError_t launch_stream_k(const Geom& g, const Planes& p, const TmaSet& tma, int src, const StepConsts& c, int lr0, int lr1,
                                   cudaStream_t st, uint32_t* sched) {
  using C = typename StreamCfgFor<K>::type;
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  auto kern = stream_step_kernel<C, EXT>;
  if (!attr_set[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    if (e != cudaSuccess) return e;
    attr_set[dev & 63] = true;
  }
  if (lr1 <= lr0) return cudaSuccess;
  const int dst = 1 - src;
  const size_t row0_off = (size_t)TWS_HALO_ROWS * g.pitch;
  FusedOut out;
  out.d = p.d[dst] + row0_off;
  for (int i = 0; i < 4; ++i) out.F[i] = p.F[dst][i] + row0_off;
  out.v = p.v + row0_off;
  const int nstrips = (g.W + C::OX - 1) / C::OX;
  // one persistent CTA per SM (fewer when there are fewer pieces); a piece computes rows + 4K rows in whole ring turns
  const int sms = stream_sm_count();
  const BandSched sch = band_build_schedule(lr1 - lr0, nstrips, sms, C::NW, 2 * K, 0, 0, nullptr, nullptr, nullptr);
  const int grid = sch.npieces < 1 ? 1 : (sch.npieces < sms ? sch.npieces : sms);
  const int bias = g.has_up ? TWS_HALO_ROWS : 0;
  kern<<<grid, C::NT, C::SMEM, st>>>(tma.m[0], tma.m[1], sch, out, g, c, lr0, lr1, nstrips, bias, sched);
  return cudaGetLastError();
}

// impl: 0 = ring kernel (per-row barriers), 1 = band kernel (CTA-synchronous bands)
cudaError_t launch_stream(int K, const Geom& g, const Planes& p, const TmaSet& tma, int src, const StepConsts& c, int lr0, int lr1,
                          cudaStream_t st, int impl, uint32_t* sched, int cta_budget) {
  const bool ext = c.ext_sources != 0;
  if (impl == 1) return launch_band(K, g, p, tma, src, c, lr0, lr1, st, sched, cta_budget);
  switch (K) {
    case 1: return ext ? launch_stream_k<1, true>(g, p, tma, src, c, lr0, lr1, st, sched) : launch_stream_k<1, false>(g, p, tma, src, c, lr0, lr1, st, sched);
    case 2: return ext ? launch_stream_k<2, true>(g, p, tma, src, c, lr0, lr1, st, sched) : launch_stream_k<2, false>(g, p, tma, src, c, lr0, lr1, st, sched);
    case 3: return ext ? launch_stream_k<3, true>(g, p, tma, src, c, lr0, lr1, st, sched) : launch_stream_k<3, false>(g, p, tma, src, c, lr0, lr1, st, sched);
    case 4: return ext ? launch_stream_k<4, true>(g, p, tma, src, c, lr0, lr1, st, sched) : launch_stream_k<4, false>(g, p, tma, src, c, lr0, lr1, st, sched);
    default: return cudaErrorInvalidValue;
  }
}

// Row descriptors: box = one SXW-cell row segment.  Same visibility rule as the tile kernel's maps
// (own rows plus the halo rows towards an existing neighbour; everything else zero-filled).
cudaError_t stream_build_tma(const Geom& g, const Planes& p, int side, TmaSet* out, std::string* err) {
  return build_tma_rows(g, p, side, StreamCfgFor<1>::type::SXW, out, err, 0);
}

}  // namespace tws
