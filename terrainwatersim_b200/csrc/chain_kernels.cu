// chain_kernels.cu — the step as a REGISTER PIPELINE of warps ("chain kernel", backend CHAIN_TB, sm_100a).
//
// Same arithmetic as every other step kernel (flowUpdate.comp:12-63 + flowApply.comp:14-53, contract in cell_math.cuh /
// DESIGN.md section 2) and bit-identical results.  Like the ring and band kernels it blocks K whole steps over one HBM
// round trip by skewing rows in time, but it needs NO neighbour synchronisation inside a step:
//
//   * a column strip (128 cells, one float4 group per lane) is streamed top to bottom through a CHAIN of K warps, warp s
//     applying step s+1.  A warp takes the rows one after the other and keeps three of them in registers: when row t
//     arrives it computes the new outflow of row t-1 (its vertical neighbours t-2 and t are in the same registers — no
//     shared-memory exchange, no barrier) and then the new depth of row t-2 (which needs the outflow of t-3, four
//     registers kept from the previous tick, and of t-1, just computed), and passes row t-2 on.  x neighbours are
//     adjacent lanes (two shuffles per half-pass);
//   * consecutive warps of a chain are connected by a ring of D row slots in shared memory (h, d and the four outflow
//     planes of one row: 3 KB) with a full / empty mbarrier pair per slot: producer and consumer run asynchronously, a
//     warp only ever waits for ITS OWN link, never for a group — the lock step of the band kernel (one barrier per
//     half-pass over 12 warps) and its exchange slots, parking buffers and carrier rows are gone, and with them a third of
//     the instructions;
//   * the first warp of a chain is fed by TMA (terrain: a 2-D box, the five state planes: one 3-D box per row) D rows
//     ahead into its ring; out-of-bounds zero fill is the reference's exterior rule;
//   * the last warp stores depth, outflow and the packed fp16 flow vector straight from registers;
//   * pieces (row segment x column strip) come from the same guided work list as the band kernel's (band_schedule.h),
//     taken from a device-wide counter.  A chain streams its pieces back to back: every row carries its own context
//     (row, strip, store / edge flags) down the chain, so the pipeline never drains between pieces; the first and last
//     2K rows of a piece are warm-up / feeder rows whose results are not stored, which also absorbs the mismatched
//     neighbours at a piece boundary.
//
// ONE arithmetic path serves every cell.  Cells outside the grid (zero-filled by TMA) keep an outflow of exactly +0
// without any mask: their depth is 0, so whatever raw outflow the gradient gives them is scaled by 0 (flowUpdate.comp:58-59
// with a = 0), and a cell inside the grid sees H = 0 there — the reference's out-of-range imageLoad.  Only the depth
// update needs a mask (water flowing off the map must not collect in the exterior), applied on edge rows / edge strips
// alone; the closed-boundary extension replaces the exterior neighbour's level by the cell's own before the same code runs.
#include "cell_math.cuh"
#include "band_schedule.h"

#include <algorithm>

namespace tws {

template <int K_, int NWARPS_, int D_>
struct ChainCfg {
  static constexpr int K = K_, NCH = NWARPS_ / K_, NW = NCH * K_, NT = NW * 32, D = D_;
  static constexpr int SXW = 128;                       // one float4 group per lane
  static constexpr int HX = stream_hx(K);
  static constexpr int OX = SXW - 2 * HX;
  static constexpr int HP = 2 * K;                      // warm-up rows above / feeder rows below a piece
  static constexpr int ROW = 6 * SXW;                   // floats per ring slot: h, d, F+X, F-X, F+Y, F-Y of one row
  static constexpr size_t SMEM = ((size_t)NCH * K * D + 1) * ROW * sizeof(float);      // the rings + one row of zeros
  static_assert(D >= 2 && D <= 4, "a piece (>= 4K + 1 rows) must be longer than the ring is deep");
  static_assert(SMEM <= 227 * 1024, "chain configuration does not fit in shared memory");
};

// per-row context, carried down the chain with the row
constexpr uint32_t CH_VALID = 1u, CH_STORE = 2u, CH_EDGE = 4u, CH_END = 8u;
struct ChainRow {
  float4 h, d, fx, fy, fz, fw;
  int y;                // local row (relative to the strip's first own row; halo / exterior rows are negative or >= rows)
  uint32_t meta;        // CH_* flags | column-strip index << 8
};

__device__ __forceinline__ void chain_zero(ChainRow& r) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  r.h = z; r.d = z; r.fx = z; r.fy = z; r.fz = z; r.fw = z; r.y = 0; r.meta = 0u;
}

__device__ __forceinline__ void mbar_wait_lean(uint32_t bar, uint32_t parity) {
#define TWS_CH_POLL "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra TWS_CH_DONE;\n"
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      ".reg .u32 n;\n"
      "mov.u32 n, 0;\n"
      "TWS_CH_LOOP:\n"
      TWS_CH_POLL TWS_CH_POLL TWS_CH_POLL TWS_CH_POLL
      "add.u32 n, n, 1;\n"
      "setp.gt.u32 q, n, 4194304;\n"
      "@q trap;\n"                                        // a link that never fills is a bug; a trap beats a hung GPU
      "bra TWS_CH_LOOP;\n"
      "TWS_CH_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
#undef TWS_CH_POLL
}
__device__ __forceinline__ void mbar_arrive_release(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// ---- rare paths, kept out of line: the per-tick code must stay small enough for the instruction cache ----------------
__device__ __noinline__ float chain_div(float a, float b) { return __fdiv_rn(a, b); }            // :59, a wet cell that would drain completely
__device__ __noinline__ void chain_ledger(double* acc, float area_inv, int W, int Hg, int gx0, int gy, float4 fx, float4 fy, float4 fz, float4 fw) {
  StepConsts c{};
  c.ledger = acc; c.area_inv = area_inv;
  Geom g{};
  g.W = W; g.Hg = Hg;
  ledger_add(c, g, gx0, gy, fx, fy, fz, fw);
}

// flowUpdate.comp:34-62 for the lane's 4 cells of row A; HU / HD: water level of the rows above / below.
__device__ __forceinline__ void chain_flux(ChainRow& A, float4 HC, float4 HU, float4 HD, const int gx, const Geom& g, const StepConsts& c) {
  float HL = __shfl_up_sync(0xffffffffu, HC.w, 1);       // strip-edge lanes get a wrapped value: they lie in the x halo
  float HR = __shfl_down_sync(0xffffffffu, HC.x, 1);
  if (c.closed && (A.meta & CH_EDGE)) {                   // EXT closed wall: an exterior neighbour reads as the cell itself
    const int gy = g.row0 + A.y;
    if (gy - 1 < 0) HU = HC;
    if (gy + 1 >= g.Hg) HD = HC;
    if (gx == 0) HL = HC.x;
    const int iw = g.W - 1 - gx;                          // component that holds x = W-1 (if any)
    if (iw == 3) HR = HC.w;
    else if (iw == 2) HC.w = HC.z;                        // the exterior cell right of it takes its level: H[i] - H[i+1] == +0
    else if (iw == 1) HC.z = HC.y;
    else if (iw == 0) HC.y = HC.x;
  }
  float total[4], scale[4];
  flux_raw4(HC, HU, HD, HL, HR, A.fx, A.fy, A.fz, A.fw, c, total);
  bool need = false;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float dep = comp(A.d, i);
    const bool over = total[i] > dep;                     // :58
    scale[i] = over ? 0.0f : 1.0f;                        // a == 0 -> a/total == +0 ; total <= a -> no scaling (x*1 == x)
    need = need || (over && dep != 0.0f);
  }
  if (need) {                                             // one warp-level branch for the cells that really divide
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float dep = comp(A.d, i);
      if (total[i] > dep && dep != 0.0f) scale[i] = chain_div(dep, total[i]);
    }
  }
  const f2 slo = pk(scale[0], scale[1]), shi = pk(scale[2], scale[3]);
  A.fx = cat4(mul2(lo2(A.fx), slo), mul2(hi2(A.fx), shi));
  A.fy = cat4(mul2(lo2(A.fy), slo), mul2(hi2(A.fy), shi));
  A.fz = cat4(mul2(lo2(A.fz), slo), mul2(hi2(A.fz), shi));
  A.fw = cat4(mul2(lo2(A.fw), slo), mul2(hi2(A.fw), shi));
}

// flowApply.comp:32-52 for the lane's 4 cells of row B; iy0: +Y outflow of the row above, iy1: -Y outflow of the row below.
template <bool EXT>
__device__ __forceinline__ void chain_depth(ChainRow& B, const float4& iy0, const float4& iy1, const int gx, const Geom& g, const StepConsts& c,
                                            const bool last, uint4& nv) {
  const float l = __shfl_up_sync(0xffffffffu, B.fx.w, 1);      // F(x-1,y).x (:33)
  const float r = __shfl_down_sync(0xffffffffu, B.fy.x, 1);    // F(x+1,y).y (:32)
  float4 nd;
  if (last) apply4<true>(B.d, B.fx, B.fy, B.fz, B.fw, l, r, iy1, iy0, c, EXT, nd, nv);     // warp-uniform: the last warp of a chain also derives the flow vector
  else apply4<false>(B.d, B.fx, B.fy, B.fz, B.fw, l, r, iy1, iy0, c, EXT, nd, nv);
  if (B.meta & CH_EDGE) {                                      // nothing collects outside the grid (and in the pad columns)
    const bool row_in = (unsigned)(g.row0 + B.y) < (unsigned)g.Hg;
    const bool i0 = row_in && (unsigned)(gx + 0) < (unsigned)g.W, i1 = row_in && (unsigned)(gx + 1) < (unsigned)g.W;
    const bool i2 = row_in && (unsigned)(gx + 2) < (unsigned)g.W, i3 = row_in && (unsigned)(gx + 3) < (unsigned)g.W;
    nd.x = i0 ? nd.x : 0.f; nd.y = i1 ? nd.y : 0.f; nd.z = i2 ? nd.z : 0.f; nd.w = i3 ? nd.w : 0.f;
    nv.x = i0 ? nv.x : 0u; nv.y = i1 ? nv.y : 0u; nv.z = i2 ? nv.z : 0u; nv.w = i3 ? nv.w : 0u;
  }
  B.d = nd;
}

// Everything one warp of a chain needs to know about its two links.
struct ChainLink {
  uint32_t in_rows;     // shared address of the input ring (slot 0), this lane's float4 group
  uint32_t in_full;     // shared address of full[0] of the input link
  uint32_t in_empty;    // ... empty[0] (links between warps only)
  uint32_t in_hdr;      // shared address of the header of input slot 0
  uint32_t out_rows, out_full, out_empty, out_hdr;
};

// One warp of a chain: ONE code path for every position (`first`: fed by TMA, `last`: stores to HBM; both warp-uniform).
template <class C, bool EXT>
__device__ __forceinline__ void chain_stage(const bool first, const bool last, const ChainLink& lk, const CUtensorMap* tm_h,
                                            const CUtensorMap* tm_s, const BandSched& sch, const FusedOut& out, const Geom& g,
                                            const StepConsts& c, const int lr0, const int nstrips, const int tma_y_bias, uint32_t* sched,
                                            const int vchain, const int nvchain, const int lane, float* ring0, uint64_t* full0,
                                            const uint32_t zero_row /* shared address of a row of zeros, this lane's group */) {
  constexpr int D = C::D, ROWB = C::ROW * 4, SXW = C::SXW, HX = C::HX, OX = C::OX, HP = C::HP;
  // ---- first warp: two cursors over the same piece sequence — the one TMA is issued at runs up to D rows ahead of the one
  // rows are consumed at.  A piece is longer than D rows, so at most one piece boundary lies between them: one queued id.
  int c_piece = -1, c_i = 0, c_n = 0, c_ystart = 0, c_ya = 0, c_yb = 0;      // consume cursor
  uint32_t c_meta = 0;                                  // strip << 8 | CH_EDGE of an edge strip
  int i_piece = -1, i_i = 0, i_n = 0, i_sx0 = 0, i_ystart = 0;              // issue cursor
  int queued = -2;                                      // id the issue cursor moved on to before the consume cursor (-2: none yet)
  int prefetched = -1;                                  // id fetched from the counter one piece ahead
  uint32_t issued = 0;                                  // rows issued so far (slot = issued % D)
  auto fetch_next_id = [&](int after) -> int {          // warp-uniform: the piece that follows `after` for this chain (>= npieces: none)
    if (sched == nullptr) return after + nvchain;       // static round robin
    int v = 0;
    if (lane == 0) v = nvchain + (int)atomicAdd(sched, 1u);
    return __shfl_sync(0xffffffffu, v, 0);
  };
  auto open_issue_piece = [&](int piece) {
    i_piece = piece; i_i = 0;
    if (piece < sch.npieces) {
      int lev, strip, ya, yb;
      band_decode(sch, piece, nstrips, lev, strip, ya, yb);
      i_sx0 = strip * OX - HX;
      i_ystart = lr0 + ya - HP;
      i_n = (yb - ya) + 2 * HP;
      // commit to the piece after this one now: the atomic's round trip hides behind this piece's rows
      prefetched = fetch_next_id(piece);
    } else {
      i_n = 0;
    }
  };
  auto open_consume_piece = [&](int piece) {
    c_piece = piece; c_i = 0;
    if (piece < sch.npieces) {
      int lev, strip, ya, yb;
      band_decode(sch, piece, nstrips, lev, strip, ya, yb);
      const int sx0 = strip * OX - HX;
      c_meta = ((uint32_t)strip << 8) | ((sx0 < 0 || sx0 + SXW > g.W) ? CH_EDGE : 0u);
      c_ya = lr0 + ya; c_yb = lr0 + yb;
      c_ystart = c_ya - HP;
      c_n = (yb - ya) + 2 * HP;
    } else {
      c_n = 0;
    }
  };
  auto issue_one = [&]() {                              // land the next row of the issue cursor in slot issued % D (no-op at the end of the list)
    if (i_piece >= sch.npieces) return;
    const uint32_t slot = issued % D;
    if (lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      uint64_t* bar = full0 + slot;
      float* dst = ring0 + slot * C::ROW;
      mbar_expect_tx(bar, (uint32_t)ROWB);
      const int ty = i_ystart + i_i + tma_y_bias;
      tma_load_2d(dst, tm_h, i_sx0, ty, bar);
      tma_load_3d(dst + SXW, tm_s, i_sx0, ty, 0, bar);  // d, F+X, F-X, F+Y, F-Y in one operation
    }
    ++issued;
    if (++i_i == i_n) {                                 // this piece is issued completely: move on
      const int nxt = prefetched;
      if (queued == -2) queued = nxt;
      open_issue_piece(nxt);
    }
  };
  uint32_t consumed = 0, emitted = 0;
  if (first) {
    open_issue_piece(vchain);
    open_consume_piece(vchain);
#pragma unroll 1
    for (int j = 0; j < D; ++j) issue_one();
  }

  const int gx_in_strip = lane * 4 - HX;
  const bool st_lane = lane * 4 >= HX && lane * 4 < HX + OX;
  float4 fzp = make_float4(0.f, 0.f, 0.f, 0.f);         // +Y outflow (after this warp's flux) of the row above B
  int tail = 0;                                         // first warp: rows handed out after the end of the piece list (END, then blanks)

  // ---- one tick: row t sits in the input ring; new outflow of row A (t-1); new depth of row B (t-2); B leaves and its
  // registers take row t.  The stream ends with an END row followed by blank rows (the pipeline lag of the next warp);
  // they travel and are computed like any row — their values are never stored and, by the warm-up / feeder margin of a
  // piece, never reach a stored value.  Returns false once the END row has left. ----
  auto tick = [&](ChainRow& A, ChainRow& B) -> bool {
    // -- the incoming row: only its water level is needed before B has left
    uint32_t a = lk.in_rows + (consumed % D) * ROWB;
    int ny = 0; uint32_t nmeta = 0;
    const bool from_ring = !(first && c_piece >= sch.npieces);
    if (from_ring) {
      mbar_wait_lean(lk.in_full + (consumed % D) * 8, (consumed / D) & 1u);
      if (first) {
        ny = c_ystart + c_i;
        const int gy = g.row0 + ny;
        nmeta = c_meta | ((ny >= c_ya && ny < c_yb) ? CH_STORE : 0u) | ((gy <= 0 || gy >= g.Hg - 1) ? CH_EDGE : 0u);
      } else {
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(ny), "=r"(nmeta) : "r"(lk.in_hdr + (consumed % D) * 8) : "memory");
      }
    } else {                                            // the first warp makes the tail of the stream itself: rows of zeros
      a = zero_row;
      nmeta = tail == 0 ? CH_END : 0u;
      ++tail;
    }
    const float4 HN = add4(lds4(a + SXW * 4), lds4(a));                  // a + r, flowUpdate.comp:34
    const float4 HA = add4(A.d, A.h), HB = add4(B.d, B.h);              // (B.d is still this step's input depth)
    const int gxA = (int)(A.meta >> 8) * OX + gx_in_strip;
    chain_flux(A, HA, HB, HN, gxA, g, c);
    if ((A.meta & (CH_EDGE | CH_STORE)) == (CH_EDGE | CH_STORE) && c.ledger != nullptr && st_lane && gxA < g.pitch)
      chain_ledger(c.ledger, c.area_inv, g.W, g.Hg, gxA, g.row0 + A.y, A.fx, A.fy, A.fz, A.fw);   // every sub-step, owner lanes only
    uint4 nv;
    const int gxB = (int)(B.meta >> 8) * OX + gx_in_strip;
    const float4 fz_b = B.fz;
    chain_depth<EXT>(B, fzp, A.fw, gxB, g, c, last, nv);
    fzp = fz_b;
    // -- B leaves: to HBM (last warp) or to the next warp of the chain
    const bool was_end = (B.meta & CH_END) != 0u;
    if (last) {
      if ((B.meta & CH_STORE) && st_lane && gxB < g.pitch) {
        const size_t go = (size_t)((long long)B.y * g.pitch) + gxB;
        st4(out.d + go, B.d);
        st4(out.F[0] + go, B.fx); st4(out.F[1] + go, B.fy); st4(out.F[2] + go, B.fz); st4(out.F[3] + go, B.fw);
        *reinterpret_cast<uint4*>(out.v + go) = nv;
      }
    } else {
      const uint32_t oslot = emitted % D;
      mbar_wait_lean(lk.out_empty + oslot * 8, ((emitted / D) & 1u) ^ 1u);   // the consumer has read what was in the slot (passes at once the first time round)
      const uint32_t o = lk.out_rows + oslot * ROWB;
      sts4(o, B.h); sts4(o + SXW * 4, B.d); sts4(o + 2 * SXW * 4, B.fx); sts4(o + 3 * SXW * 4, B.fy);
      sts4(o + 4 * SXW * 4, B.fz); sts4(o + 5 * SXW * 4, B.fw);
      if (lane == 0) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(lk.out_hdr + oslot * 8), "r"(B.y), "r"(B.meta) : "memory");
      __syncwarp();                                     // every lane's stores are ordered before the release
      if (lane == 0) mbar_arrive_release(lk.out_full + oslot * 8);
      ++emitted;
    }
    // -- B's registers take the incoming row
    B.h = lds4(a); B.d = lds4(a + SXW * 4); B.fx = lds4(a + 2 * SXW * 4); B.fy = lds4(a + 3 * SXW * 4);
    B.fz = lds4(a + 4 * SXW * 4); B.fw = lds4(a + 5 * SXW * 4);
    B.y = ny; B.meta = nmeta;
    if (from_ring) {
      __syncwarp();                                     // every lane has read the slot before it is refilled
      if (first) {
        issue_one();
        if (++c_i == c_n) {
          const int nxt = queued;                       // the issue cursor has been there already (a piece is longer than the ring)
          queued = -2;
          open_consume_piece(nxt);
        }
      } else if (lane == 0) {
        mbar_arrive_release(lk.in_empty + (consumed % D) * 8);
      }
      ++consumed;
    }
    return !was_end;
  };

  ChainRow r0, r1;
  chain_zero(r0); chain_zero(r1);
#pragma unroll 1
  for (;;) {
    if (!tick(r0, r1)) break;                           // the incoming row lands in the registers of the row that left:
    if (!tick(r1, r0)) break;                           // the two register sets swap roles every tick, nothing is moved
  }
  if (!last) {                                          // two blank rows behind END: the next warp needs them to push END through its own lag
#pragma unroll 1
    for (int j = 0; j < 2; ++j) {
      const uint32_t oslot = emitted % D;
      mbar_wait_lean(lk.out_empty + oslot * 8, ((emitted / D) & 1u) ^ 1u);
      if (lane == 0) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(lk.out_hdr + oslot * 8), "r"(0), "r"(0) : "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive_release(lk.out_full + oslot * 8);
      ++emitted;
    }
  }
  // the last chain to leave re-arms the counters for the next launch on this stream (every chain that had a piece has by then
  // seen its final, failing fetch, so no atomic on the counter is still in flight)
  if (first && sched != nullptr && vchain < sch.npieces && lane == 0) {
    const unsigned active = (unsigned)(nvchain < sch.npieces ? nvchain : sch.npieces);
    __threadfence();
    if (atomicAdd(sched + 1, 1u) == active - 1u) { sched[0] = 0u; sched[1] = 0u; }
  }
}

template <class C, bool EXT>
__global__ void __launch_bounds__(C::NT, 1) chain_step_kernel(const __grid_constant__ CUtensorMap tm_h, const __grid_constant__ CUtensorMap tm_s,
                                                              const __grid_constant__ BandSched sch, FusedOut out, Geom g, StepConsts c,
                                                              int lr0, int nstrips, int tma_y_bias, uint32_t* sched) {
  constexpr int K = C::K, NCH = C::NCH, D = C::D;
  extern __shared__ __align__(1024) float smem[];
  __shared__ uint64_t full_all[NCH * K * D];
  __shared__ uint64_t empty_all[NCH * K * D];
  __shared__ uint2 hdr_all[NCH * K * D];
  uint32_t tid_u;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid_u));
  // warp w is position w / NCH of chain w % NCH: the four schedulers of the SM each get a mix of positions
  const int tid = (int)tid_u, lane = tid & 31, warp = tid >> 5, stage = warp / NCH, chain = warp - stage * NCH;
  if (tid == 0) {
#pragma unroll 1
    for (int i = 0; i < NCH * K * D; ++i) { mbar_init(&full_all[i], 1); mbar_init(&empty_all[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  float* zeros = smem + (size_t)NCH * K * D * C::ROW;
  for (int i = tid; i < C::ROW; i += C::NT) zeros[i] = 0.0f;
  __syncthreads();                                       // the only CTA-wide barrier: the mbarriers and the zero row exist
  const int link_in = chain * K + stage;                 // link l of a chain feeds its warp l (link 0: TMA)
  const int link_out = stage + 1 < K ? link_in + 1 : link_in;        // (unused by the last warp)
  ChainLink lk;
  lk.in_rows = smem_u32(smem + (size_t)link_in * D * C::ROW + lane * 4);
  lk.in_full = smem_u32(&full_all[link_in * D]);
  lk.in_empty = smem_u32(&empty_all[link_in * D]);
  lk.in_hdr = smem_u32(&hdr_all[link_in * D]);
  lk.out_rows = smem_u32(smem + (size_t)link_out * D * C::ROW + lane * 4);
  lk.out_full = smem_u32(&full_all[link_out * D]);
  lk.out_empty = smem_u32(&empty_all[link_out * D]);
  lk.out_hdr = smem_u32(&hdr_all[link_out * D]);
  const int vchain = (int)blockIdx.x * NCH + chain, nvchain = (int)gridDim.x * NCH;
  chain_stage<C, EXT>(stage == 0, stage == K - 1, lk, &tm_h, &tm_s, sch, out, g, c, lr0, nstrips, tma_y_bias, sched, vchain, nvchain, lane,
                      smem + (size_t)link_in * D * C::ROW, &full_all[link_in * D], smem_u32(zeros + lane * 4));
}

// ---- host side ---------------------------------------------------------------------------
#ifndef TWS_CHAIN_WARPS
#define TWS_CHAIN_WARPS 16       // warps per CTA (all chains)
#endif
#ifndef TWS_CHAIN_D
#define TWS_CHAIN_D 4            // ring depth of a link (rows)
#endif
template <int K> struct ChainCfgFor { using type = ChainCfg<K, TWS_CHAIN_WARPS, TWS_CHAIN_D>; };

static int chain_sm_count() {
  static int cached[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!cached[dev & 63]) cudaDeviceGetAttribute(&cached[dev & 63], cudaDevAttrMultiProcessorCount, dev);
  return cached[dev & 63] > 0 ? cached[dev & 63] : 148;
}

template <int K, bool EXT>
static cudaError_t launch_chain_k(const Geom& g, const Planes& p, const TmaSet& tma, int src, const StepConsts& c, int lr0, int lr1,
                                  cudaStream_t st, uint32_t* sched) {
  using C = typename ChainCfgFor<K>::type;
  if (C::SXW != stream_strip_width()) return cudaErrorInvalidValue;      // shares the row descriptors of the ring / band kernels
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  auto kern = chain_step_kernel<C, EXT>;
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
  const int sms = chain_sm_count();
  const BandSched sch = band_build_schedule(lr1 - lr0, nstrips, sms * C::NCH, 1, C::HP, 0, 0, nullptr, nullptr, nullptr);
  const int want = (sch.npieces + C::NCH - 1) / C::NCH;
  const int grid = want < 1 ? 1 : (want < sms ? want : sms);
  const int bias = g.has_up ? TWS_HALO_ROWS : 0;
  kern<<<grid, C::NT, C::SMEM, st>>>(tma.m[0], tma.m[1], sch, out, g, c, lr0, nstrips, bias, sched);
  return cudaGetLastError();
}

cudaError_t launch_chain(int K, const Geom& g, const Planes& p, const TmaSet& tma, int src, const StepConsts& c, int lr0, int lr1,
                         cudaStream_t st, uint32_t* sched) {
  const bool ext = c.ext_sources != 0;
  switch (K) {
    case 1: return ext ? launch_chain_k<1, true>(g, p, tma, src, c, lr0, lr1, st, sched) : launch_chain_k<1, false>(g, p, tma, src, c, lr0, lr1, st, sched);
    case 2: return ext ? launch_chain_k<2, true>(g, p, tma, src, c, lr0, lr1, st, sched) : launch_chain_k<2, false>(g, p, tma, src, c, lr0, lr1, st, sched);
    case 3: return ext ? launch_chain_k<3, true>(g, p, tma, src, c, lr0, lr1, st, sched) : launch_chain_k<3, false>(g, p, tma, src, c, lr0, lr1, st, sched);
    case 4: return ext ? launch_chain_k<4, true>(g, p, tma, src, c, lr0, lr1, st, sched) : launch_chain_k<4, false>(g, p, tma, src, c, lr0, lr1, st, sched);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace tws
