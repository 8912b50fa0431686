// stream_kernels.cu — the row-streaming, warp-autonomous form of the fused step (sm_100a).
//
// Same arithmetic as every other step kernel (flowUpdate.comp:12-63 + flowApply.comp:14-53,
// contract in cell_math.cuh / DESIGN.md section 2) and bit-identical results; what differs is how K
// whole steps are blocked over one HBM round trip:
//
//   * The grid is cut into column strips SXW = 256 cells wide (OX = 256 - 2*HX of them are
//     outputs, HX = x halo of the K-step dependency cone) and every strip is marched top to
//     bottom.  A persistent CTA (one per SM) owns a contiguous piece of the (strip, row) work
//     list; all pieces have the same number of rows, so any grid size balances over 148 SMs.
//   * ONE WARP OWNS ONE ROW (256 cells: lane L holds the float4 groups L and L+32) for the row's
//     whole life: the TMA unit lands the row's h, d and four flux planes in the warp's private
//     landing buffer (6 x 1 KB, `cp.async.bulk.tensor.2d`, mbarrier completion, out-of-bounds
//     zero fill = the reference's exterior rule), the warp pulls them into registers, and the
//     2K half-passes (flux_1, depth_1, ... flux_K, depth_K) update them IN REGISTERS.  The
//     moment a landing buffer has been read the same warp issues the TMA load of ITS next row
//     (row + NW), which lands while the current row is being computed: loads never stall.
//   * Rows are skewed in time instead of recomputed: row y can run half-pass s as soon as rows
//     y-1 and y+1 have finished half-pass s-1.  Only what a vertical neighbour needs goes
//     through shared memory (the water level H = d + h, the +Y and -Y outflow: 3 KB per row
//     slot); x neighbours are adjacent lanes (warp shuffles).  Each row slot publishes a
//     monotonically increasing progress word in shared memory (st.release / ld.acquire); a
//     warp waits only for its two neighbours, never for the CTA — there is no __syncthreads
//     in the row loop, and warps in different phases overlap their shared-memory and FP work.
//   * No y-halo recomputation: each piece pays 2K warm-up rows once.  Per cell-update the
//     kernel executes ~1/0.94 (K <= 2) or ~1/0.91 (K = 3, 4) of the minimum work, against 1/0.71
//     for the tile kernel at K = 2, and moves 48/K B (+ x-halo re-reads) through HBM.
#include "cell_math.cuh"

namespace tws {

template <int K_, int NW_>
struct StreamCfg {
  static constexpr int K = K_, NW = NW_, NT = NW_ * 32;
  static constexpr int SXW = 256;                       // staged strip width: 2 float4 groups per lane
  static constexpr int HX = ((2 * K + 3) / 4) * 4;      // x halo rounded to whole float4 groups
  static constexpr int OX = SXW - 2 * HX;               // output columns per strip
  static constexpr int NHP = 2 * K + 1;                 // half-passes per row, the load included
  static constexpr int LAND = 6 * SXW;                  // floats per landing buffer (h, d, F x4)
  static constexpr int XROW = 3 * SXW;                  // floats per exchange slot (H, F+Y, F-Y)
  static constexpr size_t SMEM = (size_t)NW * (LAND + XROW) * sizeof(float);
  static_assert(NW > 2 * K, "the row ring must be deeper than the dependency cone");
};

// ---- neighbour synchronisation -----------------------------------------------------------
__device__ __forceinline__ uint32_t ld_acquire_shared(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_shared(uint32_t* p, uint32_t v) {
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

// Every lane polls both progress words (one broadcast LDS each).  A bounded spin: a wait that
// never completes is a bug, and a trap (launch failure on the host) is better than a hung GPU.
__device__ __forceinline__ void wait_rows(const uint32_t* pu, uint32_t need_u, const uint32_t* pd, uint32_t need_d) {
  uint32_t spins = 0;
  while (ld_acquire_shared(pu) < need_u || ld_acquire_shared(pd) < need_d) {
    if (++spins > (1u << 27)) __trap();
  }
  __syncwarp();
}
__device__ __forceinline__ void signal_row(uint32_t* p, uint32_t value, int lane) {
  __syncwarp();                                   // every lane's shared-memory stores are ordered before ...
  if (lane == 0) st_release_shared(p, value);     // ... the release of the new progress value
}

struct RowCtx {
  int gy;          // global row
  int gxa, gxb;    // global x of the lane's two groups
  bool row_in;     // the row exists in the global grid
  bool store;      // the row is an output row of this piece
  bool sta, stb;   // the lane's groups are output columns
  size_t goa, gob; // element offsets of the two groups in the output planes
};

// x neighbours of the lane's two groups.  `w3` / `x0`: component 3 / 0 of the value in groups
// (a, b).  left(a) = lane-1's a.w, left(b) = lane-1's b.w but lane 0 takes lane 31's a.w;
// right(a) = lane+1's a.x but lane 31 takes lane 0's b.x, right(b) = lane+1's b.x.  The two
// strip-edge cells get their own value back: they lie in the x halo, whose results are never kept.
__device__ __forceinline__ void x_neighbours(float aw, float bw, float ax, float bx, int lane, float& la, float& lb, float& ra, float& rb) {
  la = __shfl_up_sync(0xffffffffu, aw, 1);
  lb = __shfl_sync(0xffffffffu, lane == 31 ? aw : bw, (lane + 31) & 31);
  ra = __shfl_sync(0xffffffffu, lane == 0 ? bx : ax, (lane + 1) & 31);
  rb = __shfl_down_sync(0xffffffffu, bx, 1);
}

// flowUpdate.comp:34-62 for the lane's 8 cells.  Reads the neighbour rows' water level, leaves the
// new outflow in registers, publishes its +-Y components; LAST also stores the flux planes to HBM.
template <bool EDGE, bool LAST>
__device__ __forceinline__ void stream_flux(const float* __restrict__ sHup, const float* __restrict__ sHdn, float* __restrict__ sFyp,
                                            float* __restrict__ sFym, const int oa, const int ob, const float4 (&h)[2], const float4 (&d)[2],
                                            float4 (&fx)[2], float4 (&fy)[2], float4 (&fz)[2], float4 (&fw)[2], const RowCtx& rc,
                                            const FusedOut& out, const Geom& g, const StepConsts& c, const int lane) {
  float4 HC[2], HU[2], HD[2];
  HC[0] = add4(d[0], h[0]); HC[1] = add4(d[1], h[1]);                                // a + r, flowUpdate.comp:34
  HU[0] = ld4(sHup + oa); HU[1] = ld4(sHup + ob);
  HD[0] = ld4(sHdn + oa); HD[1] = ld4(sHdn + ob);
  float HL[2], HR[2];
  x_neighbours(HC[0].w, HC[1].w, HC[0].x, HC[1].x, lane, HL[0], HL[1], HR[0], HR[1]);
  float total[2][4], scale[2][4];
  bool need = false;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    float* pfx = &fx[q].x; float* pfy = &fy[q].x; float* pfz = &fz[q].x; float* pfw = &fw[q].x;
    const int gx0 = q ? rc.gxb : rc.gxa;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float Hc = comp(HC[q], i);
      float hxp = (i < 3) ? comp(HC[q], i + 1) : HR[q];
      float hxm = (i > 0) ? comp(HC[q], i - 1) : HL[q];
      float hyp = comp(HD[q], i), hym = comp(HU[q], i);
      if (EDGE && c.closed) {
        const int gx = gx0 + i;
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
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float dep = comp(d[q], i);
        if (total[q][i] > dep && dep != 0.0f) scale[q][i] = __fdiv_rn(dep, total[q][i]);   // :59
      }
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    float* pfx = &fx[q].x; float* pfy = &fy[q].x; float* pfz = &fz[q].x; float* pfw = &fw[q].x;
    const int gx0 = q ? rc.gxb : rc.gxa;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      pfx[i] = __fmul_rn(pfx[i], scale[q][i]); pfy[i] = __fmul_rn(pfy[i], scale[q][i]);
      pfz[i] = __fmul_rn(pfz[i], scale[q][i]); pfw[i] = __fmul_rn(pfw[i], scale[q][i]);
      if (EDGE && !(rc.row_in && (unsigned)(gx0 + i) < (unsigned)g.W)) { pfx[i] = 0.f; pfy[i] = 0.f; pfz[i] = 0.f; pfw[i] = 0.f; }
    }
  }
  st4(sFyp + oa, fz[0]); st4(sFyp + ob, fz[1]);
  st4(sFym + oa, fw[0]); st4(sFym + ob, fw[1]);
  if (LAST && rc.store) {
    if (rc.sta) { st4(out.F[0] + rc.goa, fx[0]); st4(out.F[1] + rc.goa, fy[0]); st4(out.F[2] + rc.goa, fz[0]); st4(out.F[3] + rc.goa, fw[0]); }
    if (rc.stb) { st4(out.F[0] + rc.gob, fx[1]); st4(out.F[1] + rc.gob, fy[1]); st4(out.F[2] + rc.gob, fz[1]); st4(out.F[3] + rc.gob, fw[1]); }
  }
}

// flowApply.comp:32-52 for the lane's 8 cells.  Reads the neighbour rows' +-Y outflow; not LAST:
// new depth stays in registers and the new water level is published; LAST: depth and the packed
// fp16 flow vector go to HBM.
template <bool EDGE, bool LAST>
__device__ __forceinline__ void stream_depth(const float* __restrict__ sFypUp, const float* __restrict__ sFymDn, float* __restrict__ sH,
                                             const int oa, const int ob, const float4 (&h)[2], float4 (&d)[2], const float4 (&fx)[2],
                                             const float4 (&fy)[2], const float4 (&fz)[2], const float4 (&fw)[2], const RowCtx& rc,
                                             const FusedOut& out, const Geom& g, const StepConsts& c, const int lane) {
  float4 iy1[2], iy0[2];
  iy1[0] = ld4(sFymDn + oa); iy1[1] = ld4(sFymDn + ob);      // F(x,y+1).w, flowApply.comp:34
  iy0[0] = ld4(sFypUp + oa); iy0[1] = ld4(sFypUp + ob);      // F(x,y-1).z, :35
  float l[2], r[2];
  // F(x-1,y).x (:33) is the left cell's +X outflow, F(x+1,y).y (:32) the right cell's -X outflow
  x_neighbours(fx[0].w, fx[1].w, fy[0].x, fy[1].x, lane, l[0], l[1], r[0], r[1]);
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int gx0 = q ? rc.gxb : rc.gxa;
    float nd[4]; uint32_t nv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float iX1 = (i < 3) ? comp(fy[q], i + 1) : r[q];
      const float iX0 = (i > 0) ? comp(fx[q], i - 1) : l[q];
      float vx, vy;
      nd[i] = apply_cell(comp(d[q], i), comp(fx[q], i), comp(fy[q], i), comp(fz[q], i), comp(fw[q], i), iX1, iX0, comp(iy1[q], i),
                         comp(iy0[q], i), c, vx, vy);
      if (LAST) nv[i] = pack_half2(vx, vy);
      if (EDGE && !(rc.row_in && (unsigned)(gx0 + i) < (unsigned)g.W)) { nd[i] = 0.f; if (LAST) nv[i] = 0u; }
    }
    if (!LAST) {
      d[q] = make_float4(nd[0], nd[1], nd[2], nd[3]);
      st4(sH + (q ? ob : oa), add4(d[q], h[q]));
    } else if (rc.store && (q ? rc.stb : rc.sta)) {
      const size_t go = q ? rc.gob : rc.goa;
      st4(out.d + go, make_float4(nd[0], nd[1], nd[2], nd[3]));
      *reinterpret_cast<uint4*>(out.v + go) = make_uint4(nv[0], nv[1], nv[2], nv[3]);
    }
  }
}

template <class C>
__global__ void __launch_bounds__(C::NT, 1) stream_step_kernel(const __grid_constant__ CUtensorMap tm_h,
                                                               const __grid_constant__ CUtensorMap tm_d,
                                                               const __grid_constant__ CUtensorMap tm_f0,
                                                               const __grid_constant__ CUtensorMap tm_f1,
                                                               const __grid_constant__ CUtensorMap tm_f2,
                                                               const __grid_constant__ CUtensorMap tm_f3,
                                                               FusedOut out, Geom g, StepConsts c, int lr0, int lr1, int nstrips,
                                                               int tma_y_bias) {
  constexpr int K = C::K, NW = C::NW, SXW = C::SXW, HX = C::HX, OX = C::OX, NHP = C::NHP, LAND = C::LAND;
  extern __shared__ __align__(1024) float smem[];
  __shared__ uint64_t full[NW];
  __shared__ uint32_t progress[NW];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wup = (warp + NW - 1) % NW, wdn = (warp + 1) % NW;

  float* land = smem + warp * LAND;                       // this warp's landing buffer
  float* xch = smem + NW * LAND;                          // exchange slots: [NW][H | F+Y | F-Y][SXW]
  float* sH_me = xch + warp * C::XROW;            float* sFyp_me = sH_me + SXW;            float* sFym_me = sH_me + 2 * SXW;
  const float* sH_up = xch + wup * C::XROW;       const float* sFyp_up = sH_up + SXW;
  const float* sH_dn = xch + wdn * C::XROW;       const float* sFym_dn = sH_dn + 2 * SXW;
  const int oa = lane * 4, ob = 128 + lane * 4;

  if (tid == 0) {
#pragma unroll 1
    for (int i = 0; i < NW; ++i) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t parity = 0;                                    // phase of this warp's landing barrier

  // This CTA's piece of the (strip, row) list: pieces are equal, contiguous, strip-major.
  const long long R = (long long)(lr1 - lr0);
  const long long TR = R * nstrips;
  const long long lin_begin = TR * blockIdx.x / gridDim.x, lin_end = TR * (blockIdx.x + 1) / gridDim.x;

#pragma unroll 1
  for (long long lin = lin_begin; lin < lin_end;) {
    const int strip = (int)(lin / R);
    const int ya = lr0 + (int)(lin - (long long)strip * R);
    const int yb = (int)((long long)ya + (lin_end - lin) < (long long)lr1 ? (long long)ya + (lin_end - lin) : (long long)lr1);
    lin += yb - ya;
    const int sx0 = strip * OX - HX;
    const int ystart = ya - 2 * K;                        // 2K warm-up rows above, 2K feeder rows below
    const int nrows = (yb - ya) + 4 * K;
    const bool xedge = sx0 < 0 || sx0 + SXW > g.W;

    __syncthreads();                                      // every warp is done with the previous piece
    if (tid < NW) progress[tid] = 0;
    __syncthreads();

    auto issue = [&](int y) {                             // one lane: land row y in this warp's buffer
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&full[warp], (uint32_t)(LAND * sizeof(float)));
      const int ty = y + tma_y_bias;
      tma_load_2d(land, &tm_h, sx0, ty, &full[warp]);
      tma_load_2d(land + SXW, &tm_d, sx0, ty, &full[warp]);
      tma_load_2d(land + 2 * SXW, &tm_f0, sx0, ty, &full[warp]);
      tma_load_2d(land + 3 * SXW, &tm_f1, sx0, ty, &full[warp]);
      tma_load_2d(land + 4 * SXW, &tm_f2, sx0, ty, &full[warp]);
      tma_load_2d(land + 5 * SXW, &tm_f3, sx0, ty, &full[warp]);
    };
    if (lane == 0 && warp < nrows) issue(ystart + warp);

    RowCtx rc;
    rc.gxa = sx0 + oa; rc.gxb = sx0 + ob;
    rc.sta = oa >= HX && oa < HX + OX && rc.gxa < g.pitch;
    rc.stb = ob >= HX && ob < HX + OX && rc.gxb < g.pitch;

#pragma unroll 1
    for (int idx = warp; idx < nrows; idx += NW) {
      const int y = ystart + idx;                         // local row
      rc.gy = g.row0 + y;
      rc.row_in = (unsigned)rc.gy < (unsigned)g.Hg;
      rc.store = y >= ya && y < yb;
      const size_t rowoff = (size_t)((long long)y * g.pitch);      // only dereferenced when rc.store (y >= 0)
      rc.goa = rowoff + rc.gxa; rc.gob = rowoff + rc.gxb;
      // rows below the piece only feed the rows above them: row yb-1+m stops after half-pass 2K-m
      const int smax = (y < yb) ? 2 * K : 2 * K - (y - yb + 1);
      const bool edge = xedge || rc.gy <= 0 || rc.gy >= g.Hg - 1;
      const uint32_t base_me = (uint32_t)idx * NHP, base_dn = base_me + NHP;
      const uint32_t base_up = idx > 0 ? base_me - NHP : 0u;      // the first row has nobody above: 0 is always reached
      const uint32_t up_on = idx > 0 ? 1u : 0u;

      // ---- half-pass 0: registers <- landing buffer; publish H; prefetch this warp's next row ----
      float4 h[2], d[2], fx[2], fy[2], fz[2], fw[2];
      mbar_wait(&full[warp], parity);
      parity ^= 1u;
      h[0] = ld4(land + oa);            h[1] = ld4(land + ob);
      d[0] = ld4(land + SXW + oa);      d[1] = ld4(land + SXW + ob);
      fx[0] = ld4(land + 2 * SXW + oa); fx[1] = ld4(land + 2 * SXW + ob);
      fy[0] = ld4(land + 3 * SXW + oa); fy[1] = ld4(land + 3 * SXW + ob);
      fz[0] = ld4(land + 4 * SXW + oa); fz[1] = ld4(land + 4 * SXW + ob);
      fw[0] = ld4(land + 5 * SXW + oa); fw[1] = ld4(land + 5 * SXW + ob);
      __syncwarp();
      if (lane == 0 && idx + NW < nrows) issue(y + NW);
      st4(sH_me + oa, add4(d[0], h[0]));
      st4(sH_me + ob, add4(d[1], h[1]));
      signal_row(&progress[warp], base_me + 1, lane);

      // ---- levels 1..K-1 (results stay on chip) --------------------------------------------------
      int s = 1;
#pragma unroll 1
      for (int t = 1; t < K && s <= smax; ++t) {
        wait_rows(&progress[wup], (base_up + s) * up_on, &progress[wdn], base_dn + s);
        if (edge) stream_flux<true, false>(sH_up, sH_dn, sFyp_me, sFym_me, oa, ob, h, d, fx, fy, fz, fw, rc, out, g, c, lane);
        else stream_flux<false, false>(sH_up, sH_dn, sFyp_me, sFym_me, oa, ob, h, d, fx, fy, fz, fw, rc, out, g, c, lane);
        signal_row(&progress[warp], base_me + s + 1, lane);
        ++s;
        if (s > smax) break;
        wait_rows(&progress[wup], (base_up + s) * up_on, &progress[wdn], base_dn + s);
        if (edge) stream_depth<true, false>(sFyp_up, sFym_dn, sH_me, oa, ob, h, d, fx, fy, fz, fw, rc, out, g, c, lane);
        else stream_depth<false, false>(sFyp_up, sFym_dn, sH_me, oa, ob, h, d, fx, fy, fz, fw, rc, out, g, c, lane);
        signal_row(&progress[warp], base_me + s + 1, lane);
        ++s;
      }
      // ---- level K: flux and depth / velocity go to HBM straight from registers -------------------
      if (s == 2 * K - 1 && s <= smax) {
        wait_rows(&progress[wup], (base_up + s) * up_on, &progress[wdn], base_dn + s);
        if (edge) stream_flux<true, true>(sH_up, sH_dn, sFyp_me, sFym_me, oa, ob, h, d, fx, fy, fz, fw, rc, out, g, c, lane);
        else stream_flux<false, true>(sH_up, sH_dn, sFyp_me, sFym_me, oa, ob, h, d, fx, fy, fz, fw, rc, out, g, c, lane);
        signal_row(&progress[warp], base_me + s + 1, lane);
        ++s;
        if (s <= smax) {
          wait_rows(&progress[wup], (base_up + s) * up_on, &progress[wdn], base_dn + s);
          if (edge) stream_depth<true, true>(sFyp_up, sFym_dn, sH_me, oa, ob, h, d, fx, fy, fz, fw, rc, out, g, c, lane);
          else stream_depth<false, true>(sFyp_up, sFym_dn, sH_me, oa, ob, h, d, fx, fy, fz, fw, rc, out, g, c, lane);
          signal_row(&progress[warp], base_me + s + 1, lane);
        }
      }
    }
  }
}

// ---- host side ---------------------------------------------------------------------------
#ifndef TWS_STREAM_NW
#define TWS_STREAM_NW 16
#endif
template <int K> struct StreamCfgFor { using type = StreamCfg<K, TWS_STREAM_NW>; };

static int stream_sm_count() {
  static int cached[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!cached[dev & 63]) cudaDeviceGetAttribute(&cached[dev & 63], cudaDevAttrMultiProcessorCount, dev);
  return cached[dev & 63] > 0 ? cached[dev & 63] : 148;
}

template <int K>
static cudaError_t launch_stream_k(const Geom& g, const Planes& p, const TmaSet& tma, int src, const StepConsts& c, int lr0, int lr1,
                                   cudaStream_t st) {
  using C = typename StreamCfgFor<K>::type;
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  auto kern = stream_step_kernel<C>;
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
  const long long total_rows = (long long)nstrips * (lr1 - lr0);
  // one persistent CTA per SM; small grids: at least ~4 ring turns of rows per CTA so the 4K warm-up rows amortise
  const long long min_rows = 4 * C::NW;
  long long want = (total_rows + min_rows - 1) / min_rows;
  const int sms = stream_sm_count();
  const int grid = (int)(want < 1 ? 1 : (want < sms ? want : sms));
  const int bias = g.has_up ? TWS_HALO_ROWS : 0;
  kern<<<grid, C::NT, C::SMEM, st>>>(tma.m[0], tma.m[1], tma.m[2], tma.m[3], tma.m[4], tma.m[5], out, g, c, lr0, lr1, nstrips, bias);
  return cudaGetLastError();
}

cudaError_t launch_stream(int K, const Geom& g, const Planes& p, const TmaSet& tma, int src, const StepConsts& c, int lr0, int lr1,
                          cudaStream_t st) {
  switch (K) {
    case 1: return launch_stream_k<1>(g, p, tma, src, c, lr0, lr1, st);
    case 2: return launch_stream_k<2>(g, p, tma, src, c, lr0, lr1, st);
    case 3: return launch_stream_k<3>(g, p, tma, src, c, lr0, lr1, st);
    case 4: return launch_stream_k<4>(g, p, tma, src, c, lr0, lr1, st);
    default: return cudaErrorInvalidValue;
  }
}

// Row descriptors: box = one 256-cell row segment.  Same visibility rule as the tile kernel's maps
// (own rows plus the halo rows towards an existing neighbour; everything else zero-filled).
cudaError_t stream_build_tma(const Geom& g, const Planes& p, int side, TmaSet* out, std::string* err) {
  return build_tma_boxes(g, p, side, StreamCfgFor<1>::type::SXW, 1, out, err);
}

}  // namespace tws
