// step_kernels.cu — the shallow-water step on sm_100a.
//
// Replaces flowUpdate.comp:12-63 + flowApply.comp:14-53 (dispatched at Terrain.cpp:255-264).
// Arithmetic contract: IEEE fp32, the shader's operation
// order, NO fma contraction (file is built with -fmad=false and the two a*b+c*d sites use
// __fmul_rn/__fadd_rn explicitly), IEEE division, max(0,x) == (x > 0 ? x : 0), exterior
// texels read as 0, velocity stored as 2 x fp16 round-to-nearest-even.  Every kernel in
// this file produces bit-identical state; tests compare them with the CPU oracle.
//
// Kernels:
//   unfused_update_kernel / unfused_apply_kernel — the two reference dispatches, planar,
//       in place, neighbours through L1/L2 (baseline, BASELINE.json config 2).
//   fused_step_kernel<K,OX,OY,NT> — K whole steps per HBM round trip.  One CTA stages an
//       (OX+2HX) x (OY+4K) tile of h, d and the four flux planes in shared memory with six
//       TMA tensor loads (out-of-bounds zero fill IS the reference's exterior rule), runs
//       2K in-place passes over a region shrinking by one cell per pass, and writes the
//       OX x OY centre straight from registers (flux after the last flux pass, depth and
//       velocity after the last depth pass).  HBM traffic per cell: 24 B read + 24 B
//       written per K cell-updates (+ halo over-fetch, mostly served by L2).
#include "cell_math.cuh"

namespace tws {

// ------------------------------------------------------------------------------------
// unfused baseline: one thread per 4 consecutive cells, float4 global access
// ------------------------------------------------------------------------------------
// All plane pointers address LOCAL row 0 (plane row TWS_HALO_ROWS).  Rows lr0..lr1-1 are
// processed; rows outside the GLOBAL grid read as 0.
__global__ void __launch_bounds__(256) unfused_update_kernel(Geom g, const float* __restrict__ h, const float* __restrict__ d,
                                                             float* Fxp, float* Fxm, float* Fyp, float* Fym,
                                                             StepConsts c, int lr0, int lr1) {
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (x >= g.pitch) return;
  for (int lr = lr0 + blockIdx.y; lr < lr1; lr += gridDim.y) {
  const int gy = g.row0 + lr;
  const long long o = (long long)lr * g.pitch + x;
  const bool up_ok = gy - 1 >= 0, dn_ok = gy + 1 < g.Hg;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 dC = ld4(d + o), hC = ld4(h + o);
  const float4 HC = add4(dC, hC);                                                    // a + r, flowUpdate.comp:34
  float4 HU = up_ok ? add4(ld4(d + o - g.pitch), ld4(h + o - g.pitch)) : z4;       // y-1
  float4 HD = dn_ok ? add4(ld4(d + o + g.pitch), ld4(h + o + g.pitch)) : z4;       // y+1
  float HL = (x > 0) ? (d[o - 1] + h[o - 1]) : 0.0f;
  float HR = (x + 4 < g.pitch) ? (d[o + 4] + h[o + 4]) : 0.0f;
  float4 fx = ld4(Fxp + o), fy = ld4(Fxm + o), fz = ld4(Fyp + o), fw = ld4(Fym + o);
  float* pfx = &fx.x; float* pfy = &fy.x; float* pfz = &fz.x; float* pfw = &fw.x;
  const float Hc[4] = {HC.x, HC.y, HC.z, HC.w};
  const float Hu[4] = {HU.x, HU.y, HU.z, HU.w};
  const float Hd[4] = {HD.x, HD.y, HD.z, HD.w};
  const float dc[4] = {dC.x, dC.y, dC.z, dC.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gx = x + i;
    float hxp = (i < 3) ? Hc[i + 1] : HR;
    float hxm = (i > 0) ? Hc[i - 1] : HL;
    float hyp = Hd[i], hym = Hu[i];
    if (c.closed) {
      if (gx + 1 >= g.W) hxp = Hc[i];
      if (gx - 1 < 0) hxm = Hc[i];
      if (!dn_ok) hyp = Hc[i];
      if (!up_ok) hym = Hc[i];
    }
    flux_cell(Hc[i], hxp, hxm, hyp, hym, dc[i], pfx[i], pfy[i], pfz[i], pfw[i], c);
    if (gx >= g.W) { pfx[i] = 0.f; pfy[i] = 0.f; pfz[i] = 0.f; pfw[i] = 0.f; }      // pad columns stay 0
  }
  st4(Fxp + o, fx); st4(Fxm + o, fy); st4(Fyp + o, fz); st4(Fym + o, fw);
  ledger_add(c, g, x, gy, fx, fy, fz, fw);
  }
}

__global__ void __launch_bounds__(256) unfused_apply_kernel(Geom g, float* d, const float* __restrict__ Fxp,
                                                            const float* __restrict__ Fxm, const float* __restrict__ Fyp,
                                                            const float* __restrict__ Fym, uint32_t* __restrict__ v,
                                                            StepConsts c, int lr0, int lr1) {
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  double src_acc = 0.0;                                 // EXT ledger of the sources (whole warps reach the flush below)
  for (int lr = lr0 + blockIdx.y; lr < lr1 && x < g.pitch; lr += gridDim.y) {
  const int gy = g.row0 + lr;
  const long long o = (long long)lr * g.pitch + x;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 fx = ld4(Fxp + o), fy = ld4(Fxm + o), fz = ld4(Fyp + o), fw = ld4(Fym + o);
  const float4 iy1 = (gy + 1 < g.Hg) ? ld4(Fym + o + g.pitch) : z4;                  // F(x,y+1).w, flowApply.comp:34
  const float4 iy0 = (gy - 1 >= 0) ? ld4(Fyp + o - g.pitch) : z4;                    // F(x,y-1).z, :35
  const float l = (x > 0) ? Fxp[o - 1] : 0.0f;                                       // F(x-1,y).x, :33
  const float r = (x + 4 < g.pitch) ? Fxm[o + 4] : 0.0f;                             // F(x+1,y).y, :32
  const float4 dC = ld4(d + o);
  float nd[4]; uint32_t nv[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float iX1 = (i < 3) ? comp(fy, i + 1) : r;
    const float iX0 = (i > 0) ? comp(fx, i - 1) : l;
    float vx, vy, ds;
    nd[i] = apply_cell_src(comp(dC, i), comp(fx, i), comp(fy, i), comp(fz, i), comp(fw, i), iX1, iX0, comp(iy1, i), comp(iy0, i), c,
                           c.ext_sources != 0, vx, vy, ds);
    nv[i] = pack_half2(vx, vy);
    if (x + i >= g.W) { nd[i] = 0.f; nv[i] = 0u; ds = 0.f; }
    src_acc += (double)ds;
  }
  st4(d + o, make_float4(nd[0], nd[1], nd[2], nd[3]));
  *reinterpret_cast<uint4*>(v + o) = make_uint4(nv[0], nv[1], nv[2], nv[3]);
  }
  if (c.ledger_src != nullptr) ledger_src_flush(c.ledger_src, src_acc);
}

// ------------------------------------------------------------------------------------
// fused / temporally blocked kernel
// ------------------------------------------------------------------------------------
// Tile configuration of the pipelined kernel.  One persistent CTA per SM owns TWO staging
// buffers: while the passes run on tile j in one buffer the TMA unit fills the other with
// tile j+1, so the SM never idles on the load and HBM never idles on the compute.
template <int K_, int OX_, int OY_, int NT_>
struct FusedCfg {
  static constexpr int K = K_, OX = OX_, OY = OY_, NT = NT_;
  static constexpr int HX = ((2 * K + 3) / 4) * 4;   // x halo rounded to whole float4 groups
  static constexpr int HY = 2 * K;
  static constexpr int SX = OX + 2 * HX, SY = OY + 2 * HY;
  static constexpr int NG = SX / 4;                  // float4 groups per staged row
  static constexpr int PLANE = SX * SY;              // floats per staged plane
  static constexpr int STAGE = 6 * PLANE;            // h, d, F x4
  static constexpr int STAGES = 2;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE * sizeof(float);
  static_assert(OX % 4 == 0 && SX <= 256 && SY <= 256, "TMA box limits");
  static_assert((PLANE * 4) % 128 == 0, "staged planes must keep 128 B alignment");
  static_assert(NT % 32 == 0, "whole warps");
};

struct TileCtx {
  int sx0;        // grid x of staged column 0
  int ly0;        // LOCAL row of staged row 0
  int gy0;        // GLOBAL row of staged row 0
};

// ---- register-resident tile update ---------------------------------------------------------
// Every thread OWNS up to IPT float4 groups ("items") of the staged tile (rows 1..SY-2, full
// staged width) for the whole tile: their terrain, depth and four flux values are pulled out
// of the TMA landing buffers into registers once and stay there across all K levels.  Only
// what a NEIGHBOUR needs goes through shared memory:
//   sH   (reuses the depth landing plane)  water level d+h, read by the rows above/below;
//   sFyp / sFym (their landing planes)     +Y / -Y outflow, read by the rows below/above;
//   x-neighbours sit in the adjacent lane (items are contiguous along x, full-width rows are
//   contiguous in memory) and are exchanged with warp shuffles; lanes 0 / 31 fall back to one
//   scalar shared-memory slot written by the neighbouring warp's edge lane.
// Per level and item that is 4 LDS.128 + 3 STS.128 instead of 17 + 5, and the flux never
// round-trips through shared memory for its owner.
// Passes run over all owned rows at every level: cells outside the shrinking valid region
// hold garbage that by construction never reaches the OX x OY output centre.
// Predicated single-instruction shared-memory accesses for the warp-edge lanes (an `if` around a
// lone LDS/STS makes ptxas emit a divergent branch + reconvergence pair instead).
__device__ __forceinline__ float lds_if(bool p, const float* addr, float otherwise) {
  float v = otherwise;
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %2, 0;\n@p ld.shared.f32 %0, [%1];\n}" : "+f"(v) : "r"(smem_u32(addr)), "r"((int)p));
  return v;
}
__device__ __forceinline__ void sts_if(bool p, float* addr, float v) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %2, 0;\n@p st.shared.f32 [%0], %1;\n}" ::"r"(smem_u32(addr)), "f"(v), "r"((int)p));
}
__device__ __forceinline__ void sts4_if(bool p, float* addr, float4 v) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %5, 0;\n@p st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n}" ::"r"(smem_u32(addr)), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w), "r"((int)p));
}

template <class C, bool EDGE, bool LAST, int IPT>
__device__ __forceinline__ void flux_level(float* __restrict__ st, const int (&o)[IPT], const bool (&valid)[IPT], const float4 (&h)[IPT],
                                           const float4 (&d)[IPT], float4 (&fx)[IPT], float4 (&fy)[IPT], float4 (&fz)[IPT], float4 (&fw)[IPT],
                                           const TileCtx& tc, const FusedOut& out, const Geom& g, const StepConsts& c, const int lane,
                                           const int M) {
  constexpr int SX = C::SX, PLANE = C::PLANE;
  const float* sH = st + PLANE;
  float* sFxp = st + 2 * PLANE; float* sFxm = st + 3 * PLANE; float* sFyp = st + 4 * PLANE; float* sFym = st + 5 * PLANE;
  float total[IPT][4], scale[IPT][4];
  bool need = false;
  // ---- stage A: neighbour water levels, raw outflow (no branches: the two items interleave) ----
#pragma unroll
  for (int q = 0; q < IPT; ++q) {
    const int oq = o[q];
    const float4 HC = add4(d[q], h[q]);                                              // a + r, flowUpdate.comp:34
    const float4 HU = ld4(sH + oq - SX), HD = ld4(sH + oq + SX);
    float HL = __shfl_up_sync(0xffffffffu, HC.w, 1);
    float HR = __shfl_down_sync(0xffffffffu, HC.x, 1);
    HL = lds_if(lane == 0, sH + oq - 1, HL);
    HR = lds_if(lane == 31, sH + oq + 4, HR);
    if (!EDGE && TWS_PACKED) {
      flux_raw4(HC, HU, HD, HL, HR, fx[q], fy[q], fz[q], fw[q], c, total[q]);
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
    const int r = oq / SX, x = oq - r * SX;
    const int gy = tc.gy0 + r, gx0 = tc.sx0 + x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float Hc = comp(HC, i);
      float hxp = (i < 3) ? comp(HC, i + 1) : HR;
      float hxm = (i > 0) ? comp(HC, i - 1) : HL;
      float hyp = comp(HD, i), hym = comp(HU, i);
      if (EDGE && c.closed) {
        const int gx = gx0 + i;
        if (gx + 1 >= g.W) hxp = Hc;
        if (gx - 1 < 0) hxm = Hc;
        if (gy + 1 >= g.Hg) hyp = Hc;
        if (gy - 1 < 0) hym = Hc;
      }
      total[q][i] = flux_raw(Hc, hxp, hxm, hyp, hym, pfx[i], pfy[i], pfz[i], pfw[i], c);
      const float dep = comp(d[q], i);
      const bool over = total[q][i] > dep;                                           // :58
      scale[q][i] = over ? 0.0f : 1.0f;              // a == 0 -> a/total == +0 ; total <= a -> no scaling (x*1 == x)
      need = need || (over && dep != 0.0f);
    }
  }
  // ---- stage B: the rare IEEE divisions (a wet cell that would drain completely this step) ----
  if (need) {
#pragma unroll
    for (int q = 0; q < IPT; ++q)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float dep = comp(d[q], i);
        if (total[q][i] > dep && dep != 0.0f) scale[q][i] = __fdiv_rn(dep, total[q][i]);   // :59
      }
  }
  // ---- stage C: scale, publish what the neighbours need, last level: write the flux to HBM ----
#pragma unroll
  for (int q = 0; q < IPT; ++q) {
    const int oq = o[q];
    float* pfx = &fx[q].x; float* pfy = &fy[q].x; float* pfz = &fz[q].x; float* pfw = &fw[q].x;
    const int r = oq / SX, x = oq - r * SX;
    const int gy = tc.gy0 + r, gx0 = tc.sx0 + x;
    const bool row_in = (unsigned)gy < (unsigned)g.Hg;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      pfx[i] = __fmul_rn(pfx[i], scale[q][i]); pfy[i] = __fmul_rn(pfy[i], scale[q][i]);
      pfz[i] = __fmul_rn(pfz[i], scale[q][i]); pfw[i] = __fmul_rn(pfw[i], scale[q][i]);
      if (EDGE && !(row_in && (unsigned)(gx0 + i) < (unsigned)g.W)) { pfx[i] = 0.f; pfy[i] = 0.f; pfz[i] = 0.f; pfw[i] = 0.f; }
    }
    // rows outside this level's margin M around the output tile feed nothing that is kept
    const bool live = valid[q] && r >= C::HY - M && r < C::HY + C::OY + M;
    sts4_if(live, sFyp + oq, fz[q]);
    sts4_if(live, sFym + oq, fw[q]);
    sts_if(live && lane == 31, sFxp + oq + 3, fx[q].w);        // for lane 0 of the next warp
    sts_if(live && lane == 0, sFxm + oq, fy[q].x);             // for lane 31 of the previous warp
    if (EDGE || LAST) {
      bool store = valid[q] && x >= C::HX && x < C::HX + C::OX && r >= C::HY && r < C::HY + C::OY;
      if (EDGE) store = store && tc.ly0 + r < g.rows && gx0 < g.pitch;
      if (EDGE && store) ledger_add(c, g, gx0, gy, fx[q], fy[q], fz[q], fw[q]);      // every sub-step, the tile's own output cells only
      if (LAST && store) {
        const size_t go = (size_t)((long long)(tc.ly0 + r) * g.pitch) + gx0;
        st4(out.F[0] + go, fx[q]); st4(out.F[1] + go, fy[q]); st4(out.F[2] + go, fz[q]); st4(out.F[3] + go, fw[q]);
      }
    }
  }
}

template <class C, bool EDGE, bool LAST, int IPT>
__device__ __forceinline__ void depth_level(float* __restrict__ st, const int (&o)[IPT], const bool (&valid)[IPT], const float4 (&h)[IPT],
                                            float4 (&d)[IPT], const float4 (&fx)[IPT], const float4 (&fy)[IPT], const float4 (&fz)[IPT],
                                            const float4 (&fw)[IPT], const TileCtx& tc, const FusedOut& out, const Geom& g,
                                            const StepConsts& c, const int lane, const int M, double& src_acc) {
  constexpr int SX = C::SX, PLANE = C::PLANE;
  float* sH = st + PLANE;
  const float* sFxp = st + 2 * PLANE; const float* sFxm = st + 3 * PLANE; const float* sFyp = st + 4 * PLANE; const float* sFym = st + 5 * PLANE;
#pragma unroll
  for (int q = 0; q < IPT; ++q) {
    const int oq = o[q];
    const float4 iy1 = ld4(sFym + oq + SX);            // F(x,y+1).w, flowApply.comp:34
    const float4 iy0 = ld4(sFyp + oq - SX);            // F(x,y-1).z, :35
    float l = __shfl_up_sync(0xffffffffu, fx[q].w, 1);       // F(x-1,y).x, :33
    float rgt = __shfl_down_sync(0xffffffffu, fy[q].x, 1);   // F(x+1,y).y, :32
    l = lds_if(lane == 0, sFxp + oq - 1, l);
    rgt = lds_if(lane == 31, sFxm + oq + 4, rgt);
    const int r = oq / SX, x = oq - r * SX;
    const int gy = tc.gy0 + r, gx0 = tc.sx0 + x;
    const bool row_in = (unsigned)gy < (unsigned)g.Hg;
    float nd[4]; uint32_t nv[4];
    float4 ds4 = make_float4(0.f, 0.f, 0.f, 0.f);                                     // EXT ledger: what the sources changed
    if (!EDGE && TWS_PACKED) {
      float4 nd4; uint4 nv4 = make_uint4(0u, 0u, 0u, 0u);
      apply4<LAST>(d[q], fx[q], fy[q], fz[q], fw[q], l, rgt, iy1, iy0, c, c.ext_sources != 0, nd4, nv4, &ds4);
      nd[0] = nd4.x; nd[1] = nd4.y; nd[2] = nd4.z; nd[3] = nd4.w;
      nv[0] = nv4.x; nv[1] = nv4.y; nv[2] = nv4.z; nv[3] = nv4.w;
    } else {
      float* pds = &ds4.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float iX1 = (i < 3) ? comp(fy[q], i + 1) : rgt;
      const float iX0 = (i > 0) ? comp(fx[q], i - 1) : l;
      float vx, vy;
      nd[i] = apply_cell_src(comp(d[q], i), comp(fx[q], i), comp(fy[q], i), comp(fz[q], i), comp(fw[q], i), iX1, iX0, comp(iy1, i),
                             comp(iy0, i), c, c.ext_sources != 0, vx, vy, pds[i]);
      if (LAST) nv[i] = pack_half2(vx, vy);
      if (EDGE && !(row_in && (unsigned)(gx0 + i) < (unsigned)g.W)) { nd[i] = 0.f; pds[i] = 0.f; if (LAST) nv[i] = 0u; }
    }
    }
    if (c.ledger_src != nullptr) {                     // every sub-step, the tile's own output cells only
      bool own = valid[q] && x >= C::HX && x < C::HX + C::OX && r >= C::HY && r < C::HY + C::OY;
      if (EDGE) own = own && tc.ly0 + r < g.rows && gx0 < g.pitch;
      if (own) src_acc += ((double)ds4.x + (double)ds4.y) + ((double)ds4.z + (double)ds4.w);
    }
    if (!LAST) {
      const bool live = valid[q] && r >= C::HY - M && r < C::HY + C::OY + M;
      d[q] = make_float4(nd[0], nd[1], nd[2], nd[3]);
      sts4_if(live, sH + oq, add4(d[q], h[q]));
    } else {
      bool store = valid[q] && x >= C::HX && x < C::HX + C::OX && r >= C::HY && r < C::HY + C::OY;
      if (EDGE) store = store && tc.ly0 + r < g.rows && gx0 < g.pitch;
      if (store) {
        const size_t go = (size_t)((long long)(tc.ly0 + r) * g.pitch) + gx0;
        st4(out.d + go, make_float4(nd[0], nd[1], nd[2], nd[3]));
        *reinterpret_cast<uint4*>(out.v + go) = make_uint4(nv[0], nv[1], nv[2], nv[3]);
      }
    }
  }
}

template <class C, bool EDGE>
__device__ __forceinline__ void run_tile(float* st, const TileCtx& tc, const FusedOut& out, const Geom& g, const StepConsts& c, const int tid,
                                         const BrushArgs& br) {
  constexpr int K = C::K, NG = C::NG, SX = C::SX, SY = C::SY, PLANE = C::PLANE, NT = C::NT;
  constexpr int ITEMS = NG * (SY - 2);                 // owned items: staged rows 1..SY-2
  constexpr int IPT = (ITEMS + NT - 1) / NT;
  static_assert(IPT <= 2, "tile too large for the register-resident scheme");
  static_assert(2 * NG <= NT, "halo rows are handled by the first 2*NG threads");
  const int lane = tid & 31;
  double src_acc = 0.0;                                // EXT ledger of the sources, flushed once per tile
  float* sh = st; float* sd = st + PLANE;
  int o[IPT]; bool valid[IPT];
  float4 h[IPT], d[IPT], fx[IPT], fy[IPT], fz[IPT], fw[IPT];
  // ---- pull the owned items out of the landing planes; publish H = d + h -----------------------
#pragma unroll
  for (int q = 0; q < IPT; ++q) {
    const int a = tid + q * NT;
    valid[q] = a < ITEMS;
    o[q] = (NG + (valid[q] ? a : ITEMS - 1)) * 4;      // row 1 starts NG float4 groups into the plane
    h[q] = ld4(sh + o[q]); d[q] = ld4(sd + o[q]);
    if (br.active) { const int r = o[q] / SX; d[q] = brush4(d[q], tc.sx0 + (o[q] - r * SX), tc.gy0 + r, br); }   // a pending brush, folded in
    fx[q] = ld4(st + 2 * PLANE + o[q]); fy[q] = ld4(st + 3 * PLANE + o[q]);
    fz[q] = ld4(st + 4 * PLANE + o[q]); fw[q] = ld4(st + 5 * PLANE + o[q]);
  }
#pragma unroll
  for (int q = 0; q < IPT; ++q)
    if (valid[q]) st4(sd + o[q], add4(d[q], h[q]));    // each position is read and rewritten by its owner only
  if (tid < 2 * NG) {                                  // staged rows 0 and SY-1 are never owned: H only
    const int oh = (tid < NG ? tid : (SY - 2) * NG + tid) * 4;
    float4 dh = ld4(sd + oh);
    if (br.active) { const int r = oh / SX; dh = brush4(dh, tc.sx0 + (oh - r * SX), tc.gy0 + r, br); }
    st4(sd + oh, add4(dh, ld4(sh + oh)));
  }
  __syncthreads();
#pragma unroll 1
  for (int t = 1; t < K; ++t) {
    flux_level<C, EDGE, false, IPT>(st, o, valid, h, d, fx, fy, fz, fw, tc, out, g, c, lane, 2 * (K - t) + 1);
    __syncthreads();
    depth_level<C, EDGE, false, IPT>(st, o, valid, h, d, fx, fy, fz, fw, tc, out, g, c, lane, 2 * (K - t), src_acc);
    __syncthreads();
  }
  flux_level<C, EDGE, true, IPT>(st, o, valid, h, d, fx, fy, fz, fw, tc, out, g, c, lane, 1);
  __syncthreads();
  depth_level<C, EDGE, true, IPT>(st, o, valid, h, d, fx, fy, fz, fw, tc, out, g, c, lane, 0, src_acc);
  if (c.ledger_src != nullptr) ledger_src_flush(c.ledger_src, src_acc);
}

template <class C>
__global__ void __launch_bounds__(C::NT, 1) fused_step_kernel(const __grid_constant__ CUtensorMap tm_h,
                                                              const __grid_constant__ CUtensorMap tm_d,
                                                              const __grid_constant__ CUtensorMap tm_f0,
                                                              const __grid_constant__ CUtensorMap tm_f1,
                                                              const __grid_constant__ CUtensorMap tm_f2,
                                                              const __grid_constant__ CUtensorMap tm_f3,
                                                              FusedOut out, Geom g, StepConsts c, int ty0, int tiles_x, int n_tiles,
                                                              int tma_y_bias, BrushArgs br) {
  constexpr int OX = C::OX, OY = C::OY, HX = C::HX, HY = C::HY, SX = C::SX, SY = C::SY, PLANE = C::PLANE, STAGE = C::STAGE;
  extern __shared__ __align__(1024) float smem[];
  __shared__ uint64_t full[2];
  const int tid = threadIdx.x;

  auto tile_ctx = [&](int tile) {
    const int by = tile / tiles_x, bx = tile - by * tiles_x;
    TileCtx tc;
    tc.sx0 = bx * OX - HX;
    tc.ly0 = (ty0 + by) * OY - HY;
    tc.gy0 = g.row0 + tc.ly0;
    return tc;
  };
  auto issue = [&](int tile, int stage) {          // one thread
    const TileCtx tc = tile_ctx(tile);
    float* st = smem + stage * STAGE;
    mbar_expect_tx(&full[stage], (uint32_t)(STAGE * sizeof(float)));
    const int ty = tc.ly0 + tma_y_bias;
    tma_load_2d(st, &tm_h, tc.sx0, ty, &full[stage]);
    tma_load_2d(st + PLANE, &tm_d, tc.sx0, ty, &full[stage]);
    tma_load_2d(st + 2 * PLANE, &tm_f0, tc.sx0, ty, &full[stage]);
    tma_load_2d(st + 3 * PLANE, &tm_f1, tc.sx0, ty, &full[stage]);
    tma_load_2d(st + 4 * PLANE, &tm_f2, tc.sx0, ty, &full[stage]);
    tma_load_2d(st + 5 * PLANE, &tm_f3, tc.sx0, ty, &full[stage]);
  };

  const int first = blockIdx.x, stride = gridDim.x;
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (first < n_tiles) issue(first, 0);
    if (first + stride < n_tiles) issue(first + stride, 1);
  }
  __syncthreads();

  int j = 0;
  for (int tile = first; tile < n_tiles; tile += stride, ++j) {
    const int stage = j & 1;
    float* st = smem + stage * STAGE;
    const TileCtx tc = tile_ctx(tile);
    mbar_wait(&full[stage], (uint32_t)((j >> 1) & 1));
    // Tiles whose staged window lies fully inside the grid and whose outputs are all owned rows
    // need no exterior masks, no boundary-mode handling and no store guards.
    const bool edge = tc.sx0 < 0 || tc.sx0 + SX > g.W || tc.gy0 < 0 || tc.gy0 + SY > g.Hg || tc.ly0 + HY + OY > g.rows;
    if (edge) run_tile<C, true>(st, tc, out, g, c, tid, br);
    else run_tile<C, false>(st, tc, out, g, c, tid, br);
    __syncthreads();                                  // every read of this stage is done
    if (tid == 0 && tile + 2 * stride < n_tiles) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(tile + 2 * stride, stage);
    }
  }
}

// ---- per-K tile configuration ---------------------------------------------------------
template <int K> struct CfgFor;
#ifndef TWS_FUSED_NT
#define TWS_FUSED_NT 512
#endif
template <> struct CfgFor<1> { using type = FusedCfg<1, 128, 28, TWS_FUSED_NT>; };
template <> struct CfgFor<2> { using type = FusedCfg<2, 128, 24, TWS_FUSED_NT>; };
template <> struct CfgFor<3> { using type = FusedCfg<3, 120, 20, TWS_FUSED_NT>; };
template <> struct CfgFor<4> { using type = FusedCfg<4, 120, 16, TWS_FUSED_NT>; };

static int sm_count_of_current_device() {
  static int cached[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!cached[dev & 63]) cudaDeviceGetAttribute(&cached[dev & 63], cudaDevAttrMultiProcessorCount, dev);
  return cached[dev & 63] > 0 ? cached[dev & 63] : 148;
}

template <int K>
static cudaError_t launch_fused_k(const Geom& g, const Planes& p, const TmaSet& tma, int src, const StepConsts& c, int ty0,
                                  int ty1, cudaStream_t st, const BrushArgs* brush) {
  using C = typename CfgFor<K>::type;
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  auto kern = fused_step_kernel<C>;
  if (!attr_set[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    if (e != cudaSuccess) return e;
    attr_set[dev & 63] = true;
  }
  if (ty1 <= ty0) return cudaSuccess;
  const int dst = 1 - src;
  const size_t row0_off = (size_t)TWS_HALO_ROWS * g.pitch;
  FusedOut out;
  out.d = p.d[dst] + row0_off;
  for (int i = 0; i < 4; ++i) out.F[i] = p.F[dst][i] + row0_off;
  out.v = p.v + row0_off;
  const int tiles_x = (g.W + C::OX - 1) / C::OX;
  const long long n_tiles = (long long)tiles_x * (ty1 - ty0);
  if (n_tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
  const int sms = sm_count_of_current_device();
  const int grid = (int)(n_tiles < sms ? n_tiles : sms);     // one persistent CTA per SM
  const int bias = g.has_up ? TWS_HALO_ROWS : 0;
  BrushArgs br{};
  if (brush != nullptr) br = *brush;
  kern<<<grid, C::NT, C::SMEM, st>>>(tma.m[0], tma.m[1], tma.m[2], tma.m[3], tma.m[4], tma.m[5], out, g, c, ty0, tiles_x, (int)n_tiles, bias, br);
  return cudaGetLastError();
}

int fused_out_rows_per_tile(int K) {
  switch (K) {
    case 1: return CfgFor<1>::type::OY;
    case 2: return CfgFor<2>::type::OY;
    case 3: return CfgFor<3>::type::OY;
    default: return CfgFor<4>::type::OY;
  }
}
int fused_tile_rows(int K, int rows) { const int oy = fused_out_rows_per_tile(K); return (rows + oy - 1) / oy; }

cudaError_t launch_fused(int K, const Geom& g, const Planes& p, const TmaSet& tma, int src, const StepConsts& c, int ty0,
                         int ty1, cudaStream_t st, const BrushArgs* brush) {
  switch (K) {
    case 1: return launch_fused_k<1>(g, p, tma, src, c, ty0, ty1, st, brush);
    case 2: return launch_fused_k<2>(g, p, tma, src, c, ty0, ty1, st, brush);
    case 3: return launch_fused_k<3>(g, p, tma, src, c, ty0, ty1, st, brush);
    case 4: return launch_fused_k<4>(g, p, tma, src, c, ty0, ty1, st, brush);
    default: return cudaErrorInvalidValue;
  }
}

// ---- TMA descriptors --------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = (EncodeTiledFn)p;
  return fn;
}

static void box_for(int K, int* sx, int* sy) {
  switch (K) {
    case 1: *sx = CfgFor<1>::type::SX; *sy = CfgFor<1>::type::SY; break;
    case 2: *sx = CfgFor<2>::type::SX; *sy = CfgFor<2>::type::SY; break;
    case 3: *sx = CfgFor<3>::type::SX; *sy = CfgFor<3>::type::SY; break;
    default: *sx = CfgFor<4>::type::SX; *sy = CfgFor<4>::type::SY; break;
  }
}

// Descriptors for reading side `side` with boxes of box_x x box_y cells.  The tensor covers the
// rows that hold real data: own rows plus the halo rows towards an existing neighbour;
// everything outside is the global exterior and is zero-filled by the TMA unit.
cudaError_t build_tma_boxes(const Geom& g, const Planes& p, int side, int box_x, int box_y, TmaSet* out, std::string* err, int l2_promotion) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { if (err) *err = "cuTensorMapEncodeTiled not available from the driver"; return cudaErrorNotSupported; }
  const int first_row = g.has_up ? 0 : TWS_HALO_ROWS;   // plane row of the first visible row
  const int vis_rows = g.rows + (g.has_up ? TWS_HALO_ROWS : 0) + (g.has_down ? TWS_HALO_ROWS : 0);
  float* bases[6] = {p.h, p.d[side], p.F[side][0], p.F[side][1], p.F[side][2], p.F[side][3]};
  for (int i = 0; i < 6; ++i) {
    cuuint64_t dims[2] = {(cuuint64_t)g.W, (cuuint64_t)vis_rows};
    cuuint64_t strides[1] = {(cuuint64_t)g.pitch * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)box_x, (cuuint32_t)box_y};
    cuuint32_t estr[2] = {1, 1};
    void* base = bases[i] + (size_t)first_row * g.pitch;
    const CUtensorMapL2promotion promo = l2_promotion == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                       : l2_promotion == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
    CUresult r = enc(&out->m[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      if (err) *err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r);
      return cudaErrorInvalidValue;
    }
  }
  return cudaSuccess;
}

// Row descriptors of the streaming kernels: m[0] = terrain (2-D, box box_x x 1), m[1] = the five state
// planes d, F+X, F-X, F+Y, F-Y of side `side` as ONE 3-D tensor (x, row, plane; box box_x x 1 x 5) —
// layout_planes() keeps them consecutive with a uniform stride.  Same row visibility as build_tma_boxes.
cudaError_t build_tma_rows(const Geom& g, const Planes& p, int side, int box_x, TmaSet* out, std::string* err, int l2_promotion) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { if (err) *err = "cuTensorMapEncodeTiled not available from the driver"; return cudaErrorNotSupported; }
  const int first_row = g.has_up ? 0 : TWS_HALO_ROWS;
  const int vis_rows = g.rows + (g.has_up ? TWS_HALO_ROWS : 0) + (g.has_down ? TWS_HALO_ROWS : 0);
  const CUtensorMapL2promotion promo = l2_promotion == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                     : l2_promotion == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
  const ptrdiff_t plane_stride = (const char*)p.F[side][0] - (const char*)p.d[side];
  for (int i = 0; i < 4; ++i)
    if ((const char*)p.F[side][i] - (const char*)p.d[side] != (i + 1) * plane_stride) {
      if (err) *err = "state planes are not uniformly strided";
      return cudaErrorInvalidValue;
    }
  CUresult r;
  {
    cuuint64_t dims[2] = {(cuuint64_t)g.W, (cuuint64_t)vis_rows};
    cuuint64_t strides[1] = {(cuuint64_t)g.pitch * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)box_x, 1};
    cuuint32_t estr[2] = {1, 1};
    r = enc(&out->m[0], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, p.h + (size_t)first_row * g.pitch, dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r == CUDA_SUCCESS) {
    cuuint64_t dims[3] = {(cuuint64_t)g.W, (cuuint64_t)vis_rows, 5};
    cuuint64_t strides[2] = {(cuuint64_t)g.pitch * sizeof(float), (cuuint64_t)plane_stride};
    cuuint32_t box[3] = {(cuuint32_t)box_x, 1, 5};
    cuuint32_t estr[3] = {1, 1, 1};
    r = enc(&out->m[1], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p.d[side] + (size_t)first_row * g.pitch, dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r);
    return cudaErrorInvalidValue;
  }
  return cudaSuccess;
}

cudaError_t fused_build_tma(int K, const Geom& g, const Planes& p, int side, TmaSet* out, std::string* err) {
  int sx, sy;
  box_for(K, &sx, &sy);
  return build_tma_boxes(g, p, side, sx, sy, out, err, 2);
}

// ---- unfused launchers -------------------------------------------------------------------
cudaError_t launch_unfused_update(const Geom& g, const Planes& p, int side, const StepConsts& c, int lr0, int lr1, cudaStream_t st) {
  if (lr1 <= lr0) return cudaSuccess;
  const size_t off = (size_t)TWS_HALO_ROWS * g.pitch;
  dim3 block(128), grid((g.pitch / 4 + 127) / 128, (lr1 - lr0) < 65535 ? (lr1 - lr0) : 65535);
  unfused_update_kernel<<<grid, block, 0, st>>>(g, p.h + off, p.d[side] + off, p.F[side][0] + off, p.F[side][1] + off,
                                                 p.F[side][2] + off, p.F[side][3] + off, c, lr0, lr1);
  return cudaGetLastError();
}

cudaError_t launch_unfused_apply(const Geom& g, const Planes& p, int side, const StepConsts& c, int lr0, int lr1, cudaStream_t st) {
  if (lr1 <= lr0) return cudaSuccess;
  const size_t off = (size_t)TWS_HALO_ROWS * g.pitch;
  dim3 block(128), grid((g.pitch / 4 + 127) / 128, (lr1 - lr0) < 65535 ? (lr1 - lr0) : 65535);
  unfused_apply_kernel<<<grid, block, 0, st>>>(g, p.d[side] + off, p.F[side][0] + off, p.F[side][1] + off, p.F[side][2] + off,
                                                p.F[side][3] + off, p.v + off, c, lr0, lr1);
  return cudaGetLastError();
}

}  // namespace tws
