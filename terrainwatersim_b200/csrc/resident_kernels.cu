// resident_kernels.cu — a whole frame of steps in ONE launch for grids that fit on chip (the reference's own
// operating point: 1024 x 1024, up to 10 steps per frame, Terrain.cpp:240-265).
//
// The six state planes of a 1024^2 grid are 25 MB; the 148 SMs of a B200 hold 33 MB of shared memory.  So the grid is cut
// into one OX x OY block per SM (1024^2: 4 x 37 blocks of 256 x 28 cells), every CTA keeps its block — terrain, depth and
// the four outflow planes, plus a halo of 2 rows / 4 columns — in shared memory for ALL n steps of the frame, and per step
// only the block's rim goes through L2: after a step a CTA stores the depth of its outer two rings and the outflow of its
// outer ring into the ping-pong planes in HBM/L2 (where the other kernels expect the state anyway), raises its flag, and
// its eight neighbours pick their halo cells up from there.  The launch is cooperative (all CTAs are resident by
// construction), the synchronisation is neighbour-to-neighbour (acquire / release flags in the control block), there is no
// grid-wide barrier.
//
// A step's passes are ordered so that the exchange overlaps the arithmetic that does not depend on it.  The compute warps run
//     depth(rim), publishing it -> A -> depth(interior) -> C -> flux(interior) of the NEXT step -> B -> flux(rim + halo ring) -> C
// while a few extra warps do the exchange between A and B: raise the block's flag, wait for the eight neighbours' flags,
// fetch the halo cells with cp.async (L2 -> shared memory, no registers; the addresses are per-thread constants).  A, B, C
// are named barriers (A, B: everybody; C: the compute warps).  Measured with %globaltimer stamps (round 2): with one
// exchange warp and the whole rim in the late pass, the chain publish -> flag -> poll -> fetch -> flux(late) -> depth(rim)
// of 7 us, not the arithmetic (5.6 us), set the step time at 1024^2.  The halo ring's outflow is recomputed locally from its exchanged old
// outflow and depth (the same redundant-halo scheme as the tile kernel, one cell deep), so one exchange per step suffices.
//
// Arithmetic: the same cell functions as every other kernel (cell_math.cuh; flowUpdate.comp:34-59, flowApply.comp:32-46),
// bit-identical to the oracle.  Exterior cells are zeros in shared memory (the reference's out-of-range imageLoad) and are
// never updated.
#include "cell_math.cuh"

#include <cstdio>

namespace tws {

namespace {

template <int OX_, int OY_, int NT_, int NX_>
struct ResCfg {
  static constexpr int OX = OX_, OY = OY_, NT = NT_;
  static constexpr int HX = 4, HY = 2;                       // halo: one float4 group / two rows on each side
  static constexpr int SX = OX + 2 * HX, SY = OY + 2 * HY;   // staged block
  static constexpr int NG = SX / 4;                          // float4 groups ("items") per staged row
  static constexpr int PLANE = SX * SY;                      // floats per plane
  static constexpr size_t SMEM = (size_t)6 * PLANE * sizeof(float);   // h, d, F x4
  static constexpr int NX = NX_;                             // threads of the exchange warps
  // item sets, as (row, group) rectangles of the staged block:
  //   interior  own items that neither a neighbour nor the halo touches: rows [4, SY-4) x groups [2, NG-2)
  //   rim       the other own items: the outer two rows and the outer group on each side (what the neighbours read)
  //   ring      the halo items next to the block: rows 1 and SY-2, groups 0 and NG-1 (their outflow is recomputed here)
  //   far       rows 0 and SY-1: only their depth is needed (water level of the ring's outer neighbours)
  // The outflow of an item needs the water level one cell around it: everything except the block's outermost row / group on
  // each side ("early": rows [3, SY-3) x groups [2, NG-2)) can be updated before the halo of the step has arrived, the
  // outermost own items and the ring ("late") after it.
  static constexpr int IN_ROWS = SY - 8, IN_COLS = NG - 4, N_IN = IN_ROWS * IN_COLS;
  static constexpr int RIM_W = NG - 2, N_RIM = 4 * RIM_W + 2 * IN_ROWS;
  static constexpr int N_EARLY = (IN_ROWS + 2) * IN_COLS;
  static constexpr int N_RING = 2 * NG + 2 * (SY - 4);
  static constexpr int N_OUTER = 2 * RIM_W + 2 * (IN_ROWS + 2);   // rows 2 and SY-3, groups 1 and NG-2 of rows [3, SY-3)
  static constexpr int N_LATE = N_OUTER + N_RING;
  static constexpr int N_FAR = 2 * NG;
  static constexpr int N_HALO = N_RING + N_FAR;
  static constexpr int IPT_IN = (N_IN + NT - 1) / NT, IPT_EARLY = (N_EARLY + NT - 1) / NT, IPT_LATE = (N_LATE + NT - 1) / NT,
                       IPT_RIM = (N_RIM + NT - 1) / NT, IPT_HALO = (N_HALO + NX - 1) / NX;
  static_assert(OX % 4 == 0 && OY >= 4 && OX >= 16, "block too small for the interior / rim split");
  static_assert(PLANE < (1 << 16), "item descriptors keep the plane offset in 16 bits");
};

// item descriptor: plane offset of the group's first cell | class << 16 | flags
constexpr int kClsSkip = 0;        // wholly outside the grid: stays zero
constexpr int kClsFast = 1;        // four cells inside, no boundary-mode handling: packed arithmetic
constexpr int kClsSlow = 2;        // partly outside (width not a multiple of 4) or on a CLOSED border: scalar, masked
constexpr int kFlagBorder = 1 << 18;   // own item with a cell on the border of the global grid (EXT ledger)
constexpr int kFlagOuter = 1 << 19;    // own item in the outermost ring: its outflow is published, too
constexpr int kFlagOwn = 1 << 20;

struct ResPlanes {                 // plane pointers at LOCAL row 0
  const float* h;
  float* d[2];
  float* F[2][4];
  uint32_t* v;
};

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src) {     // L2 -> shared memory, bypassing L1
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
template <int ID, int COUNT> __device__ __forceinline__ void bar_named() {
  asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory");
}
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void stcg4(float* p, const float4& v) { __stcg(reinterpret_cast<float4*>(p), v); }

template <class C> __device__ __forceinline__ void pos_interior(int i, int& r, int& gc) {
  r = 4 + i / C::IN_COLS; gc = 2 + i % C::IN_COLS;
}
template <class C> __device__ __forceinline__ void pos_rim(int i, int& r, int& gc) {
  if (i < 2 * C::RIM_W) { r = 2 + i / C::RIM_W; gc = 1 + i % C::RIM_W; return; }
  i -= 2 * C::RIM_W;
  if (i < 2 * C::RIM_W) { r = C::SY - 4 + i / C::RIM_W; gc = 1 + i % C::RIM_W; return; }
  i -= 2 * C::RIM_W;
  r = 4 + (i >> 1); gc = (i & 1) ? C::NG - 2 : 1;
}
template <class C> __device__ __forceinline__ void pos_early(int i, int& r, int& gc) {
  r = 3 + i / C::IN_COLS; gc = 2 + i % C::IN_COLS;
}
template <class C> __device__ __forceinline__ void pos_outer(int i, int& r, int& gc) {
  if (i < 2 * C::RIM_W) { r = i < C::RIM_W ? 2 : C::SY - 3; gc = 1 + (i < C::RIM_W ? i : i - C::RIM_W); return; }
  i -= 2 * C::RIM_W;
  r = 3 + (i >> 1); gc = (i & 1) ? C::NG - 2 : 1;
}
template <class C> __device__ __forceinline__ void pos_ring(int i, int& r, int& gc) {
  if (i < C::NG) { r = 1; gc = i; return; }
  i -= C::NG;
  if (i < C::NG) { r = C::SY - 2; gc = i; return; }
  i -= C::NG;
  r = 2 + (i >> 1); gc = (i & 1) ? C::NG - 1 : 0;
}
template <class C> __device__ __forceinline__ void pos_far(int i, int& r, int& gc) {
  r = i < C::NG ? 0 : C::SY - 1; gc = i < C::NG ? i : i - C::NG;
}

// descriptor of the item at staged (r, gc) of the block whose staged cell (0, 0) is global (gx_base, gy_base)
template <class C>
__device__ __forceinline__ int describe(int r, int gc, int gx_base, int gy_base, const Geom& g, const StepConsts& c) {
  const int gy = gy_base + r, gx0 = gx_base + 4 * gc;
  int desc = (r * C::NG + gc) * 4;
  const bool own = r >= C::HY && r < C::HY + C::OY && gc >= 1 && gc < C::NG - 1;
  if (own) desc |= kFlagOwn;
  if (own && (r == C::HY || r == C::HY + C::OY - 1 || gc == 1 || gc == C::NG - 2)) desc |= kFlagOuter;
  if ((unsigned)gy >= (unsigned)g.Hg || gx0 >= g.W || gx0 + 3 < 0) return desc | (kClsSkip << 16);
  const bool whole = gx0 >= 0 && gx0 + 3 < g.W;
  const bool border = gy == 0 || gy == g.Hg - 1 || gx0 == 0 || gx0 + 3 >= g.W - 1;
  if (own && border) desc |= kFlagBorder;
  const bool fast = whole && !(c.closed && border);
  return desc | ((fast ? kClsFast : kClsSlow) << 16);
}

// ---- flowUpdate.comp for one item, in place in shared memory ---------------------------------------------------------
template <class C>
__device__ __forceinline__ void res_flux_item(float* __restrict__ st, const int desc, const int gx_base, const int gy_base,
                                              const Geom& g, const StepConsts& c, double& out_acc) {
  constexpr int SX = C::SX, PLANE = C::PLANE;
  const int cls = (desc >> 16) & 3;
  if (cls == kClsSkip) return;
  const int o = desc & 0xffff;
  const float* sh = st; const float* sd = st + PLANE;
  float* sF0 = st + 2 * PLANE; float* sF1 = st + 3 * PLANE; float* sF2 = st + 4 * PLANE; float* sF3 = st + 5 * PLANE;
  const float4 d = ld4(sd + o);
  const float4 HC = add4(d, ld4(sh + o));                                              // a + r, flowUpdate.comp:34
  const float4 HU = add4(ld4(sd + o - SX), ld4(sh + o - SX)), HD = add4(ld4(sd + o + SX), ld4(sh + o + SX));
  const float HL = __fadd_rn(sd[o - 1], sh[o - 1]), HR = __fadd_rn(sd[o + 4], sh[o + 4]);
  float4 fx = ld4(sF0 + o), fy = ld4(sF1 + o), fz = ld4(sF2 + o), fw = ld4(sF3 + o);
  if (cls == kClsFast) {
    float total[4], s[4];
    flux_raw4(HC, HU, HD, HL, HR, fx, fy, fz, fw, c, total);
    const float dep[4] = {d.x, d.y, d.z, d.w};
    flux_scale4(total, dep, s);                                                          // :58-59
    fx = make_float4(__fmul_rn(fx.x, s[0]), __fmul_rn(fx.y, s[1]), __fmul_rn(fx.z, s[2]), __fmul_rn(fx.w, s[3]));
    fy = make_float4(__fmul_rn(fy.x, s[0]), __fmul_rn(fy.y, s[1]), __fmul_rn(fy.z, s[2]), __fmul_rn(fy.w, s[3]));
    fz = make_float4(__fmul_rn(fz.x, s[0]), __fmul_rn(fz.y, s[1]), __fmul_rn(fz.z, s[2]), __fmul_rn(fz.w, s[3]));
    fw = make_float4(__fmul_rn(fw.x, s[0]), __fmul_rn(fw.y, s[1]), __fmul_rn(fw.z, s[2]), __fmul_rn(fw.w, s[3]));
  } else {
    const int r = o / SX, x = o - r * SX;
    const int gy = gy_base + r, gx0 = gx_base + x;
    float* pfx = &fx.x; float* pfy = &fy.x; float* pfz = &fz.x; float* pfw = &fw.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gx = gx0 + i;
      const float Hc = comp(HC, i);
      float hxp = (i < 3) ? comp(HC, i + 1) : HR;
      float hxm = (i > 0) ? comp(HC, i - 1) : HL;
      float hyp = comp(HD, i), hym = comp(HU, i);
      if (c.closed) {
        if (gx + 1 >= g.W) hxp = Hc;
        if (gx - 1 < 0) hxm = Hc;
        if (gy + 1 >= g.Hg) hyp = Hc;
        if (gy - 1 < 0) hym = Hc;
      }
      flux_cell(Hc, hxp, hxm, hyp, hym, comp(d, i), pfx[i], pfy[i], pfz[i], pfw[i], c);
      if ((unsigned)gx >= (unsigned)g.W) { pfx[i] = 0.f; pfy[i] = 0.f; pfz[i] = 0.f; pfw[i] = 0.f; }
    }
  }
  st4(sF0 + o, fx); st4(sF1 + o, fy); st4(sF2 + o, fz); st4(sF3 + o, fw);
  if ((desc & kFlagBorder) && c.ledger != nullptr) {                                     // EXT: what leaves the map, every sub-step
    const int r = o / SX, x = o - r * SX;
    ledger_acc(out_acc, c, g, gx_base + x, gy_base + r, fx, fy, fz, fw);
  }
}

// ---- flowApply.comp for one own item --------------------------------------------------------------------------------
// PUBLISH: the item is part of the rim — its new depth (and, outer ring, its outflow) also goes to side `pub` of the planes
// in HBM/L2 for the neighbours.  LAST: the frame's final step — depth, outflow and the flow vector go to HBM.
template <class C, bool PUBLISH, bool LAST>
__device__ __forceinline__ void res_depth_item(float* __restrict__ st, const int desc, const int gx_base, const int gy_base,
                                               const Geom& g, const StepConsts& c, const ResPlanes& P, const int side, double& src_acc) {
  constexpr int SX = C::SX, PLANE = C::PLANE;
  const int cls = (desc >> 16) & 3;
  if (cls == kClsSkip) return;
  const int o = desc & 0xffff;
  float* sd = st + PLANE;
  const float* sF0 = st + 2 * PLANE; const float* sF1 = st + 3 * PLANE; const float* sF2 = st + 4 * PLANE; const float* sF3 = st + 5 * PLANE;
  const float4 d = ld4(sd + o);
  const float4 fx = ld4(sF0 + o), fy = ld4(sF1 + o), fz = ld4(sF2 + o), fw = ld4(sF3 + o);
  const float4 iy1 = ld4(sF3 + o + SX);              // F(x,y+1).w, flowApply.comp:34
  const float4 iy0 = ld4(sF2 + o - SX);              // F(x,y-1).z, :35
  const float l = sF0[o - 1];                        // F(x-1,y).x, :33
  const float rgt = sF1[o + 4];                      // F(x+1,y).y, :32
  const int r = o / SX, x = o - r * SX;
  const int gy = gy_base + r, gx0 = gx_base + x;
  float4 nd, ds = make_float4(0.f, 0.f, 0.f, 0.f);
  uint4 nv = make_uint4(0u, 0u, 0u, 0u);
  if (cls == kClsFast) {
    apply4<LAST>(d, fx, fy, fz, fw, l, rgt, iy1, iy0, c, c.ext_sources != 0, nd, nv, &ds);
  } else {
    float* pnd = &nd.x; float* pds = &ds.x; uint32_t* pnv = &nv.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float iX1 = (i < 3) ? comp(fy, i + 1) : rgt;
      const float iX0 = (i > 0) ? comp(fx, i - 1) : l;
      float vx, vy;
      pnd[i] = apply_cell_src(comp(d, i), comp(fx, i), comp(fy, i), comp(fz, i), comp(fw, i), iX1, iX0, comp(iy1, i), comp(iy0, i), c,
                              c.ext_sources != 0, vx, vy, pds[i]);
      pnv[i] = LAST ? pack_half2(vx, vy) : 0u;
      if ((unsigned)(gx0 + i) >= (unsigned)g.W) { pnd[i] = 0.f; pds[i] = 0.f; pnv[i] = 0u; }
    }
  }
  if (c.ledger_src != nullptr) src_acc += ((double)ds.x + (double)ds.y) + ((double)ds.z + (double)ds.w);
  if (!LAST) st4(sd + o, nd);
  if (PUBLISH || LAST) {
    const size_t go = (size_t)gy * (size_t)g.pitch + (size_t)gx0;
    stcg4(P.d[side] + go, nd);
    if (LAST || (desc & kFlagOuter)) {
      stcg4(P.F[side][0] + go, fx); stcg4(P.F[side][1] + go, fy); stcg4(P.F[side][2] + go, fz); stcg4(P.F[side][3] + go, fw);
    }
    if (LAST) *reinterpret_cast<uint4*>(P.v + go) = nv;
  }
}

// n steps of the whole grid; state side `src` -> side (src + n) & 1.  NT compute threads + NX exchange threads.
template <class C>
__global__ void __launch_bounds__(C::NT + C::NX, 1) resident_step_kernel(ResPlanes P, Geom g, StepConsts c, int nbx, int nby, int src, int n,
                                                                         uint32_t* flags, uint32_t epoch0, uint32_t* error) {
  constexpr int SY = C::SY, NG = C::NG, PLANE = C::PLANE, NT = C::NT, NX = C::NX, NALL = C::NT + C::NX;
  constexpr int BAR_C = 1, BAR_A = 2, BAR_B = 3, BAR_X = 4;
  extern __shared__ __align__(1024) float st[];
  const int tid = threadIdx.x;
  const int by = (int)blockIdx.x / nbx, bx = (int)blockIdx.x - by * nbx;
  const int gx_base = bx * C::OX - C::HX, gy_base = by * C::OY - C::HY;

  // ---- load the staged block (own cells + halo) from side src; exterior cells are zeros --------------------------------
  for (int a = tid; a < NG * SY; a += NALL) {
    const int r = a / NG, gc = a - r * NG;
    const int gy = gy_base + r, gx0 = gx_base + 4 * gc;
    const bool in = (unsigned)gy < (unsigned)g.Hg && (unsigned)gx0 < (unsigned)g.pitch;
    const size_t go = in ? (size_t)gy * (size_t)g.pitch + (size_t)gx0 : 0;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const int o = a * 4;
    st4(st + o, in ? ldcg4(P.h + go) : z);
    st4(st + PLANE + o, in ? ldcg4(P.d[src] + go) : z);
#pragma unroll
    for (int k = 0; k < 4; ++k) st4(st + (2 + k) * PLANE + o, in ? ldcg4(P.F[src][k] + go) : z);
  }
  __syncthreads();

  if (tid >= NT) {
    // ---- the exchange warps ---------------------------------------------------------------------------------------------
    const int xt = tid - NT;
    int watch = -1;                                       // thread j < 8 watches neighbour j
    if (xt < 8) {
      const int dx = (xt == 0 || xt == 3 || xt == 5) ? -1 : ((xt == 2 || xt == 4 || xt == 7) ? 1 : 0);
      const int dy = xt < 3 ? -1 : (xt < 5 ? 0 : 1);
      const int qx = bx + dx, qy = by + dy;
      if (qx >= 0 && qx < nbx && qy >= 0 && qy < nby) watch = qy * nbx + qx;
    }
    // this thread's halo items: shared-memory offset (negative: exterior or none, stays zero), plane offset, ring or far
    int ho[C::IPT_HALO]; unsigned hg[C::IPT_HALO]; bool hring[C::IPT_HALO];
#pragma unroll
    for (int q = 0; q < C::IPT_HALO; ++q) {
      const int i = xt + q * NX;
      int r = 0, gc = 0;
      hring[q] = i < C::N_RING;
      if (hring[q]) pos_ring<C>(i, r, gc); else if (i < C::N_HALO) pos_far<C>(i - C::N_RING, r, gc);
      const int gy = gy_base + r, gx0 = gx_base + 4 * gc;
      const bool in = i < C::N_HALO && (unsigned)gy < (unsigned)g.Hg && (unsigned)gx0 < (unsigned)g.pitch;
      ho[q] = in ? (r * NG + gc) * 4 : -1;
      hg[q] = in ? (unsigned)gy * (unsigned)g.pitch + (unsigned)gx0 : 0u;
    }
    bool dead = false;                                    // a wait timed out: stop waiting, the host reports the error
#pragma unroll 1
    for (int t = 1; t < n; ++t) {
      const int side = (src + t) & 1;
      bar_named<BAR_A, NALL>();                           // the rim of step t is on its way to L2; the halo cells are free
      if (xt == 0) st_release_u32(flags + blockIdx.x, epoch0 + (uint32_t)t);
      if (watch >= 0 && !dead) {
        const uint32_t want = epoch0 + (uint32_t)t;
        uint32_t spins = 0;
        while ((int32_t)(ld_acquire_u32(flags + watch) - want) < 0) {
          if (++spins > (1u << 22)) { dead = true; atomicExch(error, 1u); break; }
        }
      }
      bar_named<BAR_X, NX>();                             // the neighbours' rims of step t are in L2
#pragma unroll
      for (int q = 0; q < C::IPT_HALO; ++q) {
        if (ho[q] < 0) continue;
        cp_async16(st + PLANE + ho[q], P.d[side] + hg[q]);
        if (hring[q]) {
#pragma unroll
          for (int k = 0; k < 4; ++k) cp_async16(st + (2 + k) * PLANE + ho[q], P.F[side][k] + hg[q]);
        }
      }
      cp_async_wait_all();
      bar_named<BAR_B, NALL>();                           // the halo of step t + 1 is in shared memory
    }
    return;
  }

  // ---- the compute warps: this thread's items (the same every step) -----------------------------------------------------
  int it_in[C::IPT_IN], it_early[C::IPT_EARLY], it_late[C::IPT_LATE], it_rim[C::IPT_RIM];
#pragma unroll
  for (int q = 0; q < C::IPT_IN; ++q) {
    const int i = tid + q * NT;
    int r, gc; pos_interior<C>(i < C::N_IN ? i : 0, r, gc);
    it_in[q] = i < C::N_IN ? describe<C>(r, gc, gx_base, gy_base, g, c) : 0;
  }
#pragma unroll
  for (int q = 0; q < C::IPT_EARLY; ++q) {
    const int i = tid + q * NT;
    int r, gc; pos_early<C>(i < C::N_EARLY ? i : 0, r, gc);
    it_early[q] = i < C::N_EARLY ? describe<C>(r, gc, gx_base, gy_base, g, c) : 0;
  }
#pragma unroll
  for (int q = 0; q < C::IPT_LATE; ++q) {
    const int i = tid + q * NT;
    int r = 0, gc = 0;
    if (i < C::N_OUTER) pos_outer<C>(i, r, gc); else if (i < C::N_LATE) pos_ring<C>(i - C::N_OUTER, r, gc);
    it_late[q] = i < C::N_LATE ? describe<C>(r, gc, gx_base, gy_base, g, c) : 0;
  }
#pragma unroll
  for (int q = 0; q < C::IPT_RIM; ++q) {
    const int i = tid + q * NT;
    int r = 0, gc = 0;
    if (i < C::N_RIM) pos_rim<C>(i, r, gc);
    it_rim[q] = i < C::N_RIM ? describe<C>(r, gc, gx_base, gy_base, g, c) : 0;
  }

  double out_acc = 0.0, src_acc = 0.0;
  // ---- step 1: outflow of everything ------------------------------------------------------------------------------------
#pragma unroll
  for (int q = 0; q < C::IPT_EARLY; ++q) res_flux_item<C>(st, it_early[q], gx_base, gy_base, g, c, out_acc);
#pragma unroll
  for (int q = 0; q < C::IPT_LATE; ++q) res_flux_item<C>(st, it_late[q], gx_base, gy_base, g, c, out_acc);
  bar_named<BAR_C, NT>();

#pragma unroll 1
  for (int t = 1; t <= n; ++t) {
    const int side = (src + t) & 1;                       // where the state after step t lives
    if (t == n) {                                         // the frame's last step: everything goes to HBM
#pragma unroll
      for (int q = 0; q < C::IPT_RIM; ++q) res_depth_item<C, false, true>(st, it_rim[q], gx_base, gy_base, g, c, P, side, src_acc);
#pragma unroll
      for (int q = 0; q < C::IPT_IN; ++q) res_depth_item<C, false, true>(st, it_in[q], gx_base, gy_base, g, c, P, side, src_acc);
      break;
    }
    // rim first: the neighbours are waiting for it
#pragma unroll
    for (int q = 0; q < C::IPT_RIM; ++q) res_depth_item<C, true, false>(st, it_rim[q], gx_base, gy_base, g, c, P, side, src_acc);
    bar_named<BAR_A, NALL>();
#pragma unroll
    for (int q = 0; q < C::IPT_IN; ++q) res_depth_item<C, false, false>(st, it_in[q], gx_base, gy_base, g, c, P, side, src_acc);
    bar_named<BAR_C, NT>();                               // the block's depth of step t is complete
    // step t + 1: the outflow of all but the outermost own items needs nothing from outside the block
#pragma unroll
    for (int q = 0; q < C::IPT_EARLY; ++q) res_flux_item<C>(st, it_early[q], gx_base, gy_base, g, c, out_acc);
    bar_named<BAR_B, NALL>();
#pragma unroll
    for (int q = 0; q < C::IPT_LATE; ++q) res_flux_item<C>(st, it_late[q], gx_base, gy_base, g, c, out_acc);
    bar_named<BAR_C, NT>();
  }
  if (c.ledger_src != nullptr) ledger_src_flush(c.ledger_src, src_acc);
  ledger_src_flush(c.ledger, out_acc);
}

// Block shapes: the widest one covers 1024^2 with one block per SM of a B200 (4 x 37 = 148); the smaller ones give small
// grids more blocks (latency, not throughput, is what a 256^2 frame is bound by).
// (threads: NT compute + NX exchange; 640 threads leave 96 registers per thread)
using ResA = ResCfg<256, 28, 544, 96>;
using ResB = ResCfg<128, 28, 512, 128>;
using ResC = ResCfg<128, 12, 512, 128>;
using ResD = ResCfg<64, 12, 256, 64>;

int res_sm_count() {
  static int cached[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!cached[dev & 63]) cudaDeviceGetAttribute(&cached[dev & 63], cudaDevAttrMultiProcessorCount, dev);
  return cached[dev & 63] > 0 ? cached[dev & 63] : 148;
}

template <class C>
bool res_fits(const Geom& g, int sms, long long* staged) {
  const long long nbx = (g.W + C::OX - 1) / C::OX, nby = (g.Hg + C::OY - 1) / C::OY;
  *staged = (long long)C::PLANE;
  return nbx * nby <= sms;
}

template <class C>
cudaError_t res_launch(const Geom& g, const Planes& p, const StepConsts& c, int src, int n, uint32_t* flags, uint32_t epoch0,
                       uint32_t* error, cudaStream_t st) {
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  auto kern = resident_step_kernel<C>;
  if (!attr_set[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    if (e != cudaSuccess) return e;
    attr_set[dev & 63] = true;
  }
  const size_t row0_off = (size_t)TWS_HALO_ROWS * g.pitch;
  ResPlanes P;
  P.h = p.h + row0_off;
  for (int s = 0; s < 2; ++s) {
    P.d[s] = p.d[s] + row0_off;
    for (int i = 0; i < 4; ++i) P.F[s][i] = p.F[s][i] + row0_off;
  }
  P.v = p.v + row0_off;
  int nbx = (g.W + C::OX - 1) / C::OX, nby = (g.Hg + C::OY - 1) / C::OY;
  Geom gg = g;
  StepConsts cc = c;
  void* args[] = {&P, &gg, &cc, &nbx, &nby, &src, &n, &flags, &epoch0, &error};
  // cooperative: every CTA of the grid is resident before any of them runs, which the neighbour waits rely on
  return cudaLaunchCooperativeKernel((void*)kern, dim3((unsigned)(nbx * nby)), dim3(C::NT + C::NX), args, C::SMEM, st);
}

}  // namespace

// Which block shape runs this grid: the one with the least work per CTA among those whose grid of blocks fits the SMs
// (-1: the grid does not fit on chip, or it is a strip).
int resident_config(const Geom& g) {
  if (g.has_up || g.has_down || g.rows != g.Hg) return -1;
  const int sms = res_sm_count();
  long long best = -1, s = 0;
  int cfg = -1;
  if (res_fits<ResA>(g, sms, &s) && (best < 0 || s < best)) { best = s; cfg = 0; }
  if (res_fits<ResB>(g, sms, &s) && (best < 0 || s < best)) { best = s; cfg = 1; }
  if (res_fits<ResC>(g, sms, &s) && (best < 0 || s < best)) { best = s; cfg = 2; }
  if (res_fits<ResD>(g, sms, &s) && (best < 0 || s < best)) { best = s; cfg = 3; }
  return cfg;
}

int resident_blocks(const Geom& g, int cfg) {
  auto nb = [&](int ox, int oy) { return ((g.W + ox - 1) / ox) * ((g.Hg + oy - 1) / oy); };
  switch (cfg) {
    case 0: return nb(ResA::OX, ResA::OY);
    case 1: return nb(ResB::OX, ResB::OY);
    case 2: return nb(ResC::OX, ResC::OY);
    case 3: return nb(ResD::OX, ResD::OY);
    default: return 0;
  }
}

cudaError_t launch_resident(int cfg, const Geom& g, const Planes& p, const StepConsts& c, int src, int n, uint32_t* flags,
                            uint32_t epoch0, uint32_t* error, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  switch (cfg) {
    case 0: return res_launch<ResA>(g, p, c, src, n, flags, epoch0, error, st);
    case 1: return res_launch<ResB>(g, p, c, src, n, flags, epoch0, error, st);
    case 2: return res_launch<ResC>(g, p, c, src, n, flags, epoch0, error, st);
    case 3: return res_launch<ResD>(g, p, c, src, n, flags, epoch0, error, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace tws
