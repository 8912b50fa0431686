// resident_kernels.cu — a whole frame of steps in ONE launch for grids that fit on chip (the reference's own
// operating point: 1024 x 1024, up to 10 steps per frame, Terrain.cpp:240-265).
//
// The six state planes of a 1024^2 grid are 25 MB; the 148 SMs of a B200 hold 33 MB of shared memory.  So the grid is cut
// into one OX x OY block per SM (1024^2: 4 x 37 blocks of 256 x 28 cells), every CTA keeps its block — terrain, depth and
// the four outflow planes, plus a halo of 2 rows / 4 columns — in shared memory for ALL n steps of the frame, and per step
// only the block's rim goes through L2, neighbour to neighbour; there is no grid-wide barrier and no flag: every value a
// CTA publishes travels in an 8-byte word together with the number of the step it belongs to (the scheme NCCL calls LL),
// so a reader polls the data itself and one L2 round trip is the whole hand-off.  (Round 2 first had the rims go through
// the state planes behind acquire / release flags: publish -> fence -> flag -> poll -> fetch took 4.6 us per step and set the
// step time on every grid size; %globaltimer traces in DESIGN.md section 4.)  The launch is cooperative: all CTAs are
// resident by construction, which the polling relies on.
//
// A step's passes are ordered so that the exchange overlaps the arithmetic that does not depend on it:
//     depth(rim), publishing it -> depth(interior) -> issue the loads of the neighbours' rims -> | -> flux(early) of the NEXT
//     step -> check the tags of what arrived (re-poll the words that are not there yet), halo -> shared memory -> |
//     -> flux(late) -> |                                                                         ('|' = block barrier)
// "early" are the items whose outflow needs no halo cell (all but the block's outermost row / group on each side), "late"
// the outermost own items and the halo ring, whose outflow is recomputed locally from its exchanged old outflow and depth
// (the same redundant-halo scheme as the tile kernel, one cell deep), so one exchange per step suffices.
//
// Arithmetic: the same cell functions as every other kernel (cell_math.cuh; flowUpdate.comp:34-59, flowApply.comp:32-46),
// bit-identical to the oracle.  Exterior cells are zeros in shared memory (the reference's out-of-range imageLoad) and are
// never updated.
#include "cell_math.cuh"

#include <cstdio>

namespace tws {

namespace {

template <int OX_, int OY_, int NT_>
struct ResCfg {
  static constexpr int OX = OX_, OY = OY_, NT = NT_;
  static constexpr int HX = 4, HY = 2;                       // halo: one float4 group / two rows on each side
  static constexpr int SX = OX + 2 * HX, SY = OY + 2 * HY;   // staged block
  static constexpr int NG = SX / 4;                          // float4 groups ("items") per staged row
  static constexpr int PLANE = SX * SY;                      // floats per plane
  static constexpr size_t SMEM = (size_t)6 * PLANE * sizeof(float);   // h, d, F x4
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
  static constexpr int N_RING = 2 * NG + 2 * (SY - 4);
  static constexpr int N_OUTER = 2 * RIM_W + 2 * (IN_ROWS + 2);   // rows 2 and SY-3, groups 1 and NG-2 of rows [3, SY-3)
  static constexpr int N_FAR = 2 * NG;
  // early items beyond a whole number of rounds of the NT threads join the late pass when it has idle threads for them
  static constexpr int N_EARLY_ALL = (IN_ROWS + 2) * IN_COLS;
  static constexpr int EARLY_REM = N_EARLY_ALL % NT;
  static constexpr bool MOVE_REM = N_EARLY_ALL > NT && EARLY_REM > 0 && (N_OUTER + N_RING) % NT != 0 &&
                                   (N_OUTER + N_RING) % NT + EARLY_REM <= NT;
  static constexpr int N_EARLY = MOVE_REM ? N_EARLY_ALL - EARLY_REM : N_EARLY_ALL;
  static constexpr int N_LATE = N_OUTER + N_RING + (N_EARLY_ALL - N_EARLY);
  // halo fetch: 16-byte words {value, tag, value, tag} — ring items have 5 planes x 2 words, far items (depth only) 2 words
  static constexpr int N_WORDS = N_RING * 10 + N_FAR * 2;
  static constexpr int IPT_IN = (N_IN + NT - 1) / NT, IPT_EARLY = (N_EARLY + NT - 1) / NT, IPT_LATE = (N_LATE + NT - 1) / NT,
                       IPT_RIM = (N_RIM + NT - 1) / NT, WPT = (N_WORDS + NT - 1) / NT;
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

// Mailbox: for every cell of the grid, plane p in {depth, F+X, F-X, F+Y, F-Y} and step parity q an 8-byte word
// {fp32 value, step tag}; word index ((q * 5 + p) * cells + gy * pitch + gx).  Only rim cells are ever written.
struct ResMail {
  unsigned long long* base;
  unsigned long long plane_words;    // pitch * Hg
};
__device__ __forceinline__ unsigned long long mail_word(float v, uint32_t tag) {
  return ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
}
// two words per access: each 8-byte element is single-copy atomic, which is all the protocol needs
__device__ __forceinline__ void mail_store2(unsigned long long* p, unsigned long long a, unsigned long long b) {
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void mail_load2(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void mail_publish4(const ResMail& M, int q, int p, size_t go, const float4& v, uint32_t tag) {
  unsigned long long* w = M.base + (size_t)(q * 5 + p) * M.plane_words + go;
  mail_store2(w, mail_word(v.x, tag), mail_word(v.y, tag));
  mail_store2(w + 2, mail_word(v.z, tag), mail_word(v.w, tag));
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
// PUBLISH: the item is part of the rim — its new depth (and, outermost ring, its outflow) also goes into the mailbox for the
// neighbours, tagged with the step.  LAST: the frame's final step — depth, outflow and the flow vector go to the planes.
template <class C, bool PUBLISH, bool LAST>
__device__ __forceinline__ void res_depth_item(float* __restrict__ st, const int desc, const int gx_base, const int gy_base,
                                               const Geom& g, const StepConsts& c, const ResPlanes& P, const int side, const ResMail& M,
                                               const uint32_t tag, double& src_acc) {
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
  const size_t go = (size_t)gy * (size_t)g.pitch + (size_t)gx0;
  if (PUBLISH) {
    const int q = (int)(tag & 1u);
    mail_publish4(M, q, 0, go, nd, tag);
    if (desc & kFlagOuter) {
      mail_publish4(M, q, 1, go, fx, tag); mail_publish4(M, q, 2, go, fy, tag);
      mail_publish4(M, q, 3, go, fz, tag); mail_publish4(M, q, 4, go, fw, tag);
    }
  }
  if (LAST) {
    stcg4(P.d[side] + go, nd);
    stcg4(P.F[side][0] + go, fx); stcg4(P.F[side][1] + go, fy); stcg4(P.F[side][2] + go, fz); stcg4(P.F[side][3] + go, fw);
    *reinterpret_cast<uint4*>(P.v + go) = nv;
  }
}

// n steps of the whole grid; state side `src` -> side (src + n) & 1.  Steps are numbered epoch0 + 1 ... epoch0 + n (the tags).
template <class C>
__global__ void __launch_bounds__(C::NT, 1) resident_step_kernel(ResPlanes P, ResMail M, Geom g, StepConsts c, int nbx, int src, int n,
                                                                 uint32_t epoch0, uint32_t* error, BrushArgs br) {
  constexpr int SY = C::SY, NG = C::NG, PLANE = C::PLANE, NT = C::NT, WPT = C::WPT;
  extern __shared__ __align__(1024) float st[];
  const int tid = threadIdx.x;
  const int by = (int)blockIdx.x / nbx, bx = (int)blockIdx.x - by * nbx;
  const int gx_base = bx * C::OX - C::HX, gy_base = by * C::OY - C::HY;

  // ---- load the staged block (own cells + halo) from side src; exterior cells are zeros --------------------------------
#pragma unroll 2
  for (int a = tid; a < NG * SY; a += NT) {
    const int r = a / NG, gc = a - r * NG;
    const int gy = gy_base + r, gx0 = gx_base + 4 * gc;
    const bool in = (unsigned)gy < (unsigned)g.Hg && (unsigned)gx0 < (unsigned)g.pitch;
    const size_t go = in ? (size_t)gy * (size_t)g.pitch + (size_t)gx0 : 0;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const int o = a * 4;
    st4(st + o, in ? ldcg4(P.h + go) : z);
    float4 dd = in ? ldcg4(P.d[src] + go) : z;
    if (br.active && in) dd = brush4(dd, gx0, gy, br);     // a pending brush (waterBrush.comp), folded into the block load
    st4(st + PLANE + o, dd);
#pragma unroll
    for (int k = 0; k < 4; ++k) st4(st + (2 + k) * PLANE + o, in ? ldcg4(P.F[src][k] + go) : z);
  }

  // ---- this thread's items (the same every step) ------------------------------------------------------------------------
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
    int i = tid + q * NT;
    int r = 0, gc = 0;
    const bool any = i < C::N_LATE;
    if (i < C::N_OUTER) pos_outer<C>(i, r, gc);
    else if (i < C::N_OUTER + C::N_RING) pos_ring<C>(i - C::N_OUTER, r, gc);
    else if (any) pos_early<C>(C::N_EARLY + (i - C::N_OUTER - C::N_RING), r, gc);
    it_late[q] = any ? describe<C>(r, gc, gx_base, gy_base, g, c) : 0;
  }
#pragma unroll
  for (int q = 0; q < C::IPT_RIM; ++q) {
    const int i = tid + q * NT;
    int r = 0, gc = 0;
    if (i < C::N_RIM) pos_rim<C>(i, r, gc);
    it_rim[q] = i < C::N_RIM ? describe<C>(r, gc, gx_base, gy_base, g, c) : 0;
  }
  // ... and the mailbox words it fetches: word offset within a parity's planes, shared-memory destination (< 0: none / exterior)
  unsigned long long mw[WPT]; int mo[WPT];
#pragma unroll
  for (int q = 0; q < WPT; ++q) {
    const int u = tid + q * NT;
    int r = 0, gc = 0, p = 0, half = 0;
    if (u < C::N_RING * 10) { const int it = u / 10, rem = u - it * 10; p = rem >> 1; half = rem & 1; pos_ring<C>(it, r, gc); }
    else if (u < C::N_WORDS) { const int v = u - C::N_RING * 10; half = v & 1; pos_far<C>(v >> 1, r, gc); }
    const int gy = gy_base + r, gx0 = gx_base + 4 * gc;
    const bool in = u < C::N_WORDS && (unsigned)gy < (unsigned)g.Hg && gx0 >= 0 && gx0 < g.W;   // published by a neighbour (cf. describe)
    mo[q] = in ? (1 + p) * PLANE + (r * NG + gc) * 4 + half * 2 : -1;
    mw[q] = in ? (unsigned long long)p * M.plane_words + (unsigned long long)gy * (unsigned)g.pitch + (unsigned)(gx0 + half * 2) : 0ull;
  }
  __syncthreads();

  double out_acc = 0.0, src_acc = 0.0;
  // ---- step 1: outflow of everything ------------------------------------------------------------------------------------
#pragma unroll
  for (int q = 0; q < C::IPT_EARLY; ++q) res_flux_item<C>(st, it_early[q], gx_base, gy_base, g, c, out_acc);
#pragma unroll
  for (int q = 0; q < C::IPT_LATE; ++q) res_flux_item<C>(st, it_late[q], gx_base, gy_base, g, c, out_acc);
  __syncthreads();

  bool dead = false;                                      // a poll timed out: stop waiting, the host reports the error
#pragma unroll 1
  for (int t = 1; t <= n; ++t) {
    const int side = (src + t) & 1;                       // where the state after step t lives
    const uint32_t tag = epoch0 + (uint32_t)t;
    if (t == n) {                                         // the frame's last step: everything goes to HBM
#pragma unroll
      for (int q = 0; q < C::IPT_RIM; ++q) res_depth_item<C, false, true>(st, it_rim[q], gx_base, gy_base, g, c, P, side, M, tag, src_acc);
#pragma unroll
      for (int q = 0; q < C::IPT_IN; ++q) res_depth_item<C, false, true>(st, it_in[q], gx_base, gy_base, g, c, P, side, M, tag, src_acc);
      break;
    }
    // rim first: the neighbours are waiting for it
#pragma unroll
    for (int q = 0; q < C::IPT_RIM; ++q) res_depth_item<C, true, false>(st, it_rim[q], gx_base, gy_base, g, c, P, side, M, tag, src_acc);
#pragma unroll
    for (int q = 0; q < C::IPT_IN; ++q) res_depth_item<C, false, false>(st, it_in[q], gx_base, gy_base, g, c, P, side, M, tag, src_acc);
    // the neighbours' rims of step t: the loads fly while the early outflow of step t + 1 is computed
    const unsigned long long* mbase = M.base + (size_t)(tag & 1u) * 5u * M.plane_words;
    unsigned long long wa[WPT], wb[WPT];
#pragma unroll
    for (int q = 0; q < WPT; ++q) { wa[q] = 0ull; wb[q] = 0ull; if (mo[q] >= 0) mail_load2(mbase + mw[q], wa[q], wb[q]); }
    __syncthreads();                                      // the block's depth of step t is complete
#pragma unroll
    for (int q = 0; q < C::IPT_EARLY; ++q) res_flux_item<C>(st, it_early[q], gx_base, gy_base, g, c, out_acc);
    // words whose tag is not the step's yet are re-read — all of a thread's stale words per round trip, not one after the other
    uint32_t spins = 0;
    for (bool stale = true; stale && !dead;) {
      stale = false;
#pragma unroll
      for (int q = 0; q < WPT; ++q)
        if (mo[q] >= 0 && ((uint32_t)(wa[q] >> 32) != tag || (uint32_t)(wb[q] >> 32) != tag)) {
          mail_load2(mbase + mw[q], wa[q], wb[q]);
          stale = true;
        }
      if (stale && ++spins > (1u << 21)) { dead = true; atomicExch(error, 1u); }
    }
#pragma unroll
    for (int q = 0; q < WPT; ++q)
      if (mo[q] >= 0)
        *reinterpret_cast<float2*>(st + mo[q]) = make_float2(__uint_as_float((uint32_t)wa[q]), __uint_as_float((uint32_t)wb[q]));
    __syncthreads();                                      // the halo of step t + 1 is in shared memory
#pragma unroll
    for (int q = 0; q < C::IPT_LATE; ++q) res_flux_item<C>(st, it_late[q], gx_base, gy_base, g, c, out_acc);
    __syncthreads();
  }
  if (c.ledger_src != nullptr) ledger_src_flush(c.ledger_src, src_acc);
  ledger_src_flush(c.ledger, out_acc);
}

// Block shapes: the widest one covers 1024^2 with one block per SM of a B200 (4 x 37 = 148); the smaller ones give small
// grids more blocks (latency, not throughput, is what a 256^2 frame is bound by).
using ResA = ResCfg<256, 28, 512>;     // 1024^2: 4 x 37 = 148 blocks
using ResB = ResCfg<128, 28, 512>;
using ResC = ResCfg<128, 14, 512>;     //  512^2: 4 x 37
using ResD = ResCfg<128, 12, 512>;
using ResE = ResCfg<64, 12, 256>;
using ResF = ResCfg<64, 8, 256>;       //  256^2: 4 x 32
#define TWS_RES_CONFIGS(X) X(0, ResA) X(1, ResB) X(2, ResC) X(3, ResD) X(4, ResE) X(5, ResF)

int res_sm_count() {
  static int cached[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!cached[dev & 63]) cudaDeviceGetAttribute(&cached[dev & 63], cudaDevAttrMultiProcessorCount, dev);
  return cached[dev & 63] > 0 ? cached[dev & 63] : 148;
}

// Can this device hold every block of the grid at once?  (One CTA per SM: the shared memory opt-in must succeed and the
// occupancy calculator must grant a block per SM — a MIG slice or a smaller part says no here, not at launch time.)
template <class C>
bool res_fits(const Geom& g, int sms, long long* staged) {
  const long long nbx = (g.W + C::OX - 1) / C::OX, nby = (g.Hg + C::OY - 1) / C::OY;
  *staged = (long long)C::PLANE;
  if (nbx * nby > sms) return false;
  int dev = 0, coop = 0, per_sm = 0;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess || !coop) return false;
  auto kern = resident_step_kernel<C>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM) != cudaSuccess) { cudaGetLastError(); return false; }
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, C::NT, C::SMEM) != cudaSuccess) { cudaGetLastError(); return false; }
  return (long long)per_sm * sms >= nbx * nby;
}

template <class C>
cudaError_t res_launch(const Geom& g, const Planes& p, const StepConsts& c, int src, int n, void* mailbox, uint32_t epoch0,
                       uint32_t* error, cudaStream_t st, const BrushArgs* brush) {
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
  ResMail M;
  M.base = (unsigned long long*)mailbox;
  M.plane_words = (unsigned long long)g.pitch * (unsigned long long)g.Hg;
  int nbx = (g.W + C::OX - 1) / C::OX, nby = (g.Hg + C::OY - 1) / C::OY;
  Geom gg = g;
  StepConsts cc = c;
  BrushArgs br{};
  if (brush != nullptr) br = *brush;
  void* args[] = {&P, &M, &gg, &cc, &nbx, &src, &n, &epoch0, &error, &br};
  // cooperative: every CTA of the grid is resident before any of them runs, which the polling of the neighbours relies on
  return cudaLaunchCooperativeKernel((void*)kern, dim3((unsigned)(nbx * nby)), dim3(C::NT), args, C::SMEM, st);
}

}  // namespace

// Which block shape runs this grid: the one with the least work per CTA among those whose grid of blocks fits the SMs
// (-1: the grid does not fit on chip, or it is a strip).
int resident_config(const Geom& g) {
  if (g.has_up || g.has_down || g.rows != g.Hg) return -1;
  const int sms = res_sm_count();
  long long best = -1, s = 0;
  int cfg = -1;
#define X(I, C) if (res_fits<C>(g, sms, &s) && (best < 0 || s < best)) { best = s; cfg = I; }
  TWS_RES_CONFIGS(X)
#undef X
  return cfg;
}

int resident_blocks(const Geom& g, int cfg) {
  auto nb = [&](int ox, int oy) { return ((g.W + ox - 1) / ox) * ((g.Hg + oy - 1) / oy); };
  switch (cfg) {
#define X(I, C) case I: return nb(C::OX, C::OY);
    TWS_RES_CONFIGS(X)
#undef X
    default: return 0;
  }
}

size_t resident_mailbox_bytes(const Geom& g) { return (size_t)10 * (size_t)g.pitch * (size_t)g.Hg * sizeof(unsigned long long); }

cudaError_t launch_resident(int cfg, const Geom& g, const Planes& p, const StepConsts& c, int src, int n, void* mailbox,
                            uint32_t epoch0, uint32_t* error, cudaStream_t st, const BrushArgs* brush) {
  if (n <= 0) return cudaSuccess;
  switch (cfg) {
#define X(I, C) case I: return res_launch<C>(g, p, c, src, n, mailbox, epoch0, error, st, brush);
    TWS_RES_CONFIGS(X)
#undef X
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace tws
