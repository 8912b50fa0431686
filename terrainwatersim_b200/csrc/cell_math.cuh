// cell_math.cuh — the per-cell arithmetic of the step (flowUpdate.comp / flowApply.comp) and the
// sm_100a TMA / mbarrier helpers shared by every step kernel (step_kernels.cu, stream_kernels.cu).
// Arithmetic contract: see step_kernels.cu header and DESIGN.md section 2.
#pragma once
#include "tws_internal.h"

#include <cuda_fp16.h>

namespace tws {

// ------------------------------------------------------------------------------------
// cell arithmetic shared by all kernels
// ------------------------------------------------------------------------------------
__device__ __forceinline__ float max0(float v) { return (v > 0.0f) ? v : 0.0f; }   // flowUpdate.comp:54

// flowUpdate.comp:44-57 for one cell.  f* in: old outflow, out: clamped-at-zero new outflow
// BEFORE the "cannot drain below zero" scaling; returns total = (sum f') * areaInv (:57).
__device__ __forceinline__ float flux_raw(float Hc, float Hxp, float Hxm, float Hyp, float Hym,
                                          float& fx, float& fy, float& fz, float& fw, const StepConsts& c) {
  float nx = Hc - Hxp, ny = Hc - Hxm, nz = Hc - Hyp, nw = Hc - Hym;                  // :44-47
  nx = __fadd_rn(__fmul_rn(fx, c.friction), __fmul_rn(nx, c.accel));                 // :53
  ny = __fadd_rn(__fmul_rn(fy, c.friction), __fmul_rn(ny, c.accel));
  nz = __fadd_rn(__fmul_rn(fz, c.friction), __fmul_rn(nz, c.accel));
  nw = __fadd_rn(__fmul_rn(fw, c.friction), __fmul_rn(nw, c.accel));
  fx = max0(nx); fy = max0(ny); fz = max0(nz); fw = max0(nw);                        // :54
  return __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(fx, fy), fz), fw), c.area_inv);     // :57
}

// flowUpdate.comp:58-59 for the four cells of one float4 group: if (total > a) f' *= a / total.
// The IEEE division is needed only for a wet cell that would drain completely this step —
// rare — so the common path is branch free: scale = 1 when total <= a (x*1 == x bit for bit,
// the shader does not multiply at all there), scale = 0 when a == 0 (0/total == +0 for every
// total > 0), and ONE warp-level branch covers the cells that really divide.
__device__ __forceinline__ void flux_scale4(const float (&total)[4], const float (&depth)[4], float (&s)[4]) {
  bool need = false;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool over = total[i] > depth[i];
    const bool dry = depth[i] == 0.0f;
    s[i] = over ? 0.0f : 1.0f;
    need = need || (over && !dry);
  }
  if (need) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (total[i] > depth[i] && depth[i] != 0.0f) s[i] = __fdiv_rn(depth[i], total[i]);
  }
}

// Scalar form used by the unfused baseline.
__device__ __forceinline__ void flux_cell(float Hc, float Hxp, float Hxm, float Hyp, float Hym, float depth,
                                          float& fx, float& fy, float& fz, float& fw, const StepConsts& c) {
  const float total = flux_raw(Hc, Hxp, Hxm, Hyp, Hym, fx, fy, fz, fw, c);
  if (total > depth) {                                                               // :58
    const float s = (depth == 0.0f) ? 0.0f : __fdiv_rn(depth, total);                // :59
    fx = __fmul_rn(fx, s); fy = __fmul_rn(fy, s); fz = __fmul_rn(fz, s); fw = __fmul_rn(fw, s);
  }
}

// flowApply.comp:38-46.  Returns the new depth; vx, vy the flow vector; ds (EXT ledger) what rain / evaporation really
// changed in fp32: depth after the source terms minus depth before them (exact: the two are within a factor of two of each
// other, or the difference is below half an ulp of either and irrelevant), 0 without sources.
__device__ __forceinline__ float apply_cell_src(float depth, float fx, float fy, float fz, float fw, float iX1, float iX0,
                                                float iY1, float iY0, const StepConsts& c, const bool ext, float& vx, float& vy, float& ds) {
  const float in = __fadd_rn(__fadd_rn(__fadd_rn(iX1, iX0), iY1), iY0);              // :38
  const float out = __fadd_rn(__fadd_rn(__fadd_rn(fx, fy), fz), fw);                 // :39
  float nd = max0(__fadd_rn(depth, __fmul_rn(__fsub_rn(in, out), c.area_inv)));      // :41
  ds = 0.0f;
  if (ext) {                                                                         // EXT
    const float pre = nd;
    nd = max0(__fsub_rn(__fadd_rn(nd, c.rain_step), c.evap_step));
    ds = __fsub_rn(nd, pre);
  }
  vx = __fsub_rn(__fsub_rn(iX1, fx), __fsub_rn(iX0, fy));                            // :45
  vy = __fsub_rn(__fsub_rn(iY1, fz), __fsub_rn(iY0, fw));                            // :46
  return nd;
}
__device__ __forceinline__ float apply_cell(float depth, float fx, float fy, float fz, float fw, float iX1, float iX0,
                                            float iY1, float iY0, const StepConsts& c, float& vx, float& vy) {
  float ds;
  return apply_cell_src(depth, fx, fy, fz, fw, iX1, iX0, iY1, iY0, c, c.ext_sources != 0, vx, vy, ds);
}

// EXT ledger of the sources (Control::source_acc): per-thread fp64 partial sums of the source deltas of OWNED cells,
// flushed by whole warps (shuffle reduction, one atomicAdd per warp).
__device__ __forceinline__ void ledger_src_flush(double* dst, double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (dst != nullptr && (threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(dst, v);
}

// EXT mass ledger (Control::outflow_acc).  Called by the ONE lane that owns the four cells (gx0..gx0+3, gy) as output
// cells of its piece / tile, with their final outflow of a sub-step: adds what points out of the global grid — +X at
// x = W-1, -X at x = 0, +Y on the last global row, -Y on row 0 — times areaInv.  Interior lanes leave after four compares.
__device__ __forceinline__ void ledger_add(const StepConsts& c, const Geom& g, int gx0, int gy, const float4& fx, const float4& fy,
                                           const float4& fz, const float4& fw) {
  if (c.ledger == nullptr || (unsigned)gy >= (unsigned)g.Hg) return;
  const bool ytop = gy == 0, ybot = gy == g.Hg - 1;
  if (!(ytop || ybot || gx0 == 0 || gx0 + 3 >= g.W - 1)) return;
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int x = gx0 + i;
    if ((unsigned)x >= (unsigned)g.W) continue;
    const float px = i == 0 ? fx.x : (i == 1 ? fx.y : (i == 2 ? fx.z : fx.w));
    const float py = i == 0 ? fy.x : (i == 1 ? fy.y : (i == 2 ? fy.z : fy.w));
    const float pz = i == 0 ? fz.x : (i == 1 ? fz.y : (i == 2 ? fz.z : fz.w));
    const float pw = i == 0 ? fw.x : (i == 1 ? fw.y : (i == 2 ? fw.z : fw.w));
    if (x == g.W - 1) s += (double)px;
    if (x == 0) s += (double)py;
    if (ybot) s += (double)pz;
    if (ytop) s += (double)pw;
  }
  if (s != 0.0) atomicAdd(c.ledger, s * (double)c.area_inv);
}
// The same into a per-thread partial sum (flushed with ledger_src_flush at the end of the kernel): for the ring kernel, where a
// row of an edge column strip is the critical path of its ring turn — an atomic per row of the two edge strips showed up as
// 238 -> 185 Gcell/s at K = 3 (bisected, round 2).
__device__ __forceinline__ void ledger_acc(double& acc, const StepConsts& c, const Geom& g, int gx0, int gy, const float4& fx, const float4& fy,
                                           const float4& fz, const float4& fw) {
  const bool ytop = gy == 0, ybot = gy == g.Hg - 1;
  if ((unsigned)gy >= (unsigned)g.Hg || !(ytop || ybot || gx0 == 0 || gx0 + 3 >= g.W - 1)) return;
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int x = gx0 + i;
    if ((unsigned)x >= (unsigned)g.W) continue;
    const float px = i == 0 ? fx.x : (i == 1 ? fx.y : (i == 2 ? fx.z : fx.w));
    const float py = i == 0 ? fy.x : (i == 1 ? fy.y : (i == 2 ? fy.z : fy.w));
    const float pz = i == 0 ? fz.x : (i == 1 ? fz.y : (i == 2 ? fz.z : fz.w));
    const float pw = i == 0 ? fw.x : (i == 1 ? fw.y : (i == 2 ? fw.z : fw.w));
    if (x == g.W - 1) s += (double)px;
    if (x == 0) s += (double)py;
    if (ybot) s += (double)pz;
    if (ytop) s += (double)pw;
  }
  acc += s * (double)c.area_inv;
}

// waterBrush.comp:26-31 for the four cells of a float4 group at global (gx0, gy): the same expression as brush_kernel
// (aux_kernels.cu), applied to the cells inside the brush's bounding box (outside it the brush adds exactly 0).
__device__ __forceinline__ float4 brush4(float4 d, const int gx0, const int gy, const BrushArgs& br) {
  if (gy < br.y0 || gy > br.y1 || gx0 > br.x1 || gx0 + 3 < br.x0) return d;
  float* pd = &d.x;
  const float ty = br.cy - (float)gy;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int x = gx0 + i;
    if (x < br.x0 || x > br.x1) continue;
    const float tx = br.cx - (float)x;                                                           // :26
    const float dist = __fdiv_rn(__fadd_rn(__fmul_rn(tx, tx), __fmul_rn(ty, ty)), br.size_sq);   // :27
    float s = 1.0f - dist;                                                                       // :28
    s = (s < 0.0f) ? 0.0f : ((s > 1.0f) ? 1.0f : s);
    pd[i] = __fadd_rn(pd[i], __fmul_rn(s, br.intensity));
  }
  return d;
}

__device__ __forceinline__ uint32_t pack_half2(float x, float y) {                  // rg16f store, :52
  const __half2 h = __floats2half2_rn(x, y);
  return *reinterpret_cast<const uint32_t*>(&h);
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float comp(const float4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

// ------------------------------------------------------------------------------------
// packed fp32x2 arithmetic (sm_100a FADD2 / FMUL2): two IEEE fp32 operations per issue slot
// ------------------------------------------------------------------------------------
// The step is ~50 dependent-free fp32 adds/multiplies per cell and none of them may be an FMA, so
// the kernels are limited by instruction issue, not by the FP32 lanes.  sm_100 has packed
// add.rn.f32x2 / mul.rn.f32x2 (SASS FADD2 / FMUL2): each rounds both halves exactly like the scalar
// instruction, so pairing the cells (x, x+1) of a float4 group halves the issue slots of every add
// and multiply that has aligned operands, bit for bit identical to the scalar form.
//
// ptxas (12.9) CONTRACTS mul.rn.f32x2 feeding add/sub.rn.f32x2 into FFMA2 despite the .rn
// qualifiers and --fmad=false (a single-use packed product is fused).  Scalar add.rn.f32 /
// mul.rn.f32 are never contracted.  Rule used below: an add or subtract that consumes a PACKED
// product is always the scalar __fadd_rn; packed adds only consume loads, scalar products, maxima
// and other sums.  tests/test_abi.py checks that no FFMA2 appears in the library's SASS.
#ifndef TWS_PACKED
#define TWS_PACKED 1
#endif
typedef unsigned long long f2;       // {lo, hi} fp32 pair in one aligned 64-bit register pair

__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(f2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 lo2(const float4& v) { return pk(v.x, v.y); }
__device__ __forceinline__ f2 hi2(const float4& v) { return pk(v.z, v.w); }
__device__ __forceinline__ float4 cat4(f2 lo, f2 hi) { float4 r; upk(lo, r.x, r.y); upk(hi, r.z, r.w); return r; }
#if TWS_PACKED
__device__ __forceinline__ float4 add4(const float4& a, const float4& b) { return cat4(add2(lo2(a), lo2(b)), add2(hi2(a), hi2(b))); }
#else
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
#endif

// flowUpdate.comp:44-57 for the four cells of one float4 group, interior cells only (no exterior or
// boundary-mode handling).  HC: own water level; HU / HD: rows y-1 / y+1; HL / HR: the cells left of
// .x / right of .w.  f*: in old outflow, out max(0, .) of the new one (before the :58-59 scaling);
// total[i] = (sum f') * areaInv.  Same values as four flux_raw calls:
//  * the x differences are shared between neighbours, H[j]-H[j+1] serving cell j's +X and, negated,
//    cell j+1's -X: a-b == -(b-a) exactly, and where both are zero the sign of the zero is erased by
//    the max(0, .) that follows the sum (t + (+-0) differs only for t == +-0, and max0 maps both to +0);
//  * products are packed or scalar multiplies, each rounded once; the sum F*phi + g*alpha is a scalar
//    add (see the contraction rule above); the 4-term flux sums are packed adds in the shader's order.
__device__ __forceinline__ void flux_raw4(const float4& HC, const float4& HU, const float4& HD, float HL, float HR,
                                          float4& fx, float4& fy, float4& fz, float4& fw, const StepConsts& c, float (&total)[4]) {
  const f2 PHI = pk(c.friction, c.friction), ALPHA = pk(c.accel, c.accel), KAPPA = pk(c.area_inv, c.area_inv);
  // x: D[j] = H[j] - H[j+1], j = -1..3; P[j] = D[j] * accel
  const float Dm = __fsub_rn(HL, HC.x);
  const float D0 = __fsub_rn(HC.x, HC.y), D1 = __fsub_rn(HC.y, HC.z), D2 = __fsub_rn(HC.z, HC.w), D3 = __fsub_rn(HC.w, HR);
  const float Pm = __fmul_rn(Dm, c.accel);
  float P0, P1, P2, P3;
  upk(mul2(pk(D0, D1), ALPHA), P0, P1);
  upk(mul2(pk(D2, D3), ALPHA), P2, P3);
  // y: g = HC - HD (+Y, :46), HC - HU (-Y, :47)
  float Z0, Z1, Z2, Z3, W0, W1, W2, W3;
  upk(mul2(sub2(lo2(HC), lo2(HD)), ALPHA), Z0, Z1);
  upk(mul2(sub2(hi2(HC), hi2(HD)), ALPHA), Z2, Z3);
  upk(mul2(sub2(lo2(HC), lo2(HU)), ALPHA), W0, W1);
  upk(mul2(sub2(hi2(HC), hi2(HU)), ALPHA), W2, W3);
  // F * friction (:53)
  const float4 tx = cat4(mul2(lo2(fx), PHI), mul2(hi2(fx), PHI));
  const float4 ty = cat4(mul2(lo2(fy), PHI), mul2(hi2(fy), PHI));
  const float4 tz = cat4(mul2(lo2(fz), PHI), mul2(hi2(fz), PHI));
  const float4 tw = cat4(mul2(lo2(fw), PHI), mul2(hi2(fw), PHI));
  // sum of the two rounded products: scalar adds; max(0, .) (:54)
  fx.x = max0(__fadd_rn(tx.x, P0)); fx.y = max0(__fadd_rn(tx.y, P1)); fx.z = max0(__fadd_rn(tx.z, P2)); fx.w = max0(__fadd_rn(tx.w, P3));
  fy.x = max0(__fsub_rn(ty.x, Pm)); fy.y = max0(__fsub_rn(ty.y, P0)); fy.z = max0(__fsub_rn(ty.z, P1)); fy.w = max0(__fsub_rn(ty.w, P2));
  fz.x = max0(__fadd_rn(tz.x, Z0)); fz.y = max0(__fadd_rn(tz.y, Z1)); fz.z = max0(__fadd_rn(tz.z, Z2)); fz.w = max0(__fadd_rn(tz.w, Z3));
  fw.x = max0(__fadd_rn(tw.x, W0)); fw.y = max0(__fadd_rn(tw.y, W1)); fw.z = max0(__fadd_rn(tw.z, W2)); fw.w = max0(__fadd_rn(tw.w, W3));
  // ((fx + fy) + fz) + fw, times areaInv (:57); the product only feeds the comparison of :58
  const f2 slo = add2(add2(add2(lo2(fx), lo2(fy)), lo2(fz)), lo2(fw));
  const f2 shi = add2(add2(add2(hi2(fx), hi2(fy)), hi2(fz)), hi2(fw));
  upk(mul2(slo, KAPPA), total[0], total[1]);
  upk(mul2(shi, KAPPA), total[2], total[3]);
}

// flowApply.comp:38-46 for the four interior cells of one float4 group.  l / r: +X outflow of the
// cell left of .x / -X outflow of the cell right of .w; iy1 / iy0: -Y outflow of row y+1 / +Y outflow
// of row y-1.  The flux operands are results of scalar multiplies (the :59 scaling) or loads, so the
// sums and differences are packed; d + (in-out)*areaInv consumes a packed product and is scalar.
template <bool VEL>
__device__ __forceinline__ void apply4(const float4& d, const float4& fx, const float4& fy, const float4& fz, const float4& fw, float l,
                                       float r, const float4& iy1, const float4& iy0, const StepConsts& c, const bool ext, float4& nd,
                                       uint4& nv, float4* ds = nullptr /* EXT ledger: source delta per cell */) {
  // iX1 = F(x+1,y).y (:32), iX0 = F(x-1,y).x (:33): neighbours inside the group are misaligned pairs -> scalar
  const float a0 = __fadd_rn(fy.y, l), a1 = __fadd_rn(fy.z, fx.x), a2 = __fadd_rn(fy.w, fx.y), a3 = __fadd_rn(r, fx.z);
  const f2 inlo = add2(add2(pk(a0, a1), lo2(iy1)), lo2(iy0));                        // :38
  const f2 inhi = add2(add2(pk(a2, a3), hi2(iy1)), hi2(iy0));
  const f2 outlo = add2(add2(add2(lo2(fx), lo2(fy)), lo2(fz)), lo2(fw));             // :39
  const f2 outhi = add2(add2(add2(hi2(fx), hi2(fy)), hi2(fz)), hi2(fw));
  const f2 KAPPA = pk(c.area_inv, c.area_inv);
  float q0, q1, q2, q3;
  upk(mul2(sub2(inlo, outlo), KAPPA), q0, q1);
  upk(mul2(sub2(inhi, outhi), KAPPA), q2, q3);
  nd.x = max0(__fadd_rn(d.x, q0)); nd.y = max0(__fadd_rn(d.y, q1)); nd.z = max0(__fadd_rn(d.z, q2)); nd.w = max0(__fadd_rn(d.w, q3));   // :41
  if (ext) {                                                                         // EXT (same expression as apply_cell)
    const f2 RAIN = pk(c.rain_step, c.rain_step), EVAP = pk(c.evap_step, c.evap_step);
    float e0, e1, e2, e3;
    upk(sub2(add2(lo2(nd), RAIN), EVAP), e0, e1);
    upk(sub2(add2(hi2(nd), RAIN), EVAP), e2, e3);
    const float4 pre = nd;
    nd = make_float4(max0(e0), max0(e1), max0(e2), max0(e3));
    if (ds != nullptr) *ds = make_float4(__fsub_rn(nd.x, pre.x), __fsub_rn(nd.y, pre.y), __fsub_rn(nd.z, pre.z), __fsub_rn(nd.w, pre.w));
  } else if (ds != nullptr) {
    *ds = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (VEL) {
    // vx = (iX1 - fx) - (iX0 - fy) (:45): inner differences on misaligned neighbours are scalar
    const float b0 = __fsub_rn(fy.y, fx.x), b1 = __fsub_rn(fy.z, fx.y), b2 = __fsub_rn(fy.w, fx.z), b3 = __fsub_rn(r, fx.w);
    const float c0 = __fsub_rn(l, fy.x), c1 = __fsub_rn(fx.x, fy.y), c2 = __fsub_rn(fx.y, fy.z), c3 = __fsub_rn(fx.z, fy.w);
    float vx0, vx1, vx2, vx3, vy0, vy1, vy2, vy3;
    upk(sub2(pk(b0, b1), pk(c0, c1)), vx0, vx1);
    upk(sub2(pk(b2, b3), pk(c2, c3)), vx2, vx3);
    // vy = (iY1 - fz) - (iY0 - fw) (:46)
    upk(sub2(sub2(lo2(iy1), lo2(fz)), sub2(lo2(iy0), lo2(fw))), vy0, vy1);
    upk(sub2(sub2(hi2(iy1), hi2(fz)), sub2(hi2(iy0), hi2(fw))), vy2, vy3);
    nv = make_uint4(pack_half2(vx0, vy0), pack_half2(vx1, vy1), pack_half2(vx2, vy2), pack_half2(vx3, vy3));
  }
}

// ------------------------------------------------------------------------------------
// TMA / mbarrier helpers
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

// Shared-memory float4 access through 32-bit shared-window addresses.  The exchange-slot addresses are
// kept as opaque 32-bit values: left as generic pointers derived from the warp index, the compiler
// re-derives them (shared-window base, slot selects, multiplies: ~35 integer instructions) in every
// half-pass rather than spend three registers on them.
__device__ __forceinline__ float4 lds4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts4(uint32_t a, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t keep_u32(uint32_t v) {       // pins a value in a register (no rematerialisation)
  asm volatile("" : "+r"(v));
  return v;
}

#ifndef TWS_STREAM_HX_MIN
#define TWS_STREAM_HX_MIN 8    // tuning: widen the x halo so that the output columns of a strip start on a 64 / 128 B boundary
#endif
constexpr int stream_hx(int K) { return ((2 * K + 3) / 4) * 4 > TWS_STREAM_HX_MIN ? ((2 * K + 3) / 4) * 4 : TWS_STREAM_HX_MIN; }

struct FusedOut {          // plane pointers at local row 0
  float* d; float* F[4]; uint32_t* v;
};

}  // namespace tws
