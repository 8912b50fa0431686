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

// flowApply.comp:38-46.  Returns the new depth; vx, vy the flow vector.
__device__ __forceinline__ float apply_cell(float depth, float fx, float fy, float fz, float fw, float iX1, float iX0,
                                            float iY1, float iY0, const StepConsts& c, float& vx, float& vy) {
  const float in = __fadd_rn(__fadd_rn(__fadd_rn(iX1, iX0), iY1), iY0);              // :38
  const float out = __fadd_rn(__fadd_rn(__fadd_rn(fx, fy), fz), fw);                 // :39
  float nd = max0(__fadd_rn(depth, __fmul_rn(__fsub_rn(in, out), c.area_inv)));      // :41
  if (c.ext_sources) nd = max0(__fsub_rn(__fadd_rn(nd, c.rain_step), c.evap_step));  // EXT
  vx = __fsub_rn(__fsub_rn(iX1, fx), __fsub_rn(iX0, fy));                            // :45
  vy = __fsub_rn(__fsub_rn(iY1, fz), __fsub_rn(iY0, fw));                            // :46
  return nd;
}

__device__ __forceinline__ uint32_t pack_half2(float x, float y) {                  // rg16f store, :52
  const __half2 h = __floats2half2_rn(x, y);
  return *reinterpret_cast<const uint32_t*>(&h);
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float comp(const float4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

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

struct FusedOut {          // plane pointers at local row 0
  float* d; float* F[4]; uint32_t* v;
};

}  // namespace tws
