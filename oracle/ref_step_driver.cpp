// ref_step_driver.cpp — runs the reference's OWN simulation shaders on the CPU.
//
// The three .inc files included below are produced at build time by oracle/ref_shim/glsl_prep.py
// from /root/reference/terrainwatersim/shader/{flowUpdate,flowApply,waterBrush}.comp (and what they
// #include) — the shader statements are compiled exactly as written there, against the GLSL
// language shim oracle/ref_shim/glsl.h.  This file is the host side: it binds images and uniforms
// and dispatches work groups the way Terrain.cpp:150-168 (brush) and Terrain.cpp:250-265 (step
// loop) do, and exposes that over a C ABI so tests can check oracle/tws_oracle.cpp — the hand
// transcription every CUDA parity test compares against — bit for bit against the reference's
// own source.  TEST INFRASTRUCTURE; only built when /root/reference is present, output goes to
// oracle/_ref/libtws_ref_step.so (git-ignored, travels to the GPU box prebuilt).
//
// The reference dispatches res/16 (step) and res/32 (brush) groups with integer division
// (Terrain.cpp:167,258,264): texels beyond the last whole group are not processed.  This driver
// does the same, so it is only a statement about the reference on grids where the reference is
// defined (multiples of 16 / 32); ragged sizes stay oracle-only.
#include "glsl.h"

#include "flowUpdate.inc"
#include "flowApply.inc"
#include "waterBrush.inc"

#if defined(_OPENMP)
#include <omp.h>
#endif

namespace {

using glsl::FORMAT_RG16F;
using glsl::FORMAT_RGBA32F;
using glsl::image2D;

image2D bind_image(const void* texels, int w, int h, glsl::image_format f) {
  image2D img;
  img.texels = const_cast<void*>(texels);
  img.width = w; img.height = h; img.format = f;
  return img;
}

template <class S> void set_simulation_parameters(S& s, const float consts[3]) {   // UBO binding 5, simulationCommon.glsl:1-13
  s.FlowFriction_perStep = consts[0];
  s.WaterAcceleration_perStep = consts[1];
  s.CellAreaInv_timeScaled = consts[2];
}

}  // namespace

extern "C" {

int tws_ref_step_threads(void) {
#if defined(_OPENMP)
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// Terrain.cpp:255-258
void tws_ref_flow_update(int res_x, int res_y, const float* terrain_rgba, float* flow_rgba, const float consts[3]) {
  static_assert(glsl::shaders::flowUpdate_comp::TerrainData == 0 && glsl::shaders::flowUpdate_comp::Flow == 1, "image units");
  static_assert(glsl::shaders::flowUpdate_comp::Flow_format == FORMAT_RGBA32F, "format");
  glsl::shaders::flowUpdate_comp s{};
  set_simulation_parameters(s, consts);
  s.image_unit[0] = bind_image(terrain_rgba, res_x, res_y, FORMAT_RGBA32F);   // m_terrainData->BindImage(0, READ, GL_RGBA32F)
  s.image_unit[1] = bind_image(flow_rgba, res_x, res_y, FORMAT_RGBA32F);      // m_waterOutgoingFlow->BindImage(1, READ_WRITE, GL_RGBA32F)
  glsl::dispatch_compute(s, (unsigned)res_x / 16, (unsigned)res_y / 16);     // glDispatchCompute(res / 16, res / 16, 1)
}

// Terrain.cpp:260-264
void tws_ref_flow_apply(int res_x, int res_y, float* terrain_rgba, const float* flow_rgba, uint16_t* flowmap_rg16, const float consts[3]) {
  static_assert(glsl::shaders::flowApply_comp::TerrainData == 0 && glsl::shaders::flowApply_comp::OutgoingFlow == 1 &&
                glsl::shaders::flowApply_comp::FlowMap == 2, "image units");
  static_assert(glsl::shaders::flowApply_comp::FlowMap_format == FORMAT_RG16F, "format");
  glsl::shaders::flowApply_comp s{};
  set_simulation_parameters(s, consts);
  s.image_unit[0] = bind_image(terrain_rgba, res_x, res_y, FORMAT_RGBA32F);   // BindImage(0, READ_WRITE, GL_RGBA32F)
  s.image_unit[1] = bind_image(flow_rgba, res_x, res_y, FORMAT_RGBA32F);      // BindImage(1, READ, GL_RGBA32F)
  s.image_unit[2] = bind_image(flowmap_rg16, res_x, res_y, FORMAT_RG16F);     // m_waterFlowMap->BindImage(2, WRITE, GL_RG16F)
  glsl::dispatch_compute(s, (unsigned)res_x / 16, (unsigned)res_y / 16);
}

// the loop of Terrain.cpp:253-265, n times (update; apply)
void tws_ref_step(int res_x, int res_y, float* terrain_rgba, float* flow_rgba, uint16_t* flowmap_rg16, const float consts[3], int n) {
  for (int i = 0; i < n; ++i) {
    tws_ref_flow_update(res_x, res_y, terrain_rgba, flow_rgba, consts);
    tws_ref_flow_apply(res_x, res_y, terrain_rgba, flow_rgba, flowmap_rg16, consts);
  }
}

// Terrain.cpp:157-167: the three brush uniforms (binding 7), TerrainData on unit 0, res/32 groups.
void tws_ref_brush(int res_x, int res_y, float* terrain_rgba, float texel_x, float texel_y, float intensity, float size_sq) {
  static_assert(glsl::shaders::waterBrush_comp::TerrainData == 0, "image unit");
  glsl::shaders::waterBrush_comp s{};
  s.BrushPositionTexelCor = glsl::vec2(texel_x, texel_y);
  s.BrushIntensity = intensity;
  s.BrushSizeSq = size_sq;
  s.image_unit[0] = bind_image(terrain_rgba, res_x, res_y, FORMAT_RGBA32F);
  glsl::dispatch_compute(s, (unsigned)res_x / 32, (unsigned)res_y / 32);
}

}  // extern "C"
