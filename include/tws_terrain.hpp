// tws_terrain.hpp — header-only C++ host wrapper over the C ABI (tws.h) with the method
// names of the reference's `Terrain` class (terrainwatersim/source/scene/Terrain.h:18-39,
// 76-83), so a renderer written against the reference can switch its simulation calls to
// libtws.so by changing the type.  ezTime / ezVec2 are replaced by double seconds / two floats.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "tws.h"

namespace tws {

class Error : public std::runtime_error {
 public:
  Error(tws_status st, const std::string& msg) : std::runtime_error(msg), status(st) {}
  tws_status status;
};

class Terrain {
 public:
  // Terrain::Terrain (Terrain.cpp:21-121): defaults 1024 / 1024 / 60 / 0.98 / 10.
  explicit Terrain(const tws_params* params = nullptr) {
    tws_params p;
    if (params) p = *params; else tws_default_params(&p);
    const tws_status st = tws_create(&p, &sim_);
    if (st != TWS_OK) throw Error(st, tws_last_error(nullptr));
  }
  ~Terrain() { tws_destroy(sim_); }
  Terrain(const Terrain&) = delete;
  Terrain& operator=(const Terrain&) = delete;

  // Terrain.h:22 / Terrain.cpp:240-277.  Returns the number of steps run this frame.
  uint32_t PerformSimulationStep(double lastFrameDurationSeconds) {
    uint32_t n = 0;
    check(tws_advance(sim_, lastFrameDurationSeconds, &n));
    return n;
  }
  // Terrain.h:33 / Terrain.cpp:150-168.
  void ApplyRadialWaterBrush(float worldX, float worldZ, float strength) { check(tws_inject_brush_world(sim_, worldX, worldZ, strength)); }
  // Terrain.h:76-83 / Terrain.cpp:175-198.
  void SetSimulationStepsPerSecond(float v) { check(tws_set_steps_per_second(sim_, v)); }
  void SetFlowDamping(float v) { check(tws_set_flow_damping(sim_, v)); }
  void SetFlowAcceleration(float v) { check(tws_set_flow_acceleration(sim_, v)); }
  // Terrain.h:39 / Terrain.cpp:200-238 (+ Random::Init(seed), Application.cpp:57).
  void CreateHeightmapFromNoiseAndResetSim(uint32_t seed = 231656522u, float heightScale = 300.0f) {
    check(tws_reset_reference_scene(sim_, seed, heightScale, 2, 10, 0.43f));
  }

  void Step(int n) { check(tws_step(sim_, n)); }
  // one step with the water layer in host memory (band-pipelined upload / step / readback)
  void StepHost(const float* waterIn, float* waterOut, void* velocityOut) { check(tws_step_host(sim_, waterIn, waterOut, velocityOut)); }
  void Sync() { check(tws_sync(sim_)); }
  double TotalVolume() { double v = 0; check(tws_total_volume(sim_, &v)); return v; }
  float SimulationTimeMs() { float ms = 0; check(tws_elapsed_ms(sim_, &ms)); return ms; }   // "Simulation Time" stat, Scene.cpp:341-342
  // the reference's own way (gl::TimerQuery, read one frame late, never blocks): false while no timed batch has finished
  bool SimulationTimeMsNoWait(float* ms) { const tws_status st = tws_elapsed_ms_nowait(sim_, ms, nullptr); if (st == TWS_ERR_STATE) return false; check(st); return true; }
  // EXT mass ledger: volume that left the map through the open boundary (accumulated inside the step kernels)
  double BoundaryOutflowVolume() { double v = 0; check(tws_boundary_outflow_accumulated(sim_, &v)); return v; }
  void Upload(tws_field f, const void* host, size_t bytes) { check(tws_upload(sim_, f, host, bytes)); }
  void Readback(tws_field f, void* host, size_t bytes) { check(tws_readback(sim_, f, host, bytes)); }
  // Renderer hand-off (Terrain.cpp:272-276,288,323-330): the renderer keeps owning TerrainInfo (RGBA32F, full
  // mip chain) and FlowMap (RG16F); register their GL names once with the GL context current, publish per frame.
  void RegisterGLTextures(uint32_t terrainInfoTex, uint32_t flowMapTex) { check(tws_gl_register(sim_, terrainInfoTex, flowMapTex)); }
  void PublishToGL() { check(tws_gl_publish(sim_)); }
  // the same images in library-owned device memory: level L of TerrainInfo follows level L-1 (tws_mip_level_info)
  void* PublishMips(int32_t* levels = nullptr) { void* p = nullptr; check(tws_publish_mips(sim_, &p, levels)); return p; }
  tws_sim* handle() { return sim_; }

 private:
  void check(tws_status st) { if (st != TWS_OK) throw Error(st, tws_last_error(sim_)); }
  tws_sim* sim_ = nullptr;
};

}  // namespace tws
