// The reference's own frame at its own size through the C++ mirror of its Terrain class (include/tws_terrain.hpp), timed from
// C++ so that no interpreter sits between the host loop and the C ABI:
//   ApplyRadialWaterBrush + PerformSimulationStep(1/60 s) (= one step at 60 steps/s) + GenMipMaps of TerrainInfo
//   (Scene.cpp:356-364, Terrain.cpp:240-277: what its on-screen "Simulation Time" covers),
// and a frame that owes the per-frame maximum of 10 steps (Terrain.cpp:247).  Wall clock per frame over many frames.
//   g++ -O2 -std=c++17 -Iinclude scripts/refframe.cpp -o build/refframe -Lterrainwatersim_b200 -ltws -Wl,-rpath,$PWD/terrainwatersim_b200
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "tws_terrain.hpp"

int main(int argc, char** argv) {
  const int frames = argc > 1 ? std::atoi(argv[1]) : 2000;
  try {
    tws_params p;
    tws_default_params(&p);                              // 1024 x 1024, 60 steps/s, TWS_BACKEND_AUTO: the reference's defaults
    tws::Terrain t(&p);
    t.CreateHeightmapFromNoiseAndResetSim();
    for (int pass = 0; pass < 2; ++pass) {
      const double dt = pass == 0 ? 1.0 / 60.0 + 1e-9 : 10.0 / 60.0 + 1e-9;
      uint64_t steps = 0;
      int32_t levels = 0;
      auto frame = [&]() {
        t.ApplyRadialWaterBrush(512.0f, 512.0f, 100.0f / 60.0f);
        steps += t.PerformSimulationStep(dt);
        t.PublishMips(&levels);
      };
      for (int i = 0; i < 100; ++i) frame();
      t.Sync();
      steps = 0;
      const auto t0 = std::chrono::steady_clock::now();
      for (int i = 0; i < frames; ++i) frame();
      t.Sync();
      const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / frames;
      std::printf("refframe_cxx steps_per_frame=%.2f mip_levels=%d us_per_frame=%.2f\n", (double)steps / frames, (int)levels, us);
    }
  } catch (const tws::Error& e) {
    std::printf("error %d: %s\n", (int)e.status, e.what());
    return 2;
  }
  return 0;
}
