// Compiled by tests/test_abi.py: proves tws.h is valid C++ / tws_terrain.hpp compiles and links
// against libtws.so, prints the struct layout for comparison with the ctypes mirror and
// (argv[1] == "run", GPU box only) drives a tiny simulation through the C++ wrapper.
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <vector>
#include "tws_terrain.hpp"

int main(int argc, char** argv) {
  std::printf("sizeof_params=%zu sizeof_consts=%zu sizeof_handle=%zu off_backend=%zu off_rain=%zu abi=%d\n", sizeof(tws_params),
              sizeof(tws_step_constants), sizeof(tws_halo_handle), offsetof(tws_params, backend), offsetof(tws_params, rain_rate),
              tws_abi_version());
  if (argc > 1 && !std::strcmp(argv[1], "run")) {
    try {
      tws_params p;
      tws_default_params(&p);
      p.width = p.height = p.row_end = 128; p.world_size = 128.0f;
      tws::Terrain t(&p);
      t.CreateHeightmapFromNoiseAndResetSim();
      const double v0 = t.TotalVolume();
      t.ApplyRadialWaterBrush(64.0f, 64.0f, 1.0f);
      const uint32_t n = t.PerformSimulationStep(0.05);
      t.Sync();
      std::printf("steps=%u v0=%.6f v1=%.6f ms=%.4f\n", n, v0, t.TotalVolume(), t.SimulationTimeMs());
      float ms_nowait = -1.0f;
      const bool have = t.SimulationTimeMsNoWait(&ms_nowait);   // gl::TimerQuery semantics: last finished batch, no wait
      std::printf("nowait=%d ms=%.4f outflow=%.6f\n", (int)have, ms_nowait, t.BoundaryOutflowVolume());
      int32_t levels = 0;
      const void* chain = t.PublishMips(&levels);          // GenMipMaps of TerrainInfo, Terrain.cpp:272-276
      int32_t lw = 0, lh = 0; int64_t off = 0;
      tws_mip_level_info(128, 128, levels - 1, &lw, &lh, &off);
      std::printf("mips=%d chain=%s last=%dx%d@%lld\n", levels, chain ? "ok" : "null", lw, lh, (long long)off);
      try { t.RegisterGLTextures(1, 2); std::printf("gl=registered\n"); }     // no GL context on a compute box: a clean error
      catch (const tws::Error& e) { std::printf("gl=error %d\n", (int)e.status); }
    } catch (const tws::Error& e) {
      std::printf("error %d: %s\n", (int)e.status, e.what());
      return 2;
    }
  } else {
    tws_params p;
    tws_default_params(&p);
    p.size = 3;                                  // wrong ABI size must be rejected before touching CUDA
    tws_sim* s = nullptr;
    const tws_status st = tws_create(&p, &s);
    std::printf("bad_size_status=%d msg=%s\n", (int)st, tws_last_error(nullptr));
  }
  return 0;
}
