// ref_terrain_driver.cpp — drives the reference's OWN terrain generator
// (/root/reference/terrainwatersim/source/math/{NoiseGenerator,Random}.cpp, compiled
// unmodified where they lie; see oracle/Makefile) exactly as Application.cpp:57 and
// Terrain.cpp:208-219 do, and exposes the result over a C ABI so tests can check the
// restated generator in tws_oracle.cpp bit-for-bit.  TEST INFRASTRUCTURE — only built
// when /root/reference is present; output goes to oracle/_ref/ (git-ignored).
#include "PCH.h"
#include "math/NoiseGenerator.h"
#include "math/Random.h"
#include <cstdint>

// MSVC's <math.h> gives the global pow() a float overload (SURVEY.md §8c pin 4); make g++ pick the same one.
using std::pow;

// ezEngine's assert hook (Foundation/Basics/Assert.h) — the only engine symbol the two
// files reference; the engine library itself is not built.
bool ezFailedCheck(const char*, ezUInt32, const char*, const char*, const char*, ...) { return true; }

extern "C" void tws_ref_white_noise(uint32_t seed, float out[4096]) {
  Random::Init(seed);                                   // Application.cpp:57
  for (int i = 0; i < 4096; ++i) out[i] = Random::NextFloat();
}

extern "C" void tws_ref_create_scene(uint32_t seed, int res, float heightScale, float* rgba) {
  Random::Init(seed);                                   // Application.cpp:57
  NoiseGenerator noiseGen;                              // Terrain.cpp:208
  float mulitplier = 1.0f / static_cast<float>(res - 1);
  for (ezInt32 y = 0; y < res; ++y) {
    for (ezUInt32 x = 0; x < (ezUInt32)res; ++x) {      // body follows Terrain.cpp:216-219
      float* t = rgba + 4 * ((size_t)x + (size_t)y * res);
      t[0] = (noiseGen.GetValueNoise(ezVec3(mulitplier * x, mulitplier * y, 0.0f), 2, 10, 0.43f, true, NULL) * 0.5f + 0.5f) * heightScale;
      t[1] = 0.3f;
      t[2] = 0.3f;
      t[3] = std::max(0.0f, (0.45f - pow(ezVec2(x * mulitplier - 0.5f, y * mulitplier - 0.5f).GetLengthSquared(), 2.0f) * 800.0f) * heightScale - t[0]);
    }
  }
}
