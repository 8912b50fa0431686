// tws_oracle.cpp — CPU restatement of terrainwatersim's shallow-water step.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load it.  The product
// (terrainwatersim_b200 / libtws.so) never links, imports or calls anything here.
//
// PARITY STATUS: PINNED AGAINST THE REFERENCE'S OWN SOURCE.  The reference holds no test or
// golden vector for the step and its GLSL cannot run on a GPU here (no
// GL 4.3 context), but its shader SOURCE can be compiled: oracle/Makefile target `ref_step`
// builds /root/reference/terrainwatersim/shader/{flowUpdate,flowApply,waterBrush}.comp, from
// where they lie, through the GLSL language shim oracle/ref_shim/glsl.h into
// oracle/_ref/libtws_ref_step.so, and tests/test_oracle.py asserts this transcription
// bit-equal to it — BASELINE config 1 (walled + open, 1000 steps), config 2 (brush, 1000
// steps), random 16-multiple grids pass by pass, micro cases on work-group seams, fp16 ties —
// and tests/golden/ holds the hashes the reference shaders produced.  What stays PINNED BY
// ASSUMPTION because the GL driver, not the repository, decides it, is listed below (the shim
// makes the same four choices).  The terrain/initial-state part (create_reference_scene) is
// checked bit-for-bit against the reference's own NoiseGenerator.cpp/Random.cpp compiled from
// /root/reference (oracle/_ref/libtws_ref_terrain.so).  Ragged grid sizes (not multiples of
// 16) and the EXT modes are outside what the reference defines and are oracle-defined.
//
// Pinned arithmetic contract: IEEE binary32, every + - * / a single
// rounding evaluated in the order written in the shader, NO fused multiply-add
// (build with -ffp-contract=off), out-of-range imageLoad = (0,0,0,0), out-of-range
// imageStore dropped, max(0,x) == (0 < x ? x : 0), rg16f store = round-to-nearest-even,
// dispatches strictly sequential.
//
// Layout is the reference's own: TerrainData RGBA32F (r = terrain height, g = b = 0.3
// unused, a = water depth), Flow RGBA32F (x:+X y:-X z:+Y w:-Y outflow), FlowMap RG16F,
// all row-major with index x + y*W (Terrain.cpp:216).
//
// Extensions that the reference does not have (closed-wall boundary, uniform rain,
// evaporation) are defined HERE first and are marked EXT; with their defaults
// (boundary 0, rain 0, evaporation 0) the code path is exactly the reference's.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#if defined(_OPENMP)
#include <omp.h>
#define TWS_OMP_FOR _Pragma("omp parallel for schedule(static)")
#else
#define TWS_OMP_FOR
#endif

namespace {

struct Vec4 { float x, y, z, w; };

// float -> binary16, round-to-nearest-even (pin (3) of SURVEY.md §8c).
inline uint16_t float_to_half_rtne(float f) {
  uint32_t u; std::memcpy(&u, &f, 4);
  const uint32_t sign = (u >> 16) & 0x8000u;
  u &= 0x7fffffffu;
  if (u >= 0x7f800000u) {                       // inf / nan
    return (uint16_t)(sign | 0x7c00u | ((u > 0x7f800000u) ? (0x0200u | ((u >> 13) & 0x3ffu)) : 0u));
  }
  if (u >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);   // rounds to >= 65520 -> inf
  if (u < 0x33000001u) return (uint16_t)sign;                // <= 2^-25 -> 0 (tie to even 0)
  if (u < 0x38800000u) {                                     // half subnormal
    const int e = (int)(u >> 23);                            // biased exponent, 102..112
    const uint32_t m = (u & 0x7fffffu) | 0x800000u;          // 24-bit significand
    const int shift = 126 - e;                               // 14..24
    uint32_t h = m >> shift;
    const uint32_t rem = m & ((1u << shift) - 1u);
    const uint32_t half = 1u << (shift - 1);
    if (rem > half || (rem == half && (h & 1u))) ++h;
    return (uint16_t)(sign | h);
  }
  uint32_t h = (u - 0x38000000u) >> 13;                      // rebias 127 -> 15
  const uint32_t rem = u & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;    // carry into exponent is correct
  return (uint16_t)(sign | h);
}

inline float max0(float v) { return (0.0f < v) ? v : 0.0f; }   // GLSL max(0, v), flowUpdate.comp:54

// imageLoad with the pinned out-of-range rule.
inline Vec4 image_load(const Vec4* img, int W, int H, int x, int y) {
  if (x < 0 || y < 0 || x >= W || y >= H) return Vec4{0.f, 0.f, 0.f, 0.f};
  return img[(size_t)x + (size_t)y * (size_t)W];
}

struct StepConsts { float friction, accel, area_inv; };

// EXT boundary 1 (closed wall): an out-of-range neighbour is read as a copy of the own
// texel in pass 1 (zero gradient, so no outflow across the edge); pass 2 is unchanged
// (exterior flux is 0).  boundary 0 is the reference.
inline float water_level(const Vec4& t) { return t.w + t.x; }  // "terrainInfo.a + terrainInfo.r"

// ---- pass 1: flowUpdate.comp:14-62 ------------------------------------------------
void flow_update(int W, int H, const Vec4* terrain, Vec4* flow, StepConsts c, int boundary) {
  TWS_OMP_FOR
  for (int y = 0; y < H; ++y) {
    for (int x = 0; x < W; ++x) {
      const Vec4 terrainInfo = terrain[(size_t)x + (size_t)y * W];          // :18
      Vec4 tX1 = image_load(terrain, W, H, x + 1, y);                       // :29
      Vec4 tX0 = image_load(terrain, W, H, x - 1, y);                       // :30
      Vec4 tY1 = image_load(terrain, W, H, x, y + 1);                       // :31
      Vec4 tY0 = image_load(terrain, W, H, x, y - 1);                       // :32
      if (boundary == 1) {                                                   // EXT
        if (x + 1 >= W) tX1 = terrainInfo;
        if (x - 1 < 0)  tX0 = terrainInfo;
        if (y + 1 >= H) tY1 = terrainInfo;
        if (y - 1 < 0)  tY0 = terrainInfo;
      }
      const float own = water_level(terrainInfo);                            // :34
      const float hX1 = water_level(tX1), hX0 = water_level(tX0);            // :37-38
      const float hY1 = water_level(tY1), hY0 = water_level(tY0);            // :39-40
      Vec4 n;
      n.x = own - hX1; n.y = own - hX0; n.z = own - hY1; n.w = own - hY0;    // :44-47
      const Vec4 f = flow[(size_t)x + (size_t)y * W];                        // :50
      // :53  flowOut * friction + newFlowOut * accel  (two products, one sum, no fma)
      { float a = f.x * c.friction, b = n.x * c.accel; n.x = a + b; }
      { float a = f.y * c.friction, b = n.y * c.accel; n.y = a + b; }
      { float a = f.z * c.friction, b = n.z * c.accel; n.z = a + b; }
      { float a = f.w * c.friction, b = n.w * c.accel; n.w = a + b; }
      n.x = max0(n.x); n.y = max0(n.y); n.z = max0(n.z); n.w = max0(n.w);   // :54
      const float total = (((n.x + n.y) + n.z) + n.w) * c.area_inv;          // :57
      if (total > terrainInfo.w) {                                           // :58
        const float s = terrainInfo.w / total;                               // :59
        n.x *= s; n.y *= s; n.z *= s; n.w *= s;
      }
      flow[(size_t)x + (size_t)y * W] = n;                                   // :62
    }
  }
}

// ---- pass 2: flowApply.comp:16-52 -------------------------------------------------
// EXT rain_step / evap_step (both already multiplied by dt by the caller):
//   d' = max(0, max(0, d + (in-out)*k) + rain_step - evap_step); with both 0 the inner
//   expression is returned untouched (the EXT branch is skipped entirely).
void flow_apply(int W, int H, Vec4* terrain, const Vec4* flow, uint16_t* flowmap, StepConsts c,
                float rain_step, float evap_step) {
  const bool ext = (rain_step != 0.0f) || (evap_step != 0.0f);
  TWS_OMP_FOR
  for (int y = 0; y < H; ++y) {
    for (int x = 0; x < W; ++x) {
      const Vec4 f = flow[(size_t)x + (size_t)y * W];                        // :20
      const float fX1 = image_load(flow, W, H, x + 1, y).y;                  // :32
      const float fX0 = image_load(flow, W, H, x - 1, y).x;                  // :33
      const float fY1 = image_load(flow, W, H, x, y + 1).w;                  // :34
      const float fY0 = image_load(flow, W, H, x, y - 1).z;                  // :35
      const float in = ((fX1 + fX0) + fY1) + fY0;                            // :38
      const float out = ((f.x + f.y) + f.z) + f.w;                           // :39
      Vec4 t = terrain[(size_t)x + (size_t)y * W];                           // :40
      const float delta = (in - out) * c.area_inv;
      float nw = max0(t.w + delta);                                          // :41
      if (ext) nw = max0((nw + rain_step) - evap_step);                      // EXT
      const float vx = (fX1 - f.x) - (fX0 - f.y);                            // :45
      const float vy = (fY1 - f.z) - (fY0 - f.w);                            // :46
      t.w = nw;                                                              // :50
      terrain[(size_t)x + (size_t)y * W] = t;                                // :51
      flowmap[2 * ((size_t)x + (size_t)y * W) + 0] = float_to_half_rtne(vx); // :52 (rg16f)
      flowmap[2 * ((size_t)x + (size_t)y * W) + 1] = float_to_half_rtne(vy);
    }
  }
}

// ---- inject: waterBrush.comp:20-31 ------------------------------------------------
void water_brush(int W, int H, Vec4* terrain, float cx, float cy, float intensity, float size_sq) {
  TWS_OMP_FOR
  for (int y = 0; y < H; ++y) {
    for (int x = 0; x < W; ++x) {
      Vec4& t = terrain[(size_t)x + (size_t)y * W];                          // :23
      const float bx = cx - (float)x, by = cy - (float)y;                    // :26
      const float xx = bx * bx, yy = by * by;
      const float dist = (xx + yy) / size_sq;                                // :27 dot, then divide
      float s = 1.0f - dist;                                                 // :28 saturate (helper.glsl:68)
      s = (s < 0.0f) ? 0.0f : ((s > 1.0f) ? 1.0f : s);
      const float add = s * intensity;
      t.w = t.w + add;
    }
  }
}

// ---- Random.cpp:22-58 (custom MT19937 variant) ------------------------------------
struct RefRandom {
  static constexpr int N = 624, M = 397;
  uint32_t mt[N]; uint32_t index = 0;
  void init(uint32_t seed) {
    for (int i = 0; i < N; ++i)
      mt[i] = (i % 2) ? (seed + (uint32_t)i * 527u) : ((2135u + seed * 74111u) * (uint32_t)i);   // :27
    for (int i = 0; i < N; ++i) {                                                                 // :31-35
      const uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1 == N) ? 0 : i + 1] & 0x7fffffffu);
      mt[i] = mt[(i + M) % N] ^ (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
    }
    index = 0;
  }
  float next_float() {                                                                            // :41-58
    const uint32_t next = (index + 1 == (uint32_t)N) ? 0 : index + 1;
    uint32_t y = (mt[index] & 0x80000000u) | (mt[next] & 0x7fffffffu);
    mt[index] = mt[(index + M) % N] ^ (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
    y = mt[index];
    index = next;
    y ^= y >> 11;
    y ^= (y << 7) & 0x9D2C5680u;
    y ^= (y << 15) & 0xEFC60000u;
    y ^= y >> 18;
    return (float)(y * 4.656612874e-10 - 1.0);     // uint32 -> double product, then to float
  }
};

// ---- NoiseGenerator.cpp:12-97 (value noise; gradient output not used by the sim) ----
struct RefNoise {
  float white[4096];
  static int ifloor(float a) { int r = (int)a; return r - (int)((a < 0) && (a - r != 0.0f)); }   // NoiseGenerator.h:24
  static float smooth(float f) { return f * f * f * (f * (f * 6.0f - 15.0f) + 10.0f); }           // NoiseGenerator.h:22
  float noise3d(float cx, float cy, float cz, int period) const {                                // :40-97
    period = (int)std::min<uint32_t>((uint32_t)period, 16u);
    const int mod = period - 1;
    int x0 = ifloor(cx), y0 = ifloor(cy), z0 = ifloor(cz);
    const float fx = cx - x0, fy = cy - y0, fz = cz - z0;
    x0 = (x0 % period + period) & mod; y0 = (y0 % period + period) & mod; z0 = (z0 % period + period) & mod;
    const int x1 = (x0 + 1) & mod, y1 = (y0 + 1) & mod, z1 = (z0 + 1) & mod;
    const float s000 = white[x0 + 16 * (y0 + 16 * z0)], s100 = white[x1 + 16 * (y0 + 16 * z0)];
    const float s010 = white[x0 + 16 * (y1 + 16 * z0)], s110 = white[x1 + 16 * (y1 + 16 * z0)];
    const float s001 = white[x0 + 16 * (y0 + 16 * z1)], s101 = white[x1 + 16 * (y0 + 16 * z1)];
    const float s011 = white[x0 + 16 * (y1 + 16 * z1)], s111 = white[x1 + 16 * (y1 + 16 * z1)];
    const float u = smooth(fx), v = smooth(fy), w = smooth(fz);
    const float uv = u * v, uw = u * w, vw = v * w;
    const float k0 = s000, k1 = s100 - s000, k2 = s010 - s000, k3 = s001 - s000;
    const float k4 = s110 - s010 - k1;
    const float k5 = s000 - s010 - s001 + s011;
    const float k6 = -k1 - s001 + s101;
    const float k7 = -k4 + s001 - s101 - s011 + s111;
    return k0 + k1 * u + k2 * v + k3 * w + k4 * uv + k5 * vw + k6 * uw + k7 * uv * w;             // :96
  }
  float value_noise(float cx, float cy, float cz, int lo, int hi, float persistence, bool periodic) const {  // :12-38
    float res = 0.0f, amplitude = 1.0f, frequency = (float)(1 << lo);
    for (int i = lo; i <= hi; ++i) {
      res += amplitude * (noise3d(cx * frequency, cy * frequency, cz * frequency, periodic ? (int)frequency : 16) * 0.5f + 0.5f);
      amplitude *= persistence;
      frequency *= 2.0f;
    }
    return res * 2.0f * (1.0f - persistence) / (1.0f - amplitude) - 1.0f;
  }
};

}  // namespace

extern "C" {

int tws_oracle_threads(void) {
#if defined(_OPENMP)
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// Thread count of the OpenMP build (no effect on the single-threaded one); results do not depend on it.
void tws_oracle_set_threads(int n) {
#if defined(_OPENMP)
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

// Terrain.cpp:175-198 — the three per-step constants, including the float/double mix:
// dt = (double)(1.0f / stepsPerSecond); cellDistance = world / (float)res (float);
// friction = powf(damping, (float)dt); accel = (float)(dt * a * cellDistance) with the
// product evaluated in double; area_inv = (float)(dt / (double)(cellDistance*cellDistance)).
void tws_oracle_derive_consts(float world_size, uint32_t res, float steps_per_second, float damping,
                              float acceleration, float out[3]) {
  const double dt = (double)(1.0f / steps_per_second);
  const float cell = world_size / (float)res;
  out[0] = powf(damping, (float)dt);
  out[1] = (float)(dt * acceleration * cell);
  out[2] = (float)(dt / (cell * cell));
}

void tws_oracle_flow_update(int W, int H, const float* terrain_rgba, float* flow_rgba, const float consts[3], int boundary) {
  flow_update(W, H, (const Vec4*)terrain_rgba, (Vec4*)flow_rgba, StepConsts{consts[0], consts[1], consts[2]}, boundary);
}

void tws_oracle_flow_apply(int W, int H, float* terrain_rgba, const float* flow_rgba, uint16_t* flowmap_rg16,
                           const float consts[3], float rain_step, float evap_step) {
  flow_apply(W, H, (Vec4*)terrain_rgba, (const Vec4*)flow_rgba, flowmap_rg16, StepConsts{consts[0], consts[1], consts[2]},
             rain_step, evap_step);
}

// n x (pass 1; pass 2) — the loop body of Terrain.cpp:253-265.
void tws_oracle_step(int W, int H, float* terrain_rgba, float* flow_rgba, uint16_t* flowmap_rg16, const float consts[3],
                     int n, int boundary, float rain_step, float evap_step) {
  for (int i = 0; i < n; ++i) {
    tws_oracle_flow_update(W, H, terrain_rgba, flow_rgba, consts, boundary);
    tws_oracle_flow_apply(W, H, terrain_rgba, flow_rgba, flowmap_rg16, consts, rain_step, evap_step);
  }
}

void tws_oracle_brush(int W, int H, float* terrain_rgba, float cx, float cy, float intensity, float size_sq) {
  water_brush(W, H, (Vec4*)terrain_rgba, cx, cy, intensity, size_sq);
}

// Terrain.cpp:150-155 — world XZ -> texel coordinates of the brush centre.
void tws_oracle_brush_center(float world_x, float world_z, float world_size, uint32_t res, float out[2]) {
  float px = world_x / world_size, pz = world_z / world_size;
  px = px - truncf(px); pz = pz - truncf(pz);        // ezMath::Fraction, Math_inl.h:213-217
  out[0] = px * (float)res; out[1] = pz * (float)res;
}

// Terrain.cpp:240-247 — frame-time accumulator.  Returns the number of steps (<= 10).
uint32_t tws_oracle_advance(double* accumulator, double step_length, double frame_seconds) {
  *accumulator += frame_seconds;
  uint32_t n = (uint32_t)(*accumulator / step_length);
  *accumulator -= step_length * n;
  return std::min<uint32_t>(n, 10u);
}

// Random::Init + 4096 x NextFloat (NoiseGenerator.cpp:6-10).
void tws_oracle_white_noise(uint32_t seed, float out[4096]) {
  RefRandom r; r.init(seed);
  for (int i = 0; i < 4096; ++i) out[i] = r.next_float();
}

// Terrain.cpp:208-219 — initial TerrainData for a W x H grid (the reference is square;
// for W != H each axis uses its own 1/(n-1) multiplier).  lo/hi octave, persistence
// default 2, 10, 0.43f.
void tws_oracle_create_scene(uint32_t seed, int W, int H, float height_scale, int lo, int hi, float persistence,
                             float* terrain_rgba) {
  RefNoise* ng = new RefNoise;
  tws_oracle_white_noise(seed, ng->white);
  const float mx = 1.0f / (float)(W - 1), my = 1.0f / (float)(H - 1);
  Vec4* out = (Vec4*)terrain_rgba;
  TWS_OMP_FOR
  for (int y = 0; y < H; ++y) {
    for (int x = 0; x < W; ++x) {
      Vec4 t;
      t.x = (ng->value_noise(mx * x, my * y, 0.0f, lo, hi, persistence, true) * 0.5f + 0.5f) * height_scale;
      t.y = 0.3f; t.z = 0.3f;
      const float px = x * mx - 0.5f, py = y * my - 0.5f;
      const float l2 = px * px + py * py;                       // ezVec2::GetLengthSquared
      // Terrain.cpp:219 writes pow(l2, 2.0f) (float overload, SURVEY 8c pin 4).  PINNED to the product: g++ folds that very call
      // to l2 * l2 when it compiles the reference's own generator (oracle/_ref agrees bit for bit), the CUDA scene kernel
      // multiplies, and a library powf need not: glibc's differs from the product for 380 438 of the 1.06e9 binary32
      // arguments in [0, 0.5] (tests/test_oracle.py::test_pow_squared_is_pinned_to_the_product).
      const float p = l2 * l2;
      t.w = std::max(0.0f, (0.45f - p * 800.0f) * height_scale - t.x);
      out[(size_t)x + (size_t)y * W] = t;
    }
  }
  delete ng;
}

// Terrain.cpp:219 evaluates pow(lengthSquared, 2.0f); the CUDA scene kernel evaluates l2 * l2.  Counts the binary32 values in
// [lo_bits, hi_bits] (as bit patterns of non-negative floats) for which this libm's powf(x, 2.0f) differs from x * x.
uint64_t tws_oracle_pow2_mismatches(uint32_t lo_bits, uint32_t hi_bits) {
  uint64_t bad = 0;
  float (*volatile libm_powf)(float, float) = powf;     // through a volatile pointer: the library routine itself, not the
                                                        // compiler's pow(x, 2) -> x * x folding (which g++ applies to the
                                                        // oracle's and the reference generator's own call at Terrain.cpp:219)
#if defined(_OPENMP)
#pragma omp parallel for reduction(+ : bad) schedule(static)
#endif
  for (int64_t b = (int64_t)lo_bits; b <= (int64_t)hi_bits; ++b) {
    const uint32_t u = (uint32_t)b;
    float x; std::memcpy(&x, &u, 4);
    const float p = libm_powf(x, 2.0f), m = x * x;
    uint32_t pu, mu; std::memcpy(&pu, &p, 4); std::memcpy(&mu, &m, 4);
    bad += (pu != mu) ? 1u : 0u;
  }
  return bad;
}

uint16_t tws_oracle_float_to_half(float f) { return float_to_half_rtne(f); }

// FNV-1a-64 over raw bytes — the hash the golden files use.
uint64_t tws_oracle_fnv1a64(const void* data, uint64_t nbytes) {
  const uint8_t* p = (const uint8_t*)data;
  uint64_t h = 0xcbf29ce484222325ull;
  for (uint64_t i = 0; i < nbytes; ++i) { h ^= p[i]; h *= 0x100000001b3ull; }
  return h;
}

}  // extern "C"
