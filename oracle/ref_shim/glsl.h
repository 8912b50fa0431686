// glsl.h — the slice of GLSL 4.30 that terrainwatersim's three simulation compute shaders
// (shader/flowUpdate.comp, flowApply.comp, waterBrush.comp and what they #include) are written in,
// provided as C++ so that those shader sources can be compiled FROM WHERE THEY LIE under
// /root/reference into oracle/_ref/libtws_ref_step.so (oracle/Makefile target `ref_step`,
// recipe oracle/ref_shim/glsl_prep.py, host side oracle/ref_step_driver.cpp).
//
// TEST INFRASTRUCTURE.  Nothing here is simulation arithmetic: the arithmetic that runs is the
// reference's own shader text.  This header only supplies the language: vector types with the
// swizzles the shaders use, component-wise operators (one IEEE binary32 rounding per operator,
// the TU is built with -ffp-contract=off so no a*b+c is ever fused), the built-ins
// max/min/clamp/dot/any/equal/mix/normalize with the definitions of the GLSL 4.30 spec §8.3
// ("max(x,y): y if x < y, otherwise x"; "min(x,y): y if y < x, otherwise x";
// "clamp: min(max(x, minVal), maxVal)"), images, shared memory and barrier().
//
// Implementation-defined behaviour pinned here (same pins as oracle/tws_oracle.cpp):
//   * out-of-range imageLoad returns (0,0,0,0); out-of-range imageStore is dropped;
//   * a store to an rg16f image converts with round-to-nearest-even (done by the compiler's
//     _Float16 conversion, i.e. independently of the oracle's hand-written converter);
//   * no FMA contraction; dispatches are strictly sequential.
#pragma once
#include <cmath>
#include <csetjmp>
#include <cstdint>
#include <cstring>

namespace glsl {

typedef unsigned int uint;

struct vec2; struct vec3; struct vec4; struct ivec2; struct uvec2; struct uvec3; struct bvec2;

// ---- swizzle views (members of the unions below) ---------------------------------------------------
template <class V, class T, int N, int A, int B> struct swz2 {
  T d[N];
  operator V() const { return V(d[A], d[B]); }
  swz2& operator=(const V& v) { const T a = v.x, b = v.y; d[A] = a; d[B] = b; return *this; }
};
template <class V, class T, int N, int A, int B, int C> struct swz3 {
  T d[N];
  operator V() const { return V(d[A], d[B], d[C]); }
  swz3& operator=(const V& v) { const T a = v.x, b = v.y, c = v.z; d[A] = a; d[B] = b; d[C] = c; return *this; }
  swz3& operator/=(T s) { d[A] /= s; d[B] /= s; d[C] /= s; return *this; }
  swz3& operator*=(T s) { d[A] *= s; d[B] *= s; d[C] *= s; return *this; }
};

struct bvec2 { bool x, y; bvec2(bool a, bool b) : x(a), y(b) {} };

struct uvec2 {
  union { struct { uint x, y; }; swz2<uvec2, uint, 2, 0, 1> xy; };
  uvec2() = default;
  explicit uvec2(uint s) : x(s), y(s) {}
  uvec2(uint a, uint b) : x(a), y(b) {}
};
struct uvec3 {
  union { struct { uint x, y, z; }; swz2<uvec2, uint, 3, 0, 1> xy; };
  uvec3() = default;
  uvec3(uint a, uint b, uint c) : x(a), y(b), z(c) {}
};
struct ivec2 {
  union { struct { int x, y; }; swz2<ivec2, int, 2, 0, 1> xy; };
  ivec2() = default;
  explicit ivec2(int s) : x(s), y(s) {}
  ivec2(int a, int b) : x(a), y(b) {}
  explicit ivec2(const uvec2& u) : x((int)u.x), y((int)u.y) {}
};
inline ivec2 operator*(const ivec2& a, const ivec2& b) { return ivec2(a.x * b.x, a.y * b.y); }
inline ivec2 operator+(const ivec2& a, const ivec2& b) { return ivec2(a.x + b.x, a.y + b.y); }
inline ivec2 operator-(const ivec2& a, const ivec2& b) { return ivec2(a.x - b.x, a.y - b.y); }

struct vec2 {
  union { struct { float x, y; }; struct { float r, g; }; swz2<vec2, float, 2, 0, 1> xy; };
  vec2() = default;
  explicit vec2(float s) : x(s), y(s) {}
  vec2(float a, float b) : x(a), y(b) {}
  vec2(const ivec2& i) : x((float)i.x), y((float)i.y) {}     // GLSL implicit conversion ivec2 -> vec2 (spec §4.1.10)
};
struct vec3 {
  union { struct { float x, y, z; }; struct { float r, g, b; }; swz2<vec2, float, 3, 0, 1> xy; swz3<vec3, float, 3, 0, 1, 2> xyz; };
  vec3() = default;
  explicit vec3(float s) : x(s), y(s), z(s) {}
  vec3(float a, float b, float c) : x(a), y(b), z(c) {}
};
struct vec4 {
  union {
    struct { float x, y, z, w; };
    struct { float r, g, b, a; };
    swz2<vec2, float, 4, 0, 1> xy;
    swz2<vec2, float, 4, 2, 3> zw;
    swz3<vec3, float, 4, 0, 1, 2> xyz;
  };
  vec4() = default;
  explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
  vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
  vec4(const vec2& p, float c, float d) : x(p.x), y(p.y), z(c), w(d) {}
  vec4(const vec3& p, float d) : x(p.x), y(p.y), z(p.z), w(d) {}
};
struct mat4 { vec4 col[4]; };     // column-major like GLSL

// ---- component-wise arithmetic: every operator is one binary32 operation per component ---------------
#define GLSL_VEC_OPS(V, EXPAND)                                                                          \
  inline V operator+(const V& a, const V& b) { return EXPAND(a., +, b.); }                               \
  inline V operator-(const V& a, const V& b) { return EXPAND(a., -, b.); }                               \
  inline V operator*(const V& a, const V& b) { return EXPAND(a., *, b.); }                               \
  inline V operator/(const V& a, const V& b) { return EXPAND(a., /, b.); }                               \
  inline V operator*(const V& a, float s) { const V b(s); return EXPAND(a., *, b.); }                    \
  inline V operator*(float s, const V& b) { const V a(s); return EXPAND(a., *, b.); }                    \
  inline V operator/(const V& a, float s) { const V b(s); return EXPAND(a., /, b.); }                    \
  inline V operator+(const V& a, float s) { const V b(s); return EXPAND(a., +, b.); }                    \
  inline V operator-(const V& a, float s) { const V b(s); return EXPAND(a., -, b.); }                    \
  inline V operator-(const V& a) { return V(0.0f) - a; }                                                 \
  inline V& operator+=(V& a, const V& b) { a = a + b; return a; }                                        \
  inline V& operator-=(V& a, const V& b) { a = a - b; return a; }                                        \
  inline V& operator*=(V& a, const V& b) { a = a * b; return a; }                                        \
  inline V& operator*=(V& a, float s) { a = a * s; return a; }                                           \
  inline V& operator/=(V& a, float s) { a = a / s; return a; }
#define GLSL_E2(a, op, b) vec2(a x op b x, a y op b y)
#define GLSL_E3(a, op, b) vec3(a x op b x, a y op b y, a z op b z)
#define GLSL_E4(a, op, b) vec4(a x op b x, a y op b y, a z op b z, a w op b w)
GLSL_VEC_OPS(vec2, GLSL_E2)
GLSL_VEC_OPS(vec3, GLSL_E3)
GLSL_VEC_OPS(vec4, GLSL_E4)
#undef GLSL_VEC_OPS
#undef GLSL_E2
#undef GLSL_E3
#undef GLSL_E4
inline vec4 operator*(const mat4& m, const vec4& v) { return ((m.col[0] * v.x + m.col[1] * v.y) + m.col[2] * v.z) + m.col[3] * v.w; }

// ---- built-in functions (GLSL 4.30 spec §8.3, §8.5, §8.7) ------------------------------------------
inline float max(float x, float y) { return (x < y) ? y : x; }
inline float min(float x, float y) { return (y < x) ? y : x; }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline vec2 max(const vec2& a, const vec2& b) { return vec2(max(a.x, b.x), max(a.y, b.y)); }
inline vec3 max(const vec3& a, const vec3& b) { return vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline vec4 max(const vec4& a, const vec4& b) { return vec4(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z), max(a.w, b.w)); }
inline float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }                     // x[0]*y[0] + x[1]*y[1]
inline float dot(const vec3& a, const vec3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float dot(const vec4& a, const vec4& b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }
inline bvec2 equal(const uvec2& a, const uvec2& b) { return bvec2(a.x == b.x, a.y == b.y); }
inline bool any(const bvec2& b) { return b.x || b.y; }
inline vec3 mix(const vec3& x, const vec3& y, float a) { return x * (1.0f - a) + y * a; }
inline vec3 normalize(const vec3& v) { return v / std::sqrt(dot(v, v)); }

// ---- images (glBindImageTexture targets) --------------------------------------------------------------
enum image_format { FORMAT_NONE = 0, FORMAT_RGBA32F, FORMAT_RG16F };
struct image2D {
  void* texels = nullptr;            // level 0, row-major, index x + y*width (Terrain.cpp:216)
  int width = 0, height = 0;
  image_format format = FORMAT_NONE;
};
inline bool in_range(const image2D& img, const ivec2& p) { return p.x >= 0 && p.y >= 0 && p.x < img.width && p.y < img.height; }
inline vec4 image_load(const image2D& img, const ivec2& p) {
  if (!in_range(img, p) || img.format != FORMAT_RGBA32F) return vec4(0.0f);      // pinned: invalid loads return zero
  const float* t = static_cast<const float*>(img.texels) + 4 * ((size_t)p.x + (size_t)p.y * (size_t)img.width);
  return vec4(t[0], t[1], t[2], t[3]);
}
inline void image_store(const image2D& img, const ivec2& p, const vec4& v) {
  if (!in_range(img, p)) return;                                                  // pinned: invalid stores are dropped
  const size_t i = (size_t)p.x + (size_t)p.y * (size_t)img.width;
  if (img.format == FORMAT_RGBA32F) {
    float* t = static_cast<float*>(img.texels) + 4 * i;
    t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
  } else if (img.format == FORMAT_RG16F) {                                        // pinned: round to nearest even
    const _Float16 hx = (_Float16)v.x, hy = (_Float16)v.y;
    uint16_t* t = static_cast<uint16_t*>(img.texels) + 2 * i;
    std::memcpy(t + 0, &hx, 2); std::memcpy(t + 1, &hy, 2);
  }
}

// ---- one work group of a compute shader -----------------------------------------------------------------
// The generated shader struct derives from this; `shared` arrays become its members (one object = one
// work group).  barrier(): the invocations of a group are run one after the other in PASSES.  In pass p an
// invocation is replayed from the top of main(); its first p barriers fall through and barrier p+1 leaves
// main() (longjmp), so that all invocations reach that barrier before any passes it; an invocation whose
// main() returns is finished and is not run again (the shaders' ring threads `return` before the barrier).
// Replay equals suspension for these shaders because everything before a barrier is a pure function of
// the invocation id and of images no invocation of the dispatch writes before that barrier
// (flowUpdate.comp:14-19 reads TerrainData, flowApply.comp:16-21 reads OutgoingFlow — both `readonly`),
// so a replay re-writes the same values into shared memory.
struct compute_shader {
  uvec3 gl_WorkGroupID, gl_LocalInvocationID, gl_GlobalInvocationID;
  image2D image_unit[8];             // glBindImageTexture(unit, ...)
  std::jmp_buf at_barrier_;
  int pass_ = 0, barriers_seen_ = 0;
  void barrier() { if (++barriers_seen_ > pass_) longjmp(at_barrier_, 1); }
  // an image uniform is its binding unit (GLSL_IMAGE below), as in glBindImageTexture(unit, ...)
  vec4 imageLoad(int unit, const ivec2& p) const { return image_load(image_unit[unit], p); }
  void imageStore(int unit, const ivec2& p, const vec4& v) const { image_store(image_unit[unit], p, v); }
};

// What glsl_prep.py turns the shaders' `layout(...)` declarations into:
#define GLSL_IMAGE(unit, fmt, name) static constexpr int name = unit; static constexpr ::glsl::image_format name##_format = ::glsl::FORMAT_##fmt;
#define GLSL_UNIFORM_BLOCK_BEGIN(unit, name) static constexpr int name##_binding = unit;
#define GLSL_UNIFORM_BLOCK_END
#define GLSL_LOCAL_SIZE(sx, sy, sz) static constexpr unsigned local_size_x = sx, local_size_y = sy, local_size_z = sz;

// true when main() ran to its end, false when it stopped at a barrier
template <class S> __attribute__((noinline)) bool run_invocation(S& s) {
  s.barriers_seen_ = 0;
  if (_setjmp(s.at_barrier_) != 0) return false;
  s.main();
  return true;
}

// glDispatchCompute(groups_x, groups_y, 1) with `bound` carrying the uniforms and image bindings.
template <class S> void dispatch_compute(const S& bound, unsigned groups_x, unsigned groups_y) {
  constexpr unsigned LX = S::local_size_x, LY = S::local_size_y, NINV = LX * LY;
#pragma omp parallel
  {
    S* group = new S(bound);
    bool* finished = new bool[NINV];
#pragma omp for collapse(2) schedule(static)
    for (unsigned gy = 0; gy < groups_y; ++gy)
      for (unsigned gx = 0; gx < groups_x; ++gx) {
        group->gl_WorkGroupID = uvec3(gx, gy, 0);
        unsigned left = NINV;
        for (unsigned i = 0; i < NINV; ++i) finished[i] = false;
        for (int pass = 0; left != 0; ++pass) {
          group->pass_ = pass;
          for (unsigned ly = 0; ly < LY; ++ly)
            for (unsigned lx = 0; lx < LX; ++lx) {
              if (finished[lx + ly * LX]) continue;
              group->gl_LocalInvocationID = uvec3(lx, ly, 0);
              group->gl_GlobalInvocationID = uvec3(gx * LX + lx, gy * LY + ly, 0);
              if (run_invocation(*group)) { finished[lx + ly * LX] = true; --left; }
            }
        }
      }
    delete[] finished;
    delete group;
  }
}

}  // namespace glsl
