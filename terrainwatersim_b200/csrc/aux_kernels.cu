// aux_kernels.cu — everything around the step: inject (waterBrush.comp), create/reset
// (Terrain.cpp:200-238 on the GPU), layout packing for the reference's RGBA textures,
// the fp64 volume reduction, and the NVLink halo push / step flags of the strip exchange.
// Built with -fmad=false like step_kernels.cu (same arithmetic contract).
#include "tws_internal.h"

#include <cuda_fp16.h>

namespace tws {

// ---- inject: waterBrush.comp:20-31 over the brush's bounding box only -------------------
// Cells outside the box get saturate(1 - dist) == 0 exactly, so d + 0*intensity == d and
// the result equals the reference's whole-grid pass bit for bit.
__global__ void brush_kernel(Geom g, float* d /* local row 0 */, float cx, float cy, float intensity, float size_sq,
                             int bx0, int by0, int bw, int bh) {
  const int ix = blockIdx.x * blockDim.x + threadIdx.x, iy = blockIdx.y * blockDim.y + threadIdx.y;
  if (ix >= bw || iy >= bh) return;
  const int x = bx0 + ix, gy = by0 + iy;               // global texel
  const float tx = cx - (float)x, ty = cy - (float)gy;                               // :26
  const float dist = __fdiv_rn(__fadd_rn(__fmul_rn(tx, tx), __fmul_rn(ty, ty)), size_sq);   // :27
  float s = 1.0f - dist;                                                             // :28
  s = (s < 0.0f) ? 0.0f : ((s > 1.0f) ? 1.0f : s);
  const long long o = (long long)(gy - g.row0) * g.pitch + x;
  d[o] = __fadd_rn(d[o], __fmul_rn(s, intensity));
}

bool brush_bbox(const Geom& g, float cx, float cy, float size_sq, int* px0, int* px1, int* py0, int* py1) {
  if (!(size_sq > 0.0f) || !isfinite(cx) || !isfinite(cy)) return false;              // NaN/0 size: reference adds 0*i or NaN; we refuse upstream
  const float rad = ceilf(sqrtf(size_sq)) + 1.0f;
  // Box in global texels, clipped to the rows this strip stores (own rows + live halos).
  const double lo_x = floor((double)cx - rad), hi_x = ceil((double)cx + rad);
  const double lo_y = floor((double)cy - rad), hi_y = ceil((double)cy + rad);
  const int row_lo = g.row0 - (g.has_up ? TWS_HALO_ROWS : 0), row_hi = g.row0 + g.rows + (g.has_down ? TWS_HALO_ROWS : 0);
  // a box that misses the stored rows / columns entirely adds 0 everywhere, like the reference's whole-grid pass; decided in
  // double BEFORE any cast to int (a finite centre beyond INT_MAX must not reach the casts)
  if (lo_x > (double)g.W - 1 || hi_x < 0.0 || lo_y > (double)row_hi - 1 || hi_y < (double)row_lo) return false;
  const int x0 = (int)fmax(lo_x, 0.0), x1 = (int)fmin(hi_x, (double)g.W - 1);
  const int y0 = (int)fmax(lo_y, (double)row_lo), y1 = (int)fmin(hi_y, (double)row_hi - 1);
  if (x1 < x0 || y1 < y0) return false;
  *px0 = x0; *px1 = x1; *py0 = y0; *py1 = y1;
  return true;
}

cudaError_t launch_brush(const Geom& g, float* d, float cx, float cy, float intensity, float size_sq, cudaStream_t st, int* launched) {
  *launched = 0;
  int x0, x1, y0, y1;
  if (!brush_bbox(g, cx, cy, size_sq, &x0, &x1, &y0, &y1)) return cudaSuccess;
  const int bw = x1 - x0 + 1, bh = y1 - y0 + 1;
  dim3 block(16, 16), grid((bw + 15) / 16, (bh + 15) / 16);
  brush_kernel<<<grid, block, 0, st>>>(g, d, cx, cy, intensity, size_sq, x0, y0, bw, bh);
  *launched = 1;
  return cudaGetLastError();
}

// ---- create/reset: NoiseGenerator.cpp:12-97 + Terrain.cpp:209-219 ------------------------
__device__ __forceinline__ int noise_floor(float a) { int r = (int)a; return r - (int)((a < 0) && (a - (float)r != 0.0f)); }
__device__ __forceinline__ float noise_smooth(float f) { return f * f * f * (f * (f * 6.0f - 15.0f) + 10.0f); }

__device__ float noise3d(const float* white, float cx, float cy, float cz, int period) {
  period = period < 16 ? period : 16;                  // min<uint>(period, 16); period >= 1 here
  const int mod = period - 1;
  int x0 = noise_floor(cx), y0 = noise_floor(cy), z0 = noise_floor(cz);
  const float fx = cx - (float)x0, fy = cy - (float)y0, fz = cz - (float)z0;
  x0 = (x0 % period + period) & mod; y0 = (y0 % period + period) & mod; z0 = (z0 % period + period) & mod;
  const int x1 = (x0 + 1) & mod, y1 = (y0 + 1) & mod, z1 = (z0 + 1) & mod;
  const float s000 = white[x0 + 16 * (y0 + 16 * z0)], s100 = white[x1 + 16 * (y0 + 16 * z0)];
  const float s010 = white[x0 + 16 * (y1 + 16 * z0)], s110 = white[x1 + 16 * (y1 + 16 * z0)];
  const float s001 = white[x0 + 16 * (y0 + 16 * z1)], s101 = white[x1 + 16 * (y0 + 16 * z1)];
  const float s011 = white[x0 + 16 * (y1 + 16 * z1)], s111 = white[x1 + 16 * (y1 + 16 * z1)];
  const float u = noise_smooth(fx), v = noise_smooth(fy), w = noise_smooth(fz);
  const float uv = u * v, uw = u * w, vw = v * w;
  const float k0 = s000, k1 = s100 - s000, k2 = s010 - s000, k3 = s001 - s000;
  const float k4 = s110 - s010 - k1;
  const float k5 = s000 - s010 - s001 + s011;
  const float k6 = -k1 - s001 + s101;
  const float k7 = -k4 + s001 - s101 - s011 + s111;
  return k0 + k1 * u + k2 * v + k3 * w + k4 * uv + k5 * vw + k6 * uw + k7 * uv * w;
}

__global__ void __launch_bounds__(256) scene_kernel(Geom g, float* h, float* d /* local row 0 */, const float* __restrict__ white_g,
                                                    float height_scale, int lo, int hi, float persistence, int lr0, int lr1, int tile_h) {
  __shared__ float white[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) white[i] = white_g[i];
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= g.W) return;
  // tile_h < Hg: the scene of a W x tile_h grid repeated down the grid (weak-scaling workloads: every strip of
  // tile_h rows is the same scene); the reference case is tile_h == Hg
  const float mx = __fdiv_rn(1.0f, (float)(g.W - 1)), my = __fdiv_rn(1.0f, (float)(tile_h - 1));
  for (int lr = lr0 + blockIdx.y; lr < lr1; lr += gridDim.y) {
    const int gy = ((g.row0 + lr) % tile_h + tile_h) % tile_h;
    const float cx = mx * (float)x, cy = my * (float)gy, cz = 0.0f;
    float res = 0.0f, amplitude = 1.0f, frequency = (float)(1 << lo);
    for (int i = lo; i <= hi; ++i) {
      res += amplitude * (noise3d(white, cx * frequency, cy * frequency, cz * frequency, (int)frequency) * 0.5f + 0.5f);
      amplitude *= persistence;
      frequency *= 2.0f;
    }
    const float n = __fdiv_rn(res * 2.0f * (1.0f - persistence), (1.0f - amplitude)) - 1.0f;
    const float terrain = (n * 0.5f + 0.5f) * height_scale;
    const float px = (float)x * mx - 0.5f, py = (float)gy * my - 0.5f;
    const float l2 = px * px + py * py;
    const float p = l2 * l2;                             // pow(l2, 2.0f), PINNED to the product (g++ folds the
                                                         // reference generator's own call to this; a library powf may differ in the last bit)
    const float water = (0.45f - p * 800.0f) * height_scale - terrain;
    const long long o = (long long)lr * g.pitch + x;
    h[o] = terrain;
    d[o] = (0.0f < water) ? water : 0.0f;                // std::max(0.0f, water)
  }
}

cudaError_t launch_scene(const Geom& g, const Planes& p, int side, const float* white_dev, float height_scale, int lo, int hi,
                         float persistence, int tile_h, cudaStream_t st) {
  const size_t off = (size_t)TWS_HALO_ROWS * g.pitch;
  const int lr0 = g.has_up ? -TWS_HALO_ROWS : 0, lr1 = g.rows + (g.has_down ? TWS_HALO_ROWS : 0);
  dim3 block(256), grid((g.W + 255) / 256, (lr1 - lr0) < 4096 ? (lr1 - lr0) : 4096);
  scene_kernel<<<grid, block, 0, st>>>(g, p.h + off, p.d[side] + off, white_dev, height_scale, lo, hi, persistence, lr0, lr1, tile_h);
  return cudaGetLastError();
}

// ---- layout packing -------------------------------------------------------------------------
// Flux: planar (+X,-X,+Y,-Y) <-> the reference's RGBA32F texel (m_waterOutgoingFlow).
__global__ void pack_flux_kernel(Geom g, float* f0, float* f1, float* f2, float* f3, float4* aos, int lr0, int nrows, int to_aos) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= g.W) return;
  for (int r = blockIdx.y; r < nrows; r += gridDim.y) {
    const long long o = (long long)(lr0 + r) * g.pitch + x;
    const long long a = (long long)r * g.W + x;
    if (to_aos) aos[a] = make_float4(f0[o], f1[o], f2[o], f3[o]);
    else { const float4 v = aos[a]; f0[o] = v.x; f1[o] = v.y; f2[o] = v.z; f3[o] = v.w; }
  }
}
// TerrainInfo: (r = terrain, g = b = 0.3, a = water) <-> planar h, d  (Terrain.cpp:216-219).
__global__ void pack_info_kernel(Geom g, float* h, float* d, float4* aos, int lr0, int nrows, int to_aos) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= g.W) return;
  for (int r = blockIdx.y; r < nrows; r += gridDim.y) {
    const long long o = (long long)(lr0 + r) * g.pitch + x;
    const long long a = (long long)r * g.W + x;
    if (to_aos) aos[a] = make_float4(h[o], 0.3f, 0.3f, d[o]);
    else { const float4 v = aos[a]; h[o] = v.x; d[o] = v.w; }
  }
}

// The renderer hand-off in ONE pass over the planar state (Terrain.cpp:272-276,288,323-330): TerrainInfo level 0 (r = terrain,
// g = b = 0.3, a = water), its mip level 1 (the 2x2 box of mip_level_kernel below, same operation order) and the flow map.
// One thread per 2x2 block of cells.  Replaces pack + flow-map copy + first mip launch and the re-read of the 16 B texels
// the first mip level would need; every frame of the reference pays for this hand-off.
__global__ void __launch_bounds__(256) publish_fused_kernel(Geom g, const float* __restrict__ h, const float* __restrict__ d,
                                                            const uint32_t* __restrict__ v, float4* __restrict__ l0, float4* __restrict__ l1,
                                                            uint32_t* __restrict__ flow, int W, int H, int w1, int h1) {
  const int X = blockIdx.x * blockDim.x + threadIdx.x;
  if (2 * X >= W) return;
  for (int Y = blockIdx.y; 2 * Y < H; Y += gridDim.y) {
    const int x0 = 2 * X, x1 = min(2 * X + 1, W - 1), y0 = 2 * Y, y1 = min(2 * Y + 1, H - 1);
    const long long o00 = (long long)y0 * g.pitch + x0, o01 = (long long)y0 * g.pitch + x1;
    const long long o10 = (long long)y1 * g.pitch + x0, o11 = (long long)y1 * g.pitch + x1;
    const float4 a = make_float4(h[o00], 0.3f, 0.3f, d[o00]), b = make_float4(h[o01], 0.3f, 0.3f, d[o01]);
    const float4 c = make_float4(h[o10], 0.3f, 0.3f, d[o10]), e = make_float4(h[o11], 0.3f, 0.3f, d[o11]);
    l0[(long long)y0 * W + x0] = a;
    if (x1 != x0) l0[(long long)y0 * W + x1] = b;
    if (y1 != y0) {
      l0[(long long)y1 * W + x0] = c;
      if (x1 != x0) l0[(long long)y1 * W + x1] = e;
    }
    if (flow != nullptr) {
      flow[(long long)y0 * W + x0] = v[o00];
      if (x1 != x0) flow[(long long)y0 * W + x1] = v[o01];
      if (y1 != y0) {
        flow[(long long)y1 * W + x0] = v[o10];
        if (x1 != x0) flow[(long long)y1 * W + x1] = v[o11];
      }
    }
    if (l1 != nullptr && X < w1 && Y < h1) {
      float4 o;
      o.x = __fmul_rn(__fadd_rn(__fadd_rn(a.x, b.x), __fadd_rn(c.x, e.x)), 0.25f);
      o.y = __fmul_rn(__fadd_rn(__fadd_rn(a.y, b.y), __fadd_rn(c.y, e.y)), 0.25f);
      o.z = __fmul_rn(__fadd_rn(__fadd_rn(a.z, b.z), __fadd_rn(c.z, e.z)), 0.25f);
      o.w = __fmul_rn(__fadd_rn(__fadd_rn(a.w, b.w), __fadd_rn(c.w, e.w)), 0.25f);
      l1[(long long)Y * w1 + X] = o;
    }
  }
}
cudaError_t launch_publish_fused(const Geom& g, const Planes& p, int side, float* level0, float* level1, uint32_t* flow, cudaStream_t st) {
  const size_t off = (size_t)TWS_HALO_ROWS * g.pitch;
  const int W = g.W, H = g.rows, bx = (W + 1) / 2, by = (H + 1) / 2;
  const int w1 = W > 1 ? W >> 1 : 1, h1 = H > 1 ? H >> 1 : 1;
  dim3 block(128), grid((bx + 127) / 128, by < 2048 ? by : 2048);
  publish_fused_kernel<<<grid, block, 0, st>>>(g, p.h + off, p.d[side] + off, p.v + off, (float4*)level0, (float4*)level1, flow, W, H, w1, h1);
  return cudaGetLastError();
}

cudaError_t launch_pack_flux(const Geom& g, const Planes& p, int side, float* aos, int lr0, int nrows, bool to_aos, cudaStream_t st) {
  const size_t off = (size_t)TWS_HALO_ROWS * g.pitch;
  dim3 block(256), grid((g.W + 255) / 256, nrows < 1024 ? nrows : 1024);
  pack_flux_kernel<<<grid, block, 0, st>>>(g, p.F[side][0] + off, p.F[side][1] + off, p.F[side][2] + off, p.F[side][3] + off,
                                           (float4*)aos, lr0, nrows, to_aos ? 1 : 0);
  return cudaGetLastError();
}
cudaError_t launch_pack_info(const Geom& g, const Planes& p, int side, float* aos, int lr0, int nrows, bool to_aos, cudaStream_t st) {
  const size_t off = (size_t)TWS_HALO_ROWS * g.pitch;
  dim3 block(256), grid((g.W + 255) / 256, nrows < 1024 ? nrows : 1024);
  pack_info_kernel<<<grid, block, 0, st>>>(g, p.h + off, p.d[side] + off, (float4*)aos, lr0, nrows, to_aos ? 1 : 0);
  return cudaGetLastError();
}

// ---- mip chain of TerrainInfo ---------------------------------------------------------------------
// The reference regenerates the full mip chain of m_terrainData after every frame that stepped
// (Terrain.cpp:272-276 -> Texture2D::GenMipMaps, glEasy Texture2D.cpp:64-68: glGenerateMipmap).  GL
// leaves the filter to the driver; pinned here: level L has max(1, size >> L)
// texels per side, each the 2x2 box average ((t00 + t10) + (t01 + t11)) * 0.25 of level L-1 per
// channel, source coordinates clamped to the source (a side that has shrunk to 1 keeps averaging
// the other).  One thread per destination texel; levels >= 1 hold a third of level 0 in total.
__global__ void __launch_bounds__(256) mip_level_kernel(const float4* __restrict__ src, int sw, int sh, float4* __restrict__ dst, int dw, int dh) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= dw) return;
  for (int y = blockIdx.y; y < dh; y += gridDim.y) {
    const int x0 = min(2 * x, sw - 1), x1 = min(2 * x + 1, sw - 1);
    const long long r0 = (long long)min(2 * y, sh - 1) * sw, r1 = (long long)min(2 * y + 1, sh - 1) * sw;
    const float4 a = src[r0 + x0], b = src[r0 + x1], c = src[r1 + x0], d = src[r1 + x1];
    float4 o;
    o.x = __fmul_rn(__fadd_rn(__fadd_rn(a.x, b.x), __fadd_rn(c.x, d.x)), 0.25f);
    o.y = __fmul_rn(__fadd_rn(__fadd_rn(a.y, b.y), __fadd_rn(c.y, d.y)), 0.25f);
    o.z = __fmul_rn(__fadd_rn(__fadd_rn(a.z, b.z), __fadd_rn(c.z, d.z)), 0.25f);
    o.w = __fmul_rn(__fadd_rn(__fadd_rn(a.w, b.w), __fadd_rn(c.w, d.w)), 0.25f);
    dst[(long long)y * dw + x] = o;
  }
}
// The tail of the chain — every level from `first` on, once a level has shrunk to a few thousand texels — in ONE
// launch of one CTA: at the reference's own size (1024^2, 11 levels) the small levels are pure launch latency, and
// the renderer hand-off runs every frame.  Same arithmetic as mip_level_kernel; the levels are produced in order,
// a block barrier between them (plain loads: a level is read by the CTA that has just written it).
__global__ void __launch_bounds__(1024) mip_tail_kernel(float4* base, int W, int H, int first, int last) {
  long long off = 0;                                   // offset (texels) and size of level `first - 1`
  int w = W, h = H;
  for (int l = 0; l < first - 1; ++l) { off += (long long)w * h; w = max(1, w >> 1); h = max(1, h >> 1); }
  for (int l = first; l < last; ++l) {
    const int sw = w, sh = h;
    const float4* src = base + off;
    off += (long long)sw * sh;
    w = max(1, sw >> 1); h = max(1, sh >> 1);
    float4* dst = base + off;
    for (int i = threadIdx.x; i < w * h; i += blockDim.x) {
      const int x = i % w, y = i / w;
      const int x0 = min(2 * x, sw - 1), x1 = min(2 * x + 1, sw - 1);
      const int r0 = min(2 * y, sh - 1) * sw, r1 = min(2 * y + 1, sh - 1) * sw;
      const float4 a = src[r0 + x0], b = src[r0 + x1], c = src[r1 + x0], d = src[r1 + x1];
      float4 o;
      o.x = __fmul_rn(__fadd_rn(__fadd_rn(a.x, b.x), __fadd_rn(c.x, d.x)), 0.25f);
      o.y = __fmul_rn(__fadd_rn(__fadd_rn(a.y, b.y), __fadd_rn(c.y, d.y)), 0.25f);
      o.z = __fmul_rn(__fadd_rn(__fadd_rn(a.z, b.z), __fadd_rn(c.z, d.z)), 0.25f);
      o.w = __fmul_rn(__fadd_rn(__fadd_rn(a.w, b.w), __fadd_rn(c.w, d.w)), 0.25f);
      dst[i] = o;
    }
    __syncthreads();
  }
}
// Single-pass chain ("single-pass downsampler"): every CTA reduces its tile through six levels out of shared memory and the last
// CTA to finish (device-wide ticket) reduces what is left, down to 1 x 1.  Same 2x2 box and operation order as
// mip_level_kernel.  At the reference's own size (1024^2) the separate path took a pack launch, a copy, two mip launches and
// a single-CTA tail that went back to L2 between its eight levels (22 us alone); the renderer hand-off runs every frame.
__device__ __forceinline__ float4 mip_box(const float4& a, const float4& b, const float4& c, const float4& d) {
  float4 o;
  o.x = __fmul_rn(__fadd_rn(__fadd_rn(a.x, b.x), __fadd_rn(c.x, d.x)), 0.25f);
  o.y = __fmul_rn(__fadd_rn(__fadd_rn(a.y, b.y), __fadd_rn(c.y, d.y)), 0.25f);
  o.z = __fmul_rn(__fadd_rn(__fadd_rn(a.z, b.z), __fadd_rn(c.z, d.z)), 0.25f);
  o.w = __fmul_rn(__fadd_rn(__fadd_rn(a.w, b.w), __fadd_rn(c.w, d.w)), 0.25f);
  return o;
}
// The last CTA of a single-pass chain kernel: levels [first, nlevels) from level first-1 (cw x ch texels at texel offset soff),
// written by OTHER CTAs (visible: they fenced before taking their ticket).  Levels that fit the two shared buffers (<= 1024 and
// <= 256 texels) are reduced out of shared memory — no trip to L2 between them; larger ones go through global memory.
__device__ __forceinline__ void mip_finish_chain(float4* base, long long soff, int cw, int ch, int first, int nlevels, float4* sa, float4* sb) {
  const int tid = threadIdx.x;
  bool in_smem = false;                                 // level first-1 sits in `cur` (shared) as well as in global memory
  float4* cur = sa; float4* nxt = sb;
  for (int l = first; l < nlevels; ++l) {
    const int nw = max(1, cw >> 1), nh = max(1, ch >> 1);
    const float4* sp = base + soff;
    float4* dp = base + soff + (long long)cw * ch;
    if (!in_smem && cw * ch <= 1024 && nw * nh <= 256) {   // bring the source level into shared memory once
      for (int i = tid; i < cw * ch; i += blockDim.x) sa[i] = __ldcg(sp + i);
      cur = sa; nxt = sb; in_smem = true;
      __syncthreads();
    }
    for (int i = tid; i < nw * nh; i += blockDim.x) {
      const int x = i % nw, y = i / nw;
      const int x0 = min(2 * x, cw - 1), x1 = min(2 * x + 1, cw - 1);
      const int r0 = min(2 * y, ch - 1) * cw, r1 = min(2 * y + 1, ch - 1) * cw;
      float4 o;
      if (in_smem) { o = mip_box(cur[r0 + x0], cur[r0 + x1], cur[r1 + x0], cur[r1 + x1]); nxt[i] = o; }
      else o = mip_box(__ldcg(sp + r0 + x0), __ldcg(sp + r0 + x1), __ldcg(sp + r1 + x0), __ldcg(sp + r1 + x1));
      dp[i] = o;
    }
    if (!in_smem) __threadfence();
    __syncthreads();
    if (in_smem) { float4* t = cur; cur = nxt; nxt = t; }
    soff += (long long)cw * ch;
    cw = nw; ch = nh;
  }
}
// The whole renderer hand-off of a whole grid whose sides are multiples of 64 in ONE launch: every CTA takes a 64 x 64 tile of
// CELLS, writes its TerrainInfo level-0 texels and flow-map texels, reduces levels 1..6 of the tile (32^2 .. 1 texels; level 1
// straight from the planar state, the rest out of shared memory), and the last CTA to finish reduces the remaining levels.
__global__ void __launch_bounds__(256) publish_all_kernel(Geom g, const float* __restrict__ h, const float* __restrict__ d,
                                                          const uint32_t* __restrict__ v, float4* base, uint32_t* __restrict__ flow, int W, int H,
                                                          int nlevels, unsigned int* ticket) {
  __shared__ float4 sa[32 * 32];
  __shared__ float4 sb[16 * 16];
  __shared__ unsigned int s_last;
  const int tid = threadIdx.x, bx = blockIdx.x, by = blockIdx.y;
  float4* l0 = base;
  long long doff = (long long)W * H;                   // level 1
  {
    const int dw = W >> 1;
    float4* l1 = base + doff;
    for (int i = tid; i < 32 * 32; i += 256) {
      const int x = i & 31, y = i >> 5;
      const int cx = bx * 64 + 2 * x, cy = by * 64 + 2 * y;
      const long long o00 = (long long)cy * g.pitch + cx, o10 = o00 + g.pitch;
      const float2 h0 = *reinterpret_cast<const float2*>(h + o00), h1 = *reinterpret_cast<const float2*>(h + o10);
      const float2 d0 = *reinterpret_cast<const float2*>(d + o00), d1 = *reinterpret_cast<const float2*>(d + o10);
      const float4 a = make_float4(h0.x, 0.3f, 0.3f, d0.x), b = make_float4(h0.y, 0.3f, 0.3f, d0.y);
      const float4 c = make_float4(h1.x, 0.3f, 0.3f, d1.x), e = make_float4(h1.y, 0.3f, 0.3f, d1.y);
      const long long t00 = (long long)cy * W + cx;
      l0[t00] = a; l0[t00 + 1] = b; l0[t00 + W] = c; l0[t00 + W + 1] = e;
      *reinterpret_cast<uint2*>(flow + t00) = *reinterpret_cast<const uint2*>(v + o00);
      *reinterpret_cast<uint2*>(flow + t00 + W) = *reinterpret_cast<const uint2*>(v + o10);
      const float4 o = mip_box(a, b, c, e);
      sa[i] = o;
      l1[(long long)(by * 32 + y) * dw + bx * 32 + x] = o;
    }
  }
  __syncthreads();
  float4* cur = sa; float4* nxt = sb;
  int cw = 32, lw = W >> 1, lh = H >> 1;
  doff += (long long)lw * lh;
#pragma unroll 1
  for (int step = 2; step <= 6; ++step) {
    const int nw = cw >> 1, dw = lw >> 1;
    float4* dst = base + doff;
    if (tid < nw * nw) {
      const int x = tid % nw, y = tid / nw;
      const float4 o = mip_box(cur[(2 * y) * cw + 2 * x], cur[(2 * y) * cw + 2 * x + 1], cur[(2 * y + 1) * cw + 2 * x], cur[(2 * y + 1) * cw + 2 * x + 1]);
      nxt[y * nw + x] = o;
      dst[(long long)(by * nw + y) * dw + bx * nw + x] = o;
    }
    __syncthreads();
    float4* t = cur; cur = nxt; nxt = t;
    cw = nw; lw = dw; lh >>= 1;
    doff += (long long)lw * lh;
  }
  __threadfence();
  if (tid == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1u) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  if (tid == 0) *ticket = 0u;
  __threadfence();
  mip_finish_chain(base, doff - (long long)lw * lh, lw, lh, 7, nlevels, sa, sb);
}
bool publish_all_applicable(const Geom& g, int nlevels) {
  return !g.has_up && !g.has_down && g.W >= 64 && g.rows >= 64 && g.W % 64 == 0 && g.rows % 64 == 0 && nlevels >= 7;
}
cudaError_t launch_publish_all(const Geom& g, const Planes& p, int side, float* chain, uint32_t* flow, int nlevels, unsigned int* ticket,
                               cudaStream_t st) {
  const size_t off = (size_t)TWS_HALO_ROWS * g.pitch;
  dim3 grid(g.W / 64, g.rows / 64);
  publish_all_kernel<<<grid, 256, 0, st>>>(g, p.h + off, p.d[side] + off, p.v + off, (float4*)chain, flow, g.W, g.rows, nlevels, ticket);
  return cudaGetLastError();
}

cudaError_t launch_mip_tail(float* base, int W, int H, int first, int last, cudaStream_t st) {
  if (first >= last) return cudaSuccess;
  mip_tail_kernel<<<1, 1024, 0, st>>>((float4*)base, W, H, first, last);
  return cudaGetLastError();
}
cudaError_t launch_mip_level(const float* src, int sw, int sh, float* dst, int dw, int dh, cudaStream_t st) {
  dim3 block(256), grid((dw + 255) / 256, dh < 2048 ? dh : 2048);
  mip_level_kernel<<<grid, block, 0, st>>>((const float4*)src, sw, sh, (float4*)dst, dw, dh);
  return cudaGetLastError();
}

// ---- fp64 volume: fixed-shape two-stage reduction (deterministic) ------------------------------
__global__ void __launch_bounds__(256) volume_kernel(Geom g, const float* __restrict__ d /* local row 0 */, double* partials) {
  double acc = 0.0;
  for (int r = blockIdx.x; r < g.rows; r += gridDim.x) {
    const float* row = d + (long long)r * g.pitch;
    for (int x = threadIdx.x; x < g.W; x += blockDim.x) acc += (double)row[x];
  }
  __shared__ double sm[256];
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[blockIdx.x] = sm[0];
}
cudaError_t launch_volume(const Geom& g, const float* d, double* partials, int nblocks, cudaStream_t st) {
  volume_kernel<<<nblocks, 256, 0, st>>>(g, d + (size_t)TWS_HALO_ROWS * g.pitch, partials);
  return cudaGetLastError();
}

// ---- outflow through the global edge (mass ledger on the open boundary) ------------------------------
__global__ void __launch_bounds__(256) boundary_outflow_kernel(Geom g, const float* __restrict__ Fxp, const float* __restrict__ Fxm,
                                                               const float* __restrict__ Fyp, const float* __restrict__ Fym, double* out) {
  double acc = 0.0;
  for (int r = threadIdx.x; r < g.rows; r += blockDim.x) {
    const long long o = (long long)r * g.pitch;
    acc += (double)Fxp[o + g.W - 1] + (double)Fxm[o];
  }
  if (g.row0 == 0)
    for (int x = threadIdx.x; x < g.W; x += blockDim.x) acc += (double)Fym[x];
  if (g.row0 + g.rows == g.Hg)
    for (int x = threadIdx.x; x < g.W; x += blockDim.x) acc += (double)Fyp[(long long)(g.rows - 1) * g.pitch + x];
  __shared__ double sm[256];
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sm[0];
}
cudaError_t launch_boundary_outflow(const Geom& g, const Planes& p, int side, double* out, cudaStream_t st) {
  const size_t off = (size_t)TWS_HALO_ROWS * g.pitch;
  boundary_outflow_kernel<<<1, 256, 0, st>>>(g, p.F[side][0] + off, p.F[side][1] + off, p.F[side][2] + off, p.F[side][3] + off, out);
  return cudaGetLastError();
}

// ---- strip exchange: push edge rows into the neighbours' halo rows over NVLink ----------------
__global__ void __launch_bounds__(256) row_copy_kernel(RowCopy a, RowCopy b, int ncopies) {
  const RowCopy& c = (blockIdx.z == 0) ? a : b;
  if ((int)blockIdx.z >= ncopies) return;
  const int n4 = c.pitch / 4;
  for (int pl = blockIdx.y; pl < c.nplanes; pl += gridDim.y) {
    const float4* src = reinterpret_cast<const float4*>(c.src[pl] + (long long)c.src_row * c.pitch);
    float4* dst = reinterpret_cast<float4*>(c.dst[pl] + (long long)c.dst_row * c.pitch);
    const long long total = (long long)n4 * c.nrows;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i];
  }
  __threadfence_system();
}
cudaError_t launch_row_copy(const RowCopy& a, const RowCopy& b, int ncopies, cudaStream_t st) {
  if (ncopies <= 0) return cudaSuccess;
  const int np = a.nplanes > b.nplanes ? a.nplanes : b.nplanes;
  dim3 grid(32, np, ncopies);
  row_copy_kernel<<<grid, 256, 0, st>>>(a, b, ncopies);
  return cudaGetLastError();
}

__global__ void post_flags_kernel(volatile uint32_t* up_flag, volatile uint32_t* down_flag, uint32_t value) {
  __threadfence_system();
  if (up_flag) *up_flag = value;
  if (down_flag) *down_flag = value;
  __threadfence_system();
}
cudaError_t launch_post_flags(volatile uint32_t* up_flag, volatile uint32_t* down_flag, uint32_t value, cudaStream_t st) {
  post_flags_kernel<<<1, 1, 0, st>>>(up_flag, down_flag, value);
  return cudaGetLastError();
}

// Spin until both neighbours have published epoch >= value.  Bounded: after ~20 s the
// kernel gives up and raises ctrl->error so a dead peer cannot hang the GPU.
__global__ void wait_flags_kernel(Control* ctrl, int wait_up, int wait_down, uint32_t value) {
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    const bool ok_up = !wait_up || (int32_t)(ctrl->flag_from_up - value) >= 0;
    const bool ok_dn = !wait_down || (int32_t)(ctrl->flag_from_down - value) >= 0;
    if (ok_up && ok_dn) break;
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 20000000000ull) { ctrl->error = 1u; break; }
    __nanosleep(200);
  }
  __threadfence_system();
}
cudaError_t launch_wait_flags(Control* ctrl, int wait_up, int wait_down, uint32_t value, cudaStream_t st) {
  wait_flags_kernel<<<1, 1, 0, st>>>(ctrl, wait_up, wait_down, value);
  return cudaGetLastError();
}

}  // namespace tws
