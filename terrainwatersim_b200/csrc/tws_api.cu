// tws_api.cu — host side of libtws.so: the C ABI of include/tws.h.
// Mirrors the simulation half of the reference's `Terrain` class (Terrain.cpp:150-277):
// parameter derivation, create/reset, inject, the step loop and the frame accumulator.
// No arithmetic on simulation state happens on the host; there is no CPU fallback.
#include "tws_internal.h"

#include <unistd.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

using namespace tws;

namespace {

thread_local std::string g_create_error;

struct Peer {
  bool present = false;
  bool ipc = false;          // opened with cudaIpcOpenMemHandle (must be closed)
  uint8_t* slab = nullptr;   // peer slab base, addressable from this device
  int rows = 0;              // peer's own rows
  int device = -1;           // CUDA device the peer's slab lives on
  Planes planes{};
  Control* ctrl = nullptr;
};

}  // namespace

struct tws_sim {
  tws_params prm{};
  Geom geom{};
  StepConsts consts{};
  double step_length = 0.0;          // m_simulationStepLength (ezTime, double seconds)
  double accumulator = 0.0;          // m_timeSinceLastSimulationStep
  uint8_t* slab = nullptr;
  size_t slab_bytes = 0;
  size_t plane_elems = 0;
  Planes planes{};
  Control* ctrl = nullptr;
  int cur = 0;                       // ping-pong side holding the current d / F
  int K = 1;                         // steps per launch
  TmaSet tma[TWS_MAX_TB + 1][2];     // [k][side]
  bool tma_ready[TWS_MAX_TB + 1] = {};
  TmaSet tma_stream[2];              // row descriptors of the streaming pipeline, [side]
  bool tma_stream_ready = false;
  cudaStream_t st_main = nullptr, st_edge = nullptr;
  cudaStream_t st_h2d = nullptr, st_d2h = nullptr;      // tws_step_host: upload / readback streams (created on first use)
  std::vector<cudaEvent_t> band_ev;                     // tws_step_host: [2b] band b uploaded, [2b+1] band b computed
  // Timing of step batches, double buffered like gl::TimerQuery (TimerQuery.cpp:29-72): batch i records into pair i & 1,
  // so the pair of the batch BEFORE the most recent one can be read without waiting for the GPU.
  cudaEvent_t ev_pair[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [pair][0 = start, 1 = stop]
  uint64_t batches_timed = 0;        // batches recorded so far; the most recent one used pair (batches_timed - 1) & 1
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr, ev_main = nullptr, ev_edge = nullptr;   // ev_start / ev_stop: aliases of the current pair
  bool timed = false;
  uint64_t launches = 0;
  uint32_t epoch = 0;                // exchange blocks completed (same on all strips)
  int res_cfg = -1;                  // resident backend: block shape
  bool auto_picked = false;          // the backend was chosen by TWS_BACKEND_AUTO
  bool res_auto = false;             // TWS_BACKEND_AUTO on a grid that fits on chip: frames of >= kResidentMinSteps steps go resident
  uint32_t res_epoch = 0;            // resident backend: steps run so far (the tags of the rim exchange count them)
  void* res_mailbox = nullptr;       // resident backend: the rim exchange's mailbox (device)
  // A brush injected on a whole grid that runs the tile kernel is not launched at once: the next single-step launch applies it
  // while loading the depth (the reference's frame — brush, one step, mip chain — is then two launches).  Anything else that
  // reads or writes the state first materialises it with the ordinary brush kernel (flush_brush).
  BrushArgs pending_brush{};
  Peer up, down;
  bool connected = false;
  float* white_dev = nullptr;        // 4096-entry noise table
  double* partials = nullptr;        // volume partial sums (device)
  float* staging = nullptr;          // AoS staging for flux / terrain-info transfers
  size_t staging_bytes = 0;
  void* packed_info = nullptr;       // publish buffers
  void* packed_flow = nullptr;
  int published_levels = 0;          // mip levels of TerrainInfo currently valid in packed_info
  cudaGraphicsResource* gl_info = nullptr;   // renderer textures registered with tws_gl_register
  cudaGraphicsResource* gl_flow = nullptr;
  // Captured step batches (frame scheduler): one executable graph per (n, ping-pong side).
  struct StepGraph { int n; int cur; int cur_after; uint64_t launches; cudaGraphExec_t exec; };
  std::vector<StepGraph> graphs;
  bool use_graphs = true;
  uint64_t graph_replays = 0;
  std::string err;
};

namespace {

constexpr int kVolumeBlocks = 1024;

tws_status fail(tws_sim* s, tws_status code, const std::string& msg) {
  if (s) s->err = msg; else g_create_error = msg;
  return code;
}
tws_status cuda_fail(tws_sim* s, cudaError_t e, const char* what) {
  return fail(s, e == cudaErrorMemoryAllocation ? TWS_ERR_NOMEM : TWS_ERR_CUDA,
              std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}
#define TWS_CUDA(s, call)                                        \
  do {                                                           \
    cudaError_t e__ = (call);                                    \
    if (e__ != cudaSuccess) return cuda_fail((s), e__, #call);   \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Slab layout is a pure function of (width, rows) so a peer can address our planes.
void layout_planes(uint8_t* base, int pitch, int rows, Planes* p, size_t* plane_elems, size_t* total) {
  const size_t elems = (size_t)pitch * (size_t)(rows + 2 * TWS_HALO_ROWS);
  const size_t pb = align_up(elems * sizeof(float), 1024);
  size_t off = TWS_CTRL_BYTES;
  auto take = [&]() { uint8_t* q = base ? base + off : nullptr; off += pb; return q; };
  // the depth and the four flux planes of one ping-pong side are consecutive with a uniform stride,
  // so one 3-D TMA descriptor (x, row, plane) fetches a row of all five in a single operation
  p->h = (float*)take();
  for (int s = 0; s < 2; ++s) {
    p->d[s] = (float*)take();
    for (int i = 0; i < 4; ++i) p->F[s][i] = (float*)take();
  }
  p->v = (uint32_t*)take();
  if (plane_elems) *plane_elems = elems;
  if (total) *total = off;
}

int pitch_for(int width) { return (int)align_up((size_t)width, 64); }

// Terrain.cpp:175-198 — see tws_step_constants in tws.h.
void drop_graphs(tws_sim* s);
inline bool is_stream_backend(int b) { return b == TWS_BACKEND_STREAM_TB || b == TWS_BACKEND_BAND_TB; }
inline int stream_impl(const tws_sim* s) { return s->prm.backend == TWS_BACKEND_BAND_TB ? 1 : 0; }
void derive_constants(tws_sim* s) {
  drop_graphs(s);                                       // captured batches bake the constants in
  const tws_params& p = s->prm;
  s->step_length = (double)(1.0f / p.steps_per_second);                   // ezTime::Seconds(1.0f / sps)
  const float cell = p.world_size / (float)p.width;                       // float / uint -> float
  s->consts.friction = powf(p.flow_damping, (float)s->step_length);       // ezMath::Pow == powf
  s->consts.accel = (float)(s->step_length * p.flow_acceleration * cell); // double product
  s->consts.area_inv = (float)(s->step_length / (cell * cell));
  s->consts.rain_step = (float)(s->step_length * p.rain_rate);            // EXT
  s->consts.evap_step = (float)(s->step_length * p.evaporation_rate);     // EXT
  s->consts.ext_sources = (s->consts.rain_step != 0.0f || s->consts.evap_step != 0.0f) ? 1 : 0;
  s->consts.closed = (p.boundary == TWS_BOUNDARY_CLOSED) ? 1 : 0;
  s->consts.ledger = (s->ctrl != nullptr && !s->consts.closed) ? &s->ctrl->outflow_acc : nullptr;   // EXT mass ledger
  s->consts.ledger_src = (s->ctrl != nullptr && s->consts.ext_sources) ? &s->ctrl->source_acc : nullptr;
}

bool bad_float(float v) { return !(v == v) || std::isinf(v); }

tws_status ensure_tma(tws_sim* s, int k) {
  if (s->tma_ready[k]) return TWS_OK;
  for (int side = 0; side < 2; ++side) {
    std::string e;
    cudaError_t r = fused_build_tma(k, s->geom, s->planes, side, &s->tma[k][side], &e);
    if (r != cudaSuccess) return fail(s, TWS_ERR_CUDA, e.empty() ? std::string("building TMA descriptors failed") : e);
  }
  s->tma_ready[k] = true;
  return TWS_OK;
}

tws_status ensure_tma_stream(tws_sim* s) {
  if (s->tma_stream_ready) return TWS_OK;
  for (int side = 0; side < 2; ++side) {
    std::string e;
    cudaError_t r = stream_build_tma(s->geom, s->planes, side, &s->tma_stream[side], &e);
    if (r != cudaSuccess) return fail(s, TWS_ERR_CUDA, e.empty() ? std::string("building TMA descriptors failed") : e);
  }
  s->tma_stream_ready = true;
  return TWS_OK;
}

tws_status ensure_staging(tws_sim* s, size_t bytes) {
  if (s->staging_bytes >= bytes) return TWS_OK;
  if (s->staging) cudaFree(s->staging);
  s->staging = nullptr; s->staging_bytes = 0;
  TWS_CUDA(s, cudaMalloc(&s->staging, bytes));
  s->staging_bytes = bytes;
  return TWS_OK;
}

size_t field_elem_bytes(tws_field f) {
  switch (f) {
    case TWS_FIELD_TERRAIN: case TWS_FIELD_WATER: return 4;
    case TWS_FIELD_FLUX: case TWS_FIELD_TERRAIN_INFO: return 16;
    case TWS_FIELD_VELOCITY: return 4;
  }
  return 0;
}

// Push `nrows` edge rows of the given side (d + 4 flux planes, optionally h) into the
// neighbours' halo rows.  Returns the number of kernels launched.
tws_status push_edges(tws_sim* s, int side, bool with_h, cudaStream_t st) {
  RowCopy c[2]; int n = 0;
  const int P = TWS_HALO_ROWS;
  auto fill = [&](const Peer& peer, bool to_up) {
    RowCopy r{};
    int k = 0;
    r.src[k] = s->planes.d[side]; r.dst[k] = peer.planes.d[side]; ++k;
    for (int i = 0; i < 4; ++i) { r.src[k] = s->planes.F[side][i]; r.dst[k] = peer.planes.F[side][i]; ++k; }
    if (with_h) { r.src[k] = s->planes.h; r.dst[k] = peer.planes.h; ++k; }
    r.nplanes = k; r.nrows = P; r.pitch = s->geom.pitch;
    if (to_up) { r.src_row = TWS_HALO_ROWS; r.dst_row = TWS_HALO_ROWS + peer.rows; }          // my top rows -> its bottom halo
    else { r.src_row = TWS_HALO_ROWS + s->geom.rows - P; r.dst_row = TWS_HALO_ROWS - P; }      // my bottom rows -> its top halo
    return r;
  };
  if (s->up.present) c[n++] = fill(s->up, true);
  if (s->down.present) c[n++] = fill(s->down, false);
  if (n == 0) return TWS_OK;
  TWS_CUDA(s, launch_row_copy(c[0], c[n - 1], n, st));
  s->launches += 1;
  return TWS_OK;
}

tws_status post_and_count(tws_sim* s, cudaStream_t st) {
  s->epoch += 1;
  volatile uint32_t* uf = s->up.present ? &s->up.ctrl->flag_from_down : nullptr;
  volatile uint32_t* df = s->down.present ? &s->down.ctrl->flag_from_up : nullptr;
  if (uf || df) {
    TWS_CUDA(s, launch_post_flags(uf, df, s->epoch, st));
    s->launches += 1;
  }
  return TWS_OK;
}

// One block of k steps of the row-streaming pipeline.  A strip runs the rows that neither depend on
// a halo nor are pushed to a neighbour (all but the outer TWS_HALO_ROWS) on the main stream and the
// two edge bands on the edge stream behind the neighbours' flags, exactly like the tile engine.
// One block of k steps of a strip with the band kernel: ONE launch.  The edge bands are the first pieces of its
// work list; their warps wait for the neighbours' flags, push every output row into the neighbours' halo rows as they
// store it and post the new epoch when the last edge piece is done — all while the other warp groups of the same
// launch work on interior pieces (BandEdge, band_kernels.cu).  No edge stream, no copy kernel, no flag kernels.
tws_status run_block_band_strip(tws_sim* s, int k) {
  const Geom& g = s->geom;
  const int src = s->cur, dst = 1 - src;
  const int P = TWS_HALO_ROWS;
  int e_top = g.has_up ? P : 0, e_bot = g.has_down ? P : 0;
  if (e_top + e_bot >= g.rows) { e_top = g.rows; e_bot = 0; }
  BandEdge ed{};
  ed.wait_up = g.has_up ? &s->ctrl->flag_from_up : nullptr;
  ed.wait_down = g.has_down ? &s->ctrl->flag_from_down : nullptr;
  ed.wait_value = s->epoch;
  ed.error = &s->ctrl->error;
  s->epoch += 1;
  ed.post_up = s->up.present ? &s->up.ctrl->flag_from_down : nullptr;
  ed.post_down = s->down.present ? &s->down.ctrl->flag_from_up : nullptr;
  ed.post_value = s->epoch;
  ed.up_end = s->up.present ? std::min(P, g.rows) : 0;
  ed.down_begin = s->down.present ? std::max(0, g.rows - P) : INT32_MAX;
  auto peer_planes = [&](const Peer& peer, long long row_shift, float* (&out)[5]) {
    // own row y (element offset y * pitch from local row 0) -> peer plane row y + row_shift
    out[0] = peer.planes.d[dst] + row_shift * g.pitch;
    for (int i = 0; i < 4; ++i) out[1 + i] = peer.planes.F[dst][i] + row_shift * g.pitch;
  };
  if (s->up.present) peer_planes(s->up, (long long)TWS_HALO_ROWS + s->up.rows, ed.up);          // my top rows -> its bottom halo
  if (s->down.present) peer_planes(s->down, (long long)TWS_HALO_ROWS - g.rows, ed.down);        // my bottom rows -> its top halo
  TWS_CUDA(s, cudaStreamWaitEvent(s->st_main, s->ev_edge, 0));          // a tws_halo_refresh on the edge stream comes first
  TWS_CUDA(s, launch_band(k, g, s->planes, s->tma_stream[src], src, s->consts, 0, g.rows, s->st_main, s->ctrl->sched_main, 0, &ed, e_top, e_bot));
  s->launches += 1;
  TWS_CUDA(s, cudaEventRecord(s->ev_main, s->st_main));
  s->cur = dst;
  return TWS_OK;
}

// One block of k steps of the row-streaming pipeline.  A strip runs the rows that neither depend on
// a halo nor are pushed to a neighbour (all but the outer TWS_HALO_ROWS) on the main stream and the
// two edge bands on the edge stream behind the neighbours' flags, exactly like the tile engine
// (ring kernel; the band kernel does the whole block in one launch, see above).
tws_status run_block_stream(tws_sim* s, int k) {
  const Geom& g = s->geom;
  const bool strip = g.has_up || g.has_down;
  const int src = s->cur;
  tws_status r = ensure_tma_stream(s);
  if (r) return r;
  if (!strip) {
    TWS_CUDA(s, launch_stream(k, g, s->planes, s->tma_stream[src], src, s->consts, 0, g.rows, s->st_main, stream_impl(s), s->ctrl->sched_main));
    s->launches += 1;
    s->cur = 1 - src;
    return TWS_OK;
  }
  if (stream_impl(s) == 1) {
    // In the one-launch form every warp group that holds an edge piece polls the neighbour's flag.  When the neighbour
    // shares THIS GPU (a test configuration) and the grid is so wide that the edge pieces could occupy every resident
    // group, the neighbour's launch might never get an SM: such strips use the two-stream form below instead.
    const bool shared_gpu = (s->up.present && s->up.device == s->prm.device) || (s->down.present && s->down.device == s->prm.device);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->prm.device);
    const int nstrips_x = (g.W + 111) / 112;             // 112 output columns per column strip of the band kernel
    if (!(shared_gpu && 2 * nstrips_x >= sms)) return run_block_band_strip(s, k);
  }
  int e_top = g.has_up ? TWS_HALO_ROWS : 0, e_bot = g.has_down ? TWS_HALO_ROWS : 0;
  if (e_top + e_bot >= g.rows) { e_top = g.rows; e_bot = 0; }
  const int i0 = e_top, i1 = g.rows - e_bot;
  TWS_CUDA(s, cudaStreamWaitEvent(s->st_main, s->ev_edge, 0));
  if (i1 > i0) {
    TWS_CUDA(s, launch_stream(k, g, s->planes, s->tma_stream[src], src, s->consts, i0, i1, s->st_main, stream_impl(s), s->ctrl->sched_main));
    s->launches += 1;
  }
  TWS_CUDA(s, cudaStreamWaitEvent(s->st_edge, s->ev_main, 0));
  TWS_CUDA(s, launch_wait_flags(s->ctrl, g.has_up, g.has_down, s->epoch, s->st_edge)); s->launches++;
  if (e_top > 0) { TWS_CUDA(s, launch_stream(k, g, s->planes, s->tma_stream[src], src, s->consts, 0, e_top, s->st_edge, stream_impl(s), s->ctrl->sched_edge)); s->launches++; }
  if (e_bot > 0) { TWS_CUDA(s, launch_stream(k, g, s->planes, s->tma_stream[src], src, s->consts, i1, g.rows, s->st_edge, stream_impl(s), s->ctrl->sched_edge)); s->launches++; }
  r = push_edges(s, 1 - src, false, s->st_edge); if (r) return r;
  r = post_and_count(s, s->st_edge); if (r) return r;
  TWS_CUDA(s, cudaEventRecord(s->ev_edge, s->st_edge));
  TWS_CUDA(s, cudaEventRecord(s->ev_main, s->st_main));
  s->cur = 1 - src;
  return TWS_OK;
}

// Materialise a pending brush (see tws_sim::pending_brush) with the brush kernel.
tws_status flush_brush(tws_sim* s) {
  if (!s->pending_brush.active) return TWS_OK;
  const BrushArgs b = s->pending_brush;
  s->pending_brush.active = 0;
  int launched = 0;
  TWS_CUDA(s, launch_brush(s->geom, s->planes.d[s->cur] + (size_t)TWS_HALO_ROWS * s->geom.pitch, b.cx, b.cy, b.intensity, b.size_sq, s->st_main,
                           &launched));
  s->launches += launched;
  return TWS_OK;
}

// One block of k fused steps (or one unfused step) including the strip exchange.
tws_status run_block(tws_sim* s, int k) {
  const Geom& g = s->geom;
  const bool strip = g.has_up || g.has_down;
  const int src = s->cur;
  if (s->prm.backend == TWS_BACKEND_UNFUSED) {
    TWS_CUDA(s, launch_unfused_update(g, s->planes, src, s->consts, 0, g.rows, s->st_main));
    TWS_CUDA(s, launch_unfused_apply(g, s->planes, src, s->consts, 0, g.rows, s->st_main));
    s->launches += 2;
    return TWS_OK;
  }
  if (is_stream_backend(s->prm.backend)) return run_block_stream(s, k);
  if (s->prm.backend == TWS_BACKEND_RESIDENT) {           // k: any number of steps, one launch
    if (!s->res_mailbox) {
      TWS_CUDA(s, cudaMalloc(&s->res_mailbox, resident_mailbox_bytes(g)));
      TWS_CUDA(s, cudaMemsetAsync(s->res_mailbox, 0, resident_mailbox_bytes(g), s->st_main));
    }
    const BrushArgs brush = s->pending_brush;           // folded into the block load
    s->pending_brush.active = 0;
    TWS_CUDA(s, launch_resident(s->res_cfg, g, s->planes, s->consts, src, k, s->res_mailbox, s->res_epoch, &s->ctrl->error, s->st_main,
                                brush.active ? &brush : nullptr));
    s->res_epoch += (uint32_t)k;
    s->launches += 1;
    s->cur = (src + k) & 1;
    return TWS_OK;
  }
  tws_status r = ensure_tma(s, k);
  if (r) return r;
  const int tiles = fused_tile_rows(k, g.rows);
  if (!strip) {
    const BrushArgs brush = s->pending_brush;           // folded into this launch's loads (run_steps flushed it otherwise)
    s->pending_brush.active = 0;
    TWS_CUDA(s, launch_fused(k, g, s->planes, s->tma[k][src], src, s->consts, 0, tiles, s->st_main, brush.active ? &brush : nullptr));
    s->launches += 1;
    s->cur = 1 - src;
    return TWS_OK;
  }
  // Strip: tile rows whose outputs depend on a halo, or that the neighbour needs pushed,
  // run on the edge stream after the neighbours' flags; the interior does not wait.
  const int oy = fused_out_rows_per_tile(k);
  const int need = TWS_HALO_ROWS;                       // rows pushed / rows depending on the halo (>= 2k)
  int e_top = g.has_up ? (need + oy - 1) / oy : 0;
  int e_bot = g.has_down ? (need + oy - 1) / oy : 0;
  if (e_top + e_bot >= tiles) { e_top = tiles; e_bot = 0; }
  const int i0 = e_top, i1 = tiles - e_bot;
  // interior (main stream) — ordered after the previous block's edge work
  TWS_CUDA(s, cudaStreamWaitEvent(s->st_main, s->ev_edge, 0));
  if (i1 > i0) {
    TWS_CUDA(s, launch_fused(k, g, s->planes, s->tma[k][src], src, s->consts, i0, i1, s->st_main));
    s->launches += 1;
  }
  // edges (edge stream) — ordered after the previous block's interior
  TWS_CUDA(s, cudaStreamWaitEvent(s->st_edge, s->ev_main, 0));
  TWS_CUDA(s, launch_wait_flags(s->ctrl, g.has_up, g.has_down, s->epoch, s->st_edge)); s->launches++;
  if (e_top > 0) { TWS_CUDA(s, launch_fused(k, g, s->planes, s->tma[k][src], src, s->consts, 0, e_top, s->st_edge)); s->launches++; }
  if (e_bot > 0) { TWS_CUDA(s, launch_fused(k, g, s->planes, s->tma[k][src], src, s->consts, i1, tiles, s->st_edge)); s->launches++; }
  r = push_edges(s, 1 - src, false, s->st_edge); if (r) return r;
  r = post_and_count(s, s->st_edge); if (r) return r;
  TWS_CUDA(s, cudaEventRecord(s->ev_edge, s->st_edge));
  TWS_CUDA(s, cudaEventRecord(s->ev_main, s->st_main));
  s->cur = 1 - src;
  return TWS_OK;
}

void drop_graphs(tws_sim* s) {
  for (auto& e : s->graphs) cudaGraphExecDestroy(e.exec);
  s->graphs.clear();
}

// A batch of n steps as one CUDA graph: the reference runs up to 10 steps per frame
// (Terrain.cpp:247,253-265) with ~12 GL calls each; on small grids a step is a few microseconds of
// GPU time and the launch path dominates, so whole-grid sims replay a captured batch instead of
// re-launching its kernels one by one.  Captured once per (n, ping-pong side); parameter changes
// drop the cache (the per-step constants are baked into the kernel arguments).
tws_status run_batch_graph(tws_sim* s, int n, int K, bool* done) {
  *done = false;
  for (const auto& e : s->graphs)
    if (e.n == n && e.cur == s->cur) {
      TWS_CUDA(s, cudaGraphLaunch(e.exec, s->st_main));
      s->cur = e.cur_after; s->launches += e.launches; s->graph_replays += 1;
      *done = true;
      return TWS_OK;
    }
  // make sure nothing but kernel launches happens inside the capture
  if (is_stream_backend(s->prm.backend)) {
    tws_status r = ensure_tma_stream(s); if (r) return r;
    for (int k : {K, n % K}) if (k > 0) TWS_CUDA(s, launch_stream(k, s->geom, s->planes, s->tma_stream[0], 0, s->consts, 0, 0, s->st_main, stream_impl(s), s->ctrl->sched_main));
  } else if (s->prm.backend != TWS_BACKEND_UNFUSED) {
    for (int k : {K, n % K}) if (k > 0) {
      tws_status r = ensure_tma(s, k); if (r) return r;
      TWS_CUDA(s, launch_fused(k, s->geom, s->planes, s->tma[k][0], 0, s->consts, 0, 0, s->st_main));
    }
  }
  if (s->graphs.size() >= 16) drop_graphs(s);
  tws_sim::StepGraph e{n, s->cur, 0, 0, nullptr};
  const uint64_t l0 = s->launches;
  TWS_CUDA(s, cudaStreamBeginCapture(s->st_main, cudaStreamCaptureModeThreadLocal));
  tws_status r = TWS_OK;
  for (int left = n; left > 0 && r == TWS_OK;) {
    const int k = std::min(left, K);
    r = run_block(s, k);
    left -= k;
  }
  cudaGraph_t graph = nullptr;
  cudaError_t ce = cudaStreamEndCapture(s->st_main, &graph);
  e.cur_after = s->cur; e.launches = s->launches - l0;
  s->cur = e.cur; s->launches = l0;                      // nothing has run yet
  if (r) { if (graph) cudaGraphDestroy(graph); return r; }
  if (ce != cudaSuccess) return cuda_fail(s, ce, "cudaStreamEndCapture");
  ce = cudaGraphInstantiate(&e.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ce != cudaSuccess) return cuda_fail(s, ce, "cudaGraphInstantiate");
  s->graphs.push_back(e);
  TWS_CUDA(s, cudaGraphLaunch(e.exec, s->st_main));
  s->cur = e.cur_after; s->launches += e.launches; s->graph_replays += 1;
  *done = true;
  return TWS_OK;
}

// the next batch records into the other event pair
void next_timing_pair(tws_sim* s) {
  const int p = (int)(s->batches_timed & 1u);
  s->ev_start = s->ev_pair[p][0]; s->ev_stop = s->ev_pair[p][1];
  s->batches_timed += 1;
}

// Measured (scripts/resident_perf.py, profiles/r02_resident_frames.log): from 3 steps per call on, one resident launch matches
// or beats the captured batch of tile-kernel launches on grids that fill the SMs (1024^2: 33.1 vs 33.7 us at 3, 40 vs 42 at 4,
// 80 vs 102 at 10 steps; 512^2: 16.7 vs 19.2 at 3, 37 vs 46 at 10; 256^2: 13.3 vs 18.1 at 3, 27 vs 48 at 10); at 1 and 2 steps
// its block load / store is not amortised (1024^2: 26.9 vs 23.8 us at 2) — except on grids so small that a launch is all
// latency (256^2: 8.8 vs 11.5 us at 1, 11.0 vs 11.8 at 2).
constexpr int kResidentMinSteps = 3;
constexpr long long kResidentAlwaysCells = 100 * 1000;
inline int resident_min_steps(const tws_sim* s) {
  return (long long)s->geom.W * s->geom.rows <= kResidentAlwaysCells ? 1 : kResidentMinSteps;
}

tws_status run_steps(tws_sim* s, int n) {
  const Geom& g = s->geom;
  const bool strip = g.has_up || g.has_down;
  if (strip && !s->connected) return fail(s, TWS_ERR_STATE, "strip sim stepped before tws_halo_connect");
  next_timing_pair(s);
  TWS_CUDA(s, cudaEventRecord(s->ev_start, s->st_main));
  if (s->pending_brush.active && n > 0) {
    // only a direct tile-kernel or resident launch takes the brush along (captured batches bake their arguments in)
    const bool tile = s->prm.backend == TWS_BACKEND_FUSED || s->prm.backend == TWS_BACKEND_FUSED_TB;
    const bool resident = s->prm.backend == TWS_BACKEND_RESIDENT || (s->res_auto && n >= resident_min_steps(s));
    const bool batch = !resident && s->use_graphs && n >= 2 && n <= 64;
    if (strip || !(tile || resident) || batch) { tws_status r = flush_brush(s); if (r) return r; }
  }
  const int K = s->prm.backend == TWS_BACKEND_RESIDENT ? (1 << 20)
              : (s->prm.backend == TWS_BACKEND_FUSED_TB || is_stream_backend(s->prm.backend)) ? s->K : 1;
  if (s->res_auto && n >= resident_min_steps(s)) {
    const int keep = s->prm.backend;
    s->prm.backend = TWS_BACKEND_RESIDENT;
    const tws_status r = run_block(s, n);
    s->prm.backend = keep;
    if (r) return r;
    n = 0;
  }
  if (!strip && s->use_graphs && n >= 2 && n <= 64 && s->prm.backend != TWS_BACKEND_RESIDENT) {
    bool done = false;
    tws_status r = run_batch_graph(s, n, K, &done);
    if (r) return r;
    if (done) n = 0;
  }
  while (n > 0) {
    const int k = std::min(n, K);
    tws_status r = run_block(s, k);
    if (r) return r;
    n -= k;
  }
  if (strip) TWS_CUDA(s, cudaStreamWaitEvent(s->st_main, s->ev_edge, 0));
  TWS_CUDA(s, cudaEventRecord(s->ev_stop, s->st_main));
  s->timed = true;
  return TWS_OK;
}

tws_status sync_all(tws_sim* s) {
  { tws_status fr = flush_brush(s); if (fr) return fr; }         // whoever synchronises is about to look at the state
  TWS_CUDA(s, cudaStreamSynchronize(s->st_edge));
  TWS_CUDA(s, cudaStreamSynchronize(s->st_main));
  uint32_t e = 0;
  TWS_CUDA(s, cudaMemcpy(&e, &s->ctrl->error, sizeof(e), cudaMemcpyDeviceToHost));
  if (e) return fail(s, TWS_ERR_STATE, "halo exchange timed out waiting for a neighbouring strip (or, resident backend, a neighbouring block)");
  return TWS_OK;
}

void close_peer(Peer& p) {
  if (p.present && p.ipc && p.slab) cudaIpcCloseMemHandle(p.slab);
  p = Peer{};
}

}  // namespace

extern "C" {

const char* tws_version(void) { return "tws-b200 0.1 (sm_100a)"; }
int32_t tws_abi_version(void) { return TWS_ABI_VERSION; }

void tws_default_params(tws_params* p) {
  if (!p) return;
  std::memset(p, 0, sizeof(*p));
  p->size = sizeof(tws_params);
  p->width = 1024; p->height = 1024; p->row_begin = 0; p->row_end = 1024;      // Terrain.cpp:23
  p->world_size = 1024.0f;                                                      // Terrain.cpp:22
  p->steps_per_second = 60.0f; p->flow_damping = 0.98f; p->flow_acceleration = 10.0f;   // Terrain.cpp:28-30
  p->boundary = TWS_BOUNDARY_REFERENCE_OPEN;
  p->backend = TWS_BACKEND_AUTO;
  p->temporal_block = 1;
  p->device = 0;
}

const char* tws_last_error(const tws_sim* s) { return s ? s->err.c_str() : g_create_error.c_str(); }

tws_status tws_create(const tws_params* p, tws_sim** out) {
  if (out) *out = nullptr;
  if (!p || !out) return fail(nullptr, TWS_ERR_INVALID, "tws_create: null argument");
  if (p->size != sizeof(tws_params)) return fail(nullptr, TWS_ERR_INVALID, "tws_create: params.size does not match this ABI");
  if (p->width < 1 || p->height < 1) return fail(nullptr, TWS_ERR_INVALID, "tws_create: width/height must be >= 1");
  if (p->row_begin < 0 || p->row_end > p->height || p->row_end <= p->row_begin)
    return fail(nullptr, TWS_ERR_INVALID, "tws_create: bad strip rows");
  const bool strip = p->row_begin > 0 || p->row_end < p->height;
  if (strip && p->backend == TWS_BACKEND_UNFUSED)
    return fail(nullptr, TWS_ERR_UNSUPPORTED, "tws_create: the unfused baseline updates in place and cannot run on a strip; use a fused backend");
  if (strip && p->row_end - p->row_begin < TWS_HALO_ROWS)
    return fail(nullptr, TWS_ERR_INVALID, "tws_create: a strip needs at least 8 rows");
  if (bad_float(p->world_size) || !(p->world_size > 0.0f)) return fail(nullptr, TWS_ERR_INVALID, "tws_create: world_size must be > 0");
  if (bad_float(p->steps_per_second) || !(p->steps_per_second > 0.0f)) return fail(nullptr, TWS_ERR_INVALID, "tws_create: steps_per_second must be > 0");
  if (bad_float(p->flow_damping) || p->flow_damping < 0.0f) return fail(nullptr, TWS_ERR_INVALID, "tws_create: flow_damping must be >= 0");
  if (bad_float(p->flow_acceleration) || p->flow_acceleration < 0.0f) return fail(nullptr, TWS_ERR_INVALID, "tws_create: flow_acceleration must be >= 0");
  if (bad_float(p->rain_rate) || bad_float(p->evaporation_rate) || p->rain_rate < 0.0f || p->evaporation_rate < 0.0f)
    return fail(nullptr, TWS_ERR_INVALID, "tws_create: rain/evaporation must be >= 0");
  if (p->backend < TWS_BACKEND_AUTO || p->backend > TWS_BACKEND_RESIDENT) return fail(nullptr, TWS_ERR_INVALID, "tws_create: unknown backend");
  if (strip && p->backend == TWS_BACKEND_RESIDENT)
    return fail(nullptr, TWS_ERR_UNSUPPORTED, "tws_create: the resident backend runs whole grids only");
  if (p->boundary != TWS_BOUNDARY_REFERENCE_OPEN && p->boundary != TWS_BOUNDARY_CLOSED) return fail(nullptr, TWS_ERR_INVALID, "tws_create: unknown boundary");
  if ((p->backend == TWS_BACKEND_FUSED_TB || is_stream_backend(p->backend)) && (p->temporal_block < 1 || p->temporal_block > TWS_MAX_TB))
    return fail(nullptr, TWS_ERR_INVALID, "tws_create: temporal_block must be 1..4");

  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, TWS_ERR_CUDA, std::string("tws_create: no CUDA device (") + cudaGetErrorString(e) + "); libtws has no CPU path");
  if (p->device < 0 || p->device >= ndev) return fail(nullptr, TWS_ERR_INVALID, "tws_create: device ordinal out of range");
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, p->device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
  if (prop.major != 10) return fail(nullptr, TWS_ERR_CUDA, "tws_create: device is not sm_100 (kernels are built for sm_100a only)");

  tws_sim* s = new (std::nothrow) tws_sim;
  if (!s) return fail(nullptr, TWS_ERR_NOMEM, "tws_create: out of host memory");
  s->prm = *p;
  s->auto_picked = p->backend == TWS_BACKEND_AUTO;
  if (p->backend == TWS_BACKEND_AUTO) {
    // Measured on B200 (scripts/crossover_perf.py, profiles/r01_crossover_tile_vs_band.log): small grids are latency
    // bound and the tile kernel keeps more of the SM busy (209 vs 206 Gcell/s at 3072^2, 177 vs 165 at 2048^2); from
    // 4096^2 up the band kernel wins (255 vs 218, 313 vs 235 at 8192^2).  Strips count their own cells.
    const long long cells = (long long)p->width * (p->row_end - p->row_begin);
    const bool big = cells >= 12LL * 1000 * 1000;
    s->prm.backend = big ? TWS_BACKEND_BAND_TB : TWS_BACKEND_FUSED_TB;
    s->prm.temporal_block = big ? 4 : 2;
    p = &s->prm;
  }
  s->K = (p->backend == TWS_BACKEND_FUSED_TB || is_stream_backend(p->backend)) ? p->temporal_block : 1;
  if (const char* gv = getenv("TWS_GRAPHS")) s->use_graphs = gv[0] != '0';
  DeviceGuard guard(p->device);
  Geom& g = s->geom;
  g.W = p->width; g.Hg = p->height; g.row0 = p->row_begin; g.rows = p->row_end - p->row_begin;
  g.pitch = pitch_for(p->width);
  g.has_up = p->row_begin > 0; g.has_down = p->row_end < p->height;
  derive_constants(s);
  if (p->backend == TWS_BACKEND_FUSED_TB && s->auto_picked && !strip) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device);
    s->res_cfg = resident_config(g);
    s->res_auto = s->res_cfg >= 0 && 4 * resident_blocks(g, s->res_cfg) >= 3 * sms;     // most SMs get a block
  }
  if (p->backend == TWS_BACKEND_RESIDENT) {
    s->res_cfg = resident_config(g);
    if (s->res_cfg < 0) {
      g_create_error = "tws_create: the grid does not fit in the SMs' shared memory (resident backend: up to ~1 M cells)";
      delete s;
      return TWS_ERR_UNSUPPORTED;
    }
  }
  layout_planes(nullptr, g.pitch, g.rows, &s->planes, &s->plane_elems, &s->slab_bytes);
  tws_status rc = TWS_OK;
  auto bail = [&](tws_status code, const std::string& msg) { g_create_error = msg; tws_destroy(s); return code; };
  if ((e = cudaMalloc(&s->slab, s->slab_bytes)) != cudaSuccess)
    return bail(e == cudaErrorMemoryAllocation ? TWS_ERR_NOMEM : TWS_ERR_CUDA, std::string("tws_create: cudaMalloc of the state slab failed: ") + cudaGetErrorString(e));
  layout_planes(s->slab, g.pitch, g.rows, &s->planes, nullptr, nullptr);
  s->ctrl = (Control*)s->slab;
  derive_constants(s);                                  // now that the control block exists: the ledger pointer
  if ((e = cudaStreamCreateWithFlags(&s->st_main, cudaStreamNonBlocking)) != cudaSuccess) return bail(TWS_ERR_CUDA, cudaGetErrorString(e));
  int lo_prio = 0, hi_prio = 0;
  cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio);
  if ((e = cudaStreamCreateWithPriority(&s->st_edge, cudaStreamNonBlocking, hi_prio)) != cudaSuccess) return bail(TWS_ERR_CUDA, cudaGetErrorString(e));
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j)
      if ((e = cudaEventCreate(&s->ev_pair[i][j])) != cudaSuccess) return bail(TWS_ERR_CUDA, cudaGetErrorString(e));
  s->ev_start = s->ev_pair[0][0]; s->ev_stop = s->ev_pair[0][1];
  if ((e = cudaEventCreateWithFlags(&s->ev_main, cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&s->ev_edge, cudaEventDisableTiming)) != cudaSuccess)
    return bail(TWS_ERR_CUDA, cudaGetErrorString(e));
  if ((e = cudaMemsetAsync(s->slab, 0, s->slab_bytes, s->st_main)) != cudaSuccess) return bail(TWS_ERR_CUDA, cudaGetErrorString(e));
  if ((e = cudaMalloc(&s->partials, kVolumeBlocks * sizeof(double))) != cudaSuccess) return bail(TWS_ERR_NOMEM, cudaGetErrorString(e));
  if ((e = cudaEventRecord(s->ev_main, s->st_main)) != cudaSuccess || (e = cudaEventRecord(s->ev_edge, s->st_edge)) != cudaSuccess)
    return bail(TWS_ERR_CUDA, cudaGetErrorString(e));
  if ((e = cudaStreamSynchronize(s->st_main)) != cudaSuccess) return bail(TWS_ERR_CUDA, cudaGetErrorString(e));
  (void)rc;
  *out = s;
  return TWS_OK;
}

tws_status tws_destroy(tws_sim* s) {
  if (!s) return TWS_OK;
  DeviceGuard guard(s->prm.device);
  if (s->st_main) cudaStreamSynchronize(s->st_main);
  if (s->st_edge) cudaStreamSynchronize(s->st_edge);
  close_peer(s->up); close_peer(s->down);
  drop_graphs(s);
  for (cudaEvent_t e : s->band_ev) cudaEventDestroy(e);
  if (s->st_h2d) { cudaStreamSynchronize(s->st_h2d); cudaStreamDestroy(s->st_h2d); }
  if (s->st_d2h) { cudaStreamSynchronize(s->st_d2h); cudaStreamDestroy(s->st_d2h); }
  if (s->staging) cudaFree(s->staging);
  if (s->partials) cudaFree(s->partials);
  if (s->res_mailbox) cudaFree(s->res_mailbox);
  if (s->white_dev) cudaFree(s->white_dev);
  tws_gl_unregister(s);
  if (s->packed_info) cudaFree(s->packed_info);
  if (s->packed_flow) cudaFree(s->packed_flow);
  if (s->slab) cudaFree(s->slab);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j)
      if (s->ev_pair[i][j]) cudaEventDestroy(s->ev_pair[i][j]);
  if (s->ev_main) cudaEventDestroy(s->ev_main);
  if (s->ev_edge) cudaEventDestroy(s->ev_edge);
  if (s->st_main) cudaStreamDestroy(s->st_main);
  if (s->st_edge) cudaStreamDestroy(s->st_edge);
  delete s;
  return TWS_OK;
}

tws_status tws_set_steps_per_second(tws_sim* s, float v) {
  if (!s) return TWS_ERR_INVALID;
  if (bad_float(v) || !(v > 0.0f)) return fail(s, TWS_ERR_INVALID, "steps_per_second must be > 0");
  s->prm.steps_per_second = v; derive_constants(s);          // Terrain.cpp:175-185 resets all three
  return TWS_OK;
}
tws_status tws_set_flow_damping(tws_sim* s, float v) {
  if (!s) return TWS_ERR_INVALID;
  if (bad_float(v) || v < 0.0f) return fail(s, TWS_ERR_INVALID, "flow_damping must be >= 0");
  s->prm.flow_damping = v; derive_constants(s);
  return TWS_OK;
}
tws_status tws_set_flow_acceleration(tws_sim* s, float v) {
  if (!s) return TWS_ERR_INVALID;
  if (bad_float(v) || v < 0.0f) return fail(s, TWS_ERR_INVALID, "flow_acceleration must be >= 0");
  s->prm.flow_acceleration = v; derive_constants(s);
  return TWS_OK;
}
tws_status tws_set_sources(tws_sim* s, float rain, float evap) {
  if (!s) return TWS_ERR_INVALID;
  if (bad_float(rain) || bad_float(evap) || rain < 0.0f || evap < 0.0f) return fail(s, TWS_ERR_INVALID, "rain/evaporation must be >= 0");
  s->prm.rain_rate = rain; s->prm.evaporation_rate = evap; derive_constants(s);
  return TWS_OK;
}
tws_status tws_get_step_constants(const tws_sim* s, tws_step_constants* out) {
  if (!s || !out) return TWS_ERR_INVALID;
  out->flow_friction_per_step = s->consts.friction;
  out->water_acceleration_per_step = s->consts.accel;
  out->cell_area_inv_time_scaled = s->consts.area_inv;
  return TWS_OK;
}

static tws_status transfer(tws_sim* s, tws_field field, void* host, size_t bytes, bool upload) {
  if (!s) return TWS_ERR_INVALID;
  if (!host) return fail(s, TWS_ERR_INVALID, "null host buffer");
  const size_t eb = field_elem_bytes(field);
  if (eb == 0) return fail(s, TWS_ERR_INVALID, "unknown field");
  const Geom& g = s->geom;
  const size_t want = (size_t)g.W * g.rows * eb;
  if (bytes != want) return fail(s, TWS_ERR_INVALID, "host buffer size does not match the field (expected " + std::to_string(want) + " bytes)");
  if (upload && field == TWS_FIELD_VELOCITY) return fail(s, TWS_ERR_INVALID, "velocity is an output (m_waterFlowMap is write-only in the reference)");
  DeviceGuard guard(s->prm.device);
  tws_status r = sync_all(s);
  if (r) return r;
  const size_t off = (size_t)TWS_HALO_ROWS * g.pitch;
  if (field == TWS_FIELD_TERRAIN || field == TWS_FIELD_WATER || field == TWS_FIELD_VELOCITY) {
    void* plane = field == TWS_FIELD_TERRAIN ? (void*)(s->planes.h + off)
                : field == TWS_FIELD_WATER   ? (void*)(s->planes.d[s->cur] + off)
                                             : (void*)(s->planes.v + off);
    if (upload) TWS_CUDA(s, cudaMemcpy2DAsync(plane, (size_t)g.pitch * 4, host, (size_t)g.W * 4, (size_t)g.W * 4, g.rows, cudaMemcpyHostToDevice, s->st_main));
    else TWS_CUDA(s, cudaMemcpy2DAsync(host, (size_t)g.W * 4, plane, (size_t)g.pitch * 4, (size_t)g.W * 4, g.rows, cudaMemcpyDeviceToHost, s->st_main));
    TWS_CUDA(s, cudaStreamSynchronize(s->st_main));
    return TWS_OK;
  }
  // AoS fields go through a device staging buffer in row chunks (<= 64 MiB).
  const size_t row_bytes = (size_t)g.W * 16;
  int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)g.rows, ((size_t)64 << 20) / row_bytes));
  r = ensure_staging(s, (size_t)chunk * row_bytes);
  if (r) return r;
  for (int r0 = 0; r0 < g.rows; r0 += chunk) {
    const int n = std::min(chunk, g.rows - r0);
    uint8_t* hp = (uint8_t*)host + (size_t)r0 * row_bytes;
    if (upload) TWS_CUDA(s, cudaMemcpyAsync(s->staging, hp, (size_t)n * row_bytes, cudaMemcpyHostToDevice, s->st_main));
    if (field == TWS_FIELD_FLUX) TWS_CUDA(s, launch_pack_flux(g, s->planes, s->cur, s->staging, r0, n, !upload, s->st_main));
    else TWS_CUDA(s, launch_pack_info(g, s->planes, s->cur, s->staging, r0, n, !upload, s->st_main));
    s->launches += 1;
    if (!upload) TWS_CUDA(s, cudaMemcpyAsync(hp, s->staging, (size_t)n * row_bytes, cudaMemcpyDeviceToHost, s->st_main));
    TWS_CUDA(s, cudaStreamSynchronize(s->st_main));
  }
  return TWS_OK;
}

tws_status tws_upload(tws_sim* s, tws_field field, const void* host, size_t bytes) { return transfer(s, field, (void*)host, bytes, true); }
tws_status tws_readback(tws_sim* s, tws_field field, void* host, size_t bytes) { return transfer(s, field, host, bytes, false); }

// Random.cpp:22-58 + NoiseGenerator.cpp:6-10: the 4096-entry table, generated on the host
// (a 4096-step sequential recurrence), everything per-cell runs on the GPU.
static void white_noise_table(uint32_t seed, float* out) {
  const int N = 624, M = 397;
  std::vector<uint32_t> mt(N);
  for (int i = 0; i < N; ++i) mt[i] = (i % 2) ? (seed + (uint32_t)i * 527u) : ((2135u + seed * 74111u) * (uint32_t)i);
  auto twist = [&](int i) {
    const uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1 == N) ? 0 : i + 1] & 0x7fffffffu);
    mt[i] = mt[(i + M) % N] ^ (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
  };
  for (int i = 0; i < N; ++i) twist(i);
  int idx = 0;
  for (int k = 0; k < 4096; ++k) {
    twist(idx);
    uint32_t y = mt[idx];
    idx = (idx + 1 == N) ? 0 : idx + 1;
    y ^= y >> 11; y ^= (y << 7) & 0x9D2C5680u; y ^= (y << 15) & 0xEFC60000u; y ^= y >> 18;
    out[k] = (float)(y * 4.656612874e-10 - 1.0);
  }
}

tws_status tws_reset_reference_scene(tws_sim* s, uint32_t seed, float height_scale, int32_t lo, int32_t hi, float persistence) {
  return tws_reset_reference_scene_tiled(s, seed, height_scale, lo, hi, persistence, 0);
}

tws_status tws_reset_reference_scene_tiled(tws_sim* s, uint32_t seed, float height_scale, int32_t lo, int32_t hi, float persistence,
                                           int32_t tile_height) {
  if (!s) return TWS_ERR_INVALID;
  if (tile_height != 0 && (tile_height < 2 || tile_height > s->geom.Hg)) return fail(s, TWS_ERR_INVALID, "tile_height must be 0 or in [2, height]");
  const int tile_h = tile_height ? tile_height : s->geom.Hg;
  if (lo < 0 || hi < lo || hi > 24) return fail(s, TWS_ERR_INVALID, "octave range must satisfy 0 <= lo <= hi <= 24");
  if (bad_float(height_scale) || bad_float(persistence)) return fail(s, TWS_ERR_INVALID, "height_scale/persistence must be finite");
  if (s->geom.W < 2 || s->geom.Hg < 2) return fail(s, TWS_ERR_INVALID, "reference scene needs a grid of at least 2x2");
  DeviceGuard guard(s->prm.device);
  tws_status r = sync_all(s);
  if (r) return r;
  if ((s->geom.has_up || s->geom.has_down) && s->connected) {
    // the scene kernel also writes this strip's halo rows: the neighbours' pushes of the last block must have landed
    // first (as in tws_inject_brush).  All strips must be reset together, with no step in between.
    TWS_CUDA(s, cudaStreamWaitEvent(s->st_main, s->ev_edge, 0));
    TWS_CUDA(s, launch_wait_flags(s->ctrl, s->geom.has_up, s->geom.has_down, s->epoch, s->st_main));
    s->launches += 1;
  }
  if (!s->white_dev) TWS_CUDA(s, cudaMalloc(&s->white_dev, 4096 * sizeof(float)));
  float table[4096];
  white_noise_table(seed, table);
  TWS_CUDA(s, cudaMemcpyAsync(s->white_dev, table, sizeof(table), cudaMemcpyHostToDevice, s->st_main));
  // flux = 0 (Terrain.cpp:230-234), both sides, halos included; depth/terrain regenerated.
  for (int side = 0; side < 2; ++side)
    for (int i = 0; i < 4; ++i) TWS_CUDA(s, cudaMemsetAsync(s->planes.F[side][i], 0, s->plane_elems * sizeof(float), s->st_main));
  TWS_CUDA(s, launch_scene(s->geom, s->planes, s->cur, s->white_dev, height_scale, lo, hi, persistence, tile_h, s->st_main));
  s->launches += 1;
  s->accumulator = 0.0;
  TWS_CUDA(s, cudaMemsetAsync(&s->ctrl->outflow_acc, 0, sizeof(double), s->st_main));   // a new scene starts a new ledger
  TWS_CUDA(s, cudaMemsetAsync(&s->ctrl->source_acc, 0, sizeof(double), s->st_main));
  TWS_CUDA(s, cudaStreamSynchronize(s->st_main));
  return TWS_OK;
}

tws_status tws_inject_brush(tws_sim* s, float cx, float cy, float intensity, float size_sq) {
  if (!s) return TWS_ERR_INVALID;
  if (bad_float(cx) || bad_float(cy) || bad_float(intensity)) return fail(s, TWS_ERR_INVALID, "brush centre/intensity must be finite");
  if (bad_float(size_sq) || !(size_sq > 0.0f)) return fail(s, TWS_ERR_INVALID, "brush size_sq must be > 0");
  DeviceGuard guard(s->prm.device);
  const bool strip = s->geom.has_up || s->geom.has_down;
  if (strip) {
    // The brush also edits this strip's copy of the neighbours' edge rows (halo rows), so the
    // neighbours' pushes of the last block must have landed first: wait for their flags.
    if (!s->connected) return fail(s, TWS_ERR_STATE, "strip sim used before tws_halo_connect");
    TWS_CUDA(s, cudaStreamWaitEvent(s->st_main, s->ev_edge, 0));
    TWS_CUDA(s, launch_wait_flags(s->ctrl, s->geom.has_up, s->geom.has_down, s->epoch, s->st_main));
    s->launches += 1;
  }
  { tws_status fr = flush_brush(s); if (fr) return fr; }          // an earlier brush that is still pending goes first
  if (!strip && (s->prm.backend == TWS_BACKEND_FUSED || s->prm.backend == TWS_BACKEND_FUSED_TB || s->prm.backend == TWS_BACKEND_RESIDENT)) {
    // whole grid on the tile or the resident kernel: the next direct launch applies the brush while it loads the depth
    BrushArgs b{};
    if (brush_bbox(s->geom, cx, cy, size_sq, &b.x0, &b.x1, &b.y0, &b.y1)) {
      b.active = 1; b.cx = cx; b.cy = cy; b.intensity = intensity; b.size_sq = size_sq;
      s->pending_brush = b;
    }
    return TWS_OK;
  }
  int launched = 0;
  TWS_CUDA(s, launch_brush(s->geom, s->planes.d[s->cur] + (size_t)TWS_HALO_ROWS * s->geom.pitch, cx, cy, intensity, size_sq, s->st_main, &launched));
  s->launches += launched;
  if (strip) { TWS_CUDA(s, cudaEventRecord(s->ev_main, s->st_main)); }
  return TWS_OK;
}

tws_status tws_inject_brush_world(tws_sim* s, float wx, float wz, float strength) {
  if (!s) return TWS_ERR_INVALID;
  if (bad_float(wx) || bad_float(wz)) return fail(s, TWS_ERR_INVALID, "brush position must be finite");
  // Terrain.cpp:152-155: pos /= worldSize; Fraction(); pos *= (float)resolution.  The
  // reference grid is square; for width != height each axis scales by its own resolution.
  float px = wx / s->prm.world_size, pz = wz / s->prm.world_size;
  px = px - truncf(px); pz = pz - truncf(pz);
  px *= (float)s->prm.width; pz *= (float)s->prm.height;
  return tws_inject_brush(s, px, pz, strength, 32.0f);                                // Terrain.cpp:159
}

tws_status tws_step(tws_sim* s, int32_t n) {
  if (!s) return TWS_ERR_INVALID;
  if (n < 0) return fail(s, TWS_ERR_INVALID, "tws_step: n must be >= 0");
  DeviceGuard guard(s->prm.device);
  return run_steps(s, n);
}

// One step with the water layer in host memory, pipelined in row bands (see tws.h).  Band b's
// kernels read input rows up to one dependency cone (2 rows at k = 1) beyond the band, i.e. into
// band b+1: they wait for that upload; everything earlier is ordered by the upload stream.
tws_status tws_step_host(tws_sim* s, const float* water_in, float* water_out, void* velocity_out) {
  if (!s) return TWS_ERR_INVALID;
  DeviceGuard guard(s->prm.device);
  const Geom& g = s->geom;
  const size_t plane_bytes = (size_t)g.W * g.rows * 4;
  const bool strip = g.has_up || g.has_down;
  if (strip && !s->connected) return fail(s, TWS_ERR_STATE, "strip sim stepped before tws_halo_connect");
  { tws_status fr = flush_brush(s); if (fr) return fr; }
  if (s->prm.backend == TWS_BACKEND_UNFUSED) {
    // The unfused baseline updates in place: no band pipeline.
    tws_status r = TWS_OK;
    if (water_in) { r = transfer(s, TWS_FIELD_WATER, (void*)water_in, plane_bytes, true); if (r) return r; }
    r = run_steps(s, 1); if (r) return r;
    if (water_out) { r = transfer(s, TWS_FIELD_WATER, water_out, plane_bytes, false); if (r) return r; }
    if (velocity_out) { r = transfer(s, TWS_FIELD_VELOCITY, velocity_out, plane_bytes, false); if (r) return r; }
    return sync_all(s);
  }
  if (!s->st_h2d) TWS_CUDA(s, cudaStreamCreateWithFlags(&s->st_h2d, cudaStreamNonBlocking));
  if (!s->st_d2h) TWS_CUDA(s, cudaStreamCreateWithFlags(&s->st_d2h, cudaStreamNonBlocking));
  const bool stream_be = is_stream_backend(s->prm.backend);
  tws_status r = stream_be ? ensure_tma_stream(s) : ensure_tma(s, 1);
  if (r) return r;
  // Band height: a whole number of tile rows (tile engine) / any row count (row-streaming engine),
  // sized for ~8 MiB of upload so a copy is long enough to run at link speed.
  const int unit = stream_be ? 1 : fused_out_rows_per_tile(1);
  const int units = (g.rows + unit - 1) / unit;
  const size_t want_rows = std::max<size_t>(1, ((size_t)8 << 20) / ((size_t)g.W * 4));
  const int band_units = (int)std::max<size_t>(stream_be ? 8 : 1, (want_rows + unit - 1) / unit);
  const int nb = (units + band_units - 1) / band_units;
  while ((int)s->band_ev.size() < 2 * nb + 1) {
    cudaEvent_t e;
    TWS_CUDA(s, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    s->band_ev.push_back(e);
  }
  const int src = s->cur, dst = 1 - src;
  const size_t off = (size_t)TWS_HALO_ROWS * g.pitch;
  const size_t hp = (size_t)g.W * 4, dp = (size_t)g.pitch * 4;       // host / device row pitch in bytes
  auto band_rows = [&](int b, int* r0, int* r1) {
    *r0 = std::min(g.rows, b * band_units * unit);
    *r1 = std::min(g.rows, (b + 1) * band_units * unit);
  };
  auto upload_rows = [&](int r0, int r1) -> cudaError_t {
    if (r1 <= r0) return cudaSuccess;
    return cudaMemcpy2DAsync(s->planes.d[src] + off + (size_t)r0 * g.pitch, dp, (const uint8_t*)water_in + (size_t)r0 * hp, hp, hp,
                             (size_t)(r1 - r0), cudaMemcpyHostToDevice, s->st_h2d);
  };
  next_timing_pair(s);
  TWS_CUDA(s, cudaEventRecord(s->ev_start, s->st_main));
  // uploads and readbacks start after everything already queued on the main stream
  TWS_CUDA(s, cudaStreamWaitEvent(s->st_h2d, s->ev_start, 0));
  TWS_CUDA(s, cudaStreamWaitEvent(s->st_d2h, s->ev_start, 0));
  // Strips: the neighbours compute their edge rows from OUR new edge rows.  Those go up first and are pushed into the
  // neighbours' halo rows (peer stores over NVLink) behind a flag, so that by the time a neighbour's first / last band is
  // due its halo is long there; the rest of the water layer follows band by band.
  int e_top = 0, e_bot = 0;                               // own rows already uploaded ahead of the bands
  if (strip && water_in) {
    const int P = std::min(TWS_HALO_ROWS, g.rows);
    e_top = s->up.present ? P : 0;
    e_bot = s->down.present ? std::min(P, g.rows - e_top) : 0;
    TWS_CUDA(s, upload_rows(0, e_top));
    TWS_CUDA(s, upload_rows(g.rows - e_bot, g.rows));
    if (e_top < P && s->up.present) e_top = 0;            // (cannot happen: a strip has >= 8 rows)
    TWS_CUDA(s, cudaEventRecord(s->band_ev[2 * nb], s->st_h2d));
    TWS_CUDA(s, cudaStreamWaitEvent(s->st_edge, s->ev_main, 0));
    TWS_CUDA(s, cudaStreamWaitEvent(s->st_edge, s->band_ev[2 * nb], 0));
    // the bottom rows were uploaded as rows [rows - e_bot, rows): the push takes the outermost P rows of each side
    RowCopy c[2]; int n = 0;
    auto fill = [&](const Peer& peer, bool to_up) {
      RowCopy rc{};
      rc.src[0] = s->planes.d[src]; rc.dst[0] = peer.planes.d[src];
      rc.nplanes = 1; rc.nrows = P; rc.pitch = g.pitch;
      if (to_up) { rc.src_row = TWS_HALO_ROWS; rc.dst_row = TWS_HALO_ROWS + peer.rows; }
      else { rc.src_row = TWS_HALO_ROWS + g.rows - P; rc.dst_row = TWS_HALO_ROWS - P; }
      return rc;
    };
    if (s->up.present) c[n++] = fill(s->up, true);
    if (s->down.present) c[n++] = fill(s->down, false);
    // a short strip's two pushes may both need rows of the other side's upload: both uploads are behind the one event
    TWS_CUDA(s, launch_row_copy(c[0], c[n - 1], n, s->st_edge));
    s->launches += 1;
    r = post_and_count(s, s->st_edge);
    if (r) return r;
    TWS_CUDA(s, cudaEventRecord(s->ev_edge, s->st_edge));
  }
  if (water_in) {
    for (int b = 0; b < nb; ++b) {
      int r0, r1; band_rows(b, &r0, &r1);
      TWS_CUDA(s, upload_rows(std::max(r0, e_top), std::min(r1, g.rows - e_bot)));
      TWS_CUDA(s, cudaEventRecord(s->band_ev[2 * b], s->st_h2d));
    }
  }
  // Strips compute the bands that read a halo LAST (1, 2, ..., nb-2, then 0 and nb-1): by then the neighbours' pushes
  // have long landed, so a rank that entered the call a little later than its neighbour stalls nobody's pipeline.
  for (int ord = 0; ord < nb; ++ord) {
    int b = ord;
    if (strip && nb >= 3) b = ord < nb - 2 ? ord + 1 : (ord == nb - 2 ? 0 : nb - 1);
    int r0, r1; band_rows(b, &r0, &r1);
    if (water_in) TWS_CUDA(s, cudaStreamWaitEvent(s->st_main, s->band_ev[2 * std::min(b + 1, nb - 1)], 0));
    if (strip) {
      // the first / last band reads the halo rows: the neighbour's push (this step's upload, or the previous step's
      // output rows) must have landed
      const int wu = (b == 0 && g.has_up) ? 1 : 0, wd = (b == nb - 1 && g.has_down) ? 1 : 0;
      if (wu || wd) { TWS_CUDA(s, launch_wait_flags(s->ctrl, wu, wd, s->epoch, s->st_main)); s->launches += 1; }
    }
    const int u0 = b * band_units, u1 = std::min(units, (b + 1) * band_units);
    if (stream_be) TWS_CUDA(s, launch_stream(1, g, s->planes, s->tma_stream[src], src, s->consts, r0, r1, s->st_main, stream_impl(s), s->ctrl->sched_main));
    else TWS_CUDA(s, launch_fused(1, g, s->planes, s->tma[1][src], src, s->consts, u0, u1, s->st_main));
    s->launches += 1;
    TWS_CUDA(s, cudaEventRecord(s->band_ev[2 * b + 1], s->st_main));
    if ((water_out || velocity_out) && r1 > r0) {
      TWS_CUDA(s, cudaStreamWaitEvent(s->st_d2h, s->band_ev[2 * b + 1], 0));
      if (water_out)
        TWS_CUDA(s, cudaMemcpy2DAsync((uint8_t*)water_out + (size_t)r0 * hp, hp, s->planes.d[dst] + off + (size_t)r0 * g.pitch, dp, hp,
                                      (size_t)(r1 - r0), cudaMemcpyDeviceToHost, s->st_d2h));
      if (velocity_out)
        TWS_CUDA(s, cudaMemcpy2DAsync((uint8_t*)velocity_out + (size_t)r0 * hp, hp, s->planes.v + off + (size_t)r0 * g.pitch, dp, hp,
                                      (size_t)(r1 - r0), cudaMemcpyDeviceToHost, s->st_d2h));
    }
  }
  s->cur = dst;
  TWS_CUDA(s, cudaEventRecord(s->ev_main, s->st_main));
  if (strip) {
    // the new edge rows (depth + flux) into the neighbours' halo rows, then the flag: what tws_step's launches do themselves
    TWS_CUDA(s, cudaStreamWaitEvent(s->st_edge, s->ev_main, 0));
    r = push_edges(s, dst, false, s->st_edge); if (r) return r;
    r = post_and_count(s, s->st_edge); if (r) return r;
    TWS_CUDA(s, cudaEventRecord(s->ev_edge, s->st_edge));
  }
  TWS_CUDA(s, cudaEventRecord(s->ev_stop, s->st_main));
  s->timed = true;
  TWS_CUDA(s, cudaStreamSynchronize(s->st_h2d));
  TWS_CUDA(s, cudaStreamSynchronize(s->st_d2h));
  return sync_all(s);
}

tws_status tws_advance(tws_sim* s, double frame_seconds, uint32_t* steps_done) {
  if (steps_done) *steps_done = 0;
  if (!s) return TWS_ERR_INVALID;
  if (!(frame_seconds >= 0.0) || std::isinf(frame_seconds)) return fail(s, TWS_ERR_INVALID, "tws_advance: frame time must be finite and >= 0");
  // the reference casts the quotient straight to a 32-bit unsigned (:243), undefined beyond 2^32: refuse such a frame
  // before the accumulator is touched instead of running with a garbage count
  if (!((s->accumulator + frame_seconds) / s->step_length < 4294967296.0))
    return fail(s, TWS_ERR_INVALID, "tws_advance: frame time / step length does not fit the reference's 32-bit step counter");
  s->accumulator += frame_seconds;                                                    // Terrain.cpp:242
  uint32_t n = (uint32_t)(s->accumulator / s->step_length);                           // :243
  s->accumulator -= s->step_length * n;                                               // :244
  n = std::min<uint32_t>(n, 10u);                                                     // :247
  if (steps_done) *steps_done = n;
  if (n == 0) return TWS_OK;
  DeviceGuard guard(s->prm.device);
  return run_steps(s, (int)n);
}

tws_status tws_total_volume(tws_sim* s, double* volume) {
  if (!s || !volume) return TWS_ERR_INVALID;
  DeviceGuard guard(s->prm.device);
  tws_status r = sync_all(s);
  if (r) return r;
  TWS_CUDA(s, launch_volume(s->geom, s->planes.d[s->cur], s->partials, kVolumeBlocks, s->st_main));
  s->launches += 1;
  std::vector<double> h(kVolumeBlocks);
  TWS_CUDA(s, cudaMemcpyAsync(h.data(), s->partials, kVolumeBlocks * sizeof(double), cudaMemcpyDeviceToHost, s->st_main));
  TWS_CUDA(s, cudaStreamSynchronize(s->st_main));
  double acc = 0.0;
  for (double v : h) acc += v;
  *volume = acc;
  return TWS_OK;
}

tws_status tws_boundary_outflow(tws_sim* s, double* flux_sum) {
  if (!s || !flux_sum) return TWS_ERR_INVALID;
  DeviceGuard guard(s->prm.device);
  tws_status r = sync_all(s);
  if (r) return r;
  TWS_CUDA(s, launch_boundary_outflow(s->geom, s->planes, s->cur, s->partials, s->st_main));
  s->launches += 1;
  TWS_CUDA(s, cudaMemcpyAsync(flux_sum, s->partials, sizeof(double), cudaMemcpyDeviceToHost, s->st_main));
  TWS_CUDA(s, cudaStreamSynchronize(s->st_main));
  return TWS_OK;
}

tws_status tws_boundary_outflow_accumulated(tws_sim* s, double* volume) {
  if (!s || !volume) return TWS_ERR_INVALID;
  DeviceGuard guard(s->prm.device);
  tws_status r = sync_all(s);
  if (r) return r;
  TWS_CUDA(s, cudaMemcpy(volume, &s->ctrl->outflow_acc, sizeof(double), cudaMemcpyDeviceToHost));
  return TWS_OK;
}

tws_status tws_source_accumulated(tws_sim* s, double* volume) {
  if (!s || !volume) return TWS_ERR_INVALID;
  DeviceGuard guard(s->prm.device);
  tws_status r = sync_all(s);
  if (r) return r;
  TWS_CUDA(s, cudaMemcpy(volume, &s->ctrl->source_acc, sizeof(double), cudaMemcpyDeviceToHost));
  return TWS_OK;
}

tws_status tws_boundary_outflow_reset(tws_sim* s) {
  if (!s) return TWS_ERR_INVALID;
  DeviceGuard guard(s->prm.device);
  TWS_CUDA(s, cudaMemsetAsync(&s->ctrl->outflow_acc, 0, sizeof(double), s->st_main));
  TWS_CUDA(s, cudaMemsetAsync(&s->ctrl->source_acc, 0, sizeof(double), s->st_main));
  return TWS_OK;
}

tws_status tws_sync(tws_sim* s) {
  if (!s) return TWS_ERR_INVALID;
  DeviceGuard guard(s->prm.device);
  return sync_all(s);
}

tws_status tws_elapsed_ms(tws_sim* s, float* ms) {
  if (!s || !ms) return TWS_ERR_INVALID;
  if (!s->timed) return fail(s, TWS_ERR_STATE, "no step batch has been timed yet");
  DeviceGuard guard(s->prm.device);
  TWS_CUDA(s, cudaEventSynchronize(s->ev_stop));
  TWS_CUDA(s, cudaEventElapsedTime(ms, s->ev_start, s->ev_stop));
  return TWS_OK;
}

// gl::TimerQuery semantics (TimerQuery.cpp:61-72 with wait = false; read one frame late at Scene.cpp:337-342): never blocks.
tws_status tws_elapsed_ms_nowait(tws_sim* s, float* ms, uint64_t* batch_index) {
  if (!s || !ms) return TWS_ERR_INVALID;
  if (!s->timed) return fail(s, TWS_ERR_STATE, "no step batch has been timed yet");
  DeviceGuard guard(s->prm.device);
  // newest first: the most recent batch if the GPU is already through with it, else the one before
  for (uint64_t back = 0; back < 2 && back < s->batches_timed; ++back) {
    const uint64_t idx = s->batches_timed - 1 - back;
    const int p = (int)(idx & 1u);
    const cudaError_t q = cudaEventQuery(s->ev_pair[p][1]);
    if (q == cudaSuccess) {
      TWS_CUDA(s, cudaEventElapsedTime(ms, s->ev_pair[p][0], s->ev_pair[p][1]));
      if (batch_index) *batch_index = idx;
      return TWS_OK;
    }
    if (q != cudaErrorNotReady) return cuda_fail(s, q, "cudaEventQuery");
    (void)cudaGetLastError();
  }
  return fail(s, TWS_ERR_STATE, "no timed batch has finished on the GPU yet");
}

uint64_t tws_kernel_launches(const tws_sim* s) { return s ? s->launches : 0; }
tws_status tws_backend_in_use(const tws_sim* s, int32_t* backend, int32_t* temporal_block) {
  if (!s) return TWS_ERR_INVALID;
  if (backend) *backend = s->prm.backend;
  if (temporal_block) *temporal_block = s->K;
  return TWS_OK;
}

uint64_t tws_graph_replays(const tws_sim* s) { return s ? s->graph_replays : 0; }

tws_status tws_device_view(tws_sim* s, tws_field field, void** ptr, int64_t* pitch) {
  if (!s || !ptr || !pitch) return TWS_ERR_INVALID;
  {
    DeviceGuard guard(s->prm.device);
    tws_status fr = flush_brush(s);                      // the caller is about to read the planes with its own kernels
    if (fr) return fr;
  }
  const size_t off = (size_t)TWS_HALO_ROWS * s->geom.pitch;
  switch (field) {
    case TWS_FIELD_TERRAIN: *ptr = s->planes.h + off; break;
    case TWS_FIELD_WATER: *ptr = s->planes.d[s->cur] + off; break;
    case TWS_FIELD_VELOCITY: *ptr = s->planes.v + off; break;
    default: return fail(s, TWS_ERR_INVALID, "tws_device_view: only TERRAIN, WATER and VELOCITY are single planes");
  }
  *pitch = s->geom.pitch;
  return TWS_OK;
}

// ---- strips ---------------------------------------------------------------------------------
tws_status tws_halo_export(tws_sim* s, tws_halo_handle* out) {
  if (!s || !out) return TWS_ERR_INVALID;
  DeviceGuard guard(s->prm.device);
  std::memset(out, 0, sizeof(*out));
  static_assert(sizeof(cudaIpcMemHandle_t) <= TWS_IPC_HANDLE_BYTES, "IPC handle does not fit");
  cudaIpcMemHandle_t h;
  TWS_CUDA(s, cudaIpcGetMemHandle(&h, s->slab));
  std::memcpy(out->mem, &h, sizeof(h));
  out->slab_bytes = s->slab_bytes;
  out->row_begin = s->prm.row_begin; out->row_end = s->prm.row_end;
  out->device = s->prm.device;
  out->pid = (int32_t)getpid();
  out->local_ptr = (uint64_t)(uintptr_t)s->slab;
  return TWS_OK;
}

static tws_status open_peer(tws_sim* s, const tws_halo_handle* h, Peer* p, bool is_up) {
  close_peer(*p);
  if (!h) return TWS_OK;
  const int rows = h->row_end - h->row_begin;
  if (is_up ? (h->row_end != s->prm.row_begin) : (h->row_begin != s->prm.row_end))
    return fail(s, TWS_ERR_INVALID, "tws_halo_connect: neighbour strip is not adjacent");
  Planes pl; size_t total = 0;
  layout_planes(nullptr, s->geom.pitch, rows, &pl, nullptr, &total);
  if (total != h->slab_bytes) return fail(s, TWS_ERR_INVALID, "tws_halo_connect: neighbour slab layout mismatch (different width?)");
  uint8_t* base = nullptr;
  if (h->pid == (int32_t)getpid()) {
    if (h->device != s->prm.device) {
      int can = 0;
      TWS_CUDA(s, cudaDeviceCanAccessPeer(&can, s->prm.device, h->device));
      if (!can) return fail(s, TWS_ERR_CUDA, "tws_halo_connect: no peer access between the two devices");
      cudaError_t e = cudaDeviceEnablePeerAccess(h->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(s, e, "cudaDeviceEnablePeerAccess");
      (void)cudaGetLastError();
    }
    base = (uint8_t*)(uintptr_t)h->local_ptr;
  } else {
    cudaIpcMemHandle_t mh;
    std::memcpy(&mh, h->mem, sizeof(mh));
    void* ptr = nullptr;
    TWS_CUDA(s, cudaIpcOpenMemHandle(&ptr, mh, cudaIpcMemLazyEnablePeerAccess));
    base = (uint8_t*)ptr;
    p->ipc = true;
  }
  p->present = true; p->slab = base; p->rows = rows; p->device = h->device;
  layout_planes(base, s->geom.pitch, rows, &p->planes, nullptr, nullptr);
  p->ctrl = (Control*)base;
  return TWS_OK;
}

tws_status tws_halo_connect(tws_sim* s, const tws_halo_handle* up, const tws_halo_handle* down) {
  if (!s) return TWS_ERR_INVALID;
  if ((up != nullptr) != (s->geom.has_up != 0) || (down != nullptr) != (s->geom.has_down != 0))
    return fail(s, TWS_ERR_INVALID, "tws_halo_connect: a handle is required exactly where the strip has a neighbour");
  DeviceGuard guard(s->prm.device);
  tws_status r = open_peer(s, up, &s->up, true);
  if (r) return r;
  r = open_peer(s, down, &s->down, false);
  if (r) return r;
  // (re)connecting is the recovery path after an exchange time-out: the sticky error flag starts clean
  TWS_CUDA(s, cudaMemsetAsync(&s->ctrl->error, 0, sizeof(uint32_t), s->st_main));
  TWS_CUDA(s, cudaStreamSynchronize(s->st_main));
  s->connected = true;
  return TWS_OK;
}

tws_status tws_halo_refresh(tws_sim* s) {
  if (!s) return TWS_ERR_INVALID;
  const bool strip = s->geom.has_up || s->geom.has_down;
  if (!strip) return TWS_OK;
  if (!s->connected) return fail(s, TWS_ERR_STATE, "tws_halo_refresh before tws_halo_connect");
  DeviceGuard guard(s->prm.device);
  TWS_CUDA(s, cudaStreamWaitEvent(s->st_edge, s->ev_main, 0));
  tws_status r = push_edges(s, s->cur, true, s->st_edge);
  if (r) return r;
  r = post_and_count(s, s->st_edge);
  if (r) return r;
  TWS_CUDA(s, cudaEventRecord(s->ev_edge, s->st_edge));
  return TWS_OK;
}

// ---- renderer hand-off -------------------------------------------------------------------------
// Level count and sizes follow the reference's texture wrapper: glEasy Texture.cpp:28-41 counts
// halvings until every side is 0 (floor(log2(max side)) + 1 levels), glTexStorage2D gives level L
// max(1, side >> L) texels per side.
int32_t tws_mip_levels(int32_t width, int32_t height) {
  int32_t n = 0;
  while (width > 0 || height > 0) { width /= 2; height /= 2; ++n; }
  return n;
}

tws_status tws_mip_level_info(int32_t width, int32_t height, int32_t level, int32_t* w, int32_t* h, int64_t* offset_texels) {
  if (width < 1 || height < 1 || level < 0 || level >= tws_mip_levels(width, height)) return TWS_ERR_INVALID;
  int64_t off = 0;
  for (int32_t l = 0; l < level; ++l) off += (int64_t)std::max(1, width >> l) * std::max(1, height >> l);
  if (w) *w = std::max(1, width >> level);
  if (h) *h = std::max(1, height >> level);
  if (offset_texels) *offset_texels = off;
  return TWS_OK;
}

namespace {
size_t mip_chain_texels(int w, int h) {
  size_t n = 0;
  for (int l = 0, L = tws_mip_levels(w, h); l < L; ++l) n += (size_t)std::max(1, w >> l) * std::max(1, h >> l);
  return n;
}
// level 0 of TerrainInfo, the flow map and — with_level1 — mip level 1 into the publish buffers, in one launch; the info
// buffer has room for the whole chain
tws_status publish_level0(tws_sim* s, bool with_level1) {
  { tws_status fr = flush_brush(s); if (fr) return fr; }
  const Geom& g = s->geom;
  const size_t cells = (size_t)g.W * g.rows;
  if (!s->packed_info) TWS_CUDA(s, cudaMalloc(&s->packed_info, mip_chain_texels(g.W, g.rows) * 16));
  if (!s->packed_flow) TWS_CUDA(s, cudaMalloc(&s->packed_flow, cells * 4));
  if (g.has_up || g.has_down) TWS_CUDA(s, cudaStreamWaitEvent(s->st_main, s->ev_edge, 0));
  float* level1 = with_level1 ? (float*)s->packed_info + cells * 4 : nullptr;          // level 1 follows level 0
  TWS_CUDA(s, launch_publish_fused(g, s->planes, s->cur, (float*)s->packed_info, level1, (uint32_t*)s->packed_flow, s->st_main));
  s->launches += 1;
  s->published_levels = with_level1 ? 2 : 1;
  return TWS_OK;
}
// Levels a strip can filter from its own rows alone: level l needs the strip cut on a multiple of 2^l rows (then its
// rows of level l are exactly rows [row_begin >> l, row_end >> l) of the whole grid's level l: 2x2 boxes never straddle
// the seam and no source coordinate is clamped in y).  A whole grid has the full chain.
int strip_mip_levels(const Geom& g) {
  const int full = tws_mip_levels(g.W, g.Hg);
  if (!g.has_up && !g.has_down && g.rows == g.Hg) return full;
  int l = 0;
  while (l + 1 < full && (g.row0 % (2 << l)) == 0 && (g.rows % (2 << l)) == 0) ++l;
  return l + 1;
}
// levels 1.. of the published TerrainInfo (Terrain.cpp:272-276).  A strip publishes the levels it can filter without its
// neighbours (strip_mip_levels; plan_strips cuts on multiples of 8 rows, so at least levels 0-3); the consumer gathers
// the strips' level L rows and filters the few remaining small levels itself (tws.h).
tws_status publish_chain(tws_sim* s) {
  { tws_status fr = flush_brush(s); if (fr) return fr; }
  const Geom& g = s->geom;
  const int L = strip_mip_levels(g);
  if (publish_all_applicable(g, L)) {                   // e.g. the reference's 1024^2: the whole hand-off is one launch
    const size_t cells = (size_t)g.W * g.rows;
    if (!s->packed_info) TWS_CUDA(s, cudaMalloc(&s->packed_info, mip_chain_texels(g.W, g.rows) * 16));
    if (!s->packed_flow) TWS_CUDA(s, cudaMalloc(&s->packed_flow, cells * 4));
    TWS_CUDA(s, launch_publish_all(g, s->planes, s->cur, (float*)s->packed_info, (uint32_t*)s->packed_flow, L, &s->ctrl->mip_ticket, s->st_main));
    s->launches += 1;
    s->published_levels = L;
    return TWS_OK;
  }
  tws_status r = publish_level0(s, L > 1);
  if (r) return r;
  float* base = (float*)s->packed_info;
  // level 1 came with level 0; further large levels one launch each; from the first level of <= 16 K texels on, the whole tail in one launch
  int first_tail = L;
  for (int l = 2; l < L; ++l) {
    int32_t sw, sh, dw, dh; int64_t so, d_o;
    tws_mip_level_info(g.W, g.rows, l - 1, &sw, &sh, &so);
    tws_mip_level_info(g.W, g.rows, l, &dw, &dh, &d_o);
    if ((long long)dw * dh <= 16384) { first_tail = l; break; }
    TWS_CUDA(s, launch_mip_level(base + so * 4, sw, sh, base + d_o * 4, dw, dh, s->st_main));
    s->launches += 1;
  }
  if (first_tail < L) {
    TWS_CUDA(s, launch_mip_tail(base, g.W, g.rows, first_tail, L, s->st_main));
    s->launches += 1;
  }
  s->published_levels = L;
  return TWS_OK;
}
}  // namespace

tws_status tws_publish_packed(tws_sim* s, void** info, void** flow) {
  if (!s || !info || !flow) return TWS_ERR_INVALID;
  DeviceGuard guard(s->prm.device);
  tws_status r = publish_level0(s, false);
  if (r) return r;
  *info = s->packed_info; *flow = s->packed_flow;
  return TWS_OK;
}

tws_status tws_publish_mips(tws_sim* s, void** info_chain, int32_t* levels) {
  if (!s || !info_chain) return TWS_ERR_INVALID;
  DeviceGuard guard(s->prm.device);
  tws_status r = publish_chain(s);
  if (r) return r;
  *info_chain = s->packed_info;
  if (levels) *levels = s->published_levels;
  return TWS_OK;
}

tws_status tws_readback_mip(tws_sim* s, int32_t level, void* host, size_t bytes) {
  if (!s || !host) return TWS_ERR_INVALID;
  DeviceGuard guard(s->prm.device);
  if (level < 0 || level >= s->published_levels) return fail(s, TWS_ERR_STATE, "tws_readback_mip: level not published (call tws_publish_mips first)");
  int32_t w, h; int64_t off;
  tws_mip_level_info(s->geom.W, s->geom.rows, level, &w, &h, &off);
  if (bytes != (size_t)w * h * 16) return fail(s, TWS_ERR_INVALID, "tws_readback_mip: byte count does not match the level (w*h*16)");
  TWS_CUDA(s, cudaMemcpyAsync(host, (const float*)s->packed_info + off * 4, bytes, cudaMemcpyDeviceToHost, s->st_main));
  TWS_CUDA(s, cudaStreamSynchronize(s->st_main));
  return TWS_OK;
}

// CUDA-GL interop.  <cuda_gl_interop.h> needs <GL/gl.h>, which a headless build image does not ship; the
// one GL-specific entry point lives in libcudart itself and takes plain GL object names, so it is declared
// here with the GL typedefs spelled out (GLuint = GLenum = unsigned int).  The calling thread must have the
// renderer's GL context current (the reference is single-threaded: RenderWindow.cpp:103-108); without one
// the CUDA runtime reports an error and the call fails with TWS_ERR_CUDA.
extern "C" cudaError_t cudaGraphicsGLRegisterImage(struct cudaGraphicsResource** resource, unsigned int image, unsigned int target,
                                                   unsigned int flags);
static constexpr unsigned int kGlTexture2D = 0x0DE1;   // GL_TEXTURE_2D

tws_status tws_gl_unregister(tws_sim* s) {
  if (!s) return TWS_ERR_INVALID;
  DeviceGuard guard(s->prm.device);
  for (cudaGraphicsResource** r : {&s->gl_info, &s->gl_flow})
    if (*r) { cudaGraphicsUnregisterResource(*r); *r = nullptr; }
  return TWS_OK;
}

tws_status tws_gl_register(tws_sim* s, uint32_t terrain_info_tex, uint32_t flow_map_tex) {
  if (!s) return TWS_ERR_INVALID;
  DeviceGuard guard(s->prm.device);
  if (s->geom.has_up || s->geom.has_down) return fail(s, TWS_ERR_UNSUPPORTED, "tws_gl_register: whole grids only (the renderer samples one texture)");
  tws_gl_unregister(s);
  cudaError_t e = cudaGraphicsGLRegisterImage(&s->gl_info, terrain_info_tex, kGlTexture2D, cudaGraphicsRegisterFlagsWriteDiscard);
  if (e == cudaSuccess && flow_map_tex)
    e = cudaGraphicsGLRegisterImage(&s->gl_flow, flow_map_tex, kGlTexture2D, cudaGraphicsRegisterFlagsWriteDiscard);
  if (e != cudaSuccess) {
    cudaGetLastError();
    tws_gl_unregister(s);
    return cuda_fail(s, e, "tws_gl_register: cudaGraphicsGLRegisterImage (is the renderer's GL context current on this thread?)");
  }
  return TWS_OK;
}

// One frame's hand-off: TerrainInfo level 0 (r = terrain, g = b = 0.3, a = water) plus the mip chain the
// reference regenerates (Terrain.cpp:272-276) into the RGBA32F texture, the flow vectors into the RG16F one.
// Levels the texture does not have (it was allocated with fewer) end the copy loop quietly.
tws_status tws_gl_publish(tws_sim* s) {
  if (!s) return TWS_ERR_INVALID;
  if (!s->gl_info) return fail(s, TWS_ERR_STATE, "tws_gl_publish before tws_gl_register");
  DeviceGuard guard(s->prm.device);
  tws_status r = publish_chain(s);
  if (r) return r;
  const Geom& g = s->geom;
  cudaGraphicsResource* res[2] = {s->gl_info, s->gl_flow};
  const int nres = s->gl_flow ? 2 : 1;
  TWS_CUDA(s, cudaGraphicsMapResources(nres, res, s->st_main));
  cudaError_t e = cudaSuccess;
  for (int l = 0; l < s->published_levels && e == cudaSuccess; ++l) {
    int32_t w, h; int64_t off;
    tws_mip_level_info(g.W, g.rows, l, &w, &h, &off);
    cudaArray_t arr = nullptr;
    if (cudaGraphicsSubResourceGetMappedArray(&arr, s->gl_info, 0, (unsigned)l) != cudaSuccess) { cudaGetLastError(); break; }
    e = cudaMemcpy2DToArrayAsync(arr, 0, 0, (const float*)s->packed_info + off * 4, (size_t)w * 16, (size_t)w * 16, (size_t)h,
                                 cudaMemcpyDeviceToDevice, s->st_main);
  }
  if (e == cudaSuccess && s->gl_flow) {
    cudaArray_t arr = nullptr;
    e = cudaGraphicsSubResourceGetMappedArray(&arr, s->gl_flow, 0, 0);
    if (e == cudaSuccess)
      e = cudaMemcpy2DToArrayAsync(arr, 0, 0, s->packed_flow, (size_t)g.W * 4, (size_t)g.W * 4, (size_t)g.rows, cudaMemcpyDeviceToDevice, s->st_main);
  }
  cudaError_t eu = cudaGraphicsUnmapResources(nres, res, s->st_main);   // unmap orders the copies before the renderer's next GL use
  if (e != cudaSuccess) return cuda_fail(s, e, "tws_gl_publish: copy into the mapped texture");
  if (eu != cudaSuccess) return cuda_fail(s, eu, "tws_gl_publish: cudaGraphicsUnmapResources");
  return TWS_OK;
}

}  // extern "C"
