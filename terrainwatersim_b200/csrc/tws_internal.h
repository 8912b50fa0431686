// tws_internal.h — shared declarations of libtws.so (host side + kernel launchers).
//
// HBM layout.  Everything a strip owns lives in ONE cudaMalloc'ed slab so
// a single CUDA-IPC handle exposes it to the neighbouring GPUs:
//
//   [ control block 4 KiB | h | d[0] | d[1] | F[0][+X,-X,+Y,-Y] | F[1][..] | v ]
//
// Every plane is planar fp32 (v: 2 x fp16 packed in 32 bit), `pitch` elements per row
// (pitch = width rounded up to 64, pad columns always hold 0), `plane_rows` =
// own rows + 2*TWS_HALO_ROWS rows; own local row r is plane row r + TWS_HALO_ROWS.
// Halo rows mirror the neighbouring strip's edge rows (pushed over NVLink) or stay 0
// and are never exposed to the kernels when the strip touches the global edge.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <string>

#include "../../include/tws.h"

#define TWS_HALO_ROWS 8          // 2 * max temporal block (4)
#define TWS_MAX_TB 4
#define TWS_CTRL_BYTES 4096

namespace tws {

// Control block at the start of the slab (device memory, written by peers).
struct Control {
  // Written by the strip ABOVE / BELOW (remote store over NVLink), polled locally.
  alignas(128) volatile uint32_t flag_from_up;
  alignas(128) volatile uint32_t flag_from_down;
  alignas(128) uint32_t error;       // set by the wait kernel on timeout
  // Piece counters of the band kernel's dynamic schedule, {next piece, groups done}: one pair per stream
  // that launches step kernels (launches on one stream are ordered, the pair is re-armed by the kernel itself).
  // sched_*[2] counts the warps that have finished an edge piece of a strip launch (the last one posts the flags).
  alignas(128) uint32_t sched_main[4];
  alignas(128) uint32_t sched_edge[4];
  // EXT mass ledger: fp64 sum of (outflow across the edge of the GLOBAL grid) x areaInv over every sub-step this sim has
  // run, i.e. the volume that left the map through this strip's part of the edge (flowApply.comp:38-41 with the exterior
  // reading 0).  Accumulated by the step kernels themselves (atomicAdd from the lanes that own an edge cell), so it is
  // exact for k steps per launch, where the intermediate fluxes never reach HBM.
  alignas(128) double outflow_acc;
  // ... and what rain / evaporation REALLY added: fp64 sum over every owned cell and sub-step of (depth after the source
  // terms - depth before them) as computed in fp32.  d + rain_step - evap_step rounds the same way for every cell of a
  // binade, so the applied source differs systematically (up to ~1e-4 relative over 10 000 steps) from rate x time x area;
  // a ledger that is to close has to book what the arithmetic did.
  alignas(128) double source_acc;
  alignas(128) uint32_t mip_ticket;  // CTAs of the mip chain's single-pass kernel that have finished (re-armed by the last one)
};

static_assert(sizeof(Control) <= TWS_CTRL_BYTES, "the control block must fit in front of the planes");

struct StepConsts {
  float friction, accel, area_inv;   // simulationCommon.glsl:1-13
  float rain_step, evap_step;        // EXT, already multiplied by dt
  int   closed;                      // EXT boundary
  int   ext_sources;                 // rain_step != 0 || evap_step != 0
  double* ledger;                    // EXT: &Control::outflow_acc, or nullptr (closed boundary: nothing can leave)
  double* ledger_src;                // EXT: &Control::source_acc, or nullptr (no sources)
};

// Geometry handed to every kernel.
struct Geom {
  int W;            // global width
  int Hg;           // global height
  int row0;         // first global row of the strip
  int rows;         // own rows
  int pitch;        // elements per plane row
  int has_up;       // halo rows above are live (a neighbour strip exists)
  int has_down;
};

struct Planes {      // device pointers to plane row 0 (i.e. local row -TWS_HALO_ROWS)
  float* h;
  float* d[2];
  float* F[2][4];
  uint32_t* v;
};

// TMA descriptors of one ping-pong side: h, d, F x4 — boxes depend on the kernel config.
struct TmaSet { CUtensorMap m[6]; };

struct KernelStats { uint64_t launches = 0; };

// Strips, band kernel: ONE launch per block does the halo exchange itself.  Its last pieces are the edge rows:
// their warps wait for the neighbours' flags (local polling), compute, store every output row also into the
// neighbour's halo rows (peer stores over NVLink), and the warp that finishes the last edge piece posts the new
// epoch in the neighbours' control blocks — while the other warp groups are still working on interior pieces.
struct BandEdge {
  int nlev_edge;                        // levels of the schedule that are edge pieces (0: not a strip launch) ...
  int lev_edge0;                        // ... and the first of them (they are the last levels of the work list)
  unsigned n_edge_warps;                // warps that will finish an edge piece in this launch
  const volatile uint32_t* wait_up;     // own control block: flag written by the strip above / below (nullptr: none)
  const volatile uint32_t* wait_down;
  uint32_t wait_value;
  uint32_t* error;                      // own control block: raised when a wait times out
  volatile uint32_t* post_up;           // neighbours' control blocks
  volatile uint32_t* post_down;
  uint32_t post_value;
  int up_end;                           // own rows y < up_end are also stored into the upper neighbour's bottom halo
  int down_begin;                       // own rows y >= down_begin into the lower neighbour's top halo
  float* up[5];                         // d, F x4 of the neighbour, offset so that [y * pitch + x] is the halo cell of own (x, y)
  float* down[5];
};

// A brush (waterBrush.comp:20-31) that the next tile-kernel launch applies while it loads the depth — the separate brush
// launch of the reference's frame (brush, step, mip chain) folded into the step.  [x0, x1] x [y0, y1]: the brush's bounding
// box in global texels, clipped to the grid; outside it the brush adds exactly 0.
struct BrushArgs {
  int active;
  float cx, cy, intensity, size_sq;
  int x0, x1, y0, y1;
};
// the bounding box launch_brush uses; false: the brush misses the stored rows / columns entirely (adds 0 everywhere)
bool brush_bbox(const Geom& g, float cx, float cy, float size_sq, int* x0, int* x1, int* y0, int* y1);

// ---- launchers (step_kernels.cu) ------------------------------------------------------
// Fused K-level step over tile rows [ty0, ty1) of the strip; reads side `src`, writes 1-src.
// Returns the number of tile rows for the strip through tiles_y when called with ty1 < 0.
int  fused_tile_rows(int K, int rows);
int  fused_out_rows_per_tile(int K);
cudaError_t fused_build_tma(int K, const Geom& g, const Planes& p, int side, TmaSet* out, std::string* err);
cudaError_t launch_fused(int K, const Geom& g, const Planes& p, const TmaSet& tma, int src, const StepConsts& c,
                         int ty0, int ty1, cudaStream_t st, const BrushArgs* brush = nullptr);
// Row-streaming pipeline (stream_kernels.cu): K steps over local rows [lr0, lr1) of the strip.
cudaError_t build_tma_boxes(const Geom& g, const Planes& p, int side, int box_x, int box_y, TmaSet* out, std::string* err,
                            int l2_promotion);   // 0 none, 1 128 B, 2 256 B
cudaError_t build_tma_rows(const Geom& g, const Planes& p, int side, int box_x, TmaSet* out, std::string* err, int l2_promotion);
cudaError_t stream_build_tma(const Geom& g, const Planes& p, int side, TmaSet* out, std::string* err);
// impl 0: ring kernel (one warp per row, per-row barriers); impl 1: band kernel (CTA-synchronous skewed bands).
// sched: the Control counter pair of the launching stream (band kernel; nullptr = static round-robin pieces).
// cta_budget (band kernel): 0 = one CTA per SM, n > 0 = at most n CTAs, n < 0 = leave -n SMs free.
cudaError_t launch_stream(int K, const Geom& g, const Planes& p, const TmaSet& tma, int src, const StepConsts& c,
                          int lr0, int lr1, cudaStream_t st, int impl, uint32_t* sched, int cta_budget = 0);
// Band kernel (band_kernels.cu): the streaming schedule in lock step; same row descriptors as the ring kernel.
// edge != nullptr: a strip's whole block in one launch — rows [lr0, lr0 + e_top) and [lr1 - e_bot, lr1) are edge pieces
// (taken first), the rows between them interior pieces.
cudaError_t launch_band(int K, const Geom& g, const Planes& p, const TmaSet& tma, int src, const StepConsts& c,
                        int lr0, int lr1, cudaStream_t st, uint32_t* sched, int cta_budget, const BandEdge* edge = nullptr,
                        int e_top = 0, int e_bot = 0);
int stream_strip_width();
cudaError_t launch_unfused_update(const Geom& g, const Planes& p, int side, const StepConsts& c, int lr0, int lr1, cudaStream_t st);
cudaError_t launch_unfused_apply(const Geom& g, const Planes& p, int side, const StepConsts& c, int lr0, int lr1, cudaStream_t st);

// Resident kernel (resident_kernels.cu): n steps of a whole grid that fits on chip in one cooperative launch.
// resident_config: block shape index for this grid, -1 if it does not fit (or is a strip).
int resident_config(const Geom& g);
int resident_blocks(const Geom& g, int cfg);
// mailbox: resident_mailbox_bytes(g) of zeroed device memory (the rims travel through it, tagged with their step number);
// epoch0: steps this sim's resident launches have run before (tags must never repeat)
size_t resident_mailbox_bytes(const Geom& g);
cudaError_t launch_resident(int cfg, const Geom& g, const Planes& p, const StepConsts& c, int src, int n, void* mailbox,
                            uint32_t epoch0, uint32_t* error, cudaStream_t st, const BrushArgs* brush = nullptr);

// ---- launchers (aux_kernels.cu) -------------------------------------------------------
cudaError_t launch_brush(const Geom& g, float* d, float cx, float cy, float intensity, float size_sq, cudaStream_t st, int* launched);
cudaError_t launch_scene(const Geom& g, const Planes& p, int side, const float* white_dev, float height_scale, int lo, int hi,
                         float persistence, int tile_h, cudaStream_t st);
cudaError_t launch_pack_flux(const Geom& g, const Planes& p, int side, float* aos, int lr0, int nrows, bool to_aos, cudaStream_t st);
cudaError_t launch_pack_info(const Geom& g, const Planes& p, int side, float* aos, int lr0, int nrows, bool to_aos, cudaStream_t st);
// TerrainInfo level 0 + (level1 != nullptr) its mip level 1 + (flow != nullptr) the flow map, in one pass over the planar state
cudaError_t launch_publish_fused(const Geom& g, const Planes& p, int side, float* level0, float* level1, uint32_t* flow, cudaStream_t st);
// one level of the RGBA32F mip chain (2x2 box, see aux_kernels.cu): dst (dw x dh texels) <- src (sw x sh texels)
cudaError_t launch_mip_level(const float* src, int sw, int sh, float* dst, int dw, int dh, cudaStream_t st);
// whole grids with sides that are multiples of 64: level 0, the flow map and EVERY mip level in one launch (each CTA reduces
// its 64 x 64 cells through six levels in shared memory, the last CTA to finish — ticket counter in the control block — the rest)
bool publish_all_applicable(const Geom& g, int nlevels);
cudaError_t launch_publish_all(const Geom& g, const Planes& p, int side, float* chain, uint32_t* flow, int nlevels, unsigned int* ticket,
                               cudaStream_t st);
// levels [first, last) of the chain at `base` (level L follows level L-1) in one single-CTA launch: the small tail
cudaError_t launch_mip_tail(float* base, int W, int H, int first, int last, cudaStream_t st);
cudaError_t launch_boundary_outflow(const Geom& g, const Planes& p, int side, double* out, cudaStream_t st);
cudaError_t launch_volume(const Geom& g, const float* d, double* partials, int nblocks, cudaStream_t st);
// copy `nrows` plane rows of `nplanes` planes: dst[i] + dst_row*pitch <- src[i] + src_row*pitch
struct RowCopy { const float* src[6]; float* dst[6]; int nplanes; int src_row; int dst_row; int nrows; int pitch; };
cudaError_t launch_row_copy(const RowCopy& a, const RowCopy& b, int ncopies, cudaStream_t st);
cudaError_t launch_post_flags(volatile uint32_t* up_flag, volatile uint32_t* down_flag, uint32_t value, cudaStream_t st);
cudaError_t launch_wait_flags(Control* ctrl, int wait_up, int wait_down, uint32_t value, cudaStream_t st);

}  // namespace tws
