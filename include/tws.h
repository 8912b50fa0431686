/* tws.h — C ABI of libtws.so, the B200-native replacement for the simulation half of
 * terrainwatersim's `Terrain` class (terrainwatersim/source/scene/Terrain.{h,cpp}) and
 * the three compute shaders it dispatches (shader/flowUpdate.comp, flowApply.comp,
 * waterBrush.comp).  The reference has no FFI layer: the boundary it offers is the
 * public half of `Terrain` (Terrain.h:18-39,76-83) plus the GL textures the renderer
 * samples (Terrain.cpp:288,323-330).  Each entry point below names the reference
 * interface it replaces.  All arithmetic happens on the GPU (sm_100a); there is no CPU
 * fallback: tws_create fails with TWS_ERR_CUDA when no usable device exists.
 *
 * Conventions: plain C types only; every call returns a tws_status (0 = OK, negative =
 * error) and never throws or aborts; the message of the last failure on a sim is
 * returned by tws_last_error.  A tws_sim is not re-entrant (the reference is single
 * threaded, RenderWindow.cpp:103-108): the caller serialises calls on one sim.  Calls
 * enqueue work on library-owned CUDA streams and return without synchronising, except
 * tws_readback / tws_total_volume / tws_sync / tws_elapsed_ms.
 *
 * Grid layout seen through this ABI is the reference's: row-major, index x + y*width
 * (Terrain.cpp:216), y = 0 the first row.  A sim may own only a STRIP of rows
 * [row_begin,row_end) of a larger global grid (multi-GPU row decomposition, one
 * process or thread per GPU); host buffers passed to upload/readback then hold the
 * strip's own rows only.
 */
#ifndef TWS_H_
#define TWS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TWS_ABI_VERSION 1

typedef struct tws_sim tws_sim;

typedef enum tws_status {
  TWS_OK = 0,
  TWS_ERR_INVALID = -1,     /* bad argument / precondition                        */
  TWS_ERR_CUDA = -2,        /* CUDA runtime/driver failure, or no sm_100 device   */
  TWS_ERR_NOMEM = -3,       /* device allocation failed                           */
  TWS_ERR_STATE = -4,       /* call not valid in the sim's current state          */
  TWS_ERR_UNSUPPORTED = -5  /* feature compiled out (e.g. GL interop)             */
} tws_status;

/* Which kernels run a step.  All of them produce bit-identical state. */
typedef enum tws_backend {
  TWS_BACKEND_AUTO = 0,     /* default: chosen at tws_create from the (strip's) cell count — the
                               band kernel with 4 steps per launch from ~12 M cells up, the
                               tile kernel with 2 below (measured crossover ~3072^2 on B200);
                               on whole grids that fit on chip (see TWS_BACKEND_RESIDENT) calls of
                               3 or more steps run as one resident launch instead;
                               `temporal_block` is ignored; see tws_backend_in_use          */
  TWS_BACKEND_UNFUSED = 1,  /* two kernels per step, mirroring the two dispatches of
                               Terrain.cpp:255-264 (in-place, no shared memory)          */
  TWS_BACKEND_FUSED = 2,    /* one TMA-staged shared-memory stencil kernel per step      */
  TWS_BACKEND_FUSED_TB = 3, /* fused + temporal blocking: `temporal_block` steps per HBM
                               round trip, overlapped square tiles (halo rows recomputed) */
  TWS_BACKEND_STREAM_TB = 4,/* fused + temporal blocking as a row-streaming pipeline: one warp
                               per grid row, rows skewed in time instead of recomputed,
                               `temporal_block` steps per HBM round trip                  */
  TWS_BACKEND_BAND_TB = 5,  /* the same skewed row streaming run in lock step: bands of rows,
                               one group barrier per half-pass instead of per-row barriers,
                               dynamic piece schedule, strip exchange fused into the launch */
  TWS_BACKEND_RESIDENT = 6  /* grids that fit in the shared memory of the SMs (up to ~1 M cells: the
                               reference's 1024 x 1024): ALL n steps of a tws_step / tws_advance call
                               in one cooperative launch, one block of the grid resident per SM, rims
                               exchanged neighbour to neighbour through L2.  Whole grids only;
                               tws_create returns TWS_ERR_UNSUPPORTED for larger grids and strips */
} tws_backend;

typedef enum tws_boundary {
  TWS_BOUNDARY_REFERENCE_OPEN = 0, /* the reference: texels outside the grid read as 0
                                      (flowUpdate.comp:18 / flowApply.comp:20 with
                                      out-of-range imageLoad), water drains off the map */
  TWS_BOUNDARY_CLOSED = 1          /* EXTENSION: no flow across the grid edge            */
} tws_boundary;

/* Fields for upload / readback.  Element counts are per cell of the sim's own rows. */
typedef enum tws_field {
  TWS_FIELD_TERRAIN = 0,      /* float  x1  terrain height   = TerrainData.r (Terrain.cpp:216) */
  TWS_FIELD_WATER = 1,        /* float  x1  water depth      = TerrainData.a (Terrain.cpp:219) */
  TWS_FIELD_FLUX = 2,         /* float  x4  outflow (+X,-X,+Y,-Y) = m_waterOutgoingFlow RGBA32F
                                            (Terrain.cpp:230-234, flowUpdate.comp:44-47)        */
  TWS_FIELD_VELOCITY = 3,     /* half   x2  flow vector      = m_waterFlowMap RG16F
                                            (Terrain.cpp:236-237, flowApply.comp:45-52);
                                            readback only, undefined before the first step     */
  TWS_FIELD_TERRAIN_INFO = 4  /* float  x4  (r=terrain, g=b=0.3, a=water) = m_terrainData
                                            RGBA32F exactly as the renderer samples it;
                                            upload ignores g,b                                  */
} tws_field;

/* Creation parameters.  Set `size = sizeof(tws_params)`; use tws_default_params first. */
typedef struct tws_params {
  uint32_t size;
  int32_t  width;             /* global grid width  (m_gridResolution, Terrain.cpp:23)       */
  int32_t  height;            /* global grid height (the reference is square)                */
  int32_t  row_begin;         /* first global row owned by this sim (0 for a whole grid)     */
  int32_t  row_end;           /* one past the last owned row (height for a whole grid)       */
  float    world_size;        /* m_gridWorldSize, Terrain.cpp:22 — cellDistance = world_size/width */
  float    steps_per_second;  /* Terrain::SetSimulationStepsPerSecond, default 60            */
  float    flow_damping;      /* Terrain::SetFlowDamping,        default 0.98                */
  float    flow_acceleration; /* Terrain::SetFlowAcceleration,   default 10                  */
  int32_t  boundary;          /* tws_boundary                                                */
  int32_t  backend;           /* tws_backend                                                 */
  int32_t  temporal_block;    /* steps fused per launch for FUSED_TB / STREAM_TB / BAND_TB (1..4); ignored otherwise */
  int32_t  device;            /* CUDA device ordinal                                         */
  float    rain_rate;         /* EXTENSION: uniform depth added per second (0 = off)         */
  float    evaporation_rate;  /* EXTENSION: uniform depth removed per second (0 = off)       */
} tws_params;

/* The three per-step scalars of the `SimulationParameters` UBO (simulationCommon.glsl:1-13). */
typedef struct tws_step_constants {
  float flow_friction_per_step;      /* powf(damping, (float)dt)                  Terrain.cpp:190 */
  float water_acceleration_per_step; /* (float)(dt * acceleration * cellDistance) Terrain.cpp:197 */
  float cell_area_inv_time_scaled;   /* (float)(dt / cellDistance^2)              Terrain.cpp:184 */
} tws_step_constants;

/* Peer wiring of one strip for multi-GPU runs (see tws_halo_* below). */
#define TWS_IPC_HANDLE_BYTES 64
typedef struct tws_halo_handle {
  uint8_t  mem[TWS_IPC_HANDLE_BYTES];   /* cudaIpcMemHandle_t of the strip's state slab   */
  uint64_t slab_bytes;
  int32_t  row_begin, row_end;          /* for validation by the importer                  */
  int32_t  device;
  int32_t  pid;
  uint64_t local_ptr;                   /* slab address, valid inside process `pid` only  */
} tws_halo_handle;

const char* tws_version(void);
int32_t     tws_abi_version(void);

/* Reference defaults: 1024x1024, world 1024, 60 steps/s, damping 0.98, acceleration 10
 * (Terrain.cpp:21-30), open boundary, whole grid on device 0, backend TWS_BACKEND_AUTO (resolved at tws_create
 * from the cell count — several steps per launch; see tws_backend_in_use). */
void tws_default_params(tws_params* p);

/* Replaces Terrain::Terrain + texture creation (Terrain.cpp:21-121,200-238): allocates
 * the state on the device, zero-initialised (flux = 0 as Terrain.cpp:230-234). */
tws_status tws_create(const tws_params* p, tws_sim** out);
tws_status tws_destroy(tws_sim* s);
const char* tws_last_error(const tws_sim* s);   /* s may be NULL: last tws_create failure */

/* Replaces Set{SimulationStepsPerSecond,FlowDamping,FlowAcceleration} (Terrain.cpp:175-198).
 * A negative/NaN argument is TWS_ERR_INVALID.  Takes effect at the next step, like the
 * UBO upload at the next BindBuffer (UniformBuffer.cpp:122-153). */
tws_status tws_set_steps_per_second(tws_sim* s, float steps_per_second);
tws_status tws_set_flow_damping(tws_sim* s, float damping);
tws_status tws_set_flow_acceleration(tws_sim* s, float acceleration);
tws_status tws_set_sources(tws_sim* s, float rain_rate, float evaporation_rate);   /* EXTENSION */
tws_status tws_get_step_constants(const tws_sim* s, tws_step_constants* out);

/* Replaces Texture2D::SetData uploads (Terrain.cpp:223,233).  `host` holds the sim's own
 * rows, tightly packed.  Uploading TERRAIN/WATER/FLUX does not touch the other fields. */
tws_status tws_upload(tws_sim* s, tws_field field, const void* host, size_t bytes);
tws_status tws_readback(tws_sim* s, tws_field field, void* host, size_t bytes);

/* Replaces Terrain::CreateHeightmapFromNoiseAndResetSim (Terrain.cpp:200-238) with
 * Random::Init(seed) (Application.cpp:57) in front: value-noise terrain (octaves
 * lo..hi, persistence; reference 2,10,0.43), central lake, flux = 0 — generated on the
 * GPU, bit-identical to the reference's CPU fill. */
tws_status tws_reset_reference_scene(tws_sim* s, uint32_t seed, float height_scale,
                                     int32_t octave_lo, int32_t octave_hi, float persistence);
/* EXTENSION: the same generator evaluated for a width x tile_height grid and repeated every
 * tile_height rows (0 = height: the reference scene).  Weak-scaling workloads use it so that
 * every strip of tile_height rows holds the identical scene. */
tws_status tws_reset_reference_scene_tiled(tws_sim* s, uint32_t seed, float height_scale,
                                           int32_t low_octave, int32_t high_octave, float persistence,
                                           int32_t tile_height);

/* Replaces Terrain::ApplyRadialWaterBrush (Terrain.cpp:150-168) + waterBrush.comp:
 * tws_inject_brush_world takes the world XZ position like the reference and derives the
 * texel centre (Terrain.cpp:152-155); tws_inject_brush takes the texel-space centre
 * (global coordinates).  Only the brush's bounding box is touched, result identical to
 * the reference's whole-grid pass.  size_sq is 32 in the reference (Terrain.cpp:159).
 * On a whole grid stepped by the tile or the resident kernel the brush costs no launch of its own: the
 * next tws_step / tws_advance applies it while loading the depth; every other call that reads or writes
 * the state applies it first, so it is always as if the brush had run at once. */
tws_status tws_inject_brush(tws_sim* s, float center_x, float center_y, float intensity, float size_sq);
tws_status tws_inject_brush_world(tws_sim* s, float world_x, float world_z, float strength);

/* n x (flowUpdate; flowApply) — the loop body of Terrain.cpp:253-265, no clamp on n. */
tws_status tws_step(tws_sim* s, int32_t n);

/* One step (the loop body of Terrain.cpp:253-265) for a caller whose water layer lives in HOST
 * memory: uploads `water_in` (width x own rows floats, tightly packed; NULL = keep the resident
 * water), runs one step, and returns the new water depth in `water_out` (floats) and the flow
 * vector in `velocity_out` (2 x fp16 per cell, the RG16F texel of m_waterFlowMap); either output
 * may be NULL.  water_out may alias water_in.  Terrain and flux stay resident on the device, as in
 * the reference whose state never leaves the GPU.  The call is pipelined in row bands over three
 * streams — band b+1 uploads while band b computes and band b-1 reads back, so both directions
 * of the PCIe link are busy at once — and returns when the outputs are in host memory.  Host
 * buffers should be page-locked (cudaHostAlloc / cudaHostRegister); pageable memory is correct
 * but serialises the copies.  Bit-identical to tws_upload + tws_step(1) + tws_readback. */
tws_status tws_step_host(tws_sim* s, const float* water_in, float* water_out, void* velocity_out);

/* Replaces Terrain::PerformSimulationStep(ezTime) (Terrain.cpp:240-277): frame-time
 * accumulator, n = (uint)(acc/dt), acc -= dt*n, n = min(n,10); runs n steps and returns
 * n through steps_done (may be NULL). */
tws_status tws_advance(tws_sim* s, double frame_seconds, uint32_t* steps_done);

/* fp64 sum of the water depth over the sim's own rows (volume check). Synchronises. */
tws_status tws_total_volume(tws_sim* s, double* volume);
tws_status tws_sync(tws_sim* s);
/* Diagnostic for mass ledgers on the open (reference) boundary: fp64 sum of the outflow currently
 * stored in the flux field that points OUT of the global grid through this sim's part of the edge
 * (+X at x = width-1, -X at x = 0, -Y on global row 0, +Y on the last global row).  Multiplied by
 * cell_area_inv_time_scaled it is the volume the last step drained off the map
 * (flowApply.comp:38-41 with the exterior reading 0).  Synchronises. */
tws_status tws_boundary_outflow(tws_sim* s, double* flux_sum);
/* EXTENSION, the ledger that works with k steps per launch: the step kernels themselves accumulate, in fp64 and in every
 * sub-step, (outflow across the edge of the global grid) x cell_area_inv_time_scaled from the lanes that own an edge
 * cell.  `volume` = the water volume that has left the map through this sim's part of the edge since creation, the last
 * tws_reset_reference_scene or tws_boundary_outflow_reset (always 0 with TWS_BOUNDARY_CLOSED).  With rain and
 * evaporation off, total_volume(t) + this == total_volume(0) up to fp32 rounding of the depth updates; SURVEY.md 8d
 * config 5 closes the ledger V0 + sources - outflow with it.  The sum order of the atomics is not fixed: equal to the
 * oracle's per-step fp64 sum to ~1e-15 relative, not bitwise.  tws_boundary_outflow_accumulated synchronises. */
tws_status tws_boundary_outflow_accumulated(tws_sim* s, double* volume);
/* The other side of the ledger: the volume rain and evaporation REALLY changed since the same point in time — the fp64 sum,
 * over every cell of the sim's own rows and every sub-step, of (depth after the source terms - depth before them) as the
 * fp32 arithmetic produced it.  This is not rain_rate x time x area: d + rain_step - evap_step rounds the same way for every
 * cell of a binade, and evaporation is clamped at dry cells; over 10 000 steps the difference reaches 1e-4 of the volume.
 * total_volume(t) == total_volume(0) + tws_source_accumulated - tws_boundary_outflow_accumulated up to the (unbiased) fp32
 * rounding of the flux updates.  0 while rain and evaporation are off.  Synchronises. */
tws_status tws_source_accumulated(tws_sim* s, double* volume);
/* zeroes both accumulators */
tws_status tws_boundary_outflow_reset(tws_sim* s);

/* Device-side timing of the most recent tws_step/tws_advance batch (CUDA events on the
 * launching stream; replaces gl::TimerQuery around PerformSimulationStep,
 * Scene.cpp:362-364).  Synchronises on the end event. */
tws_status tws_elapsed_ms(tws_sim* s, float* ms);
/* The same without ever blocking — what the reference does: its gl::TimerQuery is double buffered
 * (TimerQuery.cpp:29-72) and Scene::Update reads LAST frame's result (Scene.cpp:337-342).  Batches
 * alternate between two event pairs; this returns the time of the newest batch the GPU has already
 * finished (the most recent one, else the one before) and its 0-based index through batch_index (may
 * be NULL); TWS_ERR_STATE when neither has finished yet.  Call it once per frame after tws_advance. */
tws_status tws_elapsed_ms_nowait(tws_sim* s, float* ms, uint64_t* batch_index);
/* Number of kernels launched by this sim since creation (kernels inside a replayed batch graph count). */
uint64_t   tws_kernel_launches(const tws_sim* s);
/* Frame scheduler: a whole-grid sim captures each batch size n (2..64) of tws_step / tws_advance as
 * one CUDA graph per ping-pong side and replays it (the reference issues ~12 GL calls per step,
 * Terrain.cpp:253-265, up to 10 steps per frame).  Returns how many batches ran as a graph replay.
 * Environment TWS_GRAPHS=0 disables capture (every kernel launched individually). */
uint64_t   tws_graph_replays(const tws_sim* s);
/* The backend and steps-per-launch this sim runs with (what TWS_BACKEND_AUTO resolved to). */
tws_status tws_backend_in_use(const tws_sim* s, int32_t* backend, int32_t* temporal_block);
/* Raw device pointer and row pitch (in elements) of a planar field for zero-copy
 * consumers (CUDA interop); planes: TERRAIN, WATER (current), VELOCITY. */
tws_status tws_device_view(tws_sim* s, tws_field field, void** device_ptr, int64_t* pitch_elems);

/* ---- multi-GPU strips: halo exchange over NVLink peer memory ----------------------
 * One sim per GPU owns rows [row_begin,row_end).  After creation each sim exports a
 * handle; the owner passes the handles of the strip above (rows < row_begin) and below
 * (rows >= row_end) to tws_halo_connect (NULL = global edge).  Sims in one process use
 * cudaDeviceEnablePeerAccess, sims in different processes CUDA IPC.  During tws_step
 * each strip stores its freshly computed edge rows also into the neighbours' halo rows
 * (peer stores over NVLink from inside the step kernel) and raises a step flag in the
 * neighbour's memory; the neighbour's edge rows wait on that flag while its interior
 * rows are computed.  tws_halo_refresh pushes the current state (after upload / reset /
 * inject); call it on all strips, then tws_sync + a host barrier.  Strips that share
 * one device are a test configuration: their launches wait for each other on the GPU with
 * spinning flag kernels, so every stream needs its own hardware queue — run such a process
 * with CUDA_DEVICE_MAX_CONNECTIONS=32 (each sim owns up to four streams). */
tws_status tws_halo_export(tws_sim* s, tws_halo_handle* out);
tws_status tws_halo_connect(tws_sim* s, const tws_halo_handle* up, const tws_halo_handle* down);
tws_status tws_halo_refresh(tws_sim* s);

/* ---- renderer hand-off (CUDA-GL interop) --------------------------------------------
 * Replaces the renderer's view of the simulation textures: m_terrainData (TerrainInfo, RGBA32F
 * with the full mip chain the reference regenerates after every stepped frame,
 * Terrain.cpp:272-276 + glEasy Texture2D.cpp:64-68) and m_waterFlowMap (FlowMap, RG16F, one
 * level), sampled at Terrain.cpp:288,323-330.
 *
 * tws_gl_register takes the GL names of the renderer's two textures (flow_map_tex may be 0) and
 * registers them with cudaGraphicsGLRegisterImage; the renderer's GL context must be current
 * on the calling thread (otherwise TWS_ERR_CUDA).  tws_gl_publish writes (r = terrain,
 * g = b = 0.3, a = water) plus the mip chain and the flow vectors into them.  tws_gl_* are for
 * whole grids (the renderer samples one texture; a strip returns TWS_ERR_UNSUPPORTED).
 *
 * Strips (multi-GPU): tws_publish_packed / tws_publish_mips work per strip.  A strip builds the levels
 * it can filter from its own rows alone — level l needs row_begin and the row count to be multiples
 * of 2^l (plan_strips cuts on multiples of 8 rows: levels 0..3 at least); `levels` returns how many.
 * Level l of a strip is max(1, width >> l) x (rows >> l) texels = rows [row_begin >> l, row_end >> l)
 * of the whole grid's level l, bit for bit.  The renderer's GPU gathers the strips' last common level
 * (1/64 of level 0 at l = 3) and filters the remaining small levels itself with the rule below.
 *
 * The same data without GL: tws_publish_packed packs level 0 and the flow map into device
 * buffers the library owns; tws_publish_mips additionally builds the mip chain, level L
 * following level L-1 contiguously in one buffer (sizes / offsets: tws_mip_level_info);
 * tws_readback_mip copies one published level to the host.  Mip rule (GL leaves the filter
 * to the driver; pinned): tws_mip_levels = floor(log2(max(w,h))) + 1 levels (glEasy
 * Texture.cpp:28-41), level L is max(1, w >> L) x max(1, h >> L), each texel the 2x2 box
 * average ((t00 + t10) + (t01 + t11)) * 0.25 of level L-1, source coordinates clamped. */
tws_status tws_gl_register(tws_sim* s, uint32_t terrain_info_tex, uint32_t flow_map_tex);
tws_status tws_gl_publish(tws_sim* s);
tws_status tws_gl_unregister(tws_sim* s);
tws_status tws_publish_packed(tws_sim* s, void** terrain_info_rgba32f_dev, void** flow_map_rg16f_dev);
tws_status tws_publish_mips(tws_sim* s, void** terrain_info_chain_rgba32f_dev, int32_t* levels);
tws_status tws_readback_mip(tws_sim* s, int32_t level, void* host, size_t bytes);
int32_t    tws_mip_levels(int32_t width, int32_t height);
tws_status tws_mip_level_info(int32_t width, int32_t height, int32_t level, int32_t* w, int32_t* h, int64_t* offset_texels);

#ifdef __cplusplus
}
#endif
#endif /* TWS_H_ */
