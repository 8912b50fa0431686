"""Host-side mirror of the reference's `Terrain` simulation interface over libtws.so.

Method names, argument meaning and order of operations follow
terrainwatersim/source/scene/Terrain.h:18-39,76-83 so that tests read like calls into the
reference: `PerformSimulationStep(lastFrameDuration)`, `ApplyRadialWaterBrush(worldXZ,
strength)`, `SetSimulationStepsPerSecond/SetFlowDamping/SetFlowAcceleration`,
`CreateHeightmapFromNoiseAndResetSim()`.  pythonic helpers (upload/readback as numpy
arrays, step(n)) sit beside them.  All arithmetic happens inside the CUDA library.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _abi


class TwsError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"tws status {status}: {message}")
        self.status = status


REFERENCE_SEED = 231656522       # Random::Init, Application.cpp:57
REFERENCE_HEIGHT_SCALE = 300.0   # m_heightScale, Terrain.cpp:25


class Terrain:
    """One simulation grid (or one row strip of it) resident on a B200."""

    def __init__(self, gridResolution: int = 1024, gridWorldSize: Optional[float] = None, *, height: Optional[int] = None,
                 rows: Optional[Tuple[int, int]] = None, backend: int = _abi.BACKEND_AUTO, temporal_block: int = 1,
                 boundary: int = _abi.BOUNDARY_REFERENCE_OPEN, device: int = 0, simulationStepsPerSecond: float = 60.0,
                 flowDamping: float = 0.98, flowAcceleration: float = 10.0, rain_rate: float = 0.0,
                 evaporation_rate: float = 0.0):
        self._lib = _abi.load()
        self._sim = _abi._SIM()
        p = _abi.TwsParams()
        self._lib.tws_default_params(C.byref(p))
        p.width = int(gridResolution)
        p.height = int(height if height is not None else gridResolution)
        p.row_begin, p.row_end = (0, p.height) if rows is None else (int(rows[0]), int(rows[1]))
        p.world_size = float(gridWorldSize if gridWorldSize is not None else gridResolution)   # Terrain.cpp:22-23: 1024/1024
        p.steps_per_second = simulationStepsPerSecond
        p.flow_damping = flowDamping
        p.flow_acceleration = flowAcceleration
        p.boundary = boundary
        p.backend = backend
        p.temporal_block = temporal_block
        p.device = device
        p.rain_rate = rain_rate
        p.evaporation_rate = evaporation_rate
        self.params = p
        st = self._lib.tws_create(C.byref(p), C.byref(self._sim))
        if st != _abi.TWS_OK:
            msg = self._lib.tws_last_error(None)
            self._sim = _abi._SIM()
            raise TwsError(st, msg.decode() if msg else "tws_create failed")
        self.width, self.height = p.width, p.height
        self.row_begin, self.row_end = p.row_begin, p.row_end
        self.rows = p.row_end - p.row_begin

    # ---- plumbing --------------------------------------------------------------------
    def _check(self, st: int) -> None:
        if st != _abi.TWS_OK:
            msg = self._lib.tws_last_error(self._sim)
            raise TwsError(st, msg.decode() if msg else "")

    def close(self) -> None:
        if getattr(self, "_sim", None) is not None and self._sim.value:
            self._lib.tws_destroy(self._sim)
            self._sim = _abi._SIM()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- the reference's interface (Terrain.h) ---------------------------------------------
    def PerformSimulationStep(self, lastFrameDuration: float) -> int:
        """Terrain.cpp:240-277.  Returns the number of steps run (0..10)."""
        n = C.c_uint32(0)
        self._check(self._lib.tws_advance(self._sim, float(lastFrameDuration), C.byref(n)))
        return int(n.value)

    def ApplyRadialWaterBrush(self, worldPositionXZ, strength: float) -> None:
        """Terrain.cpp:150-168 (strength = frame seconds * 100 at the call site, Scene.cpp:358)."""
        self._check(self._lib.tws_inject_brush_world(self._sim, float(worldPositionXZ[0]), float(worldPositionXZ[1]), float(strength)))

    def SetSimulationStepsPerSecond(self, v: float) -> None:
        self._check(self._lib.tws_set_steps_per_second(self._sim, float(v)))

    def SetFlowDamping(self, v: float) -> None:
        self._check(self._lib.tws_set_flow_damping(self._sim, float(v)))

    def SetFlowAcceleration(self, v: float) -> None:
        self._check(self._lib.tws_set_flow_acceleration(self._sim, float(v)))

    def CreateHeightmapFromNoiseAndResetSim(self, seed: int = REFERENCE_SEED, heightScale: float = REFERENCE_HEIGHT_SCALE,
                                            lowOctave: int = 2, highOctave: int = 10, persistence: float = 0.43,
                                            tileHeight: int = 0) -> None:
        """Terrain.cpp:200-238 preceded by Random::Init(seed) (Application.cpp:57).  tileHeight (extension): the scene
        of a width x tileHeight grid repeated every tileHeight rows (0: the reference scene over the whole grid)."""
        self._check(self._lib.tws_reset_reference_scene_tiled(self._sim, seed, heightScale, lowOctave, highOctave, persistence, int(tileHeight)))

    # ---- helpers ------------------------------------------------------------------------------
    def step(self, n: int = 1) -> None:
        self._check(self._lib.tws_step(self._sim, int(n)))

    def step_host(self, water_in: Optional[np.ndarray], water_out: Optional[np.ndarray] = None,
                  velocity_out: Optional[np.ndarray] = None) -> None:
        """tws_step_host: one step with the water layer in host memory (band-pipelined upload / step /
        readback).  Arrays must be C-contiguous (rows, width) float32 / (rows, width, 2) float16."""
        def ptr(a, dt, tail):
            if a is None:
                return None
            if a.dtype != dt or a.shape != (self.rows, self.width) + tail or not a.flags.c_contiguous:
                raise ValueError(f"expected C-contiguous {dt} array of shape {(self.rows, self.width) + tail}")
            return a.ctypes.data_as(C.c_void_p)
        self._check(self._lib.tws_step_host(self._sim, ptr(water_in, np.float32, ()), ptr(water_out, np.float32, ()),
                                            ptr(velocity_out, np.float16, (2,))))

    def step_host_raw(self, water_in_ptr: int, water_out_ptr: int, velocity_out_ptr: int) -> None:
        self._check(self._lib.tws_step_host(self._sim, C.c_void_p(water_in_ptr or None), C.c_void_p(water_out_ptr or None),
                                            C.c_void_p(velocity_out_ptr or None)))

    def inject_brush(self, cx: float, cy: float, intensity: float, size_sq: float = 32.0) -> None:
        self._check(self._lib.tws_inject_brush(self._sim, cx, cy, intensity, size_sq))

    def set_sources(self, rain_rate: float, evaporation_rate: float) -> None:
        self._check(self._lib.tws_set_sources(self._sim, rain_rate, evaporation_rate))

    def sync(self) -> None:
        self._check(self._lib.tws_sync(self._sim))

    def elapsed_ms(self) -> float:
        ms = C.c_float(0)
        self._check(self._lib.tws_elapsed_ms(self._sim, C.byref(ms)))
        return float(ms.value)

    def elapsed_ms_nowait(self):
        """(ms, batch index) of the newest step batch the GPU has finished, or None — never blocks (gl::TimerQuery)."""
        ms, idx = C.c_float(0), C.c_uint64(0)
        st = self._lib.tws_elapsed_ms_nowait(self._sim, C.byref(ms), C.byref(idx))
        if st == _abi.TWS_ERR_STATE:
            return None
        self._check(st)
        return float(ms.value), int(idx.value)

    def kernel_launches(self) -> int:
        return int(self._lib.tws_kernel_launches(self._sim))

    def backend_in_use(self) -> Tuple[int, int]:
        """(backend, steps per launch) — what BACKEND_AUTO resolved to."""
        b, k = C.c_int32(0), C.c_int32(0)
        self._check(self._lib.tws_backend_in_use(self._sim, C.byref(b), C.byref(k)))
        return int(b.value), int(k.value)

    def graph_replays(self) -> int:
        return int(self._lib.tws_graph_replays(self._sim))

    def step_constants(self) -> Tuple[float, float, float]:
        c = _abi.TwsStepConstants()
        self._check(self._lib.tws_get_step_constants(self._sim, C.byref(c)))
        return (c.flow_friction_per_step, c.water_acceleration_per_step, c.cell_area_inv_time_scaled)

    def total_volume(self) -> float:
        v = C.c_double(0)
        self._check(self._lib.tws_total_volume(self._sim, C.byref(v)))
        return float(v.value)

    def boundary_outflow(self) -> float:
        """fp64 sum of the flux leaving the global grid through this strip's edge cells."""
        v = C.c_double(0)
        self._check(self._lib.tws_boundary_outflow(self._sim, C.byref(v)))
        return float(v.value)

    def boundary_outflow_accumulated(self) -> float:
        """Volume that has left the map through this strip's part of the edge since creation / reset (fp64, accumulated
        inside the step kernels in every sub-step — works with k steps per launch)."""
        v = C.c_double(0)
        self._check(self._lib.tws_boundary_outflow_accumulated(self._sim, C.byref(v)))
        return float(v.value)

    def source_accumulated(self) -> float:
        """Volume rain and evaporation really changed since creation / reset (fp64 sum of the fp32 source deltas, accumulated
        inside the step kernels)."""
        v = C.c_double(0)
        self._check(self._lib.tws_source_accumulated(self._sim, C.byref(v)))
        return float(v.value)

    def boundary_outflow_reset(self) -> None:
        self._check(self._lib.tws_boundary_outflow_reset(self._sim))

    _SHAPES = {
        _abi.FIELD_TERRAIN: (np.float32, ()),
        _abi.FIELD_WATER: (np.float32, ()),
        _abi.FIELD_FLUX: (np.float32, (4,)),
        _abi.FIELD_VELOCITY: (np.float16, (2,)),
        _abi.FIELD_TERRAIN_INFO: (np.float32, (4,)),
    }

    def upload(self, field: int, array: np.ndarray) -> None:
        dt, tail = self._SHAPES[field]
        a = np.ascontiguousarray(array, dtype=dt)
        if a.shape != (self.rows, self.width) + tail:
            raise ValueError(f"expected shape {(self.rows, self.width) + tail}, got {a.shape}")
        self._check(self._lib.tws_upload(self._sim, field, a.ctypes.data_as(C.c_void_p), a.nbytes))

    def readback(self, field: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        dt, tail = self._SHAPES[field]
        if out is None:
            out = np.empty((self.rows, self.width) + tail, dtype=dt)
        self._check(self._lib.tws_readback(self._sim, field, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def upload_raw(self, field: int, ptr: int, nbytes: int) -> None:
        self._check(self._lib.tws_upload(self._sim, field, C.c_void_p(ptr), nbytes))

    def readback_raw(self, field: int, ptr: int, nbytes: int) -> None:
        self._check(self._lib.tws_readback(self._sim, field, C.c_void_p(ptr), nbytes))

    # ---- renderer hand-off ------------------------------------------------------------------------
    def publish_packed(self) -> Tuple[int, int]:
        """Device pointers of the packed TerrainInfo (RGBA32F) and FlowMap (RG16F) level-0 images."""
        info, flow = C.c_void_p(), C.c_void_p()
        self._check(self._lib.tws_publish_packed(self._sim, C.byref(info), C.byref(flow)))
        return int(info.value), int(flow.value)

    def publish_mips(self) -> list:
        """GenMipMaps of m_terrainData (Terrain.cpp:272-276): builds the chain on the device and
        returns every level as a (h, w, 4) float32 host array."""
        base, levels = C.c_void_p(), C.c_int32(0)
        self._check(self._lib.tws_publish_mips(self._sim, C.byref(base), C.byref(levels)))
        out = []
        for l in range(levels.value):
            w, h, off = C.c_int32(), C.c_int32(), C.c_int64()
            self._check(self._lib.tws_mip_level_info(self.width, self.rows, l, C.byref(w), C.byref(h), C.byref(off)))
            a = np.empty((h.value, w.value, 4), np.float32)
            self._check(self._lib.tws_readback_mip(self._sim, l, a.ctypes.data_as(C.c_void_p), a.nbytes))
            out.append(a)
        return out

    def gl_register(self, terrain_info_tex: int, flow_map_tex: int = 0) -> None:
        self._check(self._lib.tws_gl_register(self._sim, int(terrain_info_tex), int(flow_map_tex)))

    def gl_publish(self) -> None:
        self._check(self._lib.tws_gl_publish(self._sim))

    def gl_unregister(self) -> None:
        self._check(self._lib.tws_gl_unregister(self._sim))

    # ---- strips ---------------------------------------------------------------------------------
    def halo_export(self) -> bytes:
        h = _abi.TwsHaloHandle()
        self._check(self._lib.tws_halo_export(self._sim, C.byref(h)))
        return bytes(h)

    def halo_connect(self, up: Optional[bytes], down: Optional[bytes]) -> None:
        hu = _abi.TwsHaloHandle.from_buffer_copy(up) if up is not None else None
        hd = _abi.TwsHaloHandle.from_buffer_copy(down) if down is not None else None
        self._check(self._lib.tws_halo_connect(self._sim, C.byref(hu) if hu is not None else None,
                                               C.byref(hd) if hd is not None else None))

    def halo_refresh(self) -> None:
        self._check(self._lib.tws_halo_refresh(self._sim))
