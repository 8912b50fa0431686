"""ctypes wrapper of oracle/libtws_oracle*.so (see tws_oracle.cpp for the contract).

TEST INFRASTRUCTURE — the checker, never the thing shipped.  State uses the reference's
own texture layouts: TerrainData (H,W,4) float32 with r = terrain, a = water; Flow (H,W,4)
float32 (+X,-X,+Y,-Y); FlowMap (H,W,2) float16.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path
from typing import Optional, Tuple

import numpy as np

ORACLE_DIR = Path(__file__).resolve().parent
REFERENCE_SEED = 231656522


def build(force: bool = False) -> None:
    """make -C oracle (also builds oracle/_ref when /root/reference is present)."""
    need = force or not (ORACLE_DIR / "libtws_oracle.so").exists() or not (ORACLE_DIR / "libtws_oracle_omp.so").exists() \
        or (ORACLE_DIR / "tws_oracle.cpp").stat().st_mtime > (ORACLE_DIR / "libtws_oracle.so").stat().st_mtime
    ref_missing = Path("/root/reference/terrainwatersim/source/math").is_dir() and not (
        (ORACLE_DIR / "_ref" / "libtws_ref_terrain.so").exists() and (ORACLE_DIR / "_ref" / "libtws_ref_step.so").exists()
        and (ORACLE_DIR / "_ref" / "libtws_ref_step.so").stat().st_mtime >= max(
            (ORACLE_DIR / f).stat().st_mtime for f in ("ref_step_driver.cpp", "ref_shim/glsl.h", "ref_shim/glsl_prep.py")))
    if need or ref_missing:
        res = subprocess.run(["make", "-C", str(ORACLE_DIR)] + (["-B"] if force else []), capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("building the oracle failed:\n" + res.stdout + res.stderr)


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self, openmp: bool = False):
        build()
        self.lib = C.CDLL(str(ORACLE_DIR / ("libtws_oracle_omp.so" if openmp else "libtws_oracle.so")))
        L = self.lib
        L.tws_oracle_threads.restype = C.c_int
        L.tws_oracle_fnv1a64.restype = C.c_uint64
        L.tws_oracle_fnv1a64.argtypes = [C.c_void_p, C.c_uint64]
        L.tws_oracle_float_to_half.restype = C.c_uint16
        L.tws_oracle_float_to_half.argtypes = [C.c_float]
        L.tws_oracle_advance.restype = C.c_uint32
        L.tws_oracle_advance.argtypes = [C.POINTER(C.c_double), C.c_double, C.c_double]
        L.tws_oracle_derive_consts.argtypes = [C.c_float, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.tws_oracle_step.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float]
        L.tws_oracle_flow_update.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.tws_oracle_flow_apply.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float]
        L.tws_oracle_brush.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float]
        L.tws_oracle_brush_center.argtypes = [C.c_float, C.c_float, C.c_float, C.c_uint32, C.c_void_p]
        L.tws_oracle_white_noise.argtypes = [C.c_uint32, C.c_void_p]
        L.tws_oracle_create_scene.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_float, C.c_void_p]

    @property
    def threads(self) -> int:
        return int(self.lib.tws_oracle_threads())

    def set_threads(self, n: int) -> None:
        self.lib.tws_oracle_set_threads(int(n))

    def derive_consts(self, world_size=1024.0, res=1024, steps_per_second=60.0, damping=0.98, acceleration=10.0) -> np.ndarray:
        out = np.zeros(3, np.float32)
        self.lib.tws_oracle_derive_consts(world_size, res, steps_per_second, damping, acceleration, _fp(out))
        return out

    def step(self, terrain: np.ndarray, flow: np.ndarray, flowmap: np.ndarray, consts: np.ndarray, n: int = 1, boundary: int = 0,
             rain_step: float = 0.0, evap_step: float = 0.0) -> None:
        H, W = terrain.shape[:2]
        assert terrain.dtype == np.float32 and flow.dtype == np.float32 and flowmap.dtype == np.float16
        assert terrain.flags.c_contiguous and flow.flags.c_contiguous and flowmap.flags.c_contiguous
        c = np.ascontiguousarray(consts, np.float32)
        self.lib.tws_oracle_step(W, H, _fp(terrain), _fp(flow), _fp(flowmap), _fp(c), n, boundary, rain_step, evap_step)

    def flow_update(self, terrain, flow, consts, boundary=0):
        H, W = terrain.shape[:2]
        c = np.ascontiguousarray(consts, np.float32)
        self.lib.tws_oracle_flow_update(W, H, _fp(terrain), _fp(flow), _fp(c), boundary)

    def flow_apply(self, terrain, flow, flowmap, consts, rain_step=0.0, evap_step=0.0):
        H, W = terrain.shape[:2]
        c = np.ascontiguousarray(consts, np.float32)
        self.lib.tws_oracle_flow_apply(W, H, _fp(terrain), _fp(flow), _fp(flowmap), _fp(c), rain_step, evap_step)

    def brush(self, terrain: np.ndarray, cx: float, cy: float, intensity: float, size_sq: float = 32.0) -> None:
        H, W = terrain.shape[:2]
        self.lib.tws_oracle_brush(W, H, _fp(terrain), cx, cy, intensity, size_sq)

    def brush_center(self, world_x, world_z, world_size, res) -> Tuple[float, float]:
        out = np.zeros(2, np.float32)
        self.lib.tws_oracle_brush_center(world_x, world_z, world_size, res, _fp(out))
        return float(out[0]), float(out[1])

    def advance(self, accumulator: float, step_length: float, frame_seconds: float) -> Tuple[int, float]:
        acc = C.c_double(accumulator)
        n = self.lib.tws_oracle_advance(C.byref(acc), step_length, frame_seconds)
        return int(n), float(acc.value)

    def white_noise(self, seed: int = REFERENCE_SEED) -> np.ndarray:
        out = np.zeros(4096, np.float32)
        self.lib.tws_oracle_white_noise(seed, _fp(out))
        return out

    def create_scene(self, W: int, H: Optional[int] = None, seed: int = REFERENCE_SEED, height_scale: float = 300.0, lo: int = 2,
                     hi: int = 10, persistence: float = 0.43) -> np.ndarray:
        H = W if H is None else H
        out = np.zeros((H, W, 4), np.float32)
        self.lib.tws_oracle_create_scene(seed, W, H, height_scale, lo, hi, persistence, _fp(out))
        return out

    def fnv1a64(self, a: np.ndarray) -> int:
        a = np.ascontiguousarray(a)
        return int(self.lib.tws_oracle_fnv1a64(_fp(a), a.nbytes))

    def float_to_half_bits(self, f: float) -> int:
        return int(self.lib.tws_oracle_float_to_half(f))


class RefStep:
    """ctypes wrapper of oracle/_ref/libtws_ref_step.so: the reference's OWN flowUpdate.comp / flowApply.comp /
    waterBrush.comp compiled from /root/reference through the GLSL shim (oracle/ref_step_driver.cpp).  Same
    texture layouts and call shapes as Oracle, minus the extensions the reference does not have.  Defined where
    the reference is: it dispatches res/16 (step) and res/32 (brush) whole groups (Terrain.cpp:167,258,264)."""

    PATH = ORACLE_DIR / "_ref" / "libtws_ref_step.so"

    def __init__(self):
        build()
        if not self.PATH.exists():
            raise FileNotFoundError(str(self.PATH))
        self.lib = C.CDLL(str(self.PATH))
        L = self.lib
        L.tws_ref_step_threads.restype = C.c_int
        L.tws_ref_step.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.tws_ref_flow_update.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.tws_ref_flow_apply.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.tws_ref_brush.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float]

    def step(self, terrain, flow, flowmap, consts, n=1):
        H, W = terrain.shape[:2]
        assert terrain.dtype == np.float32 and flow.dtype == np.float32 and flowmap.dtype == np.float16
        assert terrain.flags.c_contiguous and flow.flags.c_contiguous and flowmap.flags.c_contiguous
        c = np.ascontiguousarray(consts, np.float32)
        self.lib.tws_ref_step(W, H, _fp(terrain), _fp(flow), _fp(flowmap), _fp(c), n)

    def flow_update(self, terrain, flow, consts):
        H, W = terrain.shape[:2]
        c = np.ascontiguousarray(consts, np.float32)
        self.lib.tws_ref_flow_update(W, H, _fp(terrain), _fp(flow), _fp(c))

    def flow_apply(self, terrain, flow, flowmap, consts):
        H, W = terrain.shape[:2]
        c = np.ascontiguousarray(consts, np.float32)
        self.lib.tws_ref_flow_apply(W, H, _fp(terrain), _fp(flow), _fp(flowmap), _fp(c))

    def brush(self, terrain, cx, cy, intensity, size_sq=32.0):
        H, W = terrain.shape[:2]
        self.lib.tws_ref_brush(W, H, _fp(terrain), cx, cy, intensity, size_sq)


def new_state(terrain_h: np.ndarray, water_d: np.ndarray):
    """(TerrainData, Flow, FlowMap) in the reference layout from planar h, d."""
    H, W = terrain_h.shape
    t = np.empty((H, W, 4), np.float32)
    t[..., 0] = terrain_h
    t[..., 1] = 0.3
    t[..., 2] = 0.3
    t[..., 3] = water_d
    return t, np.zeros((H, W, 4), np.float32), np.zeros((H, W, 2), np.float16)


def dam_break(W: int = 256, H: Optional[int] = None, rim: bool = True):
    """BASELINE config 1: ramp h = 20*x/(W-1), optional 1-cell rim of height 1000,
    40-deep water column for x < W/4 (0 on the rim).  Deterministic, no RNG."""
    H = W if H is None else H
    x = np.arange(W, dtype=np.float32)
    h = np.broadcast_to((np.float32(20.0) * x / np.float32(W - 1)).astype(np.float32), (H, W)).copy()
    d = np.zeros((H, W), np.float32)
    d[:, : W // 4] = 40.0
    if rim:
        for a in (h,):
            a[0, :] = 1000.0; a[-1, :] = 1000.0; a[:, 0] = 1000.0; a[:, -1] = 1000.0
        d[0, :] = 0; d[-1, :] = 0; d[:, 0] = 0; d[:, -1] = 0
    return h, d


def mip_levels(W: int, H: int) -> int:
    """Level count of the reference's texture wrapper (glEasy Texture.cpp:28-41): halve every side until all are 0."""
    n = 0
    while W > 0 or H > 0:
        W //= 2; H //= 2; n += 1
    return n


def mip_chain(level0: np.ndarray) -> list:
    """Restatement of `m_terrainData->GenMipMaps()` (Terrain.cpp:272-276 -> glEasy Texture2D.cpp:64-68,
    glGenerateMipmap) for an (H, W, C) float32 image.  GL leaves the filter to the driver, so it is PINNED
    (DESIGN.md section 5), not taken from the reference: level L has max(1, side >> L) texels per side
    (the glTexStorage2D rule), each the 2x2 box average ((t00 + t10) + (t01 + t11)) * 0.25 of level L-1 in
    binary32, source coordinates clamped to the source."""
    assert level0.dtype == np.float32 and level0.ndim == 3
    out = [level0]
    H, W = level0.shape[:2]
    for _ in range(1, mip_levels(W, H)):
        src = out[-1]
        sh, sw = src.shape[:2]
        dh, dw = max(1, sh >> 1), max(1, sw >> 1)
        y0 = np.minimum(2 * np.arange(dh), sh - 1); y1 = np.minimum(2 * np.arange(dh) + 1, sh - 1)
        x0 = np.minimum(2 * np.arange(dw), sw - 1); x1 = np.minimum(2 * np.arange(dw) + 1, sw - 1)
        top = src[y0][:, x0] + src[y0][:, x1]          # float32 adds, rounded once each
        bot = src[y1][:, x0] + src[y1][:, x1]
        out.append(((top + bot) * np.float32(0.25)).astype(np.float32))
    return out
