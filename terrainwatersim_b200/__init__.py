"""terrainwatersim_b200 — B200-native heightfield shallow-water step.

The product is libtws.so (hand-written sm_100a CUDA behind include/tws.h); this package
is the thin host-side mirror of the reference's `Terrain` interface plus the strip
decomposition helpers used by bench.py and the tests.
"""
from . import _abi
from ._abi import (BACKEND_AUTO, BACKEND_FUSED, BACKEND_FUSED_TB, BACKEND_STREAM_TB, BACKEND_UNFUSED, BACKEND_BAND_TB, BACKEND_RESIDENT, BOUNDARY_CLOSED, BOUNDARY_REFERENCE_OPEN, FIELD_FLUX,
                   FIELD_TERRAIN, FIELD_TERRAIN_INFO, FIELD_VELOCITY, FIELD_WATER)
from .strips import StripPlan, connect_strips, plan_strips
from .terrain import REFERENCE_HEIGHT_SCALE, REFERENCE_SEED, Terrain, TwsError

__all__ = [
    "Terrain", "TwsError", "StripPlan", "plan_strips", "connect_strips", "REFERENCE_SEED", "REFERENCE_HEIGHT_SCALE",
    "BACKEND_AUTO", "BACKEND_UNFUSED", "BACKEND_FUSED", "BACKEND_FUSED_TB", "BACKEND_STREAM_TB", "BACKEND_BAND_TB", "BACKEND_RESIDENT", "BOUNDARY_REFERENCE_OPEN", "BOUNDARY_CLOSED",
    "FIELD_TERRAIN", "FIELD_WATER", "FIELD_FLUX", "FIELD_VELOCITY", "FIELD_TERRAIN_INFO",
]
