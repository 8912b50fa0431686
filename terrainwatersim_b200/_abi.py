"""ctypes binding of include/tws.h — the same stub a host application would write.

Nothing here computes: every call goes into libtws.so (CUDA, sm_100a).  If the library
is missing, loading raises — there is no Python or CPU implementation behind it.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "libtws.so"

TWS_OK = 0
TWS_ERR_INVALID = -1
TWS_ERR_CUDA = -2
TWS_ERR_NOMEM = -3
TWS_ERR_STATE = -4
TWS_ERR_UNSUPPORTED = -5

BACKEND_AUTO = 0
BACKEND_UNFUSED = 1
BACKEND_FUSED = 2
BACKEND_FUSED_TB = 3
BACKEND_STREAM_TB = 4
BACKEND_BAND_TB = 5
BACKEND_RESIDENT = 6

BOUNDARY_REFERENCE_OPEN = 0
BOUNDARY_CLOSED = 1

FIELD_TERRAIN = 0
FIELD_WATER = 1
FIELD_FLUX = 2
FIELD_VELOCITY = 3
FIELD_TERRAIN_INFO = 4


class TwsParams(C.Structure):
    _fields_ = [
        ("size", C.c_uint32),
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("row_begin", C.c_int32),
        ("row_end", C.c_int32),
        ("world_size", C.c_float),
        ("steps_per_second", C.c_float),
        ("flow_damping", C.c_float),
        ("flow_acceleration", C.c_float),
        ("boundary", C.c_int32),
        ("backend", C.c_int32),
        ("temporal_block", C.c_int32),
        ("device", C.c_int32),
        ("rain_rate", C.c_float),
        ("evaporation_rate", C.c_float),
    ]


class TwsStepConstants(C.Structure):
    _fields_ = [
        ("flow_friction_per_step", C.c_float),
        ("water_acceleration_per_step", C.c_float),
        ("cell_area_inv_time_scaled", C.c_float),
    ]


class TwsHaloHandle(C.Structure):
    _fields_ = [
        ("mem", C.c_uint8 * 64),
        ("slab_bytes", C.c_uint64),
        ("row_begin", C.c_int32),
        ("row_end", C.c_int32),
        ("device", C.c_int32),
        ("pid", C.c_int32),
        ("local_ptr", C.c_uint64),
    ]


# name -> (restype, argtypes); every symbol include/tws.h declares.
_SIM = C.c_void_p
SYMBOLS = {
    "tws_version": (C.c_char_p, []),
    "tws_abi_version": (C.c_int32, []),
    "tws_default_params": (None, [C.POINTER(TwsParams)]),
    "tws_create": (C.c_int, [C.POINTER(TwsParams), C.POINTER(_SIM)]),
    "tws_destroy": (C.c_int, [_SIM]),
    "tws_last_error": (C.c_char_p, [_SIM]),
    "tws_set_steps_per_second": (C.c_int, [_SIM, C.c_float]),
    "tws_set_flow_damping": (C.c_int, [_SIM, C.c_float]),
    "tws_set_flow_acceleration": (C.c_int, [_SIM, C.c_float]),
    "tws_set_sources": (C.c_int, [_SIM, C.c_float, C.c_float]),
    "tws_get_step_constants": (C.c_int, [_SIM, C.POINTER(TwsStepConstants)]),
    "tws_upload": (C.c_int, [_SIM, C.c_int, C.c_void_p, C.c_size_t]),
    "tws_readback": (C.c_int, [_SIM, C.c_int, C.c_void_p, C.c_size_t]),
    "tws_reset_reference_scene": (C.c_int, [_SIM, C.c_uint32, C.c_float, C.c_int32, C.c_int32, C.c_float]),
    "tws_reset_reference_scene_tiled": (C.c_int, [_SIM, C.c_uint32, C.c_float, C.c_int32, C.c_int32, C.c_float, C.c_int32]),
    "tws_inject_brush": (C.c_int, [_SIM, C.c_float, C.c_float, C.c_float, C.c_float]),
    "tws_inject_brush_world": (C.c_int, [_SIM, C.c_float, C.c_float, C.c_float]),
    "tws_step": (C.c_int, [_SIM, C.c_int32]),
    "tws_step_host": (C.c_int, [_SIM, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tws_advance": (C.c_int, [_SIM, C.c_double, C.POINTER(C.c_uint32)]),
    "tws_total_volume": (C.c_int, [_SIM, C.POINTER(C.c_double)]),
    "tws_sync": (C.c_int, [_SIM]),
    "tws_boundary_outflow": (C.c_int, [_SIM, C.POINTER(C.c_double)]),
    "tws_boundary_outflow_accumulated": (C.c_int, [_SIM, C.POINTER(C.c_double)]),
    "tws_source_accumulated": (C.c_int, [_SIM, C.POINTER(C.c_double)]),
    "tws_boundary_outflow_reset": (C.c_int, [_SIM]),
    "tws_elapsed_ms": (C.c_int, [_SIM, C.POINTER(C.c_float)]),
    "tws_elapsed_ms_nowait": (C.c_int, [_SIM, C.POINTER(C.c_float), C.POINTER(C.c_uint64)]),
    "tws_kernel_launches": (C.c_uint64, [_SIM]),
    "tws_graph_replays": (C.c_uint64, [_SIM]),
    "tws_backend_in_use": (C.c_int, [_SIM, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "tws_device_view": (C.c_int, [_SIM, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "tws_halo_export": (C.c_int, [_SIM, C.POINTER(TwsHaloHandle)]),
    "tws_halo_connect": (C.c_int, [_SIM, C.POINTER(TwsHaloHandle), C.POINTER(TwsHaloHandle)]),
    "tws_halo_refresh": (C.c_int, [_SIM]),
    "tws_gl_register": (C.c_int, [_SIM, C.c_uint32, C.c_uint32]),
    "tws_gl_publish": (C.c_int, [_SIM]),
    "tws_gl_unregister": (C.c_int, [_SIM]),
    "tws_publish_packed": (C.c_int, [_SIM, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "tws_publish_mips": (C.c_int, [_SIM, C.POINTER(C.c_void_p), C.POINTER(C.c_int32)]),
    "tws_readback_mip": (C.c_int, [_SIM, C.c_int32, C.c_void_p, C.c_size_t]),
    "tws_mip_levels": (C.c_int32, [C.c_int32, C.c_int32]),
    "tws_mip_level_info": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
}

_lib = None


def load() -> C.CDLL:
    """Load libtws.so and attach prototypes.  Raises if the CUDA library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    import os
    path = Path(os.environ.get("TWS_LIB", str(LIB_PATH)))      # TWS_LIB: tuning builds of the same CUDA library
    if not path.exists():
        raise RuntimeError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(terrainwatersim_b200 has no CPU or pure-Python path)")
    lib = C.CDLL(str(path))
    for name, (res, args) in SYMBOLS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            if "TWS_LIB" in os.environ:       # an older tuning build may predate an entry point; the product library may not
                continue
            raise
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
