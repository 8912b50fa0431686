import sys; sys.path.insert(0,'.')
import ctypes as C
import terrainwatersim_b200 as tws
with tws.Terrain(1024, backend=2, temporal_block=1) as sim:
    sim.CreateHeightmapFromNoiseAndResetSim()
    lib, h = sim._lib, sim._sim
    n = C.c_uint32(0); base = C.c_void_p(); lv = C.c_int32(0)
    for _ in range(6):
        lib.tws_inject_brush_world(h, C.c_float(512.0), C.c_float(512.0), C.c_float(100.0 / 60.0))
        lib.tws_advance(h, C.c_double(1.0 / 60.0 + 1e-9), C.byref(n))
        lib.tws_publish_mips(h, C.byref(base), C.byref(lv))
    sim.sync()
