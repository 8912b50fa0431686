"""Launch-bound regime: (1) frames of 10 steps (the reference's per-frame maximum, Terrain.cpp:247) on the small
BASELINE grids, with batch graphs; (2) the reference's own frame at its own size — 1024^2, brush, one step at
60 steps/s, mip chain of TerrainInfo (Terrain.cpp:240-277, what its on-screen "Simulation Time" measures; BASELINE.md
quotes 1.9-5.9 ms per frame from the screenshots, unknown 2013 GPU).  Wall clock per frame includes the host launch
path, which is the point."""
import sys, time; sys.path.insert(0, '.')
import ctypes as C
import terrainwatersim_b200 as tws

only = sys.argv[1].split(',') if len(sys.argv) > 1 else None
for W in (256, 512, 1024, 2048):
    for name, b, k in (("unfused", 1, 1), ("tile k=1", 2, 1), ("tile k=2", 3, 2), ("tile k=3", 3, 3), ("band k=2", 5, 2), ("band k=4", 5, 4),
                       ("resident", 6, 1)):
        if (b == 6 and W > 1024) or (only and name.split()[0] not in only):
            continue
        with tws.Terrain(W, backend=b, temporal_block=k) as sim:
            sim.CreateHeightmapFromNoiseAndResetSim()
            for _ in range(20): sim.step(10)
            sim.sync()
            t0 = time.perf_counter(); frames = 400
            for _ in range(frames): sim.step(10)
            sim.sync(); dt = time.perf_counter() - t0
            gpu_us = sim.elapsed_ms() * 1e3           # device time of the LAST 10-step batch alone (CUDA events)
            print(f"small {W:5d} {name:10s} graphs={sim.graph_replays() > 0} {dt / frames * 1e6:8.1f} us/frame(10 steps) wall, {gpu_us:7.1f} us on the device, {W * W * 10 * frames / dt / 1e9:8.2f} Gcell/s launches={sim.kernel_launches()}", flush=True)

# the reference's frame: ApplyRadialWaterBrush + PerformSimulationStep(1/60 s) (one step) + GenMipMaps
for name, b, k in (("unfused", 1, 1), ("fused", 2, 1), ("band k=1", 5, 1), ("resident", 6, 1)):
    if only and name.split()[0] not in only:
        continue
    with tws.Terrain(1024, backend=b, temporal_block=k) as sim:
        sim.CreateHeightmapFromNoiseAndResetSim()
        lib, h = sim._lib, sim._sim
        n = C.c_uint32(0); base = C.c_void_p(); lv = C.c_int32(0)
        def frame():
            lib.tws_inject_brush_world(h, C.c_float(512.0), C.c_float(512.0), C.c_float(100.0 / 60.0))
            lib.tws_advance(h, C.c_double(1.0 / 60.0 + 1e-9), C.byref(n))
            lib.tws_publish_mips(h, C.byref(base), C.byref(lv))
        for _ in range(50): frame()
        sim.sync()
        frames = 1000; steps = 0
        t0 = time.perf_counter()
        for _ in range(frames):
            frame(); steps += n.value
        sim.sync(); dt = time.perf_counter() - t0
        print(f"refframe 1024 {name:10s} {dt / frames * 1e6:8.1f} us/frame (brush + {steps / frames:.2f} steps + {lv.value}-level mip chain), launches/frame {sim.kernel_launches() / (frames + 50):.1f}", flush=True)
