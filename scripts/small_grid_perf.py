"""Launch-bound regime: frames of 10 steps (the reference's per-frame maximum, Terrain.cpp:247) on
the small BASELINE grids, with and without batch graphs (TWS_GRAPHS=0).  Wall-clock per frame
includes the host launch path, which is the point."""
import sys, time; sys.path.insert(0, '.')
import terrainwatersim_b200 as tws
for W in (256, 1024, 2048):
    for name, b, k in (("unfused", 1, 1), ("tile k=2", 3, 2), ("stream k=2", 4, 2)):
        with tws.Terrain(W, backend=b, temporal_block=k) as sim:
            sim.CreateHeightmapFromNoiseAndResetSim()
            for _ in range(20): sim.step(10)
            sim.sync()
            t0 = time.perf_counter(); frames = 400
            for _ in range(frames): sim.step(10)
            sim.sync(); dt = time.perf_counter() - t0
            print(f"small {W:5d} {name:10s} graphs={sim.graph_replays() > 0} {dt / frames * 1e6:8.1f} us/frame(10 steps) {W * W * 10 * frames / dt / 1e9:8.2f} Gcell/s launches={sim.kernel_launches()}", flush=True)
