"""Where the band kernel overtakes the tile kernel: device-resident throughput (120 steps, CUDA events) per grid size."""
import sys; sys.path.insert(0, '.')
import terrainwatersim_b200 as tws
for W in (1024, 2048, 3072, 4096, 6144):
    for name, b, k in (("tile k=2", 3, 2), ("tile k=3", 3, 3), ("band k=2", 5, 2), ("band k=3", 5, 3), ("band k=4", 5, 4)):
        with tws.Terrain(W, backend=b, temporal_block=k) as sim:
            sim._lib.tws_sync(sim._sim)
            sim.CreateHeightmapFromNoiseAndResetSim()
            sim.step(120); sim.sync()
            best = 0.0
            for _ in range(3):
                sim.step(120); sim.sync()
                best = max(best, W * W * 120 / sim.elapsed_ms() / 1e6)
            print(f"cross {W:5d} {name:9s} {best:7.1f} Gcell/s", flush=True)
