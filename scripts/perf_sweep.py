"""Quick device-resident throughput sweep over backends (not the bench contract; for tuning)."""
import sys; sys.path.insert(0, '.')
import terrainwatersim_b200 as tws
W = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
for b, k in ((1, 1), (2, 1), (3, 2), (3, 3), (3, 4)):
    with tws.Terrain(W, backend=b, temporal_block=k) as sim:
        sim.CreateHeightmapFromNoiseAndResetSim()
        sim.step(24); sim.sync()
        best = 0
        for _ in range(3):
            sim.step(120); sim.sync(); ms = sim.elapsed_ms()
            best = max(best, W * W * 120 / ms / 1e6)
        print('perf', W, 'backend', b, 'k', k, round(best, 1), 'Gcell/s', flush=True)
