"""Quick device-resident throughput sweep over backends (not the bench contract; for tuning).
usage: perf_sweep.py [W] [H] [backend:k,backend:k,...]"""
import sys; sys.path.insert(0, '.')
import terrainwatersim_b200 as tws
W = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
H = int(sys.argv[2]) if len(sys.argv) > 2 else W
cfgs = [tuple(int(v) for v in c.split(':')) for c in sys.argv[3].split(',')] if len(sys.argv) > 3 else \
    [(1, 1), (2, 1), (3, 2), (3, 3), (3, 4), (4, 1), (4, 2), (4, 3), (4, 4)]
for b, k in cfgs:
    with tws.Terrain(W, height=H, backend=b, temporal_block=k) as sim:
        sim.CreateHeightmapFromNoiseAndResetSim()
        n = max(12, int(24 * (8192 * 8192) / (W * H))) // 12 * 12
        sim.step(n); sim.sync()
        best = 0
        for _ in range(3):
            sim.step(5 * n); sim.sync(); ms = sim.elapsed_ms()
            best = max(best, W * H * 5 * n / ms / 1e6)
        print('perf', W, H, 'backend', b, 'k', k, round(best, 1), 'Gcell/s', flush=True)
