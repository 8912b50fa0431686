"""Resident backend: device time of one tws_step(n) call against n (slope = per-step cost, intercept = block load / store),
next to the tile kernel's captured batch of the same n.  Tuning helper."""
import sys; sys.path.insert(0, '.')
import terrainwatersim_b200 as tws

sizes = [int(x) for x in sys.argv[1].split(',')] if len(sys.argv) > 1 else [256, 512, 1024]
ns = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [1, 2, 4, 10, 20, 40]
for W in sizes:
    for name, b, k in (("resident", 6, 1), ("tile k=2", 3, 2)):
        with tws.Terrain(W, backend=b, temporal_block=k) as sim:
            sim.CreateHeightmapFromNoiseAndResetSim()
            row = []
            for n in ns:
                for _ in range(10): sim.step(n)
                sim.sync()
                best = 1e9
                for _ in range(20):
                    sim.step(n); sim.sync(); best = min(best, sim.elapsed_ms() * 1e3)
                row.append(f"n={n}: {best:6.1f}")
            print(f"frame {W:5d} {name:9s} us on the device  " + "  ".join(row), flush=True)
