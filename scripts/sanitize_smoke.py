"""One short run of every backend (256x256 and a ragged 130x77, open boundary with rain + evaporation so the ledgers run,
plus two strips on one GPU with the band kernel's fused exchange) — the workload for compute-sanitizer:
    compute-sanitizer --tool memcheck  python scripts/sanitize_smoke.py
    compute-sanitizer --tool racecheck python scripts/sanitize_smoke.py
    compute-sanitizer --tool synccheck python scripts/sanitize_smoke.py
"""
import os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, '.')
import numpy as np
import terrainwatersim_b200 as tws
from oracle.oracle_py import dam_break

only = {int(x) for x in sys.argv[1].split(',')} if len(sys.argv) > 1 else None      # backend filter, e.g. "6"
for W, H in ((256, 256), (130, 77)):
    h, d = dam_break(W, H, rim=False)
    for backend, k in ((1, 1), (2, 1), (3, 2), (3, 4), (4, 3), (5, 1), (5, 4), (6, 1)):
        if only and backend not in only:
            continue
        with tws.Terrain(W, height=H, backend=backend, temporal_block=k, rain_rate=0.5, evaporation_rate=0.2) as sim:
            sim.upload(tws.FIELD_TERRAIN, h); sim.upload(tws.FIELD_WATER, d)
            sim.inject_brush(W / 2, H / 2, 1.0, 32.0)
            sim.step(2 * k + 1)
            sim.publish_mips()
            v = sim.total_volume(); o = sim.boundary_outflow_accumulated(); s = sim.source_accumulated()
        print("ok", W, H, backend, k, round(v, 3), round(o, 6), round(s, 6), flush=True)
if only:
    sys.exit(0)
W, H = 300, 64
h, d = dam_break(W, H, rim=False)
sims = [tws.Terrain(W, height=H, rows=(i * 32, (i + 1) * 32), backend=5, temporal_block=4) for i in range(2)]
hd = [s.halo_export() for s in sims]
sims[0].halo_connect(None, hd[1]); sims[1].halo_connect(hd[0], None)
for i, s in enumerate(sims):
    s.upload(tws.FIELD_TERRAIN, h[i * 32:(i + 1) * 32]); s.upload(tws.FIELD_WATER, d[i * 32:(i + 1) * 32])
for s in sims: s.halo_refresh()
for s in sims: s.sync()
for _ in range(2):
    for s in sims: s.step(4)
for s in sims: s.sync()
print("ok strips", sum(s.total_volume() for s in sims), flush=True)
for s in sims: s.close()
