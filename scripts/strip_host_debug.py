"""debug: tws_step_host on same-GPU strips, one host thread per strip"""
import os, sys, threading, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, '.')
import numpy as np, torch
import terrainwatersim_b200 as tws
from oracle.oracle_py import Oracle, new_state
o = Oracle(openmp=True)

def run(nstrips, backend, k, rows, pinned, W=2100, steps=3):
    H = rows * nstrips
    rng = np.random.default_rng(31)
    h = (rng.random((H, W)) * 8).astype(np.float32); d = (rng.random((H, W)) * 4 * (rng.random((H, W)) > 0.4)).astype(np.float32)
    c = o.derive_consts(float(W), W); t, f, v = new_state(h, d)
    bounds = [(i * rows, (i + 1) * rows) for i in range(nstrips)]
    sims = [tws.Terrain(W, height=H, rows=bounds[i], backend=backend, temporal_block=k) for i in range(nstrips)]
    hs = [s.halo_export() for s in sims]
    for i, s in enumerate(sims): s.halo_connect(hs[i - 1] if i > 0 else None, hs[i + 1] if i + 1 < nstrips else None)
    for i, s in enumerate(sims):
        r0, r1 = bounds[i]; s.upload(tws.FIELD_TERRAIN, h[r0:r1]); s.upload(tws.FIELD_WATER, d[r0:r1])
    for s in sims: s.halo_refresh()
    for s in sims: s.sync()
    if pinned:
        wt = [torch.from_numpy(np.ascontiguousarray(d[r0:r1])).pin_memory() for r0, r1 in bounds]
        vt = [torch.zeros((r1 - r0, W, 2), dtype=torch.float16).pin_memory() for r0, r1 in bounds]
        water = [x.numpy() for x in wt]; vel = [x.numpy() for x in vt]
    else:
        water = [np.ascontiguousarray(d[r0:r1]) for r0, r1 in bounds]; vel = [np.zeros((r1 - r0, W, 2), np.float16) for r0, r1 in bounds]
    ok = True
    for step in range(steps):
        errs = []
        def go(i):
            t0 = time.perf_counter()
            try: sims[i].step_host(water[i], water[i], vel[i])
            except Exception as e: errs.append((i, str(e), time.perf_counter() - t0))
        th = [threading.Thread(target=go, args=(i,)) for i in range(nstrips)]
        [x.start() for x in th]; [x.join() for x in th]
        o.step(t, f, v, c, 1)
        same = np.array_equal(np.concatenate(water).view(np.uint32), np.ascontiguousarray(t[..., 3]).view(np.uint32))
        print(f"  step {step}: errs={errs} same={same}", flush=True)
        if errs or not same: ok = False; break
    for s in sims: s.close()
    return ok

for nstrips, backend, k, rows in ((3, 2, 1, 200), (3, 5, 1, 200), (3, 3, 2, 40), (2, 2, 1, 200), (3, 4, 1, 64)):
    for pinned in (True, False):
        print("case", nstrips, backend, k, rows, "pinned" if pinned else "pageable", flush=True)
        print("  ->", run(nstrips, backend, k, rows, pinned), flush=True)
