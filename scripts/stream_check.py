"""Quick GPU check of the row-streaming backend: parity against the oracle on small grids, then
device-resident throughput at 8192^2 (tuning helper, not the bench contract)."""
import sys; sys.path.insert(0, '.')
import numpy as np
import terrainwatersim_b200 as tws
from oracle.oracle_py import Oracle, dam_break, new_state

o = Oracle(openmp=True)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
KS = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [1, 2, 3, 4]
do_parity = '--no-parity' not in sys.argv


def run(W, H, n, backend, k, rim=True):
    h, d = dam_break(W, H, rim)
    c = o.derive_consts(float(W), W)
    t, f, v = new_state(h, d); o.step(t, f, v, c, n)
    with tws.Terrain(W, height=H, backend=backend, temporal_block=k) as sim:
        sim.upload(tws.FIELD_TERRAIN, h); sim.upload(tws.FIELD_WATER, d)
        sim.step(n)
        gd = sim.readback(tws.FIELD_WATER); gf = sim.readback(tws.FIELD_FLUX); gv = sim.readback(tws.FIELD_VELOCITY)
    ok = (np.array_equal(gd.view(np.uint32), t[..., 3].view(np.uint32)), np.array_equal(gf.view(np.uint32), f.view(np.uint32)),
          np.array_equal(gv.view(np.uint16), v.view(np.uint16)))
    bad = np.argwhere(gd.view(np.uint32) != t[..., 3].view(np.uint32))
    print(W, H, n, backend, k, rim, ok, 'nbad', len(bad), bad[:4].tolist(), flush=True)
    return all(ok)


allok = True
if do_parity:
    for W, H in ((256, 256), (250, 190), (37, 5), (1, 1), (130, 29), (640, 333), (1024, 1024)):
        for k in KS:
            for rim in (True, False):
                n = 50 if W < 1024 else 20
                try:
                    allok &= run(W, H, n, B, k, rim)
                except Exception as e:
                    print('ERR', W, H, B, k, e, flush=True); allok = False
    print('PARITY', 'OK' if allok else 'FAILED', flush=True)
for W in (8192,):
    for k in KS:
        with tws.Terrain(W, backend=B, temporal_block=k) as sim:
            sim.CreateHeightmapFromNoiseAndResetSim()
            sim.step(24); sim.sync()
            best = 0
            for _ in range(3):
                sim.step(120); sim.sync(); ms = sim.elapsed_ms()
                best = max(best, W * W * 120 / ms / 1e6)
            print('perf', W, 'backend', B, 'k', k, round(best, 1), 'Gcell/s', flush=True)
