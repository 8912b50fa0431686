import sys, time; sys.path.insert(0,'.')
import numpy as np
import terrainwatersim_b200 as tws
from oracle.oracle_py import Oracle, dam_break, new_state
o=Oracle(openmp=True)
def run(W,H,n,backend,k,rim=True):
    h,d=dam_break(W,H,rim)
    c=o.derive_consts(float(W),W)
    t,f,v=new_state(h,d); o.step(t,f,v,c,n)
    with tws.Terrain(W,height=H,backend=backend,temporal_block=k) as sim:
        sim.upload(tws.FIELD_TERRAIN,h); sim.upload(tws.FIELD_WATER,d)
        sim.step(n)
        gd=sim.readback(tws.FIELD_WATER); gf=sim.readback(tws.FIELD_FLUX); gv=sim.readback(tws.FIELD_VELOCITY)
    ok=(np.array_equal(gd.view(np.uint32),t[...,3].view(np.uint32)), np.array_equal(gf.view(np.uint32),f.view(np.uint32)), np.array_equal(gv.view(np.uint16),v.view(np.uint16)))
    print(W,H,n,backend,k,rim,ok, 'maxdiff d', np.abs(gd-t[...,3]).max(), 'nbad', (gd!=t[...,3]).sum(), flush=True)
for W,H in ((256,256),(250,190),(37,5),(1,1),(130,29),(1024,1024)):
    for b,k in ((1,1),(2,1),(3,2),(3,3),(3,4)):
        for rim in (True,False):
            n = 50 if W<1024 else 20
            try: run(W,H,n,b,k,rim)
            except Exception as e: print('ERR',W,H,b,k,e, flush=True)
run(256,256,1000,3,4,True)
# perf quick
for W,b,k in ((8192,1,1),(8192,2,1),(8192,3,2),(8192,3,3),(8192,3,4)):
    with tws.Terrain(W,backend=b,temporal_block=k) as sim:
        sim.CreateHeightmapFromNoiseAndResetSim()
        sim.step(12); sim.sync()
        sim.step(48); sim.sync(); ms=sim.elapsed_ms()
        print('perf',W,b,k, ms/48,'ms/step', W*W*48/ms/1e6,'Gcell/s', flush=True)
