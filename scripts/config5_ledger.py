"""BASELINE config 5: long-horizon rain + evaporation run on the open (reference) boundary with a
mass ledger, strip-decomposed over the ranks (launch with torchrun, one rank per GPU).

    V(t) = V0 + cells * steps * (rain_step - evap_step) - sum_t boundary_outflow(t) * areaInv

rain_rate > evaporation_rate so the evaporation clamp max(0, .) never bites and the net source per
cell-step is known exactly; the outflow through the grid edge is read from the flux field after every
step (k = 1 so every step's flux is in HBM).  Everything is accumulated in fp64; the simulation itself
is fp32, so the closure is limited by fp32 rounding of d + delta (measured, not assumed).
"""
import argparse, json, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import torch.distributed as dist
import terrainwatersim_b200 as tws

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=65536)
ap.add_argument("--steps", type=int, default=1000)
ap.add_argument("--rain", type=float, default=0.6)
ap.add_argument("--evap", type=float, default=0.3)
ap.add_argument("--fast-steps", type=int, default=0, help="extra steps with temporal blocking and no per-step ledger (long horizon, invariants only)")
ap.add_argument("--backend", default="band", choices=["band", "tile"], help="band: band kernel (k=1 ledger, k=4 long run); tile: tile kernel (k=1 / k=2)")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W = a.size
plan = tws.plan_strips(W, world)


def allsum(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t)
    return float(t.item())


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


res = {}
B1, BK, KK = (tws.BACKEND_BAND_TB, tws.BACKEND_BAND_TB, 4) if a.backend == "band" else (tws.BACKEND_FUSED, tws.BACKEND_FUSED_TB, 2)
with tws.Terrain(W, rows=plan.rows(rank), backend=B1, temporal_block=1, device=local, rain_rate=a.rain, evaporation_rate=a.evap) as sim:
    tws.connect_strips(sim, plan, rank)
    sim.CreateHeightmapFromNoiseAndResetSim()
    barrier()
    c = sim.step_constants()
    dt = float(np.float32(1.0) / np.float32(60.0))
    rs, es = np.float64(np.float32(dt * np.float32(a.rain))), np.float64(np.float32(dt * np.float32(a.evap)))
    V0 = allsum(sim.total_volume())
    out = 0.0
    t0 = time.perf_counter()
    for _ in range(a.steps):
        sim.step(1)
        out += sim.boundary_outflow() * float(c[2])
    barrier()
    wall = time.perf_counter() - t0
    out = allsum(out)
    V1 = allsum(sim.total_volume())
    expect = V0 + float(W) * W * a.steps * float(rs - es) - out
    res = {"grid": [W, W], "gpus": world, "backend": a.backend, "long_run_temporal_block": KK, "steps": a.steps, "V0": V0, "V1": V1, "boundary_outflow": out,
           "net_source": float(W) * W * a.steps * float(rs - es), "closure_rel": (V1 - expect) / V1,
           "ledger_wall_s": wall, "ledger_Gcell_per_s": float(W) * W * a.steps / wall / 1e9}
if a.fast_steps:
    with tws.Terrain(W, rows=plan.rows(rank), backend=BK, temporal_block=KK, device=local, rain_rate=a.rain,
                     evaporation_rate=a.evap) as sim:
        tws.connect_strips(sim, plan, rank)
        sim.CreateHeightmapFromNoiseAndResetSim()
        barrier()
        sim.step(a.fast_steps)
        sim.sync()
        barrier()
        ms = sim.elapsed_ms()
        V = allsum(sim.total_volume())
        dmin = float(np.min(sim.readback(tws.FIELD_WATER)[:64]))
        res.update({"long_steps": a.fast_steps, "long_ms_per_step": ms / a.fast_steps, "long_Gcell_per_s": float(W) * W * a.fast_steps / ms / 1e6,
                    "long_final_volume": V, "long_finite": bool(np.isfinite(V)), "long_min_depth_sample": dmin})
if rank == 0:
    print("CONFIG5 " + json.dumps(res), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
