#!/usr/bin/env python
"""bench.py — Gcell-updates/s of the shallow-water step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference ...                           # the reference's algorithm on the host cores
    torchrun ... bench.py --gpus N ...                             # one rank per GPU (row strips)

A "step" is one simulation step (flowUpdate + flowApply of the reference, Terrain.cpp:253-265)
over the whole grid.  Workload: BASELINE config 3 — the reference scene generator
(value-noise fBm terrain + central lake, seed 231656522) at 8192 x 8192 per GPU; at N GPUs the
grid is 8192 x (8192*N), cut into N row strips (weak scaling), halo rows exchanged over NVLink
by the library itself.  `--size S --strong` runs an S x S grid split over the ranks instead
(BASELINE configs 4/5).  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

ALGO_BYTES_PER_CELL_UPDATE = 48          # R h4 + R d4 + R F16 + W d4 + W F16 + W v4
HBM_FALLBACK_GBS = 6650.0                # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1200)
    ap.add_argument("--warmup", type=int, default=60)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=8192, help="grid width (and per-GPU rows unless --strong)")
    ap.add_argument("--strong", action="store_true", help="fixed size x size grid split over the ranks (strong scaling)")
    ap.add_argument("--backend", default="band", choices=["unfused", "fused", "tb", "stream", "band"],
                    help="step kernel: band (default, lock-step row streaming), stream (ring row streaming), tb (overlapped tiles), fused (k=1 tiles), unfused")
    ap.add_argument("--tb", type=int, default=4, help="temporal block (steps per launch, 1..4) for --backend band / stream / tb")
    ap.add_argument("--e2e-steps", type=int, default=-1, help="steps of the end-to-end leg (default min(steps, 40))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--verify-strips", action="store_true", help="also compare the strips of the BENCH grid with a whole-grid run on rank 0 (small --size only)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling sub-record (BASELINE config 4: 32768^2 split over the ranks)")
    ap.add_argument("--no-frame", action="store_true", help="skip the 1024^2 frame-time sub-record (the reference's own operating point)")
    ap.add_argument("--strong-size", type=int, default=32768)
    ap.add_argument("--strong-steps", type=int, default=48)
    return ap.parse_args()


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch(kernel_tag: str):
    """dram bytes per launch of the dominant kernel from the committed ncu summary, or None."""
    p = ROOT / "profiles" / "traffic.json"
    if not p.exists():
        return None
    try:
        return json.loads(p.read_text()).get(kernel_tag)
    except Exception:
        return None


class _NvmlCtypes:
    """The five NVML calls the sampler needs, straight from libnvidia-ml.so.1 (used when the pynvml module is missing)."""
    NVML_CLOCK_SM = 1

    def __init__(self):
        import ctypes as C
        self.C = C
        self.lib = C.CDLL("libnvidia-ml.so.1")
        if self.lib.nvmlInit_v2() != 0:
            raise RuntimeError("nvmlInit_v2 failed")

    def handle(self, index):
        h = self.C.c_void_p()
        if self.lib.nvmlDeviceGetHandleByIndex_v2(self.C.c_uint(index), self.C.byref(h)) != 0:
            raise RuntimeError("nvmlDeviceGetHandleByIndex_v2 failed")
        return h

    def _uint(self, fn, h, *args):
        v = self.C.c_uint(0)
        if fn(h, *args, self.C.byref(v)) != 0:
            raise RuntimeError("nvml call failed")
        return int(v.value)

    def max_sm(self, h):
        return self._uint(self.lib.nvmlDeviceGetMaxClockInfo, h, self.C.c_int(self.NVML_CLOCK_SM))

    def sm(self, h):
        return self._uint(self.lib.nvmlDeviceGetClockInfo, h, self.C.c_int(self.NVML_CLOCK_SM))

    def reasons(self, h):
        v = self.C.c_ulonglong(0)
        fn = getattr(self.lib, "nvmlDeviceGetCurrentClocksEventReasons", None) or self.lib.nvmlDeviceGetCurrentClocksThrottleReasons
        if fn(h, self.C.byref(v)) != 0:
            raise RuntimeError("nvml reasons failed")
        return int(v.value)


class _NvmlModule:
    def __init__(self):
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()

    def handle(self, index):
        return self.nv.nvmlDeviceGetHandleByIndex(index)

    def max_sm(self, h):
        return int(self.nv.nvmlDeviceGetMaxClockInfo(h, self.nv.NVML_CLOCK_SM))

    def sm(self, h):
        return int(self.nv.nvmlDeviceGetClockInfo(h, self.nv.NVML_CLOCK_SM))

    def reasons(self, h):
        try:
            return int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(h))
        except Exception:
            return int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))


def _physical_gpu_index(local: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            pass
    return local


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML (pynvml, else ctypes on libnvidia-ml.so.1) while the GPU is under
    the bench load.  `mark(t0, t1)` windows are the regions whose samples count."""

    def __init__(self, device_index: int, period_s: float = 0.004):
        super().__init__(daemon=True)
        self.period = period_s
        self.samples = []           # (t, sm_mhz, reasons_bitmask)
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        self.via = None
        self.windows = []
        for name, cls in (("pynvml", _NvmlModule), ("ctypes:libnvidia-ml.so.1", _NvmlCtypes)):
            try:
                self.nv = cls()
                self.h = self.nv.handle(_physical_gpu_index(device_index))
                self.max_mhz = self.nv.max_sm(self.h)
                self.nv.sm(self.h)
                self.ok, self.via = True, name
                break
            except Exception:
                continue

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append((time.perf_counter(), self.nv.sm(self.h), self.nv.reasons(self.h)))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()

    def mark(self, t0: float, t1: float):
        self.windows.append((t0, t1))

    def inside(self):
        return [s for s in self.samples if any(a <= s[0] <= b for a, b in self.windows)]

    def summary(self, note=None):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"], "note": "NVML not reachable (neither pynvml nor libnvidia-ml.so.1)"}
        inside = self.inside()
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"], "note": "no NVML sample fell into the loaded region", "via": self.via}
        names = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
                 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
        mask = 0
        for smp in inside:
            mask |= smp[2]
        reasons = sorted(n for b, n in names.items() if mask & b and n != "gpu_idle")
        out = {"sm_mhz": statistics.median(smp[1] for smp in inside), "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(inside), "via": self.via}
        if note:
            out["note"] = note
        return out


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores (the ONLY place bench.py uses oracle/)
# --------------------------------------------------------------------------------------------------
def host_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_oracle_throughput(size: int, steps: int, warmup: int, budget_s: float, openmp: bool = True):
    """Times the OpenMP oracle (bit-identical to the single-threaded one) on a bounded sample of
    the workload: a `rows`-row slab of the size-wide reference scene, `rows` chosen so the whole
    run fits the budget.  Returns (Gcell/s, cores, sample description, ms per step)."""
    import numpy as np
    from oracle.oracle_py import Oracle
    o = Oracle(openmp=openmp)
    if openmp:
        # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: set the thread count explicitly so the host
        # baseline uses every core this process may run on, whatever the launcher exported
        o.set_threads(host_cores())
    cores = o.threads if openmp else 1
    consts = o.derive_consts(float(size), size)

    scene = (o if openmp else Oracle(openmp=True)).create_scene(size, size)      # scene generation is not what is timed

    def slab(rows):                       # central rows: the lake sits in the middle of the map
        a = (size - rows) // 2
        return scene[a:a + rows].copy(), np.zeros((rows, size, 4), np.float32), np.zeros((rows, size, 2), np.float16)

    # calibrate on a small slab
    rows = min(size, 256)
    t, f, v = slab(rows)
    o.step(t, f, v, consts, 1)
    t0 = time.perf_counter()
    o.step(t, f, v, consts, 2)
    per_row_step = (time.perf_counter() - t0) / 2 / rows
    total_steps = max(1, steps + warmup)
    rows_fit = int(budget_s / (per_row_step * total_steps))
    rows = max(64, min(size, rows_fit // 64 * 64))
    if rows != t.shape[0]:
        t, f, v = slab(rows)
    o.step(t, f, v, consts, warmup)
    t0 = time.perf_counter()
    o.step(t, f, v, consts, steps)
    dt = time.perf_counter() - t0
    gcells = rows * size * steps / dt / 1e9
    sample = (f"{size}x{rows} central-row slab of the {size}x{size} reference scene, {steps} steps after {warmup} warm-up, "
              + (f"OpenMP oracle on {cores} threads" if openmp else "single-threaded oracle"))
    return gcells, cores, sample, dt / steps * 1e3


def reference_frame_times(tws, device: int):
    """Device time (CUDA events inside the library, best of 20 after 10 warm-up calls) of one tws_step(n) call on the reference's
    1024 x 1024 default scene: n = 10 (its per-frame maximum, Terrain.cpp:247) and n = 1, through TWS_BACKEND_AUTO (calls of
    >= 3 steps run as ONE resident launch), next to the tile kernel's captured batch of the same n; plus the reference's whole
    frame (brush + one step + full mip chain of TerrainInfo: what its on-screen 'Simulation Time' covers), wall clock."""
    import ctypes as C
    out = {"grid": [1024, 1024], "unit": "us", "timing": "device (CUDA events), best of 20 calls"}

    def best(sim, n):
        for _ in range(10):
            sim.step(n)
        sim.sync()
        b = 1e30
        for _ in range(20):
            sim.step(n); sim.sync()
            b = min(b, sim.elapsed_ms() * 1e3)
        return b

    with tws.Terrain(1024, device=device) as sim:                       # AUTO
        sim.CreateHeightmapFromNoiseAndResetSim()
        l0 = sim.kernel_launches(); sim.step(10); sim.sync()
        out["frame10_launches"] = int(sim.kernel_launches() - l0)
        out["frame10_auto"] = best(sim, 10)
        out["frame1_auto"] = best(sim, 1)
        lib, h = sim._lib, sim._sim
        n = C.c_uint32(0); base = C.c_void_p(); lv = C.c_int32(0)

        def frame():
            lib.tws_inject_brush_world(h, C.c_float(512.0), C.c_float(512.0), C.c_float(100.0 / 60.0))
            lib.tws_advance(h, C.c_double(1.0 / 60.0 + 1e-9), C.byref(n))
            lib.tws_publish_mips(h, C.byref(base), C.byref(lv))
        for _ in range(50):
            frame()
        sim.sync()
        t0 = time.perf_counter()
        for _ in range(500):
            frame()
        sim.sync()
        out["reference_frame_wall"] = (time.perf_counter() - t0) / 500 * 1e6
        out["reference_frame"] = "brush + 1 step + %d-level mip chain through the C ABI, wall clock per frame over 500 frames" % lv.value
    with tws.Terrain(1024, backend=tws.BACKEND_FUSED_TB, temporal_block=2, device=device) as sim:
        sim.CreateHeightmapFromNoiseAndResetSim()
        out["frame10_tile_k2"] = best(sim, 10)
    return out


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    os.environ["OMP_NUM_THREADS"] = str(host_cores())      # before libgomp initialises (torchrun exports 1)
    val, cores, sample, ms = cpu_oracle_throughput(args.size, args.steps, args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": "Gcell-updates/s per sim step (fp32)", "value": val, "unit": "Gcell-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, max(1, args.gpus)),
        "cpu_baseline": {"value": val, "unit": "Gcell-updates/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Gcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference has no CPU path and its GLSL cannot run here; this is the line-for-line C++ oracle port (oracle/tws_oracle.cpp), ms_per_step is for the sampled slab",
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    W = args.size
    Hg = args.size if args.strong else args.size * world
    tiled = "" if (args.strong or world == 1) else f" — the {W}x{W} scene repeated {world}x along y, one copy per GPU"
    return {
        "workload": f"{W}x{Hg} reference-scene fBm terrain (value noise octaves 2..10, persistence 0.43, seed 231656522) + central lake, "
                    f"open (reference) boundary, no sources; BASELINE config {'4/5 (strong)' if args.strong else '3 per GPU'}{tiled}",
        "grid": [W, Hg], "backend": args.backend, "temporal_block": args.tb if args.backend in ("tb", "stream", "band") else 1,
        "decomposition": f"{world} row strip(s); edge rows are stored into the neighbours' halo rows by the step kernel itself (peer stores over NVLink, flag handshake, no collective)",
        "kernel_blocking": "column strips of 128 cells marched top to bottom, rows skewed in time (no halo rows recomputed), k steps per HBM round trip" if args.backend in ("band", "stream") else "overlapped tiles",
        "l2": "state >= 3.2 GB per GPU, far larger than the 126 MB L2 (inputs larger than L2, no flush needed)",
    }


# --------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry

    rank, world, local = dist_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if world == 1 and args.gpus > 1:
        raise SystemExit("for --gpus N > 1 launch with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: terrainwatersim_b200 has no CPU path (use --impl reference for the host baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        entry.build()
    if world > 1:
        dist.barrier()
    import terrainwatersim_b200 as tws

    W = args.size
    Hg = args.size if args.strong else args.size * world
    plan = tws.plan_strips(Hg, world)
    backend = {"unfused": tws.BACKEND_UNFUSED, "fused": tws.BACKEND_FUSED, "tb": tws.BACKEND_FUSED_TB,
               "stream": tws.BACKEND_STREAM_TB, "band": tws.BACKEND_BAND_TB}[args.backend]
    k = args.tb if args.backend in ("tb", "stream", "band") else 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def timed_steps(sim, steps, warmup):
        """warm-up, then exactly `steps` steps between barrier + synchronize; device time (CUDA events on the library's
        launching stream), max over ranks.  Returns (ms, local launches, wall t0, wall t1)."""
        barrier()
        sim.step(warmup)
        sim.sync()
        barrier()
        l0 = sim.kernel_launches()
        t0 = time.perf_counter()
        sim.step(steps)
        sim.sync()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        ms_dev = sim.elapsed_ms()
        launches_local = sim.kernel_launches() - l0
        barrier()
        return max_over_ranks(ms_dev), ms_dev, launches_local, t0, t1

    sim = tws.Terrain(W, height=Hg, rows=plan.rows(rank), backend=backend, temporal_block=k, device=local)
    tws.connect_strips(sim, plan, rank)
    # weak scaling: every GPU's strip is the SAME W x W scene (the generator's terrain is periodic, so the tiles join
    # seamlessly) — per-GPU work is then fixed as N grows, which a scene stretched over W x (W*N) is not
    tile = 0 if args.strong else W
    sim.CreateHeightmapFromNoiseAndResetSim(tileHeight=tile)
    sim.sync()
    cells_global = W * Hg
    cells_local = W * sim.rows

    sampler = ClockSampler(local)
    sampler.start()
    # ---- device-resident throughput: warm-up, then exactly K steps ---------------------------------
    ms, ms_dev, launches_local, t_wall0, t_wall1 = timed_steps(sim, args.steps, args.warmup)
    sampler.mark(t_wall0, t_wall1)
    launches = int(sum_over_ranks(float(launches_local)))
    value = cells_global * args.steps / (ms * 1e-3) / 1e9
    clock_note = None
    if ms < 40.0:
        # the timed region was shorter than a handful of NVML samples (e.g. --steps 20 = 4 ms): keep the SAME kernel running,
        # untimed, for ~60 ms so that the sampler sees the GPU under this load.  The count derives from `ms` (max over ranks),
        # so every rank runs the same number of steps — strips step in lock step with their neighbours.
        extra = int(min(4000, max(4 * k, 60.0 / max(ms / args.steps, 1e-3))))
        extra = (extra + k - 1) // k * k
        barrier()
        t0 = time.perf_counter()
        sim.step(extra)
        sim.sync()
        sampler.mark(t0, time.perf_counter())
        barrier()
        clock_note = f"timed region ({ms:.1f} ms) shorter than a handful of NVML samples: clocks sampled over it and over {extra} further untimed steps of the same kernel run right after it"
    clocks = sampler.summary(clock_note)

    # ---- roofline of the dominant kernel (the step kernel; all launches in the region are it) -------
    peak, peak_src = measured_peak_gbs()
    launch_ms = ms_dev / max(1, (args.steps + k - 1) // k)
    algo_bytes_per_launch = ALGO_BYTES_PER_CELL_UPDATE * cells_local * min(k, args.steps)
    achieved = algo_bytes_per_launch / (launch_ms * 1e-3) / 1e9
    tag = {"unfused": "unfused", "fused": "fused_k1", "tb": f"fused_k{k}", "stream": f"stream_k{k}", "band": f"band_k{k}"}[args.backend]
    traffic = ncu_traffic_per_launch(tag) if (W == 8192 and sim.rows == 8192) else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": "profiles/traffic.json (ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel on this grid, committed; not re-measured in this run)" if traffic else None,
                "dram_frac": (traffic / (launch_ms * 1e-3) / 1e9 / peak) if traffic else None,
                "kernel": {"unfused": "unfused_update_kernel+unfused_apply_kernel", "stream": "stream_step_kernel", "band": "band_step_kernel"}.get(args.backend, "fused_step_kernel"),
                "peak_source": peak_src, "algorithmic_bytes_per_cell_update": ALGO_BYTES_PER_CELL_UPDATE,
                "cell_updates_per_launch": cells_local * min(k, args.steps), "avg_launch_ms": launch_ms,
                "note": "frac = 48 B x cell-updates per launch / launch time / peak (SURVEY 8d basis); with temporal blocking (k steps per HBM round trip) "
                        "real DRAM traffic is ~48/k B per cell-update, so frac may exceed 1; dram_frac = ncu DRAM bytes per launch / launch time / peak is the "
                        "share of the HBM bandwidth the kernel really uses"}

    # ---- end to end through the C ABI with host buffers ---------------------------------------------
    e2e = None
    if not args.no_e2e:
        n_e2e = args.e2e_steps if args.e2e_steps > 0 else min(args.steps, 40)
        d_host = torch.empty((sim.rows, W), dtype=torch.float32, pin_memory=True)
        v_host = torch.empty((sim.rows, W, 2), dtype=torch.float16, pin_memory=True)
        sim.readback_raw(tws.FIELD_WATER, d_host.data_ptr(), d_host.numel() * 4)

        def serial_step():
            # the same frame as four separate calls (no overlap between the copies and the kernels)
            sim.upload_raw(tws.FIELD_WATER, d_host.data_ptr(), d_host.numel() * 4)
            if world > 1:
                sim.halo_refresh()
            sim.step(1)
            sim.readback_raw(tws.FIELD_WATER, d_host.data_ptr(), d_host.numel() * 4)
            sim.readback_raw(tws.FIELD_VELOCITY, v_host.data_ptr(), v_host.numel() * 2)

        def host_step():
            # host-driven frame through ONE C-ABI call: the host owns the water layer (pinned), the
            # library uploads it, runs one step and returns the new water layer and the flow map,
            # band-pipelined so upload, kernels and readback overlap (tws_step_host).
            sim.step_host_raw(d_host.data_ptr(), d_host.data_ptr(), v_host.data_ptr())

        timings = {}
        for name, fn in (("serial", serial_step), ("host", host_step)):
            for _ in range(3):
                fn()
            barrier()
            l_e = sim.kernel_launches()
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                fn()
            torch.cuda.synchronize()
            timings[name] = max_over_ranks(time.perf_counter() - t0)
            timings[name + "_launches"] = sim.kernel_launches() - l_e
            barrier()
        dt = timings["host"]
        e2e = {"value": cells_global * n_e2e / dt / 1e9, "unit": "Gcell-updates/s", "h2d_bytes_per_step": int(sum_over_ranks(cells_local * 4.0)),
               "d2h_bytes_per_step": int(sum_over_ranks(cells_local * 8.0)), "steps": n_e2e, "ms_per_step": dt / n_e2e * 1e3,
               "gpu_launches": int(timings["host_launches"]),
               "serial_calls_value": cells_global * n_e2e / timings["serial"] / 1e9,
               "pcie_gbs_per_gpu": {"h2d": cells_local * 4.0 * n_e2e / dt / 1e9, "d2h": cells_local * 8.0 * n_e2e / dt / 1e9},
               "what": "per step ONE call tws_step_host(water_in, water_out, velocity_out) with pinned host buffers: 4 B/cell up, 8 B/cell down, "
                       "uploaded / stepped / read back in row bands on three streams so both PCIe directions overlap the kernels (strips included: the "
                       "uploaded edge rows are pushed to the neighbours first, the output edge rows last); terrain and flux "
                       "stay device-resident as in the reference (its state never leaves the GPU). serial_calls_value = the same frame as "
                       "tws_upload + tws_step(1) + 2 x tws_readback (no overlap)"}
        del d_host, v_host

    def strips_match_whole_grid(s, pl, Wc, Hc, tile_c, n_chk):
        """resets the strips `s` (one per rank) to the reference scene, puts a brush across the first seam, steps n_chk and
        compares depth, flux and velocity bit for bit with a whole-grid run of the UNFUSED kernels on rank 0."""
        barrier()                                    # strips must be quiescent around a reset
        s.CreateHeightmapFromNoiseAndResetSim(tileHeight=tile_c)
        barrier()
        s.inject_brush(Wc / 2 + 0.25, pl.rows(1)[0] - 0.5, 5.0, 64.0)
        s.step(n_chk)
        s.sync()
        mine = (s.readback(tws.FIELD_WATER), s.readback(tws.FIELD_FLUX), s.readback(tws.FIELD_VELOCITY).view(np.uint16))
        parts = [None] * world if rank == 0 else None
        dist.gather_object(mine, parts, dst=0)
        ok = True
        if rank == 0:
            with tws.Terrain(Wc, height=Hc, backend=tws.BACKEND_UNFUSED, device=local) as whole:
                whole.CreateHeightmapFromNoiseAndResetSim(tileHeight=tile_c)
                whole.inject_brush(Wc / 2 + 0.25, pl.rows(1)[0] - 0.5, 5.0, 64.0)
                whole.step(n_chk)
                ok = (np.array_equal(np.concatenate([p[0] for p in parts]).view(np.uint32), whole.readback(tws.FIELD_WATER).view(np.uint32))
                      and np.array_equal(np.concatenate([p[1] for p in parts]).view(np.uint32), whole.readback(tws.FIELD_FLUX).view(np.uint32))
                      and np.array_equal(np.concatenate([p[2] for p in parts]), whole.readback(tws.FIELD_VELOCITY).view(np.uint16)))
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.broadcast(flag, src=0)
        return bool(flag.item())

    # ---- strips vs whole grid on the bench grid itself (tests, small --size) ---------------------------
    if args.verify_strips and world > 1:
        ok = strips_match_whole_grid(sim, plan, W, Hg, tile, 12)
        if rank == 0:
            print("STRIPS_VERIFIED" if ok else "STRIPS_MISMATCH", flush=True)
        if not ok:
            raise SystemExit(3)

    sim.close()
    del sim
    torch.cuda.empty_cache()

    # ---- multi-GPU parity, every run with N > 1: 1024 x 1024*N, band kernel k = 4, brush across a seam -----
    strips_verified = None
    if world > 1:
        Wv, Hv = 1024, 1024 * world
        planv = tws.plan_strips(Hv, world)
        simv = tws.Terrain(Wv, height=Hv, rows=planv.rows(rank), backend=tws.BACKEND_BAND_TB, temporal_block=4, device=local)
        tws.connect_strips(simv, planv, rank)
        strips_verified = strips_match_whole_grid(simv, planv, Wv, Hv, 0, 24)
        simv.close()
        if rank == 0 and not strips_verified:
            print("STRIPS_MISMATCH (1024 x 1024*N parity run)", flush=True)
        if not strips_verified:
            raise SystemExit(3)

    # ---- strong scaling, BASELINE config 4: the SAME 32768^2 grid on 1 GPU (T1, rank 0) and split over the N ranks (TN) ----
    strong = None
    if not args.no_strong and not args.strong:
        S = args.strong_size
        ks = 4
        need_gb = 48.0 * S * S / 1e9 + 4.0
        free_gb = -max_over_ranks(-torch.cuda.mem_get_info()[0] / 1e9)     # the least free memory of any rank: one decision for all
        if free_gb < need_gb:
            strong = {"skipped": f"{S}x{S} needs {need_gb:.0f} GB on one GPU for T1, {free_gb:.0f} GB free"}
        else:
            t1_ms = None
            if rank == 0:
                with tws.Terrain(S, backend=tws.BACKEND_BAND_TB, temporal_block=ks, device=local) as whole:
                    whole.CreateHeightmapFromNoiseAndResetSim()
                    whole.sync()
                    whole.step(8); whole.sync()
                    whole.step(args.strong_steps); whole.sync()
                    t1_ms = whole.elapsed_ms() / args.strong_steps
                torch.cuda.empty_cache()
            if world > 1:
                t = torch.tensor([t1_ms or 0.0], dtype=torch.float64, device="cuda")
                dist.broadcast(t, src=0)
                t1_ms = float(t.item())
                plans = tws.plan_strips(S, world)
                sims = tws.Terrain(S, height=S, rows=plans.rows(rank), backend=tws.BACKEND_BAND_TB, temporal_block=ks, device=local)
                tws.connect_strips(sims, plans, rank)
                sims.CreateHeightmapFromNoiseAndResetSim()
                sims.sync()
                tn_total, _, _, _, _ = timed_steps(sims, args.strong_steps, 8)
                tn_ms = tn_total / args.strong_steps
                sims.close()
            else:
                tn_ms = t1_ms
            strong = {"grid": [S, S], "workload": f"BASELINE config 4: {S}x{S} reference scene, band kernel k={ks}, {world} row strip(s)",
                      "steps": args.strong_steps, "t1_ms_per_step": t1_ms, "tn_ms_per_step": tn_ms,
                      "value": S * S / (tn_ms * 1e-3) / 1e9, "unit": "Gcell-updates/s", "t1_value": S * S / (t1_ms * 1e-3) / 1e9,
                      "efficiency": t1_ms / (world * tn_ms),
                      "definition": "T1 / (N x TN), both measured in THIS run: T1 = the whole grid on rank 0's GPU alone, TN = the same grid as N strips, device time, max over ranks (SURVEY 8d)"}

    # ---- the reference's own operating point (BASELINE configs[0/1] grid, Terrain.cpp:240-277): 1024^2, frames of up to 10 steps
    frame = None
    if rank == 0 and not args.no_frame:
        try:
            frame = reference_frame_times(tws, local)
        except Exception as e:                          # an extra figure must never cost the bench line
            frame = {"error": str(e)}

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        val, cores, sample, _ = cpu_oracle_throughput(W, 4, 1, budget_s=20.0)
        cpu = {"value": val, "unit": "Gcell-updates/s", "cores": cores, "kind": "port", "sample": sample}
        # the parity reference itself, one core
        try:
            v1, _, s1, _ = cpu_oracle_throughput(W, 2, 1, budget_s=6.0, openmp=False)
            cpu["single_thread"] = {"value": v1, "unit": "Gcell-updates/s", "cores": 1, "sample": s1}
        except Exception as e:                      # an extra figure must never cost the bench line
            cpu["single_thread"] = {"error": str(e)}

    if rank == 0:
        line = {
            "metric": "Gcell-updates/s per sim step (fp32)", "value": value, "unit": "Gcell-updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.strong else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, world), "roofline": roofline,
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "per_gpu_value": value / world, "strips_verified": strips_verified, "strong": strong, "frame_1024": frame,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")     # one hardware queue per stream (set before CUDA initialises)
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
