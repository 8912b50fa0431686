"""Strip decomposition on real GPUs: several strip sims wired together with tws_halo_connect
(same process: direct pointers / peer access) must reproduce the whole-grid result bitwise,
with the NVLink push + flag protocol doing the exchange.  Multi-process CUDA-IPC wiring is
covered by test_two_process_ipc (needs >= 2 GPUs; skipped otherwise)."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle.oracle_py import new_state

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def bumpy(W, H, seed=5):
    rng = np.random.default_rng(seed)
    h = (rng.random((H, W)) * 8).astype(np.float32)
    d = (rng.random((H, W)) * 4 * (rng.random((H, W)) > 0.4)).astype(np.float32)
    return h, d


def n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("nstrips,k,backend", [(2, 1, 2), (2, 4, 3), (3, 2, 3), (4, 3, 3), (2, 4, 4), (3, 1, 4), (4, 3, 4), (2, 4, 5), (3, 2, 5), (4, 3, 5)])
@pytest.mark.parametrize("spread", [False, True], ids=["one-gpu", "multi-gpu"])
def test_strips_in_one_process_match_whole_grid(tws, oracle_omp, nstrips, k, backend, spread):
    if spread and n_gpus() < 2:
        pytest.skip("needs >= 2 GPUs")
    W, H, steps = 300, 40 * nstrips + 8, 48
    h, d = bumpy(W, H)
    c = oracle_omp.derive_consts(float(W), W)
    t, f, v = new_state(h, d)
    plan = tws.plan_strips(H, nstrips)
    sims = [tws.Terrain(W, height=H, rows=plan.rows(i), backend=backend, temporal_block=k, device=(i % n_gpus()) if spread else 0)
            for i in range(nstrips)]
    try:
        handles = [s.halo_export() for s in sims]
        for i, s in enumerate(sims):
            s.halo_connect(handles[i - 1] if i > 0 else None, handles[i + 1] if i + 1 < nstrips else None)
        for i, s in enumerate(sims):
            r0, r1 = plan.rows(i)
            s.upload(tws.FIELD_TERRAIN, h[r0:r1]); s.upload(tws.FIELD_WATER, d[r0:r1])
        for s in sims:
            s.halo_refresh()
        for s in sims:
            s.sync()
        done = 0
        while done < steps:                         # small batches: one host thread feeds all strips
            n = min(4, steps - done)
            oracle_omp.brush(t, 150.25, float(plan.rows(1)[0]) - 0.5, 0.75, 32.0)      # brush straddling a strip boundary
            for s in sims:
                s.inject_brush(150.25, float(plan.rows(1)[0]) - 0.5, 0.75, 32.0)
            for s in sims:
                s.step(n)
            done += n
        oracle_omp.step  # (oracle stepped below in the same batch pattern)
        t2, f2, v2 = new_state(h, d)
        done = 0
        while done < steps:
            n = min(4, steps - done)
            oracle_omp.brush(t2, 150.25, float(plan.rows(1)[0]) - 0.5, 0.75, 32.0)
            oracle_omp.step(t2, f2, v2, c, n)
            done += n
        for s in sims:
            s.sync()
        gd = np.concatenate([s.readback(tws.FIELD_WATER) for s in sims])
        gf = np.concatenate([s.readback(tws.FIELD_FLUX) for s in sims])
        gv = np.concatenate([s.readback(tws.FIELD_VELOCITY) for s in sims])
        assert np.array_equal(gd.view(np.uint32), t2[..., 3].view(np.uint32))
        assert np.array_equal(gf.view(np.uint32), f2.view(np.uint32))
        assert np.array_equal(gv.view(np.uint16), v2.view(np.uint16))
        vol = sum(s.total_volume() for s in sims)
        assert vol == pytest.approx(float(t2[..., 3].sum(dtype=np.float64)), rel=1e-12)
    finally:
        for s in sims:
            s.close()


@pytest.mark.parametrize("rows_per_strip,nstrips,k,backend", [(40, 3, 4, 5), (10, 4, 4, 5), (9, 3, 2, 5), (16, 3, 3, 5), (12, 3, 3, 4), (40, 3, 2, 3)])
def test_strips_closed_boundary_with_sources(tws, oracle_omp, rows_per_strip, nstrips, k, backend):
    """Extensions across strip seams: closed boundary (the wall rule applies at the GLOBAL edge only), rain and
    evaporation, a wide grid (three column strips of the streaming kernels), and strips so short that a strip is
    nothing but edge rows pushed to BOTH neighbours (the band kernel's exchange is fused into its step launch)."""
    W, H, steps = 300, rows_per_strip * nstrips, 20
    h, d = bumpy(W, H, seed=21)
    c = oracle_omp.derive_consts(float(W), W)
    dt = float(np.float32(1.0) / np.float32(60.0))
    rain, evap = 0.8, 0.3
    rs, es = float(np.float32(dt * np.float32(rain))), float(np.float32(dt * np.float32(evap)))
    t, f, v = new_state(h, d)
    oracle_omp.step(t, f, v, c, steps, boundary=1, rain_step=rs, evap_step=es)
    bounds = [(i * rows_per_strip, (i + 1) * rows_per_strip) for i in range(nstrips)]
    sims = [tws.Terrain(W, height=H, rows=bounds[i], backend=backend, temporal_block=k, device=i % n_gpus(),
                        boundary=tws.BOUNDARY_CLOSED, rain_rate=rain, evaporation_rate=evap) for i in range(nstrips)]
    try:
        handles = [s.halo_export() for s in sims]
        for i, s in enumerate(sims):
            s.halo_connect(handles[i - 1] if i > 0 else None, handles[i + 1] if i + 1 < nstrips else None)
        for i, s in enumerate(sims):
            r0, r1 = bounds[i]
            s.upload(tws.FIELD_TERRAIN, h[r0:r1]); s.upload(tws.FIELD_WATER, d[r0:r1])
        for s in sims:
            s.halo_refresh()
        for s in sims:
            s.sync()
        done = 0
        while done < steps:
            n = min(k, steps - done)               # one block per call: one host thread feeds all strips
            for s in sims:
                s.step(n)
            done += n
        for s in sims:
            s.sync()
        gd = np.concatenate([s.readback(tws.FIELD_WATER) for s in sims])
        gf = np.concatenate([s.readback(tws.FIELD_FLUX) for s in sims])
        gv = np.concatenate([s.readback(tws.FIELD_VELOCITY) for s in sims])
        assert np.array_equal(gd.view(np.uint32), t[..., 3].view(np.uint32))
        assert np.array_equal(gf.view(np.uint32), f.view(np.uint32))
        assert np.array_equal(gv.view(np.uint16), v.view(np.uint16))
    finally:
        for s in sims:
            s.close()


def test_strip_scene_generation_needs_no_exchange(tws, oracle_omp):
    W, H = 256, 256
    plan = tws.plan_strips(H, 2)
    sims = [tws.Terrain(W, height=H, rows=plan.rows(i), backend=tws.BACKEND_FUSED_TB, temporal_block=4) for i in range(2)]
    try:
        hd = [s.halo_export() for s in sims]
        sims[0].halo_connect(None, hd[1]); sims[1].halo_connect(hd[0], None)
        for s in sims:
            s.CreateHeightmapFromNoiseAndResetSim()
        for _ in range(5):
            for s in sims:
                s.step(4)
        for s in sims:
            s.sync()
        s0 = oracle_omp.create_scene(W)
        t, f, v = s0, np.zeros((H, W, 4), np.float32), np.zeros((H, W, 2), np.float16)
        oracle_omp.step(t, f, v, oracle_omp.derive_consts(float(W), W), 20)
        gd = np.concatenate([s.readback(tws.FIELD_WATER) for s in sims])
        assert np.array_equal(gd.view(np.uint32), t[..., 3].view(np.uint32))
    finally:
        for s in sims:
            s.close()


def test_unconnected_strip_refuses_to_step(tws):
    with tws.Terrain(64, height=64, rows=(0, 32)) as s:
        with pytest.raises(tws.TwsError) as e:
            s.step(1)
        assert e.value.status == tws._abi.TWS_ERR_STATE


def test_two_process_ipc(tws, tmp_path):
    """bench.py's multi-GPU path end to end: torchrun, one process per GPU, CUDA-IPC halo wiring,
    strips compared with a single-GPU run of the same grid inside the script."""
    if n_gpus() < 2:
        pytest.skip("needs >= 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "bench.py"), "--gpus", "2", "--steps", "2", "--warmup", "1", "--size", "1024",
           "--verify-strips", "--no-strong", "--no-cpu-baseline", "--no-frame"]
    out = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "STRIPS_VERIFIED" in out.stdout


@pytest.mark.parametrize("k,backend", [(4, 5), (2, 3), (3, 4)])
def test_strip_source_ledgers_add_up(tws, oracle, k, backend):
    """Rain + evaporation on strips: every strip books the source deltas of its own rows; together they are the oracle's sum."""
    W, H, nstrips, steps = 200, 96, 3, 24
    h, d = bumpy(W, H, seed=9)
    c = oracle.derive_consts(float(W), W)
    dt = float(np.float32(1.0) / np.float32(60.0))
    rain, evap = 0.5, 1.5
    rs, es = float(np.float32(dt * np.float32(rain))), float(np.float32(dt * np.float32(evap)))
    t, f, v = new_state(h, d)
    plan = tws.plan_strips(H, nstrips)
    sims = [tws.Terrain(W, height=H, rows=plan.rows(i), backend=backend, temporal_block=k, device=0, rain_rate=rain, evaporation_rate=evap)
            for i in range(nstrips)]
    try:
        handles = [s.halo_export() for s in sims]
        for i, s in enumerate(sims):
            s.halo_connect(handles[i - 1] if i > 0 else None, handles[i + 1] if i + 1 < nstrips else None)
        for i, s in enumerate(sims):
            r0, r1 = plan.rows(i)
            s.upload(tws.FIELD_TERRAIN, h[r0:r1]); s.upload(tws.FIELD_WATER, d[r0:r1])
        for s in sims:
            s.halo_refresh()
        for s in sims:
            s.sync()
        done = 0
        while done < steps:
            n = min(4, steps - done)
            for s in sims:
                s.step(n)
            done += n
        want = 0.0
        for _ in range(steps):
            oracle.flow_update(t, f, c)
            t0 = t.copy(); v0 = v.copy()
            oracle.flow_apply(t0, f, v0, c)
            oracle.flow_apply(t, f, v, c, rs, es)
            want += float((t[..., 3].astype(np.float64) - t0[..., 3].astype(np.float64)).sum())
        assert sum(s.source_accumulated() for s in sims) == pytest.approx(want, rel=1e-11)
        got = np.concatenate([s.readback(tws.FIELD_WATER) for s in sims])
        assert np.array_equal(got.view(np.uint32), np.ascontiguousarray(t[..., 3]).view(np.uint32))
    finally:
        for s in sims:
            s.close()


@pytest.mark.parametrize("k,backend", [(4, 5), (2, 3), (3, 4)])
def test_strip_ledgers_add_up_to_the_whole_grid_ledger(tws, oracle, k, backend):
    """Each strip counts the outflow through ITS part of the global edge (both side columns, row 0 only on the first strip,
    the last row only on the last); the per-strip ledgers add up to the oracle's whole-grid sum."""
    W, H, nstrips, steps = 200, 96, 3, 40
    h, d = bumpy(W, H, seed=9)
    c = oracle.derive_consts(float(W), W)
    t, f, v = new_state(h, d)
    plan = tws.plan_strips(H, nstrips)
    sims = [tws.Terrain(W, height=H, rows=plan.rows(i), backend=backend, temporal_block=k, device=0) for i in range(nstrips)]
    try:
        handles = [s.halo_export() for s in sims]
        for i, s in enumerate(sims):
            s.halo_connect(handles[i - 1] if i > 0 else None, handles[i + 1] if i + 1 < nstrips else None)
        for i, s in enumerate(sims):
            r0, r1 = plan.rows(i)
            s.upload(tws.FIELD_TERRAIN, h[r0:r1]); s.upload(tws.FIELD_WATER, d[r0:r1])
        for s in sims:
            s.halo_refresh()
        for s in sims:
            s.sync()
        done = 0
        while done < steps:
            n = min(4, steps - done)
            for s in sims:
                s.step(n)
            done += n
        want = 0.0
        for _ in range(steps):
            oracle.flow_update(t, f, c)
            want += float(f[:, -1, 0].sum(dtype=np.float64) + f[:, 0, 1].sum(dtype=np.float64) + f[-1, :, 2].sum(dtype=np.float64)
                          + f[0, :, 3].sum(dtype=np.float64)) * float(c[2])
            oracle.flow_apply(t, f, v, c)
        parts = [s.boundary_outflow_accumulated() for s in sims]
        assert all(p > 0 for p in parts)
        assert sum(parts) == pytest.approx(want, rel=1e-12)
        got = np.concatenate([s.readback(tws.FIELD_WATER) for s in sims])
        assert np.array_equal(got.view(np.uint32), np.ascontiguousarray(t[..., 3]).view(np.uint32))
    finally:
        for s in sims:
            s.close()


@pytest.mark.parametrize("nstrips,backend,k,rows_per_strip", [(2, 5, 4, 1100), (3, 5, 1, 24), (3, 2, 1, 200), (2, 4, 3, 8), (3, 3, 2, 40), (2, 5, 1, 9), (2, 2, 1, 300)])
@pytest.mark.parametrize("spread", [False, True], ids=["one-gpu", "multi-gpu"])
def test_step_host_on_strips_is_pipelined_and_matches_the_oracle(tws, oracle_omp, nstrips, backend, k, rows_per_strip, spread):
    """tws_step_host on strips (the e2e leg of bench.py at N > 1): every step the host uploads a NEW water layer (the edge
    rows go up first and are pushed into the neighbours' halos), one step runs in row bands, water and velocity come back.
    One host thread per strip, as one process per GPU would call it.  Bit-identical to the oracle fed the same water."""
    import threading
    import torch
    if spread and n_gpus() < 2:
        pytest.skip("needs >= 2 GPUs")
    if not spread and nstrips > 2:
        # each sim drives four streams here (compute, exchange, upload, readback) and strips wait for each other with
        # spinning flag kernels: more than two such sims on ONE GPU exceed its 8 hardware queues, streams alias, and a
        # wait kernel can end up in front of the very push it waits for.  One process per GPU never gets there.
        pytest.skip("more than two strips with tws_step_host need more than one GPU")
    W, H, steps = 2100, rows_per_strip * nstrips, 5
    h, d = bumpy(W, H, seed=31)
    c = oracle_omp.derive_consts(float(W), W)
    t, f, v = new_state(h, d)
    bounds = [(i * rows_per_strip, (i + 1) * rows_per_strip) for i in range(nstrips)]
    sims = [tws.Terrain(W, height=H, rows=bounds[i], backend=backend, temporal_block=k, device=(i % n_gpus()) if spread else 0) for i in range(nstrips)]
    errors = []
    try:
        handles = [s.halo_export() for s in sims]
        for i, s in enumerate(sims):
            s.halo_connect(handles[i - 1] if i > 0 else None, handles[i + 1] if i + 1 < nstrips else None)
        for i, s in enumerate(sims):
            r0, r1 = bounds[i]
            s.upload(tws.FIELD_TERRAIN, h[r0:r1]); s.upload(tws.FIELD_WATER, d[r0:r1])
        for s in sims:
            s.halo_refresh()
        for s in sims:
            s.sync()
        rng = np.random.default_rng(3)
        # page-locked host buffers, as tws.h asks for (pageable memory makes the copies block the calling thread)
        keep = [torch.from_numpy(np.ascontiguousarray(d[r0:r1])).pin_memory() for r0, r1 in bounds]
        keepv = [torch.zeros((r1 - r0, W, 2), dtype=torch.float16).pin_memory() for r0, r1 in bounds]
        water = [x.numpy() for x in keep]
        vel = [x.numpy() for x in keepv]
        for step in range(steps):
            # the host edits the water between steps (what a host-side consumer of the ABI would do)
            bump = (rng.random((H, W)) * 0.25 * (rng.random((H, W)) > 0.7)).astype(np.float32)
            for i, (r0, r1) in enumerate(bounds):
                water[i] += bump[r0:r1]
            t[..., 3] += bump
            assert np.array_equal(np.concatenate(water).view(np.uint32), np.ascontiguousarray(t[..., 3]).view(np.uint32))

            def run(i):
                try:
                    sims[i].step_host(water[i], water[i], vel[i])
                except Exception as e:          # noqa: BLE001
                    errors.append((i, e))
            th = [threading.Thread(target=run, args=(i,)) for i in range(nstrips)]
            for x in th:
                x.start()
            for x in th:
                x.join()
            assert not errors, errors
            oracle_omp.step(t, f, v, c, 1)
            assert np.array_equal(np.concatenate(water).view(np.uint32), np.ascontiguousarray(t[..., 3]).view(np.uint32)), f"water differs after step {step}"
            assert np.array_equal(np.concatenate(vel).view(np.uint16), v.view(np.uint16)), f"velocity differs after step {step}"
        # the resident state is consistent too: plain steps continue from it
        for _ in range(2):
            for s in sims:
                s.step(k)
        oracle_omp.step(t, f, v, c, 2 * k)
        gd = np.concatenate([s.readback(tws.FIELD_WATER) for s in sims])
        gf = np.concatenate([s.readback(tws.FIELD_FLUX) for s in sims])
        assert np.array_equal(gd.view(np.uint32), np.ascontiguousarray(t[..., 3]).view(np.uint32))
        assert np.array_equal(gf.view(np.uint32), f.view(np.uint32))
    finally:
        for s in sims:
            s.close()


@pytest.mark.parametrize("W,H,nstrips", [(256, 256, 2), (300, 192, 3), (64, 512, 4)])
def test_strips_publish_their_own_mip_levels(tws, oracle, W, H, nstrips):
    """A strip filters the levels its 8-row-aligned cut allows from its own rows; stacked, they are the whole grid's levels."""
    from oracle.oracle_py import mip_chain
    h, d = bumpy(W, H, seed=2)
    info = np.empty((H, W, 4), np.float32)
    info[..., 0] = h; info[..., 1] = 0.3; info[..., 2] = 0.3; info[..., 3] = d
    want = mip_chain(info)
    plan = tws.plan_strips(H, nstrips)
    sims = [tws.Terrain(W, height=H, rows=plan.rows(i), backend=tws.BACKEND_BAND_TB, temporal_block=2) for i in range(nstrips)]
    try:
        handles = [s.halo_export() for s in sims]
        for i, s in enumerate(sims):
            s.halo_connect(handles[i - 1] if i > 0 else None, handles[i + 1] if i + 1 < nstrips else None)
        for i, s in enumerate(sims):
            r0, r1 = plan.rows(i)
            s.upload(tws.FIELD_TERRAIN, h[r0:r1]); s.upload(tws.FIELD_WATER, d[r0:r1])
        chains = [s.publish_mips() for s in sims]
        common = min(len(ch) for ch in chains)
        assert common >= 4                                   # cuts on multiples of 8 rows: levels 0..3
        for l in range(common):
            got = np.concatenate([ch[l] for ch in chains])
            assert got.shape == want[l].shape
            assert np.array_equal(got.view(np.uint32), want[l].view(np.uint32)), f"level {l}"
    finally:
        for s in sims:
            s.close()


def test_wide_band_strips_on_one_gpu_fall_back_to_the_two_stream_exchange(tws, oracle_omp):
    """Strips that share one GPU and are wider than the resident warp groups could serve (2 x column strips >= SMs) must
    not spin on each other inside one launch: they use the two-stream exchange and still match."""
    W, nstrips, rows = 8400, 2, 24
    H = rows * nstrips
    h, d = bumpy(W, H, seed=12)
    c = oracle_omp.derive_consts(float(W), W)
    t, f, v = new_state(h, d)
    sims = [tws.Terrain(W, height=H, rows=(i * rows, (i + 1) * rows), backend=tws.BACKEND_BAND_TB, temporal_block=4) for i in range(nstrips)]
    try:
        handles = [s.halo_export() for s in sims]
        sims[0].halo_connect(None, handles[1]); sims[1].halo_connect(handles[0], None)
        for i, s in enumerate(sims):
            s.upload(tws.FIELD_TERRAIN, h[i * rows:(i + 1) * rows]); s.upload(tws.FIELD_WATER, d[i * rows:(i + 1) * rows])
        for s in sims:
            s.halo_refresh()
        for s in sims:
            s.sync()
        for _ in range(3):
            for s in sims:
                s.step(4)
        for s in sims:
            s.sync()
        oracle_omp.step(t, f, v, c, 12)
        gd = np.concatenate([s.readback(tws.FIELD_WATER) for s in sims])
        assert np.array_equal(gd.view(np.uint32), np.ascontiguousarray(t[..., 3]).view(np.uint32))
    finally:
        for s in sims:
            s.close()
