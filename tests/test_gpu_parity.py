"""Parity tests proper: the CUDA path behind the C ABI against the CPU oracle, bit for bit.
north_star asks for <= 1e-5 relative per cell after 1000 steps; the design pins one
arithmetic contract on both sides (no FMA, IEEE div, same operation order), so the bar
enforced here is stricter: EXACT equality of every depth, flux and fp16 velocity bit."""
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle.oracle_py import dam_break, new_state

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
GOLD = json.loads((ROOT / "tests" / "golden" / "goldens.json").read_text())

BACKENDS = [("unfused", 1, 1), ("fused", 2, 1), ("tb2", 3, 2), ("tb3", 3, 3), ("tb4", 3, 4),
            ("stream1", 4, 1), ("stream2", 4, 2), ("stream3", 4, 3), ("stream4", 4, 4),
            ("band1", 5, 1), ("band2", 5, 2), ("band3", 5, 3), ("band4", 5, 4),
            ("resident", 6, 1)]       # grids that fit on chip only: all n steps of a call in one launch


def bits(a):
    return a.view(np.uint32) if a.dtype == np.float32 else a.view(np.uint16)


def assert_state_equal(sim, tws, t, f, v, what=""):
    gd = sim.readback(tws.FIELD_WATER)
    gf = sim.readback(tws.FIELD_FLUX)
    gv = sim.readback(tws.FIELD_VELOCITY)
    bad = int((bits(gd) != bits(np.ascontiguousarray(t[..., 3]))).sum())
    assert bad == 0, f"{what}: {bad} depth cells differ, max abs {np.abs(gd - t[..., 3]).max()}"
    assert np.array_equal(bits(gf), bits(f)), f"{what}: flux differs"
    assert np.array_equal(bits(gv), bits(v)), f"{what}: velocity differs"


def make_sim(tws, W, H, backend, k, **kw):
    return tws.Terrain(W, height=H, backend=backend, temporal_block=k, **kw)


def bumpy(W, H, seed=5, wall=False):
    rng = np.random.default_rng(seed)
    h = (rng.random((H, W)) * 8).astype(np.float32)
    d = (rng.random((H, W)) * 4 * (rng.random((H, W)) > 0.4)).astype(np.float32)
    if wall and H > 2 and W > 2:
        h[0] = h[-1] = 500; h[:, 0] = h[:, -1] = 500
        d[0] = d[-1] = 0; d[:, 0] = d[:, -1] = 0
    return h, d


@pytest.mark.parametrize("name,backend,k", BACKENDS)
@pytest.mark.parametrize("rim", [True, False], ids=["walled", "open"])
def test_config1_dam_break_1000_steps(tws, oracle_omp, name, backend, k, rim):
    """BASELINE config 1: 256x256 dam break, 1000 steps, checked at 1/10/100/1000 against the
    committed oracle goldens and against the live oracle."""
    h, d = dam_break(256, rim=rim)
    c = oracle_omp.derive_consts(256.0, 256)
    t, f, v = new_state(h, d)
    gold = GOLD["dam256_walled" if rim else "dam256_open"]
    with make_sim(tws, 256, 256, backend, k) as sim:
        assert np.array_equal(np.float32(sim.step_constants()), c)
        sim.upload(tws.FIELD_TERRAIN, h)
        sim.upload(tws.FIELD_WATER, d)
        v0 = sim.total_volume()
        done = 0
        for n in (1, 10, 100, 1000):
            sim.step(n - done)
            oracle_omp.step(t, f, v, c, n - done)
            done = n
            assert_state_equal(sim, tws, t, f, v, f"{name} after {n}")
            gd = sim.readback(tws.FIELD_WATER)
            assert f"{oracle_omp.fnv1a64(gd):016x}" == gold[str(n)]["d"]
            assert f"{oracle_omp.fnv1a64(sim.readback(tws.FIELD_FLUX)):016x}" == gold[str(n)]["F"]
            assert f"{oracle_omp.fnv1a64(sim.readback(tws.FIELD_VELOCITY).view(np.uint16)):016x}" == gold[str(n)]["v"]
        v1 = sim.total_volume()
        assert v1 == gd.sum(dtype=np.float64) or abs(v1 - gd.sum(dtype=np.float64)) < 1e-6 * v0
        if rim:
            assert abs(v1 - v0) / v0 < 1e-6          # closed domain: volume conserved to 1e-6 (north_star)


@pytest.mark.parametrize("name,backend,k", [BACKENDS[0], BACKENDS[1], BACKENDS[4], BACKENDS[6], BACKENDS[8], BACKENDS[10], BACKENDS[12]])
def test_config2_reference_scene_with_brush_1000_steps(tws, oracle_omp, name, backend, k):
    """BASELINE config 2: 1024x1024 reference default scene (generated ON THE GPU), brush at
    (512,512) with intensity 100/60 before every step (Scene.cpp:356-363), 1000 steps."""
    gold = GOLD["scene1024_brush"]
    with make_sim(tws, 1024, 1024, backend, k) as sim:
        sim.CreateHeightmapFromNoiseAndResetSim()
        assert f"{oracle_omp.fnv1a64(sim.readback(tws.FIELD_TERRAIN)):016x}" == GOLD["scene1024_ref"]["h"]
        assert f"{oracle_omp.fnv1a64(sim.readback(tws.FIELD_WATER)):016x}" == GOLD["scene1024_ref"]["d"]
        done = 0
        for n in (1, 10, 100, 1000):
            for _ in range(n - done):
                sim.inject_brush(512.0, 512.0, float(np.float32(100.0 / 60.0)), 32.0)
                sim.step(1)
            done = n
            assert f"{oracle_omp.fnv1a64(sim.readback(tws.FIELD_WATER)):016x}" == gold[str(n)]["d"], f"{name} depth after {n}"
            assert f"{oracle_omp.fnv1a64(sim.readback(tws.FIELD_FLUX)):016x}" == gold[str(n)]["F"], f"{name} flux after {n}"
            assert f"{oracle_omp.fnv1a64(sim.readback(tws.FIELD_VELOCITY).view(np.uint16)):016x}" == gold[str(n)]["v"]
        assert abs(sim.total_volume() - gold["1000"]["volume"]) < 1e-9 * gold["1000"]["volume"]


@pytest.mark.parametrize("W,H", [(1, 1), (2, 3), (5, 37), (37, 5), (130, 29), (250, 190), (257, 64), (640, 333)])
@pytest.mark.parametrize("name,backend,k", BACKENDS)
def test_ragged_sizes_open_boundary(tws, oracle, W, H, name, backend, k):
    """The reference silently needs res % 16 == 0 (Terrain.cpp:258); the replacement accepts any
    w,h >= 1 and matches the oracle on all of them (partial tiles, pad columns)."""
    h, d = bumpy(W, H)
    c = oracle.derive_consts(float(W), W)
    t, f, v = new_state(h, d)
    oracle.step(t, f, v, c, 37)
    with make_sim(tws, W, H, backend, k) as sim:
        sim.upload(tws.FIELD_TERRAIN, h)
        sim.upload(tws.FIELD_WATER, d)
        sim.step(37)
        assert_state_equal(sim, tws, t, f, v, f"{name} {W}x{H}")


@pytest.mark.parametrize("name,backend,k", BACKENDS)
def test_extensions_closed_boundary_rain_evaporation(tws, oracle, name, backend, k):
    W, H = 200, 120
    h, d = bumpy(W, H, seed=9)
    c = oracle.derive_consts(float(W), W)
    dt = float(np.float32(1.0) / np.float32(60.0))
    rain, evap = 0.8, 0.3
    rs, es = float(np.float32(dt * np.float32(rain))), float(np.float32(dt * np.float32(evap)))
    t, f, v = new_state(h, d)
    oracle.step(t, f, v, c, 30, boundary=1, rain_step=rs, evap_step=es)
    with make_sim(tws, W, H, backend, k, boundary=tws.BOUNDARY_CLOSED, rain_rate=rain, evaporation_rate=evap) as sim:
        sim.upload(tws.FIELD_TERRAIN, h)
        sim.upload(tws.FIELD_WATER, d)
        sim.step(30)
        assert_state_equal(sim, tws, t, f, v, name)
    # closed boundary without sources conserves volume
    t, f, v = new_state(h, d)
    oracle.step(t, f, v, c, 60, boundary=1)
    with make_sim(tws, W, H, backend, k, boundary=tws.BOUNDARY_CLOSED) as sim:
        sim.upload(tws.FIELD_TERRAIN, h)
        sim.upload(tws.FIELD_WATER, d)
        v0 = sim.total_volume()
        sim.step(60)
        assert_state_equal(sim, tws, t, f, v, name + " closed")
        assert abs(sim.total_volume() - v0) / v0 < 1e-6


def test_upload_readback_roundtrip_all_fields(tws):
    W, H = 70, 33
    rng = np.random.default_rng(1)
    with make_sim(tws, W, H, tws.BACKEND_FUSED, 1) as sim:
        h = rng.random((H, W)).astype(np.float32)
        d = rng.random((H, W)).astype(np.float32)
        F = rng.random((H, W, 4)).astype(np.float32)
        sim.upload(tws.FIELD_TERRAIN, h); sim.upload(tws.FIELD_WATER, d); sim.upload(tws.FIELD_FLUX, F)
        assert np.array_equal(sim.readback(tws.FIELD_TERRAIN), h)
        assert np.array_equal(sim.readback(tws.FIELD_WATER), d)
        assert np.array_equal(sim.readback(tws.FIELD_FLUX), F)
        info = sim.readback(tws.FIELD_TERRAIN_INFO)           # m_terrainData layout, Terrain.cpp:216-219
        assert np.array_equal(info[..., 0], h) and np.array_equal(info[..., 3], d)
        assert np.all(info[..., 1] == np.float32(0.3)) and np.all(info[..., 2] == np.float32(0.3))
        info2 = info.copy(); info2[..., 0] += 1; info2[..., 3] *= 2
        sim.upload(tws.FIELD_TERRAIN_INFO, info2)
        assert np.array_equal(sim.readback(tws.FIELD_TERRAIN), info2[..., 0])
        assert np.array_equal(sim.readback(tws.FIELD_WATER), info2[..., 3])
        assert sim.total_volume() == pytest.approx(float(info2[..., 3].sum(dtype=np.float64)), rel=1e-12)
        with pytest.raises(ValueError):
            sim.upload(tws.FIELD_WATER, d[:-1])
        with pytest.raises(tws.TwsError) as e:
            sim._check(sim._lib.tws_upload(sim._sim, tws.FIELD_WATER, d.ctypes.data, d.nbytes - 4))
        assert e.value.status == tws._abi.TWS_ERR_INVALID
        with pytest.raises(tws.TwsError) as e:
            sim._check(sim._lib.tws_upload(sim._sim, tws.FIELD_VELOCITY, d.ctypes.data, d.nbytes))
        assert e.value.status == tws._abi.TWS_ERR_INVALID


@pytest.mark.parametrize("W,H", [(1024, 1024), (300, 200), (64, 512)])
def test_gpu_scene_generator_matches_oracle(tws, oracle_omp, W, H):
    """a9: Terrain::CreateHeightmapFromNoiseAndResetSim on the GPU == the (reference-pinned) oracle."""
    s = oracle_omp.create_scene(W, H)
    with make_sim(tws, W, H, tws.BACKEND_FUSED, 1) as sim:
        sim.upload(tws.FIELD_FLUX, np.ones((H, W, 4), np.float32))
        sim.CreateHeightmapFromNoiseAndResetSim()
        info = sim.readback(tws.FIELD_TERRAIN_INFO)
        assert np.array_equal(bits(info), bits(s))
        assert not sim.readback(tws.FIELD_FLUX).any()          # flux reset to 0, Terrain.cpp:230-234
        sim.CreateHeightmapFromNoiseAndResetSim(seed=1234, heightScale=120.0, lowOctave=1, highOctave=6, persistence=0.5)
        s2 = oracle_omp.create_scene(W, H, seed=1234, height_scale=120.0, lo=1, hi=6, persistence=0.5)
        assert np.array_equal(bits(sim.readback(tws.FIELD_TERRAIN_INFO)), bits(s2))


@pytest.mark.parametrize("W,H,tile", [(300, 600, 200), (256, 512, 256), (64, 100, 37)])
def test_gpu_tiled_scene_is_the_small_scene_repeated(tws, oracle_omp, W, H, tile):
    """Extension used by the weak-scaling bench: tws_reset_reference_scene_tiled = the W x tile scene of the
    reference generator repeated every `tile` rows (the last copy cut off), bit for bit."""
    one = oracle_omp.create_scene(W, tile)
    want = np.concatenate([one] * ((H + tile - 1) // tile))[:H]
    with make_sim(tws, W, H, 5, 2) as sim:
        sim.CreateHeightmapFromNoiseAndResetSim(tileHeight=tile)
        got = sim.readback(tws.FIELD_TERRAIN_INFO)
        assert np.array_equal(bits(got), bits(want))
        assert np.array_equal(sim.readback(tws.FIELD_FLUX), np.zeros((H, W, 4), np.float32))
        with pytest.raises(tws.TwsError):
            sim.CreateHeightmapFromNoiseAndResetSim(tileHeight=H + 1)


def test_gpu_scene_generator_8192_golden(tws, oracle_omp):
    with make_sim(tws, 8192, 8192, tws.BACKEND_FUSED, 1) as sim:
        sim.CreateHeightmapFromNoiseAndResetSim()
        assert f"{oracle_omp.fnv1a64(sim.readback(tws.FIELD_TERRAIN)):016x}" == GOLD["scene8192"]["h"]
        d = sim.readback(tws.FIELD_WATER)
        assert f"{oracle_omp.fnv1a64(d):016x}" == GOLD["scene8192"]["d"]
        assert sim.total_volume() == pytest.approx(GOLD["scene8192"]["d_sum"], rel=1e-12)


@pytest.mark.parametrize("cx,cy,inten,size", [(512.0, 512.0, 100.0 / 60.0, 32.0), (0.0, 0.0, 1.0, 32.0), (99.75, 3.25, 0.5, 7.5),
                                             (199.9, 119.9, 2.0, 200.0), (-3.0, 50.0, 1.0, 32.0), (50.0, 500.0, 1.0, 32.0), (10.5, 10.5, -0.25, 32.0)])
def test_inject_brush_matches_whole_grid_pass(tws, oracle, cx, cy, inten, size):
    """a8: sparse-footprint brush == waterBrush.comp over the whole grid (bitwise)."""
    W, H = (1024, 1024) if cx > 300 else (200, 120)
    h, d = bumpy(W, H, seed=2)
    t, f, v = new_state(h, d)
    oracle.brush(t, cx, cy, inten, size)
    with make_sim(tws, W, H, tws.BACKEND_FUSED, 1) as sim:
        sim.upload(tws.FIELD_TERRAIN, h); sim.upload(tws.FIELD_WATER, d)
        sim.inject_brush(cx, cy, inten, size)
        assert np.array_equal(bits(sim.readback(tws.FIELD_WATER)), bits(np.ascontiguousarray(t[..., 3])))


@pytest.mark.parametrize("backend,k", [(2, 1), (3, 2), (0, 1), (6, 1)], ids=["fused", "fused_tb2", "auto", "resident"])
@pytest.mark.parametrize("W,H", [(300, 200), (1024, 1024)])
def test_brush_is_folded_into_the_next_single_step_launch(tws, oracle_omp, W, H, backend, k):
    """On a whole grid that runs the tile kernel (or the resident kernel) a brush is not launched on its own: the next direct
    step launch applies it while loading the depth (the reference's frame — ApplyRadialWaterBrush, one step, GenMipMaps — is two launches).  Same bits as
    brush-then-step on the oracle for brushes on tile seams, tile corners, the grid's corners and edges, partly and wholly off
    the grid; anything that looks at the state before the step (a second brush, a readback, the renderer hand-off, a
    multi-step call) materialises the brush first."""
    h, d = bumpy(W, H, seed=21)
    c = oracle_omp.derive_consts(float(W), W)
    t, f, v = new_state(h, d)
    spots = [(W * 0.5, H * 0.5, 1.5, 32.0), (128.0, 28.0, 0.7, 32.0), (127.5, 27.5, 0.7, 20.0), (255.9, 55.5, 1.0, 90.0), (0.0, 0.0, 2.0, 32.0),
             (W - 1.0, H - 1.0, 2.0, 32.0), (-3.0, 5.0, 1.0, 32.0), (W + 2.5, H * 0.5, 1.0, 40.0), (-500.0, -500.0, 1.0, 32.0), (64.0, H - 0.5, -0.25, 32.0)]
    with make_sim(tws, W, H, backend, k) as sim:
        sim.upload(tws.FIELD_TERRAIN, h)
        sim.upload(tws.FIELD_WATER, d)
        for cx, cy, inten, size in spots:
            sim.inject_brush(cx, cy, inten, size)
            l0 = sim.kernel_launches()
            sim.step(1)
            assert sim.kernel_launches() - l0 == 1, "the brush took a launch of its own"
            oracle_omp.brush(t, cx, cy, inten, size)
            oracle_omp.step(t, f, v, c, 1)
        assert_state_equal(sim, tws, t, f, v, "brush folded into the step")
        # two brushes before one step: the first is materialised by the second, the second rides along
        sim.inject_brush(100.0, 100.0, 1.0, 32.0); sim.inject_brush(103.0, 98.0, 0.5, 50.0)
        sim.step(1)
        oracle_omp.brush(t, 100.0, 100.0, 1.0, 32.0); oracle_omp.brush(t, 103.0, 98.0, 0.5, 50.0)
        oracle_omp.step(t, f, v, c, 1)
        assert_state_equal(sim, tws, t, f, v, "two brushes, one step")
        # a brush that is looked at before any step
        sim.inject_brush(40.0, 60.0, 3.0, 32.0)
        oracle_omp.brush(t, 40.0, 60.0, 3.0, 32.0)
        assert np.array_equal(bits(sim.readback(tws.FIELD_WATER)), bits(np.ascontiguousarray(t[..., 3])))
        # ... before a multi-step call (a captured batch or the resident kernel), and before the renderer hand-off
        for n in (2, 5):
            sim.inject_brush(200.0, 30.0, 1.0, 32.0)
            sim.step(n)
            oracle_omp.brush(t, 200.0, 30.0, 1.0, 32.0)
            oracle_omp.step(t, f, v, c, n)
        sim.inject_brush(10.0, 10.0, 1.0, 32.0)
        oracle_omp.brush(t, 10.0, 10.0, 1.0, 32.0)
        level0 = sim.publish_mips()[0]
        assert np.array_equal(bits(np.ascontiguousarray(level0[..., 3])), bits(np.ascontiguousarray(t[..., 3])))
        assert_state_equal(sim, tws, t, f, v, "after flushes")


def test_reference_interface_frame_loop(tws, oracle):
    """Scene::Update order (Scene.cpp:356-363): brush with strength dt*100 at the camera XZ, then
    PerformSimulationStep(dt) with the accumulator / 10-step clamp of Terrain.cpp:240-247."""
    W = 256
    s = oracle.create_scene(W)
    t, f, v = s.copy(), np.zeros((W, W, 4), np.float32), np.zeros((W, W, 2), np.float16)
    with tws.Terrain(W, gridWorldSize=1024.0, backend=tws.BACKEND_FUSED_TB, temporal_block=4) as sim:
        sim.CreateHeightmapFromNoiseAndResetSim()
        sim.SetFlowDamping(0.96); sim.SetFlowAcceleration(16.0); sim.SetSimulationStepsPerSecond(120.0)
        c = oracle.derive_consts(1024.0, W, 120.0, 0.96, 16.0)
        assert np.array_equal(np.float32(sim.step_constants()), c)
        step_len = float(np.float32(1.0) / np.float32(120.0))
        acc, total = 0.0, 0
        for frame_dt in [0.016, 0.0171, 0.004, 0.0, 0.25, 0.033, 0.0009, 0.0167]:
            cam = (1536.0 + 37.5, 512.0 - 20.25)               # wraps with Fraction (Terrain.cpp:152-155)
            sim.ApplyRadialWaterBrush(cam, frame_dt * 100.0)
            cx, cy = oracle.brush_center(cam[0], cam[1], 1024.0, W)
            oracle.brush(t, cx, cy, float(np.float32(frame_dt * 100.0)), 32.0)
            n = sim.PerformSimulationStep(frame_dt)
            n_ref, acc = oracle.advance(acc, step_len, frame_dt)
            assert n == n_ref
            oracle.step(t, f, v, c, n_ref)
            total += n
        assert total >= 20 and sim.elapsed_ms() > 0
        assert_state_equal(sim, tws, t, f, v, "frame loop")


def test_boundary_outflow_ledger_matches_oracle(tws, oracle):
    """Mass ledger on the open boundary: volume change per step == -(outflow through the edge) * areaInv,
    with the outflow read from the flux field exactly as the oracle sums it."""
    W, H = 96, 80
    h, d = dam_break(W, H, rim=False)
    c = oracle.derive_consts(float(W), W)
    t, f, v = new_state(h, d)
    with make_sim(tws, W, H, tws.BACKEND_FUSED, 1) as sim:
        sim.upload(tws.FIELD_TERRAIN, h); sim.upload(tws.FIELD_WATER, d)
        vol = sim.total_volume()
        lost = 0.0
        for _ in range(40):
            sim.step(1)
            oracle.step(t, f, v, c, 1)
            want = (f[:, -1, 0].sum(dtype=np.float64) + f[:, 0, 1].sum(dtype=np.float64) + f[-1, :, 2].sum(dtype=np.float64) + f[0, :, 3].sum(dtype=np.float64))
            got = sim.boundary_outflow()
            assert got == pytest.approx(float(want), rel=1e-12, abs=1e-12)
            lost += got * float(c[2])
        assert lost > 0
        assert abs((vol - sim.total_volume()) - lost) / vol < 1e-6


def _oracle_outflow_per_step(oracle, t, f, v, c, n, rain=0.0, evap=0.0):
    """n oracle steps; returns the fp64 sum over the steps of (flux pointing out of the grid) * areaInv, summed the way
    tests/test_oracle.py::test_open_volume_change_equals_boundary_outflow does (flowApply.comp:38-41 with exterior = 0)."""
    lost = 0.0
    for _ in range(n):
        oracle.flow_update(t, f, c)
        out = (f[:, -1, 0].sum(dtype=np.float64) + f[:, 0, 1].sum(dtype=np.float64) + f[-1, :, 2].sum(dtype=np.float64) + f[0, :, 3].sum(dtype=np.float64))
        oracle.flow_apply(t, f, v, c, rain, evap)
        lost += float(out) * float(c[2])
    return lost


@pytest.mark.parametrize("name,backend,k", BACKENDS)
@pytest.mark.parametrize("W,H", [(96, 80), (250, 190), (130, 29), (5, 3)])
def test_boundary_outflow_accumulated_in_kernel_matches_oracle(tws, oracle, W, H, name, backend, k):
    """The in-kernel fp64 ledger (every sub-step, k steps per launch included) against the oracle's per-step sum; the
    state stays bit-identical, the drained volume closes the open-domain balance."""
    h, d = dam_break(W, H, rim=False)
    d[:, -max(1, W // 5):] = 7.0          # water at the +X edge too
    d[: max(1, H // 6)] += 3.0            # ... and along row 0
    c = oracle.derive_consts(float(W), W)
    t, f, v = new_state(h, d)
    with make_sim(tws, W, H, backend, k) as sim:
        sim.upload(tws.FIELD_TERRAIN, h); sim.upload(tws.FIELD_WATER, d)
        vol0 = sim.total_volume()
        assert sim.boundary_outflow_accumulated() == 0.0
        want = 0.0
        for n in (1, 7, 4 * k, 26):       # batches that are and are not multiples of k
            sim.step(n)
            want += _oracle_outflow_per_step(oracle, t, f, v, c, n)
            got = sim.boundary_outflow_accumulated()
            assert got == pytest.approx(want, rel=1e-12, abs=1e-300), f"after a batch of {n}"
        assert want > 0
        assert_state_equal(sim, tws, t, f, v, name)
        assert abs((vol0 - sim.total_volume()) - got) / vol0 < 1e-6          # volume lost == volume counted out
        sim.boundary_outflow_reset()
        assert sim.boundary_outflow_accumulated() == 0.0
        sim.step(3)
        assert sim.boundary_outflow_accumulated() == pytest.approx(_oracle_outflow_per_step(oracle, t, f, v, c, 3), rel=1e-12)
    # closed boundary: nothing can leave
    with make_sim(tws, W, H, backend, k, boundary=tws.BOUNDARY_CLOSED) as sim:
        sim.upload(tws.FIELD_TERRAIN, h); sim.upload(tws.FIELD_WATER, d)
        sim.step(9)
        assert sim.boundary_outflow_accumulated() == 0.0


@pytest.mark.parametrize("name,backend,k", BACKENDS)
def test_ledger_with_rain_and_evaporation_closes(tws, oracle, name, backend, k):
    """SURVEY 8d config 5 in small: V(t) = V0 + sources - boundary outflow, BOTH sides booked inside the step kernels in
    fp64: the outflow through the open edge, and what rain / evaporation really changed in fp32 (not rate x time x area:
    d + rain_step - evap_step rounds alike for every cell of a binade, and evaporation is clamped at dry cells).  Each is
    compared with the oracle's own per-step sum; together they close the volume balance."""
    W, H = 250, 190
    h, d = dam_break(W, H, rim=False)
    c = oracle.derive_consts(float(W), W)
    rain, evap = 0.6, 0.9                 # evaporation > rain: the clamp at dry cells is exercised
    dt = float(np.float32(1.0) / np.float32(60.0))
    rs, es = float(np.float32(dt * rain)), float(np.float32(dt * evap))
    t, f, v = new_state(h, d)
    with make_sim(tws, W, H, backend, k, rain_rate=rain, evaporation_rate=evap) as sim:
        sim.upload(tws.FIELD_TERRAIN, h); sim.upload(tws.FIELD_WATER, d)
        vol0 = sim.total_volume()
        assert sim.source_accumulated() == 0.0
        want_out = want_src = 0.0
        for n in (1, 6, 4 * k, 37):
            sim.step(n)
            for _ in range(n):
                oracle.flow_update(t, f, c)
                want_out += float(f[:, -1, 0].sum(dtype=np.float64) + f[:, 0, 1].sum(dtype=np.float64) + f[-1, :, 2].sum(dtype=np.float64)
                                  + f[0, :, 3].sum(dtype=np.float64)) * float(c[2])
                t0 = t.copy(); v0 = v.copy()
                oracle.flow_apply(t0, f, v0, c)                       # the same pass without the source terms
                oracle.flow_apply(t, f, v, c, rs, es)
                want_src += float((t[..., 3].astype(np.float64) - t0[..., 3].astype(np.float64)).sum())
            assert sim.boundary_outflow_accumulated() == pytest.approx(want_out, rel=1e-12)
            assert sim.source_accumulated() == pytest.approx(want_src, rel=1e-11, abs=1e-9)
        assert_state_equal(sim, tws, t, f, v, name)
        assert want_src < 0 and abs(want_src - W * H * 50 * (rs - es)) > 1e-3 * abs(want_src)   # NOT rate x time x area
        vol = sim.total_volume()
        assert abs(vol - (vol0 + sim.source_accumulated() - sim.boundary_outflow_accumulated())) / vol0 < 1e-6
        sim.boundary_outflow_reset()
        assert sim.source_accumulated() == 0.0 and sim.boundary_outflow_accumulated() == 0.0
    with make_sim(tws, W, H, backend, k) as sim:                      # no sources: nothing is booked
        sim.upload(tws.FIELD_TERRAIN, h); sim.upload(tws.FIELD_WATER, d)
        sim.step(5)
        assert sim.source_accumulated() == 0.0


def test_elapsed_ms_nowait_reads_the_previous_batch_without_blocking(tws):
    """gl::TimerQuery semantics (TimerQuery.cpp:29-72, read one frame late at Scene.cpp:337-342): double-buffered event
    pairs; the non-blocking read returns the newest FINISHED batch and never waits for the GPU."""
    import time
    with make_sim(tws, 4096, 4096, tws.BACKEND_BAND_TB, 4) as sim:
        sim.CreateHeightmapFromNoiseAndResetSim()
        assert sim.elapsed_ms_nowait() is None               # nothing timed yet
        sim.step(4); sim.sync()
        ms0, idx0 = sim.elapsed_ms_nowait()
        assert idx0 == 0 and ms0 > 0 and ms0 == pytest.approx(sim.elapsed_ms(), rel=1e-6)
        sim.step(4); sim.sync()
        sim.step(400)                                        # ~50 ms of GPU work queued: the call below must not wait for it
        t0 = time.perf_counter()
        got = sim.elapsed_ms_nowait()
        waited = time.perf_counter() - t0
        assert got is not None
        ms, idx = got
        assert idx in (1, 2) and ms > 0
        assert waited < 0.02 or idx == 2                     # it returned while batch 2 was still running (or that one was already done)
        sim.sync()
        ms2, idx2 = sim.elapsed_ms_nowait()
        assert idx2 == 2 and ms2 == pytest.approx(sim.elapsed_ms(), rel=1e-6) and ms2 > 10 * ms0


def test_advice_r1_edge_cases(tws, oracle):
    """Brush centres beyond INT_MAX add nothing (the reference's whole-grid pass would add 0 everywhere) instead of failing;
    a frame time whose step count overflows the reference's 32-bit counter is refused before the accumulator moves."""
    with make_sim(tws, 64, 64, tws.BACKEND_FUSED, 1) as sim:
        d = np.ones((64, 64), np.float32)
        sim.upload(tws.FIELD_WATER, d)
        for cx, cy in ((3.0e9, 10.0), (-3.0e9, 10.0), (10.0, 3.0e9), (10.0, -3.0e9), (3.0e38, -3.0e38)):
            sim.inject_brush(cx, cy, 2.0, 32.0)
        assert np.array_equal(sim.readback(tws.FIELD_WATER), d)
        with pytest.raises(tws.TwsError) as e:
            sim.PerformSimulationStep(1.0e9)
        assert e.value.status == tws._abi.TWS_ERR_INVALID
        assert sim.PerformSimulationStep(1.0) == 10          # the accumulator was not poisoned by the refused frame


@pytest.mark.parametrize("name,backend,k", [BACKENDS[0], BACKENDS[4], BACKENDS[8], BACKENDS[12]])
def test_cuda_path_against_the_reference_shaders_directly(tws, name, backend, k):
    """No oracle in between: the CUDA library against the reference's OWN flowUpdate / flowApply / waterBrush shaders
    (oracle/_ref/libtws_ref_step.so, compiled from /root/reference in the build container and shipped prebuilt) — BASELINE
    config 1 for 300 steps and the 1024^2 reference scene with the brush before every step for 60 steps, bit for bit."""
    from oracle.oracle_py import RefStep
    try:
        ref = RefStep()
    except FileNotFoundError:
        pytest.skip("oracle/_ref/libtws_ref_step.so not present on this box")
    # config 1, open boundary (the reference's native behaviour)
    h, d = dam_break(256, rim=False)
    t, f, v = new_state(h, d)
    with make_sim(tws, 256, 256, backend, k) as sim:
        c = np.float32(sim.step_constants())
        sim.upload(tws.FIELD_TERRAIN, h); sim.upload(tws.FIELD_WATER, d)
        sim.step(300)
        ref.step(t, f, v, c, 300)
        assert_state_equal(sim, tws, t, f, v, name + " vs reference shaders, config 1")
    # config 2: reference scene + brush (Scene.cpp:356-363)
    with make_sim(tws, 1024, 1024, backend, k) as sim:
        sim.CreateHeightmapFromNoiseAndResetSim()
        c = np.float32(sim.step_constants())
        t = sim.readback(tws.FIELD_TERRAIN_INFO)
        f = np.zeros((1024, 1024, 4), np.float32); v = np.zeros((1024, 1024, 2), np.float16)
        inten = np.float32(100.0 / 60.0)
        for _ in range(60):
            sim.inject_brush(512.0, 512.0, float(inten), 32.0)
            sim.step(1)
            ref.brush(t, 512.0, 512.0, inten, 32.0)
            ref.step(t, f, v, c, 1)
        assert_state_equal(sim, tws, t, f, v, name + " vs reference shaders, config 2")


def test_parameter_validation_on_device(tws):
    with make_sim(tws, 64, 64, tws.BACKEND_FUSED, 1) as sim:
        for call in (lambda: sim.SetSimulationStepsPerSecond(0.0), lambda: sim.SetFlowDamping(-1.0), lambda: sim.SetFlowAcceleration(float("nan")),
                     lambda: sim.inject_brush(1.0, 1.0, 1.0, 0.0), lambda: sim.step(-1), lambda: sim.PerformSimulationStep(-0.1)):
            with pytest.raises(tws.TwsError) as e:
                call()
            assert e.value.status == tws._abi.TWS_ERR_INVALID
        with pytest.raises(tws.TwsError) as e:
            sim.elapsed_ms()
        assert e.value.status == tws._abi.TWS_ERR_STATE
    with pytest.raises(tws.TwsError):
        tws.Terrain(64, rows=(8, 40), backend=tws.BACKEND_UNFUSED)
    with pytest.raises(tws.TwsError):
        tws.Terrain(64, backend=tws.BACKEND_FUSED_TB, temporal_block=5)


def test_cxx_host_wrapper_runs(tws, tmp_path):
    exe = tmp_path / "cxx_host_check"
    cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    cmd = [cxx, "-std=c++17", "-O1", f"-I{ROOT / 'include'}", str(ROOT / "tests" / "cxx_host_check.cpp"), "-o", str(exe),
           f"-L{ROOT / 'terrainwatersim_b200'}", "-ltws", f"-Wl,-rpath,{ROOT / 'terrainwatersim_b200'}"]
    assert subprocess.run(cmd, capture_output=True, text=True).returncode == 0
    out = subprocess.run([str(exe), "run"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "steps=2" in out.stdout          # 0.05 s at 60 steps/s -> (uint)(0.05/0.016667) = 2 (Terrain.cpp:243)
    assert "mips=8 chain=ok last=1x1@" in out.stdout and "gl=" in out.stdout      # 128 -> 8 levels (Texture.cpp:28-41)


# ---- full-size properties (BASELINE config 3 grid: 8192 x 8192) ------------------------------------
def test_8192_backends_bitwise_equal_and_oracle_few_steps(tws, oracle_omp):
    """At 8192^2 the oracle only runs a few steps (8); beyond that the GPU variants are checked
    against each other bitwise (24 steps) — tiling, temporal blocking and the unfused baseline
    must all agree exactly."""
    W = 8192
    ref = {}
    for name, backend, k in BACKENDS:
        if backend == tws.BACKEND_RESIDENT:
            continue                                   # 67 M cells do not fit in shared memory
        with make_sim(tws, W, W, backend, k) as sim:
            sim.CreateHeightmapFromNoiseAndResetSim()
            sim.inject_brush(4096.0, 4100.5, 25.0, 900.0)
            sim.step(8)
            d8 = sim.readback(tws.FIELD_WATER)
            sim.step(16)
            cur = {"d8": d8, "d": sim.readback(tws.FIELD_WATER), "v": sim.readback(tws.FIELD_VELOCITY).view(np.uint16), "vol": sim.total_volume()}
            Fsum = sim.readback(tws.FIELD_FLUX)
            cur["Fh"] = oracle_omp.fnv1a64(Fsum)
            del Fsum
        if not ref:
            ref = cur
            s = oracle_omp.create_scene(W)
            t, f, v = s, np.zeros((W, W, 4), np.float32), np.zeros((W, W, 2), np.float16)
            oracle_omp.brush(t, 4096.0, 4100.5, 25.0, 900.0)
            oracle_omp.step(t, f, v, oracle_omp.derive_consts(float(W), W), 8)
            assert np.array_equal(bits(d8), bits(np.ascontiguousarray(t[..., 3]))), "8192^2: GPU differs from the oracle after 8 steps"
            del s, t, f, v
        else:
            for key in ("d8", "d", "v"):
                assert np.array_equal(bits(cur[key]) if cur[key].dtype != np.uint16 else cur[key], bits(ref[key]) if ref[key].dtype != np.uint16 else ref[key]), f"{name}: {key} differs from unfused at 8192^2"
            assert cur["Fh"] == ref["Fh"] and cur["vol"] == ref["vol"]


def test_8192_closed_domain_conserves_volume(tws):
    with make_sim(tws, 8192, 8192, tws.BACKEND_FUSED_TB, 4, boundary=tws.BOUNDARY_CLOSED) as sim:
        sim.CreateHeightmapFromNoiseAndResetSim()
        v0 = sim.total_volume()
        sim.step(200)
        v1 = sim.total_volume()
        d = sim.readback(tws.FIELD_WATER)
        assert d.min() >= 0 and np.isfinite(d).all()
        assert abs(v1 - v0) / v0 < 1e-6


@pytest.mark.parametrize("name,backend,k", [BACKENDS[0], BACKENDS[1], BACKENDS[2], BACKENDS[5], BACKENDS[8], BACKENDS[11], BACKENDS[13]])
@pytest.mark.parametrize("W,H", [(2048, 3000), (250, 190), (37, 5), (1, 1), (4100, 1100)])
def test_step_host_band_pipeline_matches_oracle(tws, oracle_omp, W, H, name, backend, k):
    """tws_step_host (band-pipelined upload / step / readback, several bands at 2048x3000 and
    4100x1100) against the oracle driven the same way: the host owns the water layer, edits it
    between steps, and gets depth + flow vector back each step."""
    if backend == tws.BACKEND_RESIDENT and W * H > 1000 * 1000:
        pytest.skip("does not fit on chip")
    h, d = bumpy(W, H, seed=11)
    c = oracle_omp.derive_consts(float(W), W)
    t, f, v = new_state(h, d)
    rng = np.random.default_rng(3)
    with make_sim(tws, W, H, backend, k) as sim:
        sim.upload(tws.FIELD_TERRAIN, h)
        water = d.copy()
        vel = np.empty((H, W, 2), np.float16)
        for step in range(4):
            if step == 2:                               # host-side edit of the water layer between steps
                edit = (rng.random((H, W)) > 0.9).astype(np.float32)
                water += edit
                t[..., 3] += edit
            if step == 3:                               # NULL input: keep the resident water; separate output buffer
                out = np.empty_like(water)
                sim.step_host(None, out, vel)
                water = out
            else:
                sim.step_host(water, water, vel)        # in place: water_out aliases water_in
            oracle_omp.step(t, f, v, c, 1)
            assert np.array_equal(bits(water), bits(np.ascontiguousarray(t[..., 3]))), f"{name} step {step}: depth differs"
            assert np.array_equal(bits(vel), bits(v)), f"{name} step {step}: velocity differs"
        assert_state_equal(sim, tws, t, f, v, f"{name} resident state after step_host")
        sim.step(3)                                     # and the resident state carries on in ordinary steps
        oracle_omp.step(t, f, v, c, 3)
        assert_state_equal(sim, tws, t, f, v, f"{name} step after step_host")
        sim.step_host(None, None, None)                 # outputs are optional
        oracle_omp.step(t, f, v, c, 1)
        assert_state_equal(sim, tws, t, f, v, f"{name} step_host without outputs")


@pytest.mark.parametrize("W,H", [(1024, 1024), (300, 200), (37, 5), (5, 37), (257, 64), (1, 1), (128, 128), (1024, 128), (256, 640), (2048, 1152)])
def test_terrain_info_mip_chain_matches_oracle(tws, oracle, W, H):
    """Renderer hand-off: TerrainInfo level 0 = (terrain, 0.3, 0.3, water) and the mip chain the
    reference regenerates after every stepped frame (Terrain.cpp:272-276), against the pinned box filter of
    oracle_py.mip_chain — every level, bit for bit, after a few steps so that the water channel is live."""
    from oracle.oracle_py import mip_chain
    h, d = bumpy(W, H, seed=11)
    t, f, v = new_state(h, d)
    c = oracle.derive_consts(float(W), W)
    oracle.step(t, f, v, c, 3)
    with make_sim(tws, W, H, 5, 3) as sim:
        sim.upload(tws.FIELD_TERRAIN, h); sim.upload(tws.FIELD_WATER, d)
        sim.step(3)
        got = sim.publish_mips()
        want = mip_chain(t)
        assert len(got) == len(want)
        for l, (g_, w_) in enumerate(zip(got, want)):
            assert g_.shape == w_.shape, f"level {l}"
            assert np.array_equal(bits(g_), bits(w_)), f"level {l} differs"
        # again (the single-pass kernel re-arms its ticket counter itself), after another step
        sim.step(2)
        oracle.step(t, f, v, c, 2)
        for l, (g_, w_) in enumerate(zip(sim.publish_mips(), mip_chain(t))):
            assert np.array_equal(bits(g_), bits(w_)), f"second publish: level {l} differs"
        # level 0 alone (tws_publish_packed) leaves the rest of the chain unpublished
        sim.step(1)
        sim.publish_packed()
        a = np.empty((max(1, H >> 1), max(1, W >> 1), 4), np.float32)
        if len(want) > 1:
            assert sim._lib.tws_readback_mip(sim._sim, 1, a.ctypes.data, a.nbytes) == tws._abi.TWS_ERR_STATE


def test_gl_interop_without_a_gl_context_fails_cleanly(tws):
    """The GL entry points are real (cudaGraphicsGLRegisterImage + mapped-array copies), but the GPU box has no
    GL context: registering must come back as an error status with a message, never crash; publish before a
    successful register is a state error; unregister is idempotent."""
    with make_sim(tws, 64, 64, 2, 1) as sim:
        with pytest.raises(tws.TwsError) as e:
            sim.gl_register(1, 2)
        assert e.value.status in (tws._abi.TWS_ERR_CUDA, tws._abi.TWS_ERR_UNSUPPORTED) and str(e.value)
        with pytest.raises(tws.TwsError) as e:
            sim.gl_publish()
        assert e.value.status == tws._abi.TWS_ERR_STATE
        sim.gl_unregister(); sim.gl_unregister()
        sim.step(2)                                           # the sim is still usable
        assert np.isfinite(sim.total_volume())


@pytest.mark.parametrize("W,H", [(1024, 1024), (1000, 1030), (1021, 700), (512, 512), (384, 900), (256, 256), (130, 29), (64, 12), (65, 13)])
@pytest.mark.parametrize("boundary", ["open", "closed"])
def test_resident_backend_frames_of_any_length_match_the_oracle(tws, oracle_omp, W, H, boundary):
    """TWS_BACKEND_RESIDENT: every tws_step(n) is ONE launch that keeps the grid in shared memory for all n steps and
    exchanges block rims through tagged words in L2.  Frames of 1, 2, 3, 10 and 25 steps back to back (the step tags run on
    across launches), a brush between frames, every block shape (256x28, 128x28, 128x12, 64x12), ragged sizes, both
    boundary modes, rain + evaporation with both ledgers."""
    closed = boundary == "closed"
    h, d = bumpy(W, H, seed=11)
    rain, evap = 0.8, 0.3
    dt = float(np.float32(1.0) / np.float32(60.0))
    rs, es = float(np.float32(dt * np.float32(rain))), float(np.float32(dt * np.float32(evap)))
    c = oracle_omp.derive_consts(float(W), W)
    t, f, v = new_state(h, d)
    with make_sim(tws, W, H, tws.BACKEND_RESIDENT, 1, boundary=tws.BOUNDARY_CLOSED if closed else tws.BOUNDARY_REFERENCE_OPEN,
                  rain_rate=rain, evaporation_rate=evap) as sim:
        assert sim.backend_in_use()[0] == tws.BACKEND_RESIDENT
        sim.upload(tws.FIELD_TERRAIN, h)
        sim.upload(tws.FIELD_WATER, d)
        for n in (1, 2, 3, 10, 25):
            sim.inject_brush(W * 0.5, H * 0.25, 0.75, 40.0)
            oracle_omp.brush(t, W * 0.5, H * 0.25, 0.75, 40.0)
            launches = sim.kernel_launches()
            sim.step(n)
            assert sim.kernel_launches() - launches == 1
            oracle_omp.step(t, f, v, c, n, boundary=1 if closed else 0, rain_step=rs, evap_step=es)
            assert_state_equal(sim, tws, t, f, v, f"resident {W}x{H} {boundary} after a frame of {n}")


def test_resident_backend_long_run_equals_the_tile_kernel_at_1024(tws, oracle_omp):
    """The reference's operating point for a long time: 1024^2 default scene, 300 frames of 10 steps with the brush before every
    frame (3000 steps, 2700 rim exchanges per block).  The resident kernel and the tile kernel must agree on every bit of depth,
    flux and flow vector at the end, and both with the oracle after the first 3 frames."""
    states = {}
    for name, backend, k in (("resident", tws.BACKEND_RESIDENT, 1), ("tile", tws.BACKEND_FUSED_TB, 2)):
        with make_sim(tws, 1024, 1024, backend, k) as sim:
            sim.CreateHeightmapFromNoiseAndResetSim()
            early = None
            for frame in range(300):
                sim.inject_brush(512.0, 512.0, float(np.float32(100.0 / 60.0)), 32.0)
                sim.step(10)
                if frame == 2:
                    early = (sim.readback(tws.FIELD_WATER), sim.readback(tws.FIELD_FLUX), sim.readback(tws.FIELD_VELOCITY))
            states[name] = (early, sim.readback(tws.FIELD_WATER), sim.readback(tws.FIELD_FLUX), sim.readback(tws.FIELD_VELOCITY),
                            sim.boundary_outflow_accumulated())
    t = oracle_omp.create_scene(1024)
    f, v = np.zeros((1024, 1024, 4), np.float32), np.zeros((1024, 1024, 2), np.float16)
    c = oracle_omp.derive_consts(1024.0, 1024)
    for _ in range(3):
        oracle_omp.brush(t, 512.0, 512.0, float(np.float32(100.0 / 60.0)), 32.0)
        oracle_omp.step(t, f, v, c, 10)
    for name in states:
        e = states[name][0]
        assert np.array_equal(bits(e[0]), bits(np.ascontiguousarray(t[..., 3]))) and np.array_equal(bits(e[1]), bits(f)) and \
            np.array_equal(bits(e[2]), bits(v)), f"{name} differs from the oracle after 30 steps"
    a, b = states["resident"], states["tile"]
    for i, what in ((1, "depth"), (2, "flux"), (3, "flow vector")):
        assert np.array_equal(bits(a[i]), bits(b[i])), f"{what} differs after 3000 steps"
    assert a[4] == pytest.approx(b[4], rel=1e-12)          # fp64 ledgers: same terms, different summation order


def test_resident_backend_refuses_what_does_not_fit(tws):
    with pytest.raises(tws.TwsError) as e:
        tws.Terrain(2048, backend=tws.BACKEND_RESIDENT)
    assert e.value.status == -5                        # TWS_ERR_UNSUPPORTED
    with pytest.raises(tws.TwsError) as e:
        tws.Terrain(1024, rows=(0, 512), backend=tws.BACKEND_RESIDENT)
    assert e.value.status == -5


def test_auto_backend_resolves_by_grid_size_and_matches_the_oracle(tws, oracle):
    """TWS_BACKEND_AUTO (the default of tws_default_params): tile kernel, 2 steps per launch, below ~12 M cells;
    band kernel, 4 steps per launch, above (measured crossover, profiles/r01_crossover_tile_vs_band.log)."""
    with tws.Terrain(1024) as sim:
        assert sim.backend_in_use() == (tws.BACKEND_FUSED_TB, 2)
    with tws.Terrain(4096) as sim:
        assert sim.backend_in_use() == (tws.BACKEND_BAND_TB, 4)
    with tws.Terrain(4096, rows=(1024, 2048)) as sim:            # a strip counts its own cells
        assert sim.backend_in_use() == (tws.BACKEND_FUSED_TB, 2)
    W, H = 250, 190
    h, d = bumpy(W, H, seed=3)
    t, f, v = new_state(h, d)
    oracle.step(t, f, v, oracle.derive_consts(float(W), W), 13)
    with tws.Terrain(W, height=H) as sim:
        sim.upload(tws.FIELD_TERRAIN, h); sim.upload(tws.FIELD_WATER, d)
        sim.step(13)
        assert_state_equal(sim, tws, t, f, v, "auto")
    # whole grids that fit on chip and fill the SMs: calls of >= 3 steps are ONE resident launch, shorter ones tile-kernel
    # launches; the two share the state and may alternate freely (the reference's frame: 1..10 steps, Terrain.cpp:247-265)
    W = 512
    h, d = bumpy(W, W, seed=4)
    t, f, v = new_state(h, d)
    c = oracle.derive_consts(float(W), W)
    with tws.Terrain(W) as sim:
        assert sim.backend_in_use() == (tws.BACKEND_FUSED_TB, 2)
        sim.upload(tws.FIELD_TERRAIN, h); sim.upload(tws.FIELD_WATER, d)
        for n, launches in ((10, 1), (3, 1), (4, 1), (1, 1), (7, 1), (2, 1), (10, 1)):
            l0 = sim.kernel_launches()
            sim.step(n)
            assert sim.kernel_launches() - l0 == launches, (n, sim.kernel_launches() - l0)
            oracle.step(t, f, v, c, n)
            assert_state_equal(sim, tws, t, f, v, f"auto, frame of {n}")
