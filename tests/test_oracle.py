"""Pins the CPU oracle (oracle/tws_oracle.cpp): against the reference's own simulation shaders
(flowUpdate.comp / flowApply.comp / waterBrush.comp compiled from /root/reference through the GLSL shim
into oracle/_ref/libtws_ref_step.so) and the reference's own terrain generator (oracle/_ref/
libtws_ref_terrain.so), bit for bit; against hand-derived micro cases of the two shader passes; against
invariants of the scheme and against the committed golden hashes (which make_goldens.py takes from the
reference-compiled shaders).  The reference has no tests for this path — these are the
replacement pins of SURVEY.md §8c."""
import ctypes as C
import json
from pathlib import Path

import numpy as np
import pytest

from oracle.oracle_py import RefStep, dam_break, new_state

ROOT = Path(__file__).resolve().parent.parent
GOLD = json.loads((ROOT / "tests" / "golden" / "goldens.json").read_text())


def f32(x):
    return np.float32(x)


# ---- a1: per-step constants (Terrain.cpp:175-198) ------------------------------------------
def test_default_constants_pinned(oracle):
    c = oracle.derive_consts(1024.0, 1024, 60.0, 0.98, 10.0)
    assert [float(x).hex() for x in c] == GOLD["consts_default_hex"]
    # 0.99966335, 0.16666667, 0.016666668
    assert np.allclose(c, [0.99966335, 0.16666667, 0.016666668], rtol=1e-7)


def test_constants_follow_float_double_mix(oracle):
    for world, res, sps, damp, acc in [(1024.0, 1024, 60.0, 0.98, 10.0), (512.0, 1024, 120.0, 0.96, 16.0), (1024.0, 256, 30.0, 0.5, 30.0)]:
        dt = float(f32(1.0) / f32(sps))                   # ezTime::Seconds(1.0f / sps) -> double
        cell = f32(world) / f32(res)
        want_acc = f32(dt * float(f32(acc)) * float(cell))
        want_area = f32(dt / float(cell * cell))
        c = oracle.derive_consts(world, res, sps, damp, acc)
        assert c[1] == want_acc and c[2] == want_area
        assert abs(float(c[0]) - float(damp) ** dt) < 1e-6


# ---- rg16f store: round-to-nearest-even --------------------------------------------------------
def test_half_conversion_matches_ieee_rtne(oracle):
    rng = np.random.default_rng(7)
    vals = np.concatenate([
        rng.standard_normal(20000).astype(np.float32) * 100,
        (rng.standard_normal(5000) * 1e-6).astype(np.float32),
        np.float32([0.0, -0.0, 65504.0, 65519.9, 65520.0, 70000.0, -70000.0, 6e-8, 5.96e-8, 2.98e-8, 2.9802322e-8, 3e-8, 1e-10,
                    6.1035156e-5, 6.0975552e-5, np.inf, -np.inf, 1.00048828125, 1.0009765625 + 0.00048828125, 2049.0, 2051.0]),
    ])
    with np.errstate(over="ignore"):
        want = vals.astype(np.float16).view(np.uint16)
    got = np.array([oracle.float_to_half_bits(float(v)) for v in vals], dtype=np.uint16)
    assert np.array_equal(got, want)
    assert np.isnan(np.uint16(oracle.float_to_half_bits(float("nan"))).view(np.float16))


# ---- a9: create/reset against the reference's own generator ----------------------------------------
def _ref_lib():
    p = ROOT / "oracle" / "_ref" / "libtws_ref_terrain.so"
    if not p.exists():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return C.CDLL(str(p))


def test_white_noise_table_equals_reference_random(oracle):
    ref = _ref_lib()
    w = np.zeros(4096, np.float32)
    ref.tws_ref_white_noise(C.c_uint32(231656522), w.ctypes.data_as(C.c_void_p))
    assert np.array_equal(w.view(np.uint32), oracle.white_noise(231656522).view(np.uint32))
    assert f"{oracle.fnv1a64(w):016x}" == GOLD["white_noise_ref"]
    assert -1.0 <= w.min() and w.max() <= 1.0


@pytest.mark.parametrize("res", [64, 777, 1024])
def test_scene_bit_equal_to_reference_compiled_generator(oracle_omp, res):
    ref = _ref_lib()
    a = np.zeros((res, res, 4), np.float32)
    ref.tws_ref_create_scene(C.c_uint32(231656522), res, C.c_float(300.0), a.ctypes.data_as(C.c_void_p))
    b = oracle_omp.create_scene(res)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_scene_golden_and_survey_statistics(oracle_omp):
    s = oracle_omp.create_scene(1024)
    h, d = np.ascontiguousarray(s[..., 0]), np.ascontiguousarray(s[..., 3])
    g = GOLD["scene1024"]
    assert f"{oracle_omp.fnv1a64(h):016x}" == g["h"] == GOLD["scene1024_ref"]["h"]
    assert f"{oracle_omp.fnv1a64(d):016x}" == g["d"] == GOLD["scene1024_ref"]["d"]
    # SURVEY.md Appendix B figures, measured on the reference-compiled generator
    assert h.min() == f32(47.4498672) and h.max() == f32(257.200928)
    assert (d > 0).sum() == 43911 and d.max() == f32(81.4746552)
    assert abs(d.sum(dtype=np.float64) - 1828432.61) < 0.01
    assert h[0, 0] == h[1023, 1023] == f32(257.151459)          # period-1 noise
    assert h[511, 511] == f32(66.4997406) and d[512, 512] == f32(68.531517)
    assert np.all(s[..., 1] == f32(0.3)) and np.all(s[..., 2] == f32(0.3))


def test_pow_squared_is_pinned_to_the_product(oracle_omp):
    """Terrain.cpp:219 calls pow(l2, 2.0f) with l2 = |p - 0.5|^2 in [0, 0.5].  The oracle and the CUDA scene kernel evaluate
    l2 * l2 — and so does the reference's own generator as g++ compiles it (the call is folded; oracle/_ref agrees bit for bit,
    test_scene_bit_equal_to_reference_compiled_generator, also at a size that is no power of two).  That this is a PIN and not
    an identity is checked exhaustively: over EVERY binary32 value in [0, 0.5] this libm's powf routine (called through a
    volatile pointer, so nothing is folded) is within rounding of the product but not always equal to it."""
    lib = oracle_omp.lib
    lib.tws_oracle_set_threads(max(1, len(__import__("os").sched_getaffinity(0))))
    lib.tws_oracle_pow2_mismatches.restype = C.c_uint64
    lib.tws_oracle_pow2_mismatches.argtypes = [C.c_uint32, C.c_uint32]
    half = int(np.float32(0.5).view(np.uint32))
    total = half + 1
    bad = lib.tws_oracle_pow2_mismatches(0, half)
    assert bad < total // 1000                      # a (nearly) correctly rounded routine: equal for > 99.9 % of the arguments
    assert lib.tws_oracle_pow2_mismatches(0, 0x00800000) == 0          # subnormal arguments: both underflow to 0


# ---- a5/a6 micro cases (hand-derived from the shader text) ----------------------------------------
def test_single_wet_cell_one_step(oracle):
    """3x3 flat terrain, centre depth d0: each direction gets accel*d0 (flowUpdate.comp:44-54);
    the clamp (:57-59) triggers iff 4*accel*d0*areaInv > d0."""
    c = np.float32([1.0, 0.125, 0.5])
    h = np.zeros((3, 3), np.float32)
    d = np.zeros((3, 3), np.float32)
    d[1, 1] = 2.0
    t, f, v = new_state(h, d)
    oracle.flow_update(t, f, c)
    assert np.array_equal(f[1, 1], np.float32([0.25, 0.25, 0.25, 0.25]))      # total*areaInv = 0.5 <= 2: no clamp
    assert np.count_nonzero(f) == 4
    oracle.flow_apply(t, f, v, c)
    assert t[1, 1, 3] == f32(2.0 - 1.0 * 0.5)
    for (y, x) in [(1, 2), (1, 0), (2, 1), (0, 1)]:
        assert t[y, x, 3] == f32(0.125)
    assert t[..., 3].sum() == f32(2.0)
    # velocity of the centre: symmetric -> 0; right neighbour (x=2): in from -X side 0.25
    assert np.array_equal(v[1, 1], np.float16([0, 0]))
    assert np.array_equal(v[1, 2], np.float16([-0.25, 0]))   # (iX1 - f.x) - (iX0 - f.y) = (0-0) - (0.25-0)
    assert np.array_equal(v[2, 1], np.float16([0, -0.25]))

    # clamp case: accel 1, areaInv 0.5 -> total = 4*2*0.5 = 4 > 2 -> scale 0.5
    c2 = np.float32([1.0, 1.0, 0.5])
    t, f, v = new_state(h, d)
    oracle.flow_update(t, f, c2)
    assert np.array_equal(f[1, 1], np.float32([1, 1, 1, 1]))
    oracle.flow_apply(t, f, v, c2)
    assert t[1, 1, 3] == 0.0 and t[..., 3].sum() == f32(2.0)


def test_open_boundary_drains_to_zero_height_exterior(oracle):
    """A border cell sees an exterior neighbour of height 0 (out-of-range imageLoad = 0)."""
    c = np.float32([1.0, 0.125, 0.5])
    h = np.full((1, 1), 3.0, np.float32)
    d = np.full((1, 1), 1.0, np.float32)
    t, f, v = new_state(h, d)
    oracle.flow_update(t, f, c)
    assert np.array_equal(f[0, 0], np.float32([0.5, 0.5, 0.5, 0.5]))          # (4 - 0) * 0.125 each way
    oracle.flow_apply(t, f, v, c)
    assert t[0, 0, 3] == 0.0                                                   # 1 + (0 - 2)*0.5 -> max(0, 0)
    # closed-wall extension: nothing leaves
    t, f, v = new_state(h, d)
    oracle.step(t, f, v, c, 5, boundary=1)
    assert t[0, 0, 3] == 1.0 and not f.any()


def test_lake_at_rest_and_dry_grid_are_fixed_points(oracle):
    rng = np.random.default_rng(3)
    c = oracle.derive_consts(64.0, 64)
    h = (rng.random((64, 64)) * 10).astype(np.float32)
    h[0, :] = h[-1, :] = h[:, 0] = h[:, -1] = 100.0
    level = f32(20.0)
    d = np.where(h < level, level - h, 0).astype(np.float32)
    # make H = d + h exactly flat where wet (choose h so that the sum is exact)
    h = np.where(d > 0, level - d, h).astype(np.float32)
    assert np.all((d + h)[d > 0] == level)
    t, f, v = new_state(h, d)
    oracle.step(t, f, v, c, 25)
    assert np.array_equal(t[..., 3], d) and not f.any()
    t, f, v = new_state(h, np.zeros_like(d))
    oracle.step(t, f, v, c, 25)
    assert not t[..., 3].any() and not f.any() and not v.view(np.uint16).any()


def test_mirror_symmetry_is_bitwise(oracle):
    rng = np.random.default_rng(11)
    c = oracle.derive_consts(48.0, 48)
    q = (rng.random((48, 24)) * 5).astype(np.float32)
    h = np.concatenate([q, q[:, ::-1]], axis=1)
    dq = (rng.random((48, 24)) * 3).astype(np.float32)
    d = np.concatenate([dq, dq[:, ::-1]], axis=1)
    t, f, v = new_state(h, d)
    oracle.step(t, f, v, c, 40)
    assert np.array_equal(t[..., 3], t[:, ::-1, 3])
    assert np.array_equal(f[..., 0], f[:, ::-1, 1]) and np.array_equal(f[..., 2], f[:, ::-1, 2])
    # Transposition / y-mirror are NOT bitwise symmetries: the shader sums ((x + y) + z) + w
    # (flowUpdate.comp:57, flowApply.comp:38-39), which is only invariant under swapping x<->y
    # components.  They hold to rounding only.
    t2, f2, v2 = new_state(h.T.copy(), d.T.copy())
    oracle.step(t2, f2, v2, c, 40)
    assert np.allclose(t2[..., 3], t[..., 3].T, rtol=1e-4, atol=1e-6)


def test_invariants_and_closed_volume(oracle):
    h, d = dam_break(128, rim=True)
    c = oracle.derive_consts(128.0, 128)
    t, f, v = new_state(h, d)
    v0 = t[..., 3].sum(dtype=np.float64)
    for _ in range(10):
        oracle.step(t, f, v, c, 50)
        assert t[..., 3].min() >= 0 and f.min() >= 0
        assert np.isfinite(t).all() and np.isfinite(f).all()
    assert abs(t[..., 3].sum(dtype=np.float64) - v0) / v0 < 1e-6      # north_star: volume conserved to 1e-6


def test_open_volume_change_equals_boundary_outflow(oracle):
    h, d = dam_break(96, rim=False)
    c = oracle.derive_consts(96.0, 96)
    t, f, v = new_state(h, d)
    vol = t[..., 3].sum(dtype=np.float64)
    lost = 0.0
    for _ in range(200):
        oracle.flow_update(t, f, c)
        out = (f[:, -1, 0].sum(dtype=np.float64) + f[:, 0, 1].sum(dtype=np.float64) + f[-1, :, 2].sum(dtype=np.float64) + f[0, :, 3].sum(dtype=np.float64))
        oracle.flow_apply(t, f, v, c)
        lost += out * float(c[2])
    now = t[..., 3].sum(dtype=np.float64)
    assert lost > 0.01 * vol                                   # the open boundary really drains (SURVEY.md §0 finding 1)
    assert abs((vol - now) - lost) / vol < 1e-6


def test_openmp_build_is_bit_identical(oracle, oracle_omp):
    h, d = dam_break(200, 120, rim=False)
    c = oracle.derive_consts(200.0, 200)
    a = new_state(h, d)
    b = new_state(h, d)
    oracle.step(*a, c, 60)
    oracle_omp.step(*b, c, 60)
    for x, y in zip(a, b):
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8))


@pytest.mark.parametrize("name,rim", [("dam256_walled", True), ("dam256_open", False)])
def test_dam_break_goldens(oracle_omp, name, rim):
    h, d = dam_break(256, rim=rim)
    c = oracle_omp.derive_consts(256.0, 256)
    t, f, v = new_state(h, d)
    done = 0
    for n in (1, 10, 100, 1000):
        oracle_omp.step(t, f, v, c, n - done)
        done = n
        g = GOLD[name][str(n)]
        assert f"{oracle_omp.fnv1a64(np.ascontiguousarray(t[..., 3])):016x}" == g["d"]
        assert f"{oracle_omp.fnv1a64(f):016x}" == g["F"]
        assert f"{oracle_omp.fnv1a64(v.view(np.uint16)):016x}" == g["v"]
    if rim:
        v0 = float(d.sum(dtype=np.float64))
        assert abs(GOLD[name]["1000"]["volume"] - v0) / v0 < 1e-6
    else:
        assert GOLD[name]["1000"]["volume"] < 0.35 * float(d.sum(dtype=np.float64))   # ~69 % drained


# ---- a7 / a8 host logic -------------------------------------------------------------------------------
def test_advance_accumulator(oracle):
    step = float(f32(1.0) / f32(60.0))
    n, acc = oracle.advance(0.0, step, 1.0 / 60.0)       # double 1/60 < (double)(1.0f/60.0f): one frame is not quite a step
    assert n == 0 and acc == 1.0 / 60.0
    n, acc = oracle.advance(acc, step, 1.0 / 60.0)
    assert n == 1 and abs(acc - (2.0 / 60.0 - step)) < 1e-15
    n, acc = oracle.advance(0.0, step, 1.0)              # 59 or 60 steps due -> clamped to 10, remainder time kept as in Terrain.cpp:244
    assert n == 10 and acc < step
    n, acc = oracle.advance(0.0, step, 0.001)
    assert n == 0 and acc == 0.001


def test_brush_center_and_footprint(oracle):
    cx, cy = oracle.brush_center(1536.5, 100.25, 1024.0, 1024)
    assert (cx, cy) == (512.5, 100.25)
    cx, cy = oracle.brush_center(-10.0, 5.0, 1024.0, 1024)
    assert cx == -10.0 and cy == 5.0                      # Fraction keeps the sign: negative coords -> no wrap
    t, f, v = new_state(np.zeros((64, 64), np.float32), np.ones((64, 64), np.float32))
    oracle.brush(t, 31.5, 20.0, 2.0, 32.0)
    yy, xx = np.mgrid[0:64, 0:64]
    dist = ((31.5 - xx) ** 2 + (20.0 - yy) ** 2) / 32.0
    assert np.array_equal(t[..., 3] != 1.0, dist < 1.0)
    assert t[20, 31, 3] == f32(1.0 + (1.0 - 0.25 / 32.0) * 2.0)


# ---- a second, independent transcription -------------------------------------------------------------
def _numpy_step(t, f, consts):
    """flowUpdate.comp:26-62 then flowApply.comp:30-52 as whole-image numpy float32 expressions, written from the
    shader text independently of oracle/tws_oracle.cpp (different structure: padded images instead of per-cell
    neighbour loads).  Out-of-range imageLoad = 0 is the zero padding.  Returns the RG16F flow map."""
    fr, ac, ai = (np.float32(x) for x in consts)
    H, W = t.shape[:2]
    hw = np.zeros((H + 2, W + 2), np.float32)
    hw[1:-1, 1:-1] = t[..., 3] + t[..., 0]                     # terrainInfo.a + terrainInfo.r  (:34-40)
    own = hw[1:-1, 1:-1]
    new = np.empty((H, W, 4), np.float32)
    new[..., 0] = own - hw[1:-1, 2:]                           # X1 = x+1  (:44)
    new[..., 1] = own - hw[1:-1, :-2]                          # X0 = x-1  (:45)
    new[..., 2] = own - hw[2:, 1:-1]                           # Y1 = y+1  (:46)
    new[..., 3] = own - hw[:-2, 1:-1]                          # Y0 = y-1  (:47)
    new = f * fr + new * ac                                    # :53 (each product and the sum rounded to binary32)
    new = np.maximum(np.float32(0), new)                       # :54
    total = (((new[..., 0] + new[..., 1]) + new[..., 2]) + new[..., 3]) * ai      # :57
    a = t[..., 3]
    over = total > a                                           # :58
    with np.errstate(divide="ignore", invalid="ignore"):       # 0/0 where the branch is not taken
        f[...] = np.where(over[..., None], new * (a / total)[..., None], new)     # :59, :62
    fp = np.zeros((H + 2, W + 2, 4), np.float32)
    fp[1:-1, 1:-1] = f
    x1 = fp[1:-1, 2:, 1]; x0 = fp[1:-1, :-2, 0]; y1 = fp[2:, 1:-1, 3]; y0 = fp[:-2, 1:-1, 2]      # flowApply.comp:32-35
    ingoing = ((x1 + x0) + y1) + y0                            # :38
    outgoing = ((f[..., 0] + f[..., 1]) + f[..., 2]) + f[..., 3]                   # :39
    t[..., 3] = np.maximum(np.float32(0), a + (ingoing - outgoing) * ai)           # :41
    v = np.empty((H, W, 2), np.float32)
    v[..., 0] = (x1 - f[..., 0]) - (x0 - f[..., 1])            # :45
    v[..., 1] = (y1 - f[..., 2]) - (y0 - f[..., 3])            # :46
    return v.astype(np.float16)                                # imageStore to rg16f: round to nearest even


@pytest.mark.parametrize("W,H,seed", [(48, 40, 1), (33, 7, 2), (5, 61, 3)])
def test_cxx_oracle_equals_an_independent_numpy_transcription(oracle, W, H, seed):
    """Two transcriptions of the two shaders written separately (C++ per-cell loops, numpy whole-image
    expressions) agree bit for bit over 40 steps of a rough random scene with dry, wet and draining cells —
    a guard against transcription slips, since the reference offers no golden vector for the step."""
    rng = np.random.default_rng(seed)
    h = (rng.random((H, W)) * 8).astype(np.float32)
    d = (rng.random((H, W)) * 4 * (rng.random((H, W)) > 0.4)).astype(np.float32)
    c = oracle.derive_consts(float(max(W, H)), max(W, H))
    t1, f1, v1 = new_state(h, d)
    t2, f2, _ = new_state(h, d)
    scaled = 0
    for step in range(40):
        oracle.step(t1, f1, v1, c, 1)
        v2 = _numpy_step(t2, f2, c)
        assert np.array_equal(t1.view(np.uint32), t2.view(np.uint32)), f"TerrainData differs at step {step}"
        assert np.array_equal(f1.view(np.uint32), f2.view(np.uint32)), f"Flow differs at step {step}"
        assert np.array_equal(v1.view(np.uint16), v2.view(np.uint16)), f"FlowMap differs at step {step}"
        scaled += int(((t2[..., 3] == 0) & (f2.sum(axis=-1) > 0)).sum())
    assert scaled > 0                      # the clamp branch (:58-59) was exercised


# ---- the step pinned against the reference's OWN shader source ----------------------------------------------------
# oracle/_ref/libtws_ref_step.so = /root/reference/terrainwatersim/shader/{flowUpdate,flowApply,waterBrush}.comp compiled
# where they lie (oracle/Makefile `ref_step`; language shim oracle/ref_shim/glsl.h; host side oracle/ref_step_driver.cpp
# dispatching res/16 x res/16 groups of 18 x 18 invocations as Terrain.cpp:258,264 and res/32 groups of 32 x 32 for the
# brush as Terrain.cpp:167).  The reference is defined on multiples of 16 (32 for the brush) only.
@pytest.fixture(scope="session")
def ref_step(built):
    try:
        return RefStep()
    except FileNotFoundError:
        pytest.skip("oracle/_ref/libtws_ref_step.so not built (needs /root/reference)")


def _bits_equal(a, b):
    return all(np.array_equal(x.view(np.uint8), y.view(np.uint8)) for x, y in zip(a, b))


def _hashes(o, t, f, v):
    return {"d": f"{o.fnv1a64(np.ascontiguousarray(t[..., 3])):016x}", "F": f"{o.fnv1a64(f):016x}", "v": f"{o.fnv1a64(v.view(np.uint16)):016x}"}


@pytest.mark.parametrize("name,rim", [("dam256_walled", True), ("dam256_open", False)])
def test_reference_shaders_config1_1000_steps(oracle_omp, ref_step, name, rim):
    """BASELINE config 1 through the reference's shaders: equal to the oracle and to the goldens at 1/10/100/1000 steps."""
    assert GOLD["step_goldens_source"].startswith("reference shaders")
    h, d = dam_break(256, rim=rim)
    c = oracle_omp.derive_consts(256.0, 256)
    a, b = new_state(h, d), new_state(h, d)
    done = 0
    for n in (1, 10, 100, 1000):
        ref_step.step(*a, c, n - done)
        oracle_omp.step(*b, c, n - done)
        done = n
        assert _bits_equal(a, b), f"oracle differs from the reference shaders after {n} steps"
        g = GOLD[name][str(n)]
        assert _hashes(oracle_omp, *a) == {k: g[k] for k in ("d", "F", "v")}


def test_reference_shaders_config2_scene_with_brush_1000_steps(oracle_omp, ref_step):
    """BASELINE config 2 (1024^2 reference scene, brush at (512,512) with 100/60 before every step, Scene.cpp:356-363)
    through the reference's brush and step shaders."""
    c = oracle_omp.derive_consts(1024.0, 1024)
    scene = oracle_omp.create_scene(1024)
    a = (scene.copy(), np.zeros((1024, 1024, 4), np.float32), np.zeros((1024, 1024, 2), np.float16))
    b = (scene.copy(), np.zeros((1024, 1024, 4), np.float32), np.zeros((1024, 1024, 2), np.float16))
    inten = np.float32(100.0 / 60.0)
    done = 0
    for n in (1, 10, 100, 1000):
        for _ in range(n - done):
            ref_step.brush(a[0], 512.0, 512.0, inten, 32.0)
            ref_step.step(*a, c, 1)
            oracle_omp.brush(b[0], 512.0, 512.0, inten, 32.0)
            oracle_omp.step(*b, c, 1)
        done = n
        assert _bits_equal(a, b), f"oracle differs from the reference shaders after {n} steps"
        g = GOLD["scene1024_brush"][str(n)]
        assert _hashes(oracle_omp, *a) == {k: g[k] for k in ("d", "F", "v")}


@pytest.mark.parametrize("W,H,seed", [(16, 16, 1), (64, 32, 2), (32, 96, 3), (160, 128, 4), (256, 64, 5)])
def test_reference_shaders_random_grids_every_pass(oracle, ref_step, W, H, seed):
    """Rough random scenes (dry, wet and draining cells, huge gradients at the open edge), random per-step constants, each of
    the two dispatches compared separately, brush with fractional / off-grid centres and negative intensity in between."""
    rng = np.random.default_rng(seed)
    h = (rng.random((H, W)) * 8).astype(np.float32)
    d = (rng.random((H, W)) * 4 * (rng.random((H, W)) > 0.4)).astype(np.float32)
    c = np.float32([0.9 + 0.1 * rng.random(), 0.05 + 0.3 * rng.random(), 0.01 + 0.3 * rng.random()])
    a, b = new_state(h, d), new_state(h, d)
    scaled = 0
    for step in range(60):
        if step % 7 == 3 and W % 32 == 0 and H % 32 == 0:
            cx, cy = float(rng.uniform(-4, W + 4)), float(rng.uniform(-4, H + 4))
            inten, size_sq = float(rng.uniform(-1, 3)), float(rng.uniform(1, 60))
            ref_step.brush(a[0], cx, cy, inten, size_sq)
            oracle.brush(b[0], cx, cy, inten, size_sq)
            assert _bits_equal(a, b), f"brush differs at step {step}"
        ref_step.flow_update(a[0], a[1], c)
        oracle.flow_update(b[0], b[1], c)
        assert _bits_equal(a, b), f"flowUpdate differs at step {step}"
        scaled += int((((a[1].sum(axis=-1) * c[2]) >= a[0][..., 3]) & (a[1].sum(axis=-1) > 0)).sum())
        ref_step.flow_apply(a[0], a[1], a[2], c)
        oracle.flow_apply(b[0], b[1], b[2], c)
        assert _bits_equal(a, b), f"flowApply differs at step {step}"
    assert scaled > 0                      # the clamp branch (flowUpdate.comp:58-59) was exercised
    assert np.isfinite(a[0]).all()


def test_reference_shaders_micro_cases_and_group_seams(oracle, ref_step):
    """Hand-checkable values through the reference's shaders: a single wet cell placed ON a work-group seam (x = 15 | 16)
    and one in the grid corner (the open boundary: out-of-range imageLoad = 0), flowUpdate.comp:44-59."""
    c = np.float32([1.0, 0.125, 0.5])
    h = np.zeros((32, 32), np.float32)
    d = np.zeros((32, 32), np.float32)
    d[15, 16] = 2.0
    d[0, 0] = 1.0
    h[0, 0] = 3.0
    t, f, v = new_state(h, d)
    ref_step.flow_update(t, f, c)
    assert np.array_equal(f[15, 16], np.float32([0.25, 0.25, 0.25, 0.25]))
    assert np.array_equal(f[0, 0], np.float32([0.5, 0.5, 0.5, 0.5]))          # (4 - 0) * 0.125 towards exterior and dry neighbours
    ref_step.flow_apply(t, f, v, c)
    assert t[15, 16, 3] == f32(1.5) and t[15, 15, 3] == t[15, 17, 3] == t[14, 16, 3] == t[16, 16, 3] == f32(0.125)
    assert t[0, 0, 3] == 0.0 and t[0, 1, 3] == f32(0.25) and t[1, 0, 3] == f32(0.25)     # half of the corner's water left the grid
    assert np.array_equal(v[15, 17], np.float16([-0.25, 0]))
    t2, f2, v2 = new_state(h, d)
    oracle.step(t2, f2, v2, c, 1)
    assert _bits_equal((t, f, v), (t2, f2, v2))


def test_reference_shader_rg16f_store_rounds_to_nearest_even(oracle, ref_step):
    """FlowMap values straddling fp16 rounding ties, subnormals and overflow, produced through flowApply.comp:45-52."""
    vals = np.float32([1.00048828125, 1.0009765625 + 0.00048828125, 2049.0, 2051.0, 65519.9, 65520.0, 70000.0, 6e-8, 2.98e-8,
                       2.9802322e-8, 3e-8, 6.1035156e-5, 6.0975552e-5, 0.1, 1e-3, 3.14159])
    W = H = 16
    t, f, v = new_state(np.zeros((H, W), np.float32), np.zeros((H, W), np.float32))
    for i, x in enumerate(vals):                  # cell (1, i%..): inflow from its -X neighbour's +X flux = x  ->  v.x = -(x)
        f[2 * (i // 4) + 1, 4 * (i % 4), 0] = x   # +X outflow of the left neighbour
    t2, f2, v2 = t.copy(), f.copy(), v.copy()
    c = np.float32([1.0, 0.0, 0.0])
    ref_step.flow_apply(t, f, v, c)
    oracle.flow_apply(t2, f2, v2, c)
    assert np.array_equal(v.view(np.uint16), v2.view(np.uint16))
    with np.errstate(over="ignore"):
        for i, x in enumerate(vals):
            assert v[2 * (i // 4) + 1, 4 * (i % 4) + 1, 0].view(np.uint16) == np.float16(-x).view(np.uint16)
