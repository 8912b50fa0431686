"""The drop-in boundary: libtws.so loads, exports every symbol include/tws.h declares, the
ctypes mirror matches the C structs, argument validation works without a GPU, and the
product fails loudly (no CPU fallback) when no CUDA device exists."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "tws.h").read_text()


def declared_symbols():
    return sorted(set(re.findall(r"^(?:tws_status|const char\*|int32_t|uint64_t|void)\s+(tws_[a-z0-9_]+)\s*\(", HEADER, flags=re.M)))


def test_library_exports_every_declared_symbol(tws):
    lib = tws._abi.load()
    names = declared_symbols()
    assert len(names) >= 28
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/tws.h but not exported by libtws.so"
    assert set(names) == set(tws._abi.SYMBOLS), "ctypes prototype table out of sync with include/tws.h"
    assert lib.tws_abi_version() == int(re.search(r"#define TWS_ABI_VERSION (\d+)", HEADER).group(1))
    assert b"sm_100a" in lib.tws_version()


def test_no_oracle_or_cpu_path_in_the_product():
    """The product never imports/links the oracle; the extension is mandatory."""
    for p in list((ROOT / "terrainwatersim_b200").rglob("*.py")) + list((ROOT / "terrainwatersim_b200" / "csrc").glob("*")):
        if p.suffix not in (".py", ".cu", ".h", ".cpp"):
            continue
        txt = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f"{p} imports the oracle"
        assert "libtws_oracle" not in txt and "oracle_py" not in txt and "tws_oracle" not in txt, f"{p} links/loads the oracle"
    out = subprocess.run(["ldd", str(ROOT / "terrainwatersim_b200" / "libtws.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


@pytest.fixture(scope="module")
def cxx_check(built, tmp_path_factory):
    exe = tmp_path_factory.mktemp("cxx") / "cxx_host_check"
    cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    cmd = [cxx, "-std=c++17", "-O1", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(ROOT / "tests" / "cxx_host_check.cpp"), "-o", str(exe),
           f"-L{ROOT / 'terrainwatersim_b200'}", "-ltws", f"-Wl,-rpath,{ROOT / 'terrainwatersim_b200'}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "c_check.c"
    src.write_text('#include "tws.h"\nint main(void){ tws_params p; (void)p; return (int)sizeof(tws_halo_handle) == 0; }\n')
    cc = "/usr/bin/gcc" if Path("/usr/bin/gcc").exists() else "gcc"
    res = subprocess.run([cc, "-std=c99", "-Wall", "-Werror", "-pedantic", f"-I{ROOT / 'include'}", "-c", str(src), "-o", str(tmp_path / "c_check.o")],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_struct_layout_matches_ctypes_and_bad_size_rejected(tws, cxx_check):
    out = subprocess.run([str(cxx_check)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    kv = dict(re.findall(r"(\w+)=(-?\d+)", out.stdout))
    A = tws._abi
    assert int(kv["sizeof_params"]) == C.sizeof(A.TwsParams)
    assert int(kv["sizeof_consts"]) == C.sizeof(A.TwsStepConstants)
    assert int(kv["sizeof_handle"]) == C.sizeof(A.TwsHaloHandle)
    assert int(kv["off_backend"]) == A.TwsParams.backend.offset
    assert int(kv["off_rain"]) == A.TwsParams.rain_rate.offset
    assert int(kv["bad_size_status"]) == A.TWS_ERR_INVALID


def test_default_params_are_the_reference_defaults(tws):
    lib = tws._abi.load()
    p = tws._abi.TwsParams()
    lib.tws_default_params(C.byref(p))
    assert (p.size, p.width, p.height, p.row_begin, p.row_end) == (C.sizeof(p), 1024, 1024, 0, 1024)      # Terrain.cpp:22-23
    assert (p.world_size, p.steps_per_second, p.flow_acceleration) == (1024.0, 60.0, 10.0)               # Terrain.cpp:22,28,30
    assert abs(p.flow_damping - 0.98) < 1e-7                                                             # Terrain.cpp:29
    assert p.boundary == tws.BOUNDARY_REFERENCE_OPEN


@pytest.mark.parametrize("field,value", [("width", 0), ("height", -4), ("row_end", 2000), ("world_size", 0.0), ("steps_per_second", -1.0),
                                         ("flow_damping", float("nan")), ("backend", 9), ("boundary", 5), ("rain_rate", -1.0)])
def test_create_validates_arguments_before_cuda(tws, field, value):
    lib = tws._abi.load()
    p = tws._abi.TwsParams()
    lib.tws_default_params(C.byref(p))
    setattr(p, field, value)
    sim = C.c_void_p()
    assert lib.tws_create(C.byref(p), C.byref(sim)) == tws._abi.TWS_ERR_INVALID
    assert not sim.value and lib.tws_last_error(None)


def test_null_handles_are_rejected_not_crashing(tws):
    lib = tws._abi.load()
    assert lib.tws_step(None, 1) == tws._abi.TWS_ERR_INVALID
    assert lib.tws_sync(None) == tws._abi.TWS_ERR_INVALID
    assert lib.tws_destroy(None) == tws._abi.TWS_OK
    assert lib.tws_kernel_launches(None) == 0


def test_fails_loudly_without_a_gpu(tws):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device path is exercised on the CPU box")
    with pytest.raises(tws.TwsError) as e:
        tws.Terrain(64)
    assert e.value.status == tws._abi.TWS_ERR_CUDA and "no CPU path" in str(e.value)


def test_no_fused_multiply_add_in_the_step_kernels(tws):
    """Arithmetic contract: no FMA contraction anywhere in the step.  ptxas fuses a
    packed mul.rn.f32x2 feeding a packed add into FFMA2 despite .rn / --fmad=false, so the packed cell
    math never lets a packed product feed a packed add (cell_math.cuh); this checks the SASS.  Scalar
    FFMA is allowed only inside the IEEE division sequence (MUFU.RCP refinement) and its slow-path
    subroutine — never next to the flux / depth arithmetic — so it must not outnumber the divisions."""
    import shutil
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    from terrainwatersim_b200.build import LIB_PATH
    sass = subprocess.run([cuobjdump, "-sass", str(LIB_PATH)], capture_output=True, text=True).stdout
    assert "Function :" in sass
    assert "FFMA2" not in sass, "a packed multiply was contracted into a packed FMA"
    for name, body in re.findall(r"Function : (\S*(?:band_step|fused_step|stream_step|resident_step|unfused_update|unfused_apply)\S*)(.*?)(?=Function :|\Z)", sass, re.S):
        n_ffma = len(re.findall(r"\bFFMA\b", body))
        n_rcp = len(re.findall(r"\bMUFU\.RCP\b", body))
        n_call = len(re.findall(r"\bCALL\.REL", body))
        # 5 FFMA per inlined division fast path; the out-of-line slow path (one copy per kernel) uses a few dozen
        assert n_ffma <= 5 * n_rcp + 60, f"{name}: {n_ffma} FFMA for {n_rcp} divisions"
        assert n_rcp > 0 or n_ffma == 0, name
        assert n_call >= 0


def test_mip_level_rule_matches_the_reference_texture_wrapper(tws):
    """tws_mip_levels / tws_mip_level_info are pure host functions (no GPU needed): level count of glEasy
    Texture.cpp:28-41, glTexStorage2D level sizes, contiguous level offsets."""
    from oracle.oracle_py import mip_levels
    lib = tws._abi.load()
    for W, H in [(1024, 1024), (1, 1), (37, 5), (5, 37), (257, 64), (8192, 8192), (3, 2)]:
        L = lib.tws_mip_levels(W, H)
        assert L == mip_levels(W, H)
        off = 0
        for l in range(L):
            w, h, o = C.c_int32(), C.c_int32(), C.c_int64()
            assert lib.tws_mip_level_info(W, H, l, C.byref(w), C.byref(h), C.byref(o)) == 0
            assert (w.value, h.value, o.value) == (max(1, W >> l), max(1, H >> l), off)
            off += w.value * h.value
        assert lib.tws_mip_level_info(W, H, L, None, None, None) != 0
        assert lib.tws_mip_level_info(W, H, -1, None, None, None) != 0
    assert lib.tws_mip_levels(1024, 1024) == 11
