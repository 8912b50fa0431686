"""Build libtws.so (the sm_100a CUDA library behind include/tws.h) in-tree with nvcc.

The library is the product: there is no Python/CPU implementation to fall back to.
`-fmad=false` is part of the arithmetic contract, not a tuning flag.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libtws.so"
SOURCES = ["tws_api.cu", "step_kernels.cu", "stream_kernels.cu", "band_kernels.cu", "resident_kernels.cu", "aux_kernels.cu"]
HEADERS = [CSRC / "tws_internal.h", CSRC / "cell_math.cuh", CSRC / "band_schedule.h", PKG_DIR.parent / "include" / "tws.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-fast-math", "-Xcompiler", "-ffp-contract=off",
    "-shared", "-cudart", "static",
]


def find_nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found; libtws.so cannot be built")
    return cand


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + HEADERS
    return any(d.stat().st_mtime > t for d in deps)


def build_variant(out: Path, defines, verbose: bool = False) -> Path:
    """Tuning helper: build a copy of the library with extra -D defines (not used by the product)."""
    host_cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else None
    cmd = [find_nvcc(), *NVCC_FLAGS] + [f"-D{d}" for d in defines]
    if host_cxx:
        cmd += ["-ccbin", host_cxx]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", str(out), *[str(CSRC / s) for s in SOURCES]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return out


def build_libtws(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB_PATH
    host_cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else None
    cmd = [find_nvcc(), *NVCC_FLAGS]
    if host_cxx:
        cmd += ["-ccbin", host_cxx]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", str(LIB_PATH), *[str(CSRC / s) for s in SOURCES]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build_libtws(force=True, verbose="-v" in sys.argv))
