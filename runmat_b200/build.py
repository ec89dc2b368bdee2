"""Builds librm_accel_b200.so (CUDA, sm_100a) in-tree, plus the CPU oracle used by the tests.

nvcc cross-compiles without a GPU, so this runs on the CPU box (driver's `build()` check) and the
resulting .so travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
BUILD = PKG / "_build"
LIB = PKG / "librm_accel_b200.so"
ORACLE_DIR = ROOT / "oracle"
ORACLE_LIB = ORACLE_DIR / "librm_oracle.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CUDA_LIB = "/usr/local/cuda/lib64"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # Rust never contracts a*b+c; keep the AOT kernels on the same rounding as the CPU builtins.
    "-fmad=false",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unused-function",
    "-cudart", "static",
]


def _sources() -> list[Path]:
    return sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cpp")))


def _stamp(src: Path) -> str:
    h = hashlib.sha256()
    h.update(src.read_bytes())
    for hdr in sorted(CSRC.glob("*.h")) + sorted((ROOT / "include").glob("*.h")):
        h.update(hdr.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(src: Path, verbose: bool) -> Path:
    obj = BUILD / (src.name + ".o")
    stamp_file = BUILD / (src.name + ".stamp")
    stamp = _stamp(src)
    if obj.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return obj
    cmd = [NVCC, *NVCC_FLAGS, "-x", "cu", "-c", str(src), "-o", str(obj), "-I", str(ROOT / "include")]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    if verbose and r.stderr:
        (BUILD / (src.name + ".ptxas.log")).write_text(r.stderr)
    stamp_file.write_text(stamp)
    return obj


def build_library(verbose: bool = False, force: bool = False) -> Path:
    if not shutil.which(NVCC) and not Path(NVCC).exists():
        raise RuntimeError("nvcc not found; the CUDA extension cannot be built")
    BUILD.mkdir(exist_ok=True)
    if force:
        for f in BUILD.glob("*.stamp"):
            f.unlink()
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile_one(s, verbose), srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if LIB.exists() and LIB.stat().st_mtime >= newest and not force:
        return LIB
    cmd = [NVCC, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
           "-o", str(LIB), *map(str, objs), f"-L{CUDA_LIB}", "-lnvrtc", "-ldl", "-lpthread",
           "-Xlinker", f"-rpath={CUDA_LIB}", "-Xlinker", "--no-undefined"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


def build_oracle() -> Path:
    """Compiles the CPU oracle (test infrastructure). Building the checker is not using it."""
    r = subprocess.run(["make", "-C", str(ORACLE_DIR), "-s"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"oracle build failed:\n{r.stdout}\n{r.stderr}")
    return ORACLE_LIB


if __name__ == "__main__":
    verbose = "-v" in sys.argv
    print(build_library(verbose=verbose, force="-f" in sys.argv))
    print(build_oracle())
