"""In-tree build of librvae_b200.so (sm_100a only) with nvcc.

The shared library is a plain C-ABI object (include/rvae_b200.h); it is built next to this file so that it
travels with the repo snapshot to the GPU box and shows up as an in-tree native module when loaded.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
BUILD_DIR = PKG_DIR / "_build"
LIB_PATH = PKG_DIR / "librvae_b200.so"
SOURCES = ["gemm_host.cu", "elementwise.cu", "capi.cu"]
HEADERS = ["ptx.cuh", "gemm.cuh", "common.h", "../../include/rvae_b200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]
if os.environ.get("RVAE_EXPERIMENTS", "0") not in ("", "0"):
    NVCC_FLAGS.append("-DRVAE_EXPERIMENTS=1")   # measured-and-rejected paths (csrc/gemm.cuh); part of the fingerprint


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found: librvae_b200.so cannot be built (no CPU fallback exists)")
    return cand


def _fingerprint() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        h.update((CSRC / name).read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def needs_build() -> bool:
    stamp = BUILD_DIR / "fingerprint"
    return not (LIB_PATH.exists() and stamp.exists() and stamp.read_text() == _fingerprint())


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a and link librvae_b200.so. Returns the library path.
    Safe under torchrun: an inter-process file lock serialises concurrent first-use builds (the ranks that lose the
    race find an up-to-date library when they get the lock) and the library is linked to a temporary name and moved
    into place atomically, so nobody can dlopen a half-written file."""
    if not force and not needs_build():
        return LIB_PATH
    import fcntl
    BUILD_DIR.mkdir(exist_ok=True)
    with open(BUILD_DIR / ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():
                return LIB_PATH
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool) -> Path:
    nvcc = _nvcc()

    def compile_one(src: str) -> Path:
        obj = BUILD_DIR / (Path(src).stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        (BUILD_DIR / (Path(src).stem + ".ptxas.log")).write_text(res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            print(res.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = LIB_PATH.with_suffix(f".so.tmp{os.getpid()}")
    cmd = [nvcc, "-shared", "-o", str(tmp), *map(str, objs), "-cudart", "static", "-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    os.replace(tmp, LIB_PATH)
    (BUILD_DIR / "fingerprint").write_text(_fingerprint())
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose=True)
    print(path)
