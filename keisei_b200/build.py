"""In-tree nvcc build of libkeisei_b200.so (sm_100a only).

The library is a plain C-ABI shared object (see include/keisei_b200.h); it links against
libcudart and libcuda only — no torch, no cuBLAS/cuDNN. Objects are rebuilt when their
source (or a shared header) is newer, so `build()` is cheap to call repeatedly.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
BUILD_DIR = PKG_DIR / "csrc" / "build"
LIB_PATH = PKG_DIR / "libkeisei_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: keisei_b200 needs the CUDA 12.9 toolchain to build")


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _headers_mtime() -> float:
    hs = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list((PKG_DIR.parent / "include").glob("*.h"))
    return max((h.stat().st_mtime for h in hs), default=0.0)


def _compile_one(src: Path, obj: Path, log: Path) -> tuple[Path, int, str]:
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", str(CSRC), "-I", str(PKG_DIR.parent / "include"),
           "-c", str(src), "-o", str(obj)]
    p = subprocess.run(cmd, capture_output=True, text=True)
    log.write_text(" ".join(cmd) + "\n" + p.stdout + p.stderr)
    return src, p.returncode, p.stdout + p.stderr


def build(force: bool = False, verbose: bool = False) -> Path:
    BUILD_DIR.mkdir(parents=True, exist_ok=True)
    hdr_m = _headers_mtime()
    jobs = []
    objs = []
    for src in _sources():
        obj = BUILD_DIR / (src.stem + ".o")
        objs.append(obj)
        stale = force or not obj.exists() or obj.stat().st_mtime < max(src.stat().st_mtime, hdr_m)
        if stale:
            jobs.append((src, obj, BUILD_DIR / (src.stem + ".log")))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            results = list(ex.map(lambda j: _compile_one(*j), jobs))
        for src, rc, out in results:
            if verbose or rc != 0:
                print(f"--- nvcc {src.name} (rc={rc})\n{out}", file=sys.stderr)
            if rc != 0:
                raise RuntimeError(f"nvcc failed on {src}")
    need_link = bool(jobs) or not LIB_PATH.exists() or any(
        o.stat().st_mtime > LIB_PATH.stat().st_mtime for o in objs)
    if need_link:
        cmd = [_nvcc(), "-shared", "-o", str(LIB_PATH), *map(str, objs),
               "-gencode", "arch=compute_100a,code=sm_100a", "-lcuda", "-Xcompiler", "-fPIC"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            print(p.stdout + p.stderr, file=sys.stderr)
            raise RuntimeError("link of libkeisei_b200.so failed")
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
