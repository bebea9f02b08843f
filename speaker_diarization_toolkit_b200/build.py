"""In-tree build of the CUDA extension (nvcc, sm_100a only).  No torch involved: the product is a plain
C-ABI shared library, `speaker_diarization_toolkit_b200/libsdk_b200.so`, loaded with ctypes."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OUT = PKG / "libsdk_b200.so"
SOURCES = ["api.cu", "normalize.cu", "exact.cu", "select.cu", "poolgemm.cu", "poolacc.cu", "gemv.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def _stale(obj: Path, src: Path) -> bool:
    if not obj.exists():
        return True
    t = obj.stat().st_mtime
    deps = [src, CSRC / "common.cuh", CSRC / "tcgen05.cuh", PKG.parent / "include" / "sdk_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    bdir = PKG / "build"
    bdir.mkdir(exist_ok=True)
    cc = nvcc()
    jobs = []
    for s in SOURCES:
        src, obj = CSRC / s, bdir / (s + ".o")
        if force or _stale(obj, src):
            cmd = [cc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas")
                cmd.insert(2, "-v")
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(6, len(jobs))) as ex:
            for log in ex.map(run, jobs):
                if verbose and log:
                    print(log, file=sys.stderr)
    objs = [str(bdir / (s + ".o")) for s in SOURCES]
    if force or jobs or not OUT.exists():
        cmd = [cc, "-shared", "-o", str(OUT), *objs, "-cudart", "shared", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed: " + r.stdout + r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
