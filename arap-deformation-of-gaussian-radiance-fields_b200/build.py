"""In-tree build of libarapgs.so (sm_100a only).

nvcc cross-compiles without a GPU.  Every translation unit is built with
-fmad=false: float expressions must round exactly like the reference's host
code (no FMA contraction); fused multiply-adds are written explicitly where
wanted.  The .so stays in-tree (git-ignored, but it travels with gpurun).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OBJ = HERE / "_build"
SO = HERE / "libarapgs.so"
CLI = HERE / "arap_replay"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = [*ARCH, "-lineinfo", "-O3", "-std=c++17", "-fmad=false", "-Xcompiler", "-fPIC"]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off"]

CU = ["apply.cu", "knn.cu", "solve.cu", "solve_smem.cu", "solve_pipe.cu", "grid.cu", "session.cu"]
CPP = ["host_io.cpp", "mcast.cpp"]
HEADERS = ["common.cuh", "device_math.cuh", "sh_fast.cuh", "solve_dev.h", "solve_smem_dev.cuh", "kernels.h", "session.h", "mcast.h", "../../include/arapgs.h", "../../include/arapgs_kernels.h"]


def _stale(src: Path, obj: Path) -> bool:
    if not obj.exists():
        return True
    t = obj.stat().st_mtime
    deps = [src] + [(CSRC / h).resolve() for h in HEADERS]
    return any(d.stat().st_mtime > t for d in deps if d.exists())


def _compile(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: " + " ".join(map(str, cmd)) + "\n" + r.stdout + r.stderr)


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    jobs, objs = [], []
    for f in CU:
        o = OBJ / (Path(f).stem + ".o")
        objs.append(o)
        if force or _stale(CSRC / f, o):
            jobs.append([NVCC, *NVCC_FLAGS, "-c", str(CSRC / f), "-o", str(o)])
    for f in CPP:
        o = OBJ / (Path(f).stem + ".o")
        objs.append(o)
        if force or _stale(CSRC / f, o):
            jobs.append(["g++", *CXX_FLAGS, "-I/usr/local/cuda/include", "-c", str(CSRC / f), "-o", str(o)])
    if verbose:
        for j in jobs:
            print(" ".join(map(str, j)), file=sys.stderr)
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 4))) as ex:
        list(ex.map(_compile, jobs))
    if jobs or not SO.exists():
        _compile([NVCC, *ARCH, "-shared", "-o", str(SO), *map(str, objs), "-lcudart", "-ldl"])
    # headless replay CLI (tools/arap_replay.cpp): plain C++ over the C ABI, linked against the library next to it
    cli_src = HERE.parent / "tools" / "arap_replay.cpp"
    if cli_src.exists() and (jobs or not CLI.exists() or cli_src.stat().st_mtime > CLI.stat().st_mtime):
        _compile(["g++", *CXX_FLAGS, str(cli_src), "-o", str(CLI), f"-L{HERE}", "-larapgs", "-Wl,-rpath,$ORIGIN",
                  "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-lcudart"])
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
