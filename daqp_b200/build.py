"""Build the CUDA extension in-tree: daqp_b200/libdaqp_b200.so (sm_100a only; nvcc cross-compiles without a GPU).

The translation units are compiled side by side (one nvcc process each) into daqp_b200/build/*.o and linked into the
shared library; objects whose sources did not change are reused."""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libdaqp_b200.so")
SOURCES = ["daqp_b200.cu", "team_launch.cu", "setup2_launch.cu", "dropin.cu"]
HEADERS = ["common.cuh", "ldp_kernel.cuh", "setup_kernel.cuh", "update_kernel.cuh", "minrep_kernel.cuh",
           "warmstart_kernel.cuh", "team_ops.cuh", "setup2_kernel.cuh", "engine.h", os.path.join("..", "..", "include", "daqp_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _newest_dep() -> float:
    deps = [os.path.normpath(os.path.join(CSRC, s)) for s in _sources() + HEADERS]
    return max(os.path.getmtime(d) for d in deps if os.path.exists(d))


def is_stale() -> bool:
    return not os.path.exists(LIB) or _newest_dep() > os.path.getmtime(LIB)


def _compile(src: str, force: bool, verbose: bool, extra) -> str:
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
    path = os.path.join(CSRC, src)
    hdr_t = max(os.path.getmtime(os.path.normpath(os.path.join(CSRC, h))) for h in HEADERS
                if os.path.exists(os.path.normpath(os.path.join(CSRC, h))))
    if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(path), hdr_t):
        return obj
    cmd = [_nvcc()] + NVCC_FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, path]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return obj


def build(force: bool = False, verbose: bool = False, extra=()) -> str:
    if not force and not is_stale():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, verbose, extra), srcs))
    r = subprocess.run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs,
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
