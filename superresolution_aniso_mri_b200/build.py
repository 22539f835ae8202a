"""In-tree build of libaesr_b200.so (nvcc, sm_100a only).  No JIT cache: the .so travels with the repo snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libaesr_b200.so")
LIB_PROBE = os.path.join(LIB_DIR, "libaesr_b200_probe.so")      # diagnostic build (-DAESR_WITH_PROBES), tools only
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-shared"]


def sources():
    deps = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "aesr_b200.h"))
    return deps


def is_stale(lib: str = LIB) -> bool:
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    return any(os.path.getmtime(s) > t for s in sources())


def build_library(force: bool = False, verbose: bool = False, probes: bool = False) -> str:
    """``probes``: the diagnostic library with the aesr_probe_* micro-benchmarks (never loaded by the product path)."""
    lib = LIB_PROBE if probes else LIB
    if not force and not is_stale(lib):
        return lib
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + (["-DAESR_WITH_PROBES"] if probes else []) + (["-Xptxas", "-v"] if verbose else []) + \
        [os.path.join(CSRC, "api.cu"), "-o", lib]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), res.stderr))
    if verbose:
        print(res.stderr)
    return lib


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, probes="--probes" in sys.argv))
