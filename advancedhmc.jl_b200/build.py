"""build.py -- compile libahmc_b200.so IN-TREE with nvcc for sm_100a (no JIT cache: the .so travels with
the repo snapshot to the GPU box).  Usage: python advancedhmc.jl_b200/build.py [--force] [--verbose]"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.environ.get("AHMC_OBJ_DIR", "/tmp/ahmc_b200_obj")  # objects stay out of the repo snapshot
LIB = os.path.join(HERE, "libahmc_b200.so")
SOURCES = ["ahmc_api.cu", "ahmc_leapfrog.cu", "ahmc_nuts.cu", "ahmc_nuts_var.cu", "ahmc_nuts_adapt.cu", "ahmc_adapt.cu", "ahmc_multinomial.cu", "ahmc_dense.cu", "ahmc_pooled.cu"]
MB_LIB = os.path.join(HERE, "libahmc_microbench.so")  # bench.py's measurement helpers; not part of the C ABI
MB_SOURCE = "ahmc_microbench.cu"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v" if os.environ.get("AHMC_PTXAS_V") else "-O3"]


def _digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(f.encode())
                    h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    stamp = LIB + ".digest"  # travels with the .so (the _obj/ directory is not shipped to the GPU box)
    dg = _digest()
    if (not force and os.path.exists(LIB) and os.path.exists(MB_LIB) and os.path.exists(stamp)
            and open(stamp).read() == dg):
        return LIB
    if not os.path.exists(NVCC):
        raise RuntimeError(f"nvcc not found at {NVCC}: cannot build libahmc_b200.so (no CPU fallback exists)")
    r = subprocess.run([NVCC, *FLAGS, "-shared", os.path.join(CSRC, MB_SOURCE), "-o", MB_LIB, "-cudart", "shared"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {MB_SOURCE}:\n{r.stdout}\n{r.stderr}")

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "shared", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(dg)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
