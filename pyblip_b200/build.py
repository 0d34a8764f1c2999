"""Build libblipgpu.so (hand-written CUDA for sm_100a) in-tree with nvcc.

    python -m pyblip_b200.build [--force] [--verbose]

The shared library lands in ``pyblip_b200/_C/libblipgpu.so`` (git-ignored; it
travels to the GPU box with the snapshot).  nvcc cross-compiles without a GPU.
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_C")
LIB = os.path.join(OUT_DIR, "libblipgpu.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

SOURCES = ["api.cu", "gram.cu", "gibbs.cu", "samples.cu", "counting.cu", "nprior.cu", "gridcount.cu", "probe.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "--fmad=false",            # scalar decision math must not be contracted; dots use fma() explicitly
    "-Xcompiler", "-fPIC,-O2,-Wall",
    "-I", INCLUDE,
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libblipgpu cannot be built")
    return exe


def _fingerprint():
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for root in (CSRC, INCLUDE):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, name), "rb") as f:
                    h.update(name.encode())
                    h.update(f.read())
    return h.hexdigest()


def is_current():
    stamp = os.path.join(OUT_DIR, "build.stamp")
    if not (os.path.exists(LIB) and os.path.exists(stamp)):
        return False
    with open(stamp) as f:
        return f.read().strip() == _fingerprint()


def build(force=False, verbose=False, extra_flags=()):
    """Compile every CUDA source for sm_100a and link libblipgpu.so."""
    if not force and is_current():
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = _nvcc()
    objs = []

    def compile_one(src):
        obj = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *extra_flags, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        if verbose and (res.stdout or res.stderr):
            print(res.stdout, res.stderr, flush=True)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    with open(os.path.join(OUT_DIR, "build.stamp"), "w") as f:
        f.write(_fingerprint())
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv,
                 extra_flags=("-Xptxas", "-v") if "--ptxas" in sys.argv else ())
    print(path)
