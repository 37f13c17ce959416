"""Build recipes for the native pieces (in-tree, so the built .so files travel with the snapshot).

* libdabstar_b200.so  -- the product: CUDA kernels + C-ABI (nvcc, sm_100a only)
* libdab_synth.so     -- bundled synthetic Mode-I transmitter (gcc + OpenMP), a test-signal source
* oracle/libdab_oracle.so and oracle/_ref/libdabref.so -- test infrastructure (see oracle/)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_CUDA = os.path.join(PKG_DIR, "libdabstar_b200.so")
LIB_SYNTH = os.path.join(PKG_DIR, "libdab_synth.so")
LIB_ORACLE = os.path.join(REPO, "oracle", "libdab_oracle.so")
LIB_REF = os.path.join(REPO, "oracle", "_ref", "libdabref.so")
LIB_REF_FAST = os.path.join(REPO, "oracle", "_ref", "libdabref_fast.so")  # timing build of the same sources (bench.py's CPU arm)
REFERENCE_ROOT = "/root/reference"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-DDABSTAR_NO_FAST_MATH",  # IEEE division/sqrt: parity with the float reference matters (no --use_fast_math)
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-cudart", "shared",
]
OBJ_DIR = os.path.join(CSRC, "_obj")  # per-source objects (git-ignored): only what changed is recompiled, in parallel


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _run(cmd: list[str], cwd: str | None = None) -> None:
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))


def cuda_sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def build_cuda(force: bool = False, verbose_ptxas: bool = False) -> str:
    from concurrent.futures import ThreadPoolExecutor
    srcs = cuda_sources()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(REPO, "include", "dabstar_b200.h"))
    if not force and _newer(LIB_CUDA, srcs + headers):
        return LIB_CUDA
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(OBJ_DIR, exist_ok=True)
    base = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose_ptxas else []) + ["-I", os.path.join(REPO, "include"), "-I", CSRC]
    objs, todo = [], []
    for src in srcs:
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or verbose_ptxas or not _newer(obj, [src] + headers):
            todo.append(base + ["-c", "-o", obj, src])
    with ThreadPoolExecutor(max_workers=max(1, min(8, len(todo)))) as pool:
        list(pool.map(_run, todo))
    _run([nvcc, "-shared", "-cudart", "shared", "-o", LIB_CUDA] + objs)
    return LIB_CUDA


def build_synth(force: bool = False) -> str:
    src = os.path.join(PKG_DIR, "synth", "dab_synth.c")
    if not force and _newer(LIB_SYNTH, [src]):
        return LIB_SYNTH
    _run(["gcc", "-std=gnu11", "-O2", "-fPIC", "-fopenmp", "-shared", "-o", LIB_SYNTH, src, "-lm"])
    return LIB_SYNTH


def build_oracle(force: bool = False) -> str:
    odir = os.path.join(REPO, "oracle")
    if force or not _newer(LIB_ORACLE, [os.path.join(odir, f) for f in ("dab_oracle.c", "dab_outer.c", "dab_tii.c", "dab_oracle.h")]):
        _run(["make", "-C", odir, "-B" if force else "-s", "libdab_oracle.so"])
    return LIB_ORACLE


def build_ref(force: bool = False) -> str | None:
    """oracle/_ref from the reference's own sources; only possible where /root/reference exists."""
    if not os.path.isdir(REFERENCE_ROOT):
        return LIB_REF if os.path.exists(LIB_REF) else None
    _run(["make", "-C", os.path.join(REPO, "oracle", "ref_build"), "-j8"] + (["-B"] if force else []))
    _run(["make", "-C", os.path.join(REPO, "oracle", "ref_build"), "-j8", "fast"] + (["-B"] if force else []))
    return LIB_REF


def build_all(force: bool = False) -> None:
    build_cuda(force)
    build_synth(force)
    build_oracle(force)
    build_ref(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built:", LIB_CUDA, LIB_SYNTH, LIB_ORACLE, LIB_REF if os.path.exists(LIB_REF) else "(no oracle/_ref)")
