"""In-tree build of the CUDA library: nvcc -> facemap_b200/lib/libfacemap_b200.so (sm_100a only).

    python -m facemap_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  nvcc cross-compiles
without a GPU, so this also runs in the CPU-only build container.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIBPATH = os.path.join(LIBDIR, "libfacemap_b200.so")
SOURCES = ["api.cu", "bin_motion.cu", "svd_helpers.cu", "gemm_tf32x3.cu", "eigh.cu", "eigh_cluster.cu", "roi_procs.cu", "motion_ref.cu", "gemm_i8.cu", "project_i8.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "numpy_sum.cuh"), os.path.join(CSRC, "pupil_math.cuh"),
           os.path.join(CSRC, "slice_math.cuh"), os.path.join(CSRC, "i8_mma.cuh"),
           os.path.join(os.path.dirname(PKG), "include", "facemap_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def _nvcc() -> str:
    cand = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    if os.path.exists(cand):
        return cand
    found = shutil.which("nvcc")
    if not found:
        raise RuntimeError("nvcc not found: cannot build libfacemap_b200.so")
    return found


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    objs, procs = [], []
    for src in SOURCES:
        spath = os.path.join(CSRC, src)
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [spath] + HEADERS):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", spath, "-o", obj]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, pr in procs:
        out, _ = pr.communicate()
        log.append(f"==== {src}\n{out}")
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    if procs:
        with open(os.path.join(LIBDIR, "build.log"), "w") as f:
            f.write("\n".join(log))
        if verbose:
            print("\n".join(log))
    if force or procs or _stale(LIBPATH, objs):
        cuda_lib = os.path.join(os.path.dirname(os.path.dirname(nvcc)), "lib64")
        cmd = [nvcc, "-shared", "-o", LIBPATH] + objs + [
            "-gencode", "arch=compute_100a,code=sm_100a", "-L" + cuda_lib, "-lcusolver", "-lcudart",
            "-Xlinker", "-rpath," + cuda_lib]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout)
    return LIBPATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
