#!/usr/bin/env python
"""Builds `oracle/_ref/liblgrngn_ref_cuda.so`: the reference's OWN Thrust/CUDA back-end (src/lib_cuda.cu, unmodified,
compiled in place for sm_100) behind the same flat C binding as the CPU oracle.

Purpose: a like-for-like GPU baseline - `bench.py --impl reference-cuda` times it on the same B200 and the same cfg4 slab
as the product ("the bar on the GPU side is the reference's generic Thrust/CUB path compiled for
sm_100").  Informative arm only: no parity test compares against it (its random stream is cuRAND MTGP32).

Flags follow the reference's release configuration for CUDA sources (CMakeLists.txt:283: -DNDEBUG -O3 -use_fast_math,
host side -Ofast) without -march=native (the object must run on the GPU box's CPU).  The single nvcc invocation takes
about 25 minutes; it needs no GPU.  Outputs only under oracle/_ref/ (git-ignored, shipped by gpurun).
"""
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("LCX_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref", "cuda")
LIB = os.path.join(HERE, "_ref", "liblgrngn_ref_cuda.so")
NVCC = os.environ.get("LCX_NVCC", "/usr/local/cuda/bin/nvcc")
CXX = os.environ.get("LCX_CXX", "/usr/bin/g++")

INC = ["-I", os.path.join(HERE, "boost_shim"), "-I", os.path.join(REF, "include"), "-I", os.path.join(REF, "src")]


def build(force=False, verbose=True):
    if not os.path.isdir(os.path.join(REF, "src")):
        if os.path.exists(LIB):
            return LIB
        raise RuntimeError("reference sources not found at %s and no prebuilt %s" % (REF, LIB))
    if os.path.exists(LIB) and not force:
        return LIB
    import importlib.util
    spec = importlib.util.spec_from_file_location("build_ref", os.path.join(HERE, "build_ref.py"))
    cpu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cpu)
    cpu.build(verbose=verbose, fast=True)          # serial / OpenMP objects and the binding are shared with the -Ofast CPU build
    os.makedirs(OUT, exist_ok=True)
    cu_obj = os.path.join(OUT, "lib_cuda.o")
    if force or not os.path.exists(cu_obj):
        t0 = time.time()
        subprocess.run([NVCC, "-std=c++17", "-O3", "-DNDEBUG", "-use_fast_math", "-gencode", "arch=compute_100,code=sm_100",
                        "--expt-relaxed-constexpr", "--extended-lambda", "-w", "-Xcompiler", "-fPIC,-Ofast,-fopenmp", "-ccbin", CXX]
                       + INC + ["-c", os.path.join(REF, "src", "lib_cuda.cu"), "-o", cu_obj], check=True)
        if verbose:
            print("[oracle cuda] lib_cuda.o %.0f s" % (time.time() - t0), flush=True)
    lib_obj = os.path.join(OUT, "lib.o")
    subprocess.run([CXX, "-std=c++17", "-Ofast", "-DNDEBUG", "-fPIC", "-w", "-fopenmp", "-DCUDA_FOUND"] + INC
                   + ["-I", "/usr/local/cuda/include", "-c", os.path.join(REF, "src", "lib.cpp"), "-o", lib_obj], check=True)
    fast = os.path.join(HERE, "_ref", "fast")
    objs = [os.path.join(fast, u[0]) for u in cpu.UNITS if u[0] != "lib.o"] + [lib_obj, cu_obj]
    subprocess.run([CXX, "-shared", "-fopenmp", "-o", LIB] + objs
                   + ["-L/usr/local/cuda/lib64", "-lcudart_static", "-lcurand", "-ldl", "-lrt", "-lpthread",
                      "-Wl,-Bsymbolic", "-Wl,--exclude-libs,ALL"], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
