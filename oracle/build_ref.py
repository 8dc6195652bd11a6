#!/usr/bin/env python
"""Builds the parity oracle `oracle/_ref/liblgrngn_ref.so` from the UNMODIFIED reference sources.

What is compiled (all with g++ -std=c++17 -O2 -DNDEBUG -ffp-contract=off, no -ffast-math, so that
floating point is plain IEEE and comparable with `nvcc -fmad=false`):
  * /root/reference/src/lib_cpp.cpp   - the reference's serial back-end (Thrust "cpp" system)
  * /root/reference/src/lib_omp.cpp   - the reference's OpenMP back-end (-fopenmp)
  * /root/reference/src/lib.cpp       - the reference's factory
  * libcloudphxx_b200/bindings/lgrngn_capi.cpp, with -DLGC_REFERENCE_BUILD and the reference's
    headers: the same flat C binding the product ships, so Python drives both identically; twice, for the
    reference's double and float instantiations (lgc_* / lgcf_*)
  * oracle/ref_internals.cpp (x2) + ref_internals_glue.cpp - read-only access to private state
The reference needs Boost (absent from this image); oracle/boost_shim/ supplies the few names it uses.
Thrust comes from the CUDA toolkit (host back-ends only: no GPU code in the oracle).

Outputs go only to oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun).  The reference
tree is read in place and never copied.  On the GPU box /root/reference does not exist: the script
then leaves a previously built library alone.
"""
import os
import subprocess
import sys
import time
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("LCX_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT, "liblgrngn_ref.so")

CXX = os.environ.get("LCX_CXX", "/usr/bin/g++")
BASE = [CXX, "-std=c++17", "-O2", "-DNDEBUG", "-ffp-contract=off", "-fPIC", "-w",
        "-I", os.path.join(HERE, "boost_shim"), "-I", os.path.join(REF, "include"),
        "-I", "/usr/local/cuda/include", "-I", os.path.join(REF, "src")]

UNITS = [  # (object name, source, extra flags)
    ("lib_cpp.o", os.path.join(REF, "src", "lib_cpp.cpp"), []),
    ("lib_omp.o", os.path.join(REF, "src", "lib_omp.cpp"), ["-fopenmp"]),
    ("lib.o", os.path.join(REF, "src", "lib.cpp"), ["-fopenmp"]),
    ("capi.o", os.path.join(REPO, "libcloudphxx_b200", "bindings", "lgrngn_capi.cpp"),
     ["-DLGC_REFERENCE_BUILD", "-I", os.path.join(REPO, "libcloudphxx_b200", "bindings")]),
    ("capi_f32.o", os.path.join(REPO, "libcloudphxx_b200", "bindings", "lgrngn_capi.cpp"),
     ["-DLGC_REFERENCE_BUILD", "-DLGC_FLOAT", "-I", os.path.join(REPO, "libcloudphxx_b200", "bindings")]),
    ("internals_serial.o", os.path.join(HERE, "ref_internals.cpp"), []),
    ("internals_serial_f32.o", os.path.join(HERE, "ref_internals.cpp"), ["-DLGC_INTERNALS_F32"]),
    ("internals_omp_f32.o", os.path.join(HERE, "ref_internals.cpp"), ["-fopenmp", "-DLGC_INTERNALS_OMP", "-DLGC_INTERNALS_F32"]),
    ("internals_omp.o", os.path.join(HERE, "ref_internals.cpp"), ["-fopenmp", "-DLGC_INTERNALS_OMP"]),
    ("internals_glue.o", os.path.join(HERE, "ref_internals_glue.cpp"), []),
]


def newest_input():
    t = 0.0
    for _, src, _ in UNITS:
        if os.path.exists(src):
            t = max(t, os.path.getmtime(src))
    for root, _, files in os.walk(os.path.join(HERE, "boost_shim")):
        for f in files:
            t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


# A second build of the same sources with the optimisation flags of the reference's own portable release configuration
# (CMakeLists.txt:126, RelWithDebInfoPortable: -Ofast; the default Release adds -march=native, which would tie the library to the
# CPU of the build container).  TIMING ONLY: bench.py's CPU baseline uses it so that the reference is not handicapped by the
# IEEE-strict flags the parity oracle needs; no test compares results of this library.
LIB_FAST = os.path.join(OUT, "liblgrngn_ref_fast.so")
FAST_FLAGS = ["-Ofast", "-DNDEBUG"]


def build(force=False, verbose=True, fast=False):
    lib = LIB_FAST if fast else LIB
    out = os.path.join(OUT, "fast") if fast else OUT
    base = [f for f in BASE if f not in ("-O2", "-DNDEBUG", "-ffp-contract=off")] + FAST_FLAGS if fast else BASE
    if not os.path.isdir(os.path.join(REF, "src")):
        if os.path.exists(lib):
            return lib
        raise RuntimeError("reference sources not found at %s and no prebuilt %s" % (REF, lib))
    if not force and os.path.exists(lib) and os.path.getmtime(lib) >= newest_input():
        return lib
    os.makedirs(out, exist_ok=True)

    def compile_one(unit):
        obj, src, extra = unit
        objp = os.path.join(out, obj)
        if not force and os.path.exists(objp) and os.path.getmtime(objp) >= max(os.path.getmtime(src), 0):
            if not src.startswith(REPO) or os.path.getmtime(objp) >= newest_input():
                return obj, 0.0
        t0 = time.time()
        subprocess.run(base + extra + ["-c", src, "-o", objp], check=True)
        return obj, time.time() - t0

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for obj, dt in ex.map(compile_one, UNITS):
            if verbose:
                print("[oracle%s] %-20s %6.1f s" % (" -Ofast" if fast else "", obj, dt), flush=True)
    subprocess.run([CXX, "-shared", "-fopenmp", "-o", lib] + [os.path.join(out, u[0]) for u in UNITS]
                   + ["-Wl,-Bsymbolic", "-Wl,--exclude-libs,ALL"], check=True)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
    print(build(force="--force" in sys.argv, fast=True))
