"""In-tree build of the B200 back-end (no JIT cache: the .so files travel with the repository snapshot).

  lib/liblcx_b200.so     CUDA engine, hand-written sm_100a kernels behind the C ABI of include/lcx_b200.h (double precision)
  lib/liblcx_b200_f32.so the same sources compiled with -DLCX_F32: real = float, every entry point suffixed _f32
  lib/liblgrngn_b200.so  C++ host layer (lgrngn::factory / particles_proto_t) + flat C binding for Python

nvcc cross-compiles without a GPU.  Flags: -gencode arch=compute_100a,code=sm_100a -lineinfo; -fmad=false keeps
the reference's order of floating-point operations (bit-exact positions / collision decisions).
"""
import os
import subprocess
import sys
import time
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
BIND = os.path.join(HERE, "bindings")
# LCX_BUILD_TAG=<tag> builds a tuning variant into lib_<tag>/ (objects in build_<tag>/) next to the product; the loaders pick
# it up through LCX_B200_LIBDIR.  Experiments only: the product is lib/.
_TAG = os.environ.get("LCX_BUILD_TAG", "")
LIBDIR = os.path.join(HERE, "lib" + ("_" + _TAG if _TAG else ""))
OBJDIR = os.path.join(HERE, "build" + ("_" + _TAG if _TAG else ""))

NVCC = os.environ.get("LCX_NVCC", "/usr/local/cuda/bin/nvcc")
CXX = os.environ.get("LCX_CXX", "/usr/bin/g++")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC", "-ccbin", CXX,
              "-I", os.path.join(REPO, "include")]
CXX_FLAGS = ["-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-fopenmp", "-Wall", "-Wno-unused-function",
             "-I", os.path.join(REPO, "include"), "-I", os.path.join(HOST, "include"), "-I", CSRC, "-I", BIND,
             "-I", "/usr/local/cuda/include"]

CU_SOURCES = ["lcx_api.cu", "lcx_sort.cu", "lcx_cells.cu", "lcx_diag.cu", "lcx_cond.cu", "lcx_cond_pp.cu", "lcx_coal.cu",
              "lcx_transport.cu", "lcx_layout.cu", "lcx_init.cu"]
CU_HEADERS = ["lcx_engine.cuh", "lcx_physics.h", os.path.join(REPO, "include", "lcx_b200_f32_names.h")]
# translation units whose results are tolerance-class anyway (condensation root solve): FMA contraction allowed
FMAD_OK = {"lcx_cond.cu"}


def _mtime(p):
    return os.path.getmtime(p) if os.path.exists(p) else 0.0


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def build_engine(force=False, verbose=True, extra=(), f32=False):
    """the double-precision engine, or (f32) the single-precision one: same sources, -DLCX_F32, own object directory"""
    objdir = os.path.join(OBJDIR, "f32") if f32 else OBJDIR
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(objdir, exist_ok=True)
    hdr_t = max([_mtime(os.path.join(CSRC, h)) for h in CU_HEADERS] + [_mtime(os.path.join(REPO, "include", "lcx_b200.h"))])
    lib = os.path.join(LIBDIR, "liblcx_b200_f32.so" if f32 else "liblcx_b200.so")
    if f32:
        extra = list(extra) + ["-DLCX_F32"]

    def one(src):
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        if not force and _mtime(o) >= max(_mtime(s), hdr_t):
            return src, 0.0, ""
        t0 = time.time()
        flags = list(NVCC_FLAGS)
        if src in FMAD_OK and os.environ.get("LCX_COND_FMAD", "0") == "1":
            flags[flags.index("-fmad=false")] = "-fmad=true"
        if src == "lcx_cond.cu":
            flags += [f for f in os.environ.get("LCX_COND_DEFS", "").split() if f]
        flags += [f for f in os.environ.get("LCX_DEFS_" + src.split(".")[0].upper(), "").split() if f]   # tuning experiments
        out = _run([NVCC] + flags + list(extra) + ["-c", s, "-o", o])
        return src, time.time() - t0, out

    objs = []
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for src, dt, out in ex.map(one, CU_SOURCES):
            objs.append(os.path.join(objdir, src.replace(".cu", ".o")))
            if verbose and dt:
                print("[build] nvcc %-18s %5.1f s%s" % (src, dt, "  (f32)" if f32 else ""), flush=True)
            if verbose and out.strip():
                print(out)
    if force or _mtime(lib) < max(_mtime(o) for o in objs):
        # -Bsymbolic: the two engines define the same C++ helpers; each must bind to its own
        _run([NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-Xlinker", "-Bsymbolic"])
    return lib


def build_host(force=False, verbose=True):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    lib = os.path.join(LIBDIR, "liblgrngn_b200.so")
    srcs = [os.path.join(HOST, "particles_b200.cpp"), os.path.join(BIND, "lgrngn_capi.cpp")]
    deps = srcs + [os.path.join(REPO, "include", "lcx_b200.h"), os.path.join(CSRC, "lcx_physics.h"), os.path.join(HOST, "lcx_api_select.hpp"),
                   os.path.join(HOST, "include", "libcloudph++", "lgrngn", "lgrngn_b200_api.hpp"),
                   os.path.join(BIND, "lgrngn_capi.h")]
    if not force and _mtime(lib) >= max(_mtime(d) for d in deps) and _mtime(lib) >= _mtime(os.path.join(LIBDIR, "liblcx_b200.so")) \
            and _mtime(lib) >= _mtime(os.path.join(LIBDIR, "liblcx_b200_f32.so")):
        return lib
    # the flat binding twice: particles_proto_t<double> (lgc_*) and particles_proto_t<float> (lgcf_*)
    units = [(srcs[0], "particles_b200.o", []), (srcs[1], "lgrngn_capi.o", []), (srcs[1], "lgrngn_capi_f32.o", ["-DLGC_FLOAT"])]

    def one(u):
        s, name, extra = u
        o = os.path.join(OBJDIR, name)
        t0 = time.time()
        _run([CXX] + CXX_FLAGS + extra + ["-c", s, "-o", o])
        return o, name, time.time() - t0

    objs = []
    with ThreadPoolExecutor(max_workers=3) as ex:
        for o, name, dt in ex.map(one, units):
            if verbose:
                print("[build] g++  %-18s %5.1f s" % (name, dt), flush=True)
            objs.append(o)
    _run([CXX, "-shared", "-fopenmp", "-o", lib] + objs + ["-L", LIBDIR, "-llcx_b200", "-llcx_b200_f32", "-Wl,-rpath,$ORIGIN", "-Wl,-Bsymbolic"])
    return lib


def build_all(force=False, verbose=True):
    return build_engine(force, verbose), build_engine(force, verbose, f32=True), build_host(force, verbose)


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv))
