"""CPU-only checks of the drop-in boundary: the shared libraries load without a GPU, export every symbol the headers
declare, refuse to run without a device (no CPU fallback), and reproduce the reference's error behaviour for API misuse."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from libcloudphxx_b200 import engine as E
from libcloudphxx_b200 import lgrngn as L
from tests import support as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:lcx|lgc|lgrngn_b200)_[a-z0-9_]+)\s*\(", text)))


def test_engine_abi_exports_every_declared_symbol():
    lib = E.lib()
    names = declared_functions(os.path.join(ROOT, "include", "lcx_b200.h"))
    assert len(names) > 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert b"sm_100a" in lib.lcx_version()


def test_single_precision_engine_exports_the_same_abi_with_suffix():
    """liblcx_b200_f32.so: every entry point of include/lcx_b200.h under its _f32 name (include/lcx_b200_f32_names.h), and the
    generated name list is in step with the header"""
    names = declared_functions(os.path.join(ROOT, "include", "lcx_b200.h"))
    cdll = C.CDLL(E.LCX_F32_LIB_PATH)
    missing = [n for n in names if not hasattr(cdll, n + "_f32")]
    assert not missing, missing
    listed = re.findall(r"#define (lcx_[A-Za-z0-9_]+) \1_f32", open(os.path.join(ROOT, "include", "lcx_b200_f32_names.h")).read())
    assert sorted(set(listed) - {"lcx_engine"}) == sorted(set(re.findall(r"\b(lcx_[A-Za-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "lcx_b200.h")).read(), flags=re.S))))
    assert b"sm_100a" in E.lib("f32").lcx_version()
    # the two engines must not share C++ symbols (each binds its own: -Bsymbolic + namespaces lcx / lcx_f32)
    assert not hasattr(cdll, "lcx_create")


def test_float_binding_is_exported():
    """factory<float> through the flat binding: the lgcf_* twins of every lgc_* entry point"""
    lib = L.b200("f32")
    names = [n for n in declared_functions(os.path.join(ROOT, "libcloudphxx_b200", "bindings", "lgrngn_capi.h")) if n.startswith("lgc_")]
    missing = [n for n in names if not hasattr(lib.cdll, "lgcf_" + n[4:])]
    assert names and not missing, missing
    assert lib.lib.lgc_impl_name() == b"b200" and lib.dtype == np.float32


def test_host_library_exports_binding_and_extras():
    lib = L.b200().lib
    for header in ("libcloudphxx_b200/bindings/lgrngn_capi.h", "libcloudphxx_b200/host/particles_b200.h"):
        names = declared_functions(os.path.join(ROOT, header))
        assert names
        missing = [n for n in names if not hasattr(lib, n)]
        assert not missing, (header, missing)


@pytest.mark.parametrize("header", ["include/lcx_b200.h", "libcloudphxx_b200/bindings/lgrngn_capi.h", "libcloudphxx_b200/host/particles_b200.h"])
def test_boundary_headers_are_plain_c(header, tmp_path):
    """the drop-in boundary is a C ABI: every header a foreign-function binding would read compiles as C99 (and as C++) on its own"""
    import subprocess
    for lang, std in (("c", "-std=c99"), ("c++", "-std=c++11")):
        r = subprocess.run(["gcc", std, "-Wall", "-Wextra", "-pedantic", "-Werror", "-x", lang, "-fsyntax-only", "-"],
                           input='#include "%s"\n' % os.path.join(ROOT, header), capture_output=True, text=True)
        assert r.returncode == 0, (lang, r.stderr)


def gpu_present():
    return E.lib().lcx_device_count() > 0


@pytest.mark.skipif("gpu_present()")
def test_no_cpu_fallback():
    """without a CUDA device the product refuses to run instead of silently computing on the host"""
    lib = L.b200()
    oi, o, f = S.box_golovin(lib, n_sd=64)
    p = lib.factory(L.backend_t.CUDA, oi)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        p.init(f["th"], f["rv"], f["rhod"])


def test_cpu_backends_are_not_part_of_the_product():
    lib = L.b200()
    oi, _, _ = S.box_golovin(lib, n_sd=64)
    for backend, name in ((L.backend_t.serial, "serial"), (L.backend_t.OpenMP, "OpenMP")):
        with pytest.raises(RuntimeError, match="%s backend was not compiled" % name):
            lib.factory(backend, oi)
    with pytest.raises(RuntimeError, match="unknown backend"):
        lib.factory(L.backend_t.undefined, oi)


@pytest.mark.parametrize("field,value,message", [
    ("chem_switch", 1, "chemistry"), ("ice_switch", 1, "ice"), ("turb_coal_switch", 1, "turbulence"),
    ("turb_cond_switch", 1, "turbulence")])
def test_out_of_scope_options_are_refused_loudly(field, value, message):
    lib = L.b200()
    oi, _, _ = S.box_golovin(lib, n_sd=64)
    setattr(oi, field, value)
    with pytest.raises(RuntimeError, match=message):
        lib.factory(L.backend_t.CUDA, oi)


def test_call_order_and_argument_errors_match_the_reference(ref):
    """same std::runtime_error texts as the reference (src/particles_step.ipp:44-47,169-170,343-344; init_sanity_check.ipp)"""
    new = L.b200()

    def messages(lib, backend):
        out = []
        oi, o, f = S.box_golovin(lib, n_sd=64)
        p = lib.factory(backend, oi)
        for call in (lambda: p.step_sync(o, f["th"], f["rv"], f["rhod"]),        # before init
                     lambda: p.init(None, f["rv"], f["rhod"]),                    # th missing
                     ):
            try:
                call()
                out.append(None)
            except RuntimeError as ex:
                out.append(str(ex))
        return out
    a = messages(ref, L.backend_t.serial)
    b = messages(new, L.backend_t.CUDA)
    assert a == b and all(a), (a, b)

    # kernel parameter validation happens before any device work
    oi, o, f = S.box_golovin(new, n_sd=64)
    oi.kernel_parameters = []
    p = new.factory(L.backend_t.CUDA, oi)
    with pytest.raises(RuntimeError, match="Golovin kernel accepts exactly one parameter"):
        p.init(f["th"], f["rv"], f["rhod"])
    oi, o, f = S.box_golovin(new, n_sd=64)
    oi.dt = 0
    p = new.factory(L.backend_t.CUDA, oi)
    with pytest.raises(RuntimeError, match="please specify opts_init.dt"):
        p.init(f["th"], f["rv"], f["rhod"])


def test_multi_cuda_constructor_checks():
    lib = L.b200()
    oi, _, _ = S.box_golovin(lib, n_sd=64)
    with pytest.raises(RuntimeError, match="multi_CUDA doesn't work for 0D setup"):
        lib.factory(L.backend_t.multi_CUDA, oi)


def test_efficiency_tables_are_shipped():
    for name in ("hall", "hall_davis_no_waals", "vohl_davis_no_waals", "hall_pinsky_stratocumulus", "hall_pinsky_cumulonimbus", "hall_pinsky_1000mb_grav"):
        t = np.fromfile(os.path.join(ROOT, "libcloudphxx_b200", "data", name + ".f64"))
        assert t.size == 1 + 201 * 202 // 2 and t[0] == 1100.0
        assert 0.0 <= t[1:].min() and t[1:].max() < 1e2


# ---- binary compatibility with callers compiled against the reference's own headers -------------------------------------------
REF_ROOT = "/root/reference"
GOLDEN_LAYOUT = os.path.join(ROOT, "tests", "golden", "abi_layout_reference.txt")
OWN_INC = os.path.join(ROOT, "libcloudphxx_b200", "host", "include")


def build_probe(tmp_path, include_dirs, name):
    import subprocess
    exe = str(tmp_path / name)
    cmd = ["g++", "-std=c++17", "-w"] + [a for d in include_dirs for a in ("-I", d)] + ["-I", OWN_INC, os.path.join(ROOT, "tests", "cpp", "abi_probe_main.cpp"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return subprocess.run([exe], capture_output=True, text=True, check=True).stdout


def test_api_layout_equals_the_reference_layout(tmp_path):
    """sizes, member offsets, enum values and v-table slots seen through THIS library's headers == those recorded from the
    reference's headers (tests/golden/abi_layout_reference.txt): a model compiled against either can link to liblgrngn_b200.so"""
    own = build_probe(tmp_path, [], "probe_own")
    assert own == open(GOLDEN_LAYOUT).read()
    assert "diag_sd_conc=5" in own and "outbuf=" in own


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF_ROOT, "include")), reason="reference headers only exist in the development container")
def test_recorded_reference_layout_is_current(tmp_path):
    ref = build_probe(tmp_path, [os.path.join(ROOT, "oracle", "boost_shim"), os.path.join(REF_ROOT, "include"), "/usr/local/cuda/include"], "probe_ref")
    assert ref == open(GOLDEN_LAYOUT).read()


def test_library_reports_the_layout_it_was_built_with():
    lib = L.b200().lib
    lib.lgrngn_b200_abi_layout.restype = C.c_long
    lib.lgrngn_b200_abi_layout.argtypes = [C.c_char_p, C.c_long]
    n = lib.lgrngn_b200_abi_layout(None, 0)
    buf = C.create_string_buffer(n + 1)
    lib.lgrngn_b200_abi_layout(buf, n + 1)
    assert buf.value.decode() == open(GOLDEN_LAYOUT).read()


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF_ROOT, "include")), reason="reference headers only exist in the development container")
def test_caller_built_with_reference_headers_links_and_calls(tmp_path):
    """a translation unit that only ever saw the REFERENCE's headers, linked to liblgrngn_b200.so: factory() (options passed by
    value across the boundary) answers with the reference's error for a back-end that is not compiled in, and - where no GPU
    exists - refuses the CUDA back-end instead of computing on the host"""
    import subprocess
    src = tmp_path / "caller.cpp"
    src.write_text(r'''
#include <libcloudph++/lgrngn/factory.hpp>
#include <iostream>
using namespace libcloudphxx::lgrngn;
int main()
{
  opts_init_t<double> oi;
  oi.dt = 1; oi.sd_conc = 8; oi.n_sd_max = 8; oi.kernel = kernel_t::golovin; oi.kernel_parameters = {1500.};
  oi.terminal_velocity = vt_t::beard77fast;
  int ok = 0;
  try { factory<double>(serial, oi); } catch (const std::runtime_error &e) { std::cout << "serial: " << e.what() << "\n"; ++ok; }
  try
  {
    particles_proto_t<double> *p = factory<double>(CUDA, oi);
    std::cout << "cuda: created n_sd_max=" << p->opts_init->n_sd_max << " kernel=" << int(p->opts_init->kernel) << "\n";
    try { p->diag_sd_conc(); } catch (const std::exception &e) { std::cout << "diag: " << e.what() << "\n"; }
    delete p;
    ++ok;
  }
  catch (const std::runtime_error &e) { std::cout << "cuda: " << e.what() << "\n"; ++ok; }
  return ok == 2 ? 0 : 1;
}
''')
    exe = str(tmp_path / "caller")
    libdir = os.path.join(ROOT, "libcloudphxx_b200", "lib")
    cmd = ["g++", "-std=c++17", "-w", "-I", os.path.join(ROOT, "oracle", "boost_shim"), "-I", os.path.join(REF_ROOT, "include"),
           "-I", "/usr/local/cuda/include", str(src), "-o", exe, "-L", libdir, "-llgrngn_b200", "-llcx_b200", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "serial: libcloudph++: serial backend was not compiled" in out.stdout
    assert "cuda: created n_sd_max=8 kernel=2" in out.stdout, out.stdout                    # options survived the trip by value
    assert "diag: libcloudph++: please call init() before asking for diagnostics" in out.stdout, out.stdout      # v-table slot 5 is diag_sd_conc
