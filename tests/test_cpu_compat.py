"""The Python surface of the reference (`from libcloudphxx import lgrngn, common`) re-created on the flat C binding.

Here (no GPU) the compat package is pointed at the parity oracle - the reference's own CPU back-ends - and must run the
reference's own test scripts UNCHANGED (tests/python/{unit,physics}/*.py): that pins names,
argument order, defaults, enum spelling, buffer protocol and dict conventions.  The same package serves the B200 library
on the GPU box (tests/test_gpu_compat.py).  The scripts are read from /root/reference, so this part only runs where the
reference is mounted."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "libcloudphxx_b200", "compat")
REF_TESTS = "/root/reference/tests/python"
FIXTURES = os.path.join(ROOT, "tests", "golden", "ref_scripts")      # verbatim copies, so the scripts also run where the reference is absent

SCRIPTS = ["unit/uniform_init.py", "unit/lgrngn_adve.py", "unit/terminal_velocities.py", "unit/multiple_kappas.py",
           "unit/adve_scheme.py", "unit/lgrngn_subsidence.py", "physics/test_coal.py", "physics/lgrngn_cond.py", "physics/puddle.py"]


def env(impl):
    e = dict(os.environ)
    e["PYTHONPATH"] = ROOT + os.pathsep + COMPAT + os.pathsep + e.get("PYTHONPATH", "")
    if impl == "reference":      # the oracle exports the same flat binding: point the package at it (test infrastructure only)
        e["LIBCLOUDPHXX_COMPAT_LIBRARY"] = os.path.join(ROOT, "oracle", "_ref", "liblgrngn_ref.so")
    e.setdefault("OMP_NUM_THREADS", "8")
    return e


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="reference test scripts are not mounted here")
@pytest.mark.parametrize("script", SCRIPTS)
def test_fixture_copy_is_the_reference_script(script):
    assert open(os.path.join(FIXTURES, script), "rb").read() == open(os.path.join(REF_TESTS, script), "rb").read()


@pytest.mark.parametrize("script", SCRIPTS)
def test_reference_script_runs_unchanged_on_the_compat_package(script):
    path = os.path.join(FIXTURES, script)
    r = subprocess.run([sys.executable, os.path.basename(path)], cwd=os.path.dirname(path), env=env("reference"),
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]


def test_common_constants_and_known_answers():
    sys.path.insert(0, COMPAT)
    try:
        from libcloudphxx import common
    finally:
        sys.path.remove(COMPAT)
    assert common.p_vs(273.16) == pytest.approx(611.73, rel=1e-12)          # tests/unit/test_common_pvs.cpp:7
    assert common.R_d == pytest.approx(8.3144621 / 0.02897)
    assert common.th_std2dry(common.th_dry2std(300., 0.01), 0.01) == pytest.approx(300., rel=1e-14)
    assert common.T(300., 1.1) == pytest.approx(300. * common.exner(common.p(1.1, 0., common.T(300., 1.1))), rel=1e-12)
    assert common.S_cr(1e-24, 0.61, 283.) > 1 and common.rw3_cr(1e-24, 0.61, 283.) > 1e-24
    assert common.p_hydro(0., 300., 0.01, 0., 1e5) == pytest.approx(1e5, rel=1e-12)
    with pytest.raises(AttributeError):
        common._call("no_such_function", 1.0)


def test_surface_names_defaults_and_errors():
    os.environ.pop("LIBCLOUDPHXX_COMPAT_LIBRARY", None)
    sys.path.insert(0, COMPAT)
    try:
        from libcloudphxx import lgrngn
    finally:
        sys.path.remove(COMPAT)
    oi = lgrngn.opts_init_t()
    assert (oi.nx, oi.ny, oi.nz, oi.sd_conc, oi.sstp_cond, oi.sstp_coal) == (0, 0, 0, 0, 1, 1)      # opts_init.hpp:194-249
    assert oi.RH_max == pytest.approx(.95) and oi.rng_seed == 44 and oi.th_dry and not oi.const_p
    assert str(lgrngn.kernel_t.long) == "kernel_t.long" and str(oi.kernel) == "kernel_t.undefined"
    oi.kernel = lgrngn.kernel_t.hall_davis_no_waals
    assert oi.kernel == lgrngn.kernel_t.hall_davis_no_waals
    o = lgrngn.opts_t()
    assert o.adve and o.sedi and o.cond and o.coal and not o.rcyc and o.RH_max == 44 and o.dt == -1    # opts.hpp:50-66
    with pytest.raises(RuntimeError, match="getter"):
        oi.dry_distros
    oi.dry_distros = {0.61: lambda lnr: 1.0}
    oi.dt, oi.sd_conc, oi.n_sd_max = 1, 8, 8
    for backend, name in ((lgrngn.backend_t.serial, "serial"), (lgrngn.backend_t.OpenMP, "OpenMP")):
        with pytest.raises(RuntimeError, match="%s backend was not compiled" % name):
            lgrngn.factory(backend, oi)
    o.chem_rct = True
    with pytest.raises(RuntimeError, match="chemistry was switched off"):
        o._check()
