"""`libcloudphxx.common` as the reference's Python binding exposes it (bindings/python/lib.cpp:40-120, common.hpp:20-170):
constants and scalar thermodynamic helpers.  Evaluated by the same C++ formulae the back-end uses on the host
(csrc/lcx_physics.h through lgrngn_b200_common); no GPU is needed for these."""
import ctypes as _C

from libcloudphxx_b200 import lgrngn as _L

_lib = _C.CDLL(_L.B200_LIB_PATH, mode=_C.RTLD_LOCAL)
_lib.lgrngn_b200_common.restype = _C.c_double
_lib.lgrngn_b200_common.argtypes = [_C.c_char_p] + [_C.c_double] * 5


def _call(name, *args):
    a = [float(x) for x in args] + [0.0] * (5 - len(args))
    v = _lib.lgrngn_b200_common(name.encode(), *a)
    if v != v and all(x == x for x in a):
        raise AttributeError("libcloudphxx.common.%s is not provided by the B200 package" % name)
    return v


for _c in ("R_d", "R_v", "c_pd", "c_pv", "c_pw", "g", "p_1000", "eps", "rho_stp", "rho_w", "T_tri", "p_tri", "l_tri"):
    globals()[_c] = _call(_c)


def th_dry2std(th_dry, r): return _call("th_dry2std", th_dry, r)
def th_std2dry(th_std, r): return _call("th_std2dry", th_std, r)
def exner(p): return _call("exner", p)
def p_v(p, r): return _call("p_v", p, r)
def p_vs(T): return _call("p_vs", T)
def p_vs_tet(T): return _call("p_vs_tet", T)
def r_vs(T, p): return _call("r_vs", T, p)
def l_v(T): return _call("l_v", T)
def T(th, rhod): return _call("T", th, rhod)
def p(rhod, r, T): return _call("p", rhod, r, T)
def visc(T): return _call("visc", T)
def rw3_cr(rd3, kappa, T): return _call("rw3_cr", rd3, kappa, T)
def S_cr(rd3, kappa, T): return _call("S_cr", rd3, kappa, T)
def p_hydro(z, th_0, r_0, z_0, p_0): return _call("p_hydro", z, th_0, r_0, z_0, p_0)
def rhod(p, th_std, r_v): return _call("rhod", p, th_std, r_v)
