"""CPU restatement of libcloudph++'s Lagrangian super-droplet step (serial back-end), in numpy / plain Python.

TEST INFRASTRUCTURE - not part of the product.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import this module; libcloudphxx_b200/ never does.

Parity status: PINNED.  tests/test_cpu_oracle.py runs this port beside the reference's own serial back-end built from the
unmodified sources (oracle/_ref, see oracle/build_ref.py) on the same seeded inputs and requires bit-identical super-droplet
state (multiplicities, radii, positions) and per-cell fields; tests/golden/ holds vectors generated from that reference
build (tools/make_golden.py) so the pin also holds where /root/reference is absent.  Scalar libm calls go through
Python's math module (the same glibc the reference's CPU back-end uses), numpy is used only for + - * / sqrt and for
integer work, so that agreement can be exact.

Covered (each pinned bit for bit by a test in tests/test_cpu_oracle.py): every initialisation flavour; sort / shuffle / count;
coalescence with the Golovin, geometric (with and without multiplier), Long and tabulated-efficiency kernels incl. kappa mixing;
per-cell and per-particle (mixing, no mixing, adaptive, activation sub-stepping) condensation with TOMS 748, the four RH
formulae, dry theta with diagnosed pressure or standard theta with prescribed pressure; the five fall-speed formulae; implicit,
Euler and predictor-corrector advection, sedimentation, subsidence; periodic / open walls and the puddle; removal and recycling;
selectors, moments, SD concentration, precipitation flux, largest radius.  0-D, 2-D and 3-D.
x-slab decomposition and migration (SlabParticles) are restated too but UNPINNED: no reference run of them is possible here (multi_CUDA
needs a GPU, MPI is not built); the tests check what the reference's own distributed test demands instead.
Not covered: chemistry, ice, turbulence, sources / relaxation (outside the hot path).

Every function cites the reference code it restates (paths relative to the reference repository root).
"""
import math

import numpy as np

# ---- constants: include/libcloudph++/common/{moist_air.hpp:26-45,104,110; const_cp.hpp:27-31; earth.hpp:17-22} ---------
c_pd, c_pv, c_pw = 1005.0, 1850.0, 4218.0
M_d = 0.02897
M_v = 1 * 1e-3 + 17 * 1e-3
eps = M_v / M_d
kaBoNA = 8.3144621
R_d, R_v = kaBoNA / M_d, kaBoNA / M_v
rho_w = 1e3
D_0, K_0 = 2.26e-5, 2.4e-2
p_1000 = 100000.0
p_tri, T_tri, l_tri = 611.73, 273.16, 2.5e6
g_acc = 9.81
p_stp, T_stp = 101325.0, 273.15 + 15
rho_stp = p_stp / T_stp / R_d
PI = math.pi

N_ITER = 100                     # src/detail/config.hpp:17
EPS_TOL = max(math.ldexp(1.0, 1 - 16), 4 * np.finfo(np.float64).eps)   # eps_tolerance(sizeof(double)*8/4): toms748.hpp:262-286
VT0_N_BIN, VT0_LN_R_MIN, VT0_LN_R_MAX = 10000, math.log(5e-7), math.log(3e-3)   # config.hpp:27-44


# ---- random stream: src/detail/urand.hpp:19-88 (one std::mt19937 per object; libstdc++ distributions) -------------------
class HostRNG:
    def __init__(self, seed):
        self.rs = np.random.RandomState(seed)          # init_genrand(seed) == std::mt19937(seed)

    def reseed(self, seed):
        self.rs = np.random.RandomState(seed)

    def raw(self, n):
        return self.rs.randint(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint64)

    def u01(self, n):
        # std::uniform_real_distribution<double> = generate_canonical<double,53>: two 32-bit draws, low word first
        r = self.raw(2 * n)
        v = (r[0::2].astype(np.float64) + r[1::2].astype(np.float64) * 4294967296.0) / 18446744073709551616.0
        return np.where(v >= 1.0, np.nextafter(1.0, 0.0), v)

    def un(self, n):
        return self.raw(n).astype(np.uint32)           # std::uniform_int_distribution<unsigned>(0, UINT_MAX)


# ---- per-cell thermodynamics: src/impl/housekeeping/particles_impl_hskpng_Tpr.ipp:219-305 ------------------------------
def T_of_th_dry(th, rhod):                              # common/theta_dry.hpp:24-35
    return math.pow(th * math.pow(rhod * R_d / p_1000, R_d / c_pd), c_pd / (c_pd - R_d))


def p_vs_cc(T):                                         # common/const_cp.hpp:34-43
    return p_tri * math.exp((l_tri + (c_pw - c_pv) * T_tri) / R_v * (1.0 / T_tri - 1.0 / T) - (c_pw - c_pv) / R_v * math.log(T / T_tri))


def RH_pv_cc(p, rv, T):                                 # hskpng_Tpr.ipp:71-78, common/moist_air.hpp:88-95
    return (p * rv / (rv + eps)) / p_vs_cc(T)


def r_vs_cc(T, p):                                      # common/const_cp.hpp:57-63
    return eps / (p / p_vs_cc(T) - 1)


def p_vs_tet(T):                                        # common/tetens.hpp:15-24
    Tc = T - 273.15
    return 6.1078e2 * math.exp((17.27 * Tc) / (Tc + 237.3))


def r_vs_tet(T, p):                                     # common/tetens.hpp:26-35
    Tc = T - 273.15
    return 380. / (p * math.exp(-17.2693882 * Tc / (T - 35.86)) - 610.9)


def RH_of(formula, p, rv, T):                           # hskpng_Tpr.ipp:71-103,141-161
    if formula == "pv_cc":
        return RH_pv_cc(p, rv, T)
    if formula == "rv_cc":
        return rv / r_vs_cc(T, p)
    if formula == "pv_tet":
        return (p * rv / (rv + eps)) / p_vs_tet(T)
    if formula == "rv_tet":
        return rv / r_vs_tet(T, p)
    raise ValueError(formula)


def visc(T):                                            # common/vterm.hpp:22-31
    q = T / T_tri
    return (1.72 * 1e-5) * (393.0 / (T + 120.0)) * (q * math.sqrt(q))


def l_v(T):                                             # common/const_cp.hpp:82-87
    return l_tri + (c_pv - c_pw) * (T - T_tri)


def kelvin_A(T):                                        # common/kelvin_term.hpp:23-41
    return 2.0 * (0.07275 * (1.0 - 0.002 * (T - 291.0))) / R_v / T / rho_w


def a_w(rw3, rd3, kappa):                               # common/kappa_koehler.hpp:46-54
    return (rw3 - rd3) / (rw3 - rd3 * (1.0 - kappa))


# ---- TOMS 748: include/libcloudph++/common/detail/toms748.hpp:291-454 ----------------------------------------------------
DBL_EPS, DBL_MIN, DBL_MAX = np.finfo(np.float64).eps, np.finfo(np.float64).tiny, np.finfo(np.float64).max


def _tol(a, b):
    return abs(a - b) <= EPS_TOL * min(abs(a), abs(b))


def _safe_div(num, den, r):
    if abs(den) < 1 and abs(den * DBL_MAX) <= abs(num):
        return r
    return num / den


def _secant(a, b, fa, fb):
    tol = DBL_EPS * 5
    c = a - (fa / (fb - fa)) * (b - a)
    if c <= a + abs(a) * tol or c >= b - abs(b) * tol:
        return (a + b) / 2
    return c


def _quadratic(a, b, d, fa, fb, fd, count):
    B = _safe_div(fb - fa, b - a, DBL_MAX)
    A = _safe_div(fd - fb, d - b, DBL_MAX)
    A = _safe_div(A - B, d - a, 0.0)
    if A == 0:
        return _secant(a, b, fa, fb)
    c = a if math.copysign(1.0, A * fa) > 0 else b
    for _ in range(count):
        c -= _safe_div(fa + (B + A * (c - b)) * (c - a), B + A * (2 * c - a - b), 1 + c - a)
    if c <= a or c >= b:
        c = _secant(a, b, fa, fb)
    return c


def _cubic(a, b, d, e, fa, fb, fd, fe):
    with np.errstate(all="ignore"):
        q11 = np.float64(d - e) * fd / np.float64(fe - fd)
        q21 = np.float64(b - d) * fb / np.float64(fd - fb)
        q31 = np.float64(a - b) * fa / np.float64(fb - fa)
        d21 = np.float64(b - d) * fd / np.float64(fd - fb)
        d31 = np.float64(a - b) * fb / np.float64(fb - fa)
        q22 = (d21 - q11) * fb / np.float64(fe - fb)
        q32 = (d31 - q21) * fa / np.float64(fd - fa)
        d32 = (d31 - q21) * fd / np.float64(fd - fa)
        q33 = (d32 - q22) * fa / np.float64(fe - fa)
        c = float(q31 + q32 + q33 + a)
    if not (c > a and c < b):
        c = _quadratic(a, b, d, fa, fb, fd, 3)
    return c


def _prof(fa, fb, fd, fe):
    m = DBL_MIN * 32
    return (abs(fa - fb) < m or abs(fa - fd) < m or abs(fa - fe) < m or abs(fb - fd) < m or abs(fb - fe) < m or abs(fd - fe) < m)


def toms748(f, a, b, fa, fb, max_iter=N_ITER):
    count = max_iter
    st = dict(a=a, b=b, fa=fa, fb=fb, d=0.0, fd=0.0)

    def bracket(c):
        a, b = st["a"], st["b"]
        tol = DBL_EPS * 2
        if (b - a) < 2 * tol * a:
            c = a + (b - a) / 2
        elif c <= a + abs(a) * tol:
            c = a + abs(a) * tol
        elif c >= b - abs(b) * tol:
            c = b - abs(a) * tol
        fc = f(c)
        if fc == 0:
            st.update(a=c, fa=0.0, d=0.0, fd=0.0)
            return
        if math.copysign(1.0, st["fa"] * fc) < 0:
            st.update(d=st["b"], fd=st["fb"], b=c, fb=fc)
        else:
            st.update(d=st["a"], fd=st["fa"], a=c, fa=fc)

    def done():
        if st["fa"] == 0:
            st["b"] = st["a"]
        elif st["fb"] == 0:
            st["a"] = st["b"]
        return (st["a"] + st["b"]) / 2

    if _tol(a, b) or fa == 0 or fb == 0:
        return done()
    e = fe = 1e5
    st["fd"] = 1e5
    if st["fa"] != 0:
        bracket(_secant(st["a"], st["b"], st["fa"], st["fb"]))
        count -= 1
        if count and st["fa"] != 0 and not _tol(st["a"], st["b"]):
            c = _quadratic(st["a"], st["b"], st["d"], st["fa"], st["fb"], st["fd"], 2)
            e, fe = st["d"], st["fd"]
            bracket(c)
            count -= 1
    while count and st["fa"] != 0 and not _tol(st["a"], st["b"]):
        a0, b0 = st["a"], st["b"]
        if _prof(st["fa"], st["fb"], st["fd"], fe):
            c = _quadratic(st["a"], st["b"], st["d"], st["fa"], st["fb"], st["fd"], 2)
        else:
            c = _cubic(st["a"], st["b"], st["d"], e, st["fa"], st["fb"], st["fd"], fe)
        e, fe = st["d"], st["fd"]
        bracket(c)
        count -= 1
        if count == 0 or st["fa"] == 0 or _tol(st["a"], st["b"]):
            break
        if _prof(st["fa"], st["fb"], st["fd"], fe):
            c = _quadratic(st["a"], st["b"], st["d"], st["fa"], st["fb"], st["fd"], 3)
        else:
            c = _cubic(st["a"], st["b"], st["d"], e, st["fa"], st["fb"], st["fd"], fe)
        bracket(c)
        count -= 1
        if count == 0 or st["fa"] == 0 or _tol(st["a"], st["b"]):
            break
        if abs(st["fa"]) < abs(st["fb"]):
            u, fu = st["a"], st["fa"]
        else:
            u, fu = st["b"], st["fb"]
        c = u - 2 * (fu / (st["fb"] - st["fa"])) * (st["b"] - st["a"])
        if abs(c - u) > (st["b"] - st["a"]) / 2:
            c = st["a"] + (st["b"] - st["a"]) / 2
        e, fe = st["d"], st["fd"]
        bracket(c)
        count -= 1
        if count == 0 or st["fa"] == 0 or _tol(st["a"], st["b"]):
            break
        if (st["b"] - st["a"]) < 0.5 * (b0 - a0):
            continue
        e, fe = st["d"], st["fd"]
        bracket(st["a"] + (st["b"] - st["a"]) / 2)
        count -= 1
    return done()


# ---- equilibrium wet radius: common/kappa_koehler.hpp:58-146, src/impl/initialization/particles_impl_init_wet.ipp:18-74 --
def rw3_eq(rd3, kappa, RH, T):
    if kappa == 0:
        return rd3
    A = kelvin_A(T)
    f = lambda rw3: RH - a_w(rw3, rd3, kappa) * math.exp(A / math.cbrt(rw3))
    lo, hi = rd3, rd3 * (1 - RH * (1 - kappa)) / (1 - RH)
    return toms748(f, lo, hi, f(lo), f(hi))


def brent_find_minima(f, lo, hi, max_iter=N_ITER):
    """Brent's minimiser as Boost.Math implements it (boost/math/tools/minima.hpp, bits capped at half the mantissa; the golden
    section constant is a single-precision literal there) - called by init_dist_analysis.ipp:95.  The reference build of the
    oracle uses oracle/boost_shim's restatement of the same routine: parity with real Boost is unpinned (no Boost here)."""
    tol = math.ldexp(1.0, 1 - 53 // 2)
    golden = float(np.float32(0.3819660))
    x = w = v = hi
    fw = fv = fx = f(x)
    delta2 = delta = 0.0
    for _ in range(max_iter):
        mid = (lo + hi) / 2
        fract1 = tol * abs(x) + tol / 4
        fract2 = 2 * fract1
        if abs(x - mid) <= (fract2 - (hi - lo) / 2):
            break
        parabolic = False
        if abs(delta2) > fract1:
            r = (x - w) * (fx - fv)
            q = (x - v) * (fx - fw)
            pnum = (x - v) * q - (x - w) * r
            q = 2 * (q - r)
            if q > 0:
                pnum = -pnum
            q = abs(q)
            td = delta2
            delta2 = delta
            if not (abs(pnum) >= abs(q * td / 2) or pnum <= q * (lo - x) or pnum >= q * (hi - x)):
                delta = pnum / q
                u = x + delta
                if (u - lo) < fract2 or (hi - u) < fract2:
                    delta = -abs(fract1) if (mid - x) < 0 else abs(fract1)
                parabolic = True
        if not parabolic:
            delta2 = (lo - x) if x >= mid else (hi - x)
            delta = golden * delta2
        u = (x + delta) if abs(delta) >= fract1 else ((x + abs(fract1)) if delta > 0 else (x - abs(fract1)))
        fu = f(u)
        if fu <= fx:
            if u >= x:
                lo = x
            else:
                hi = x
            v, w, x = w, x, u
            fv, fw, fx = fw, fx, fu
        else:
            if u < x:
                lo = u
            else:
                hi = u
            if fu <= fw or w == x:
                v, w = w, u
                fv, fw = fw, fu
            elif fu <= fv or v == x or v == w:
                v, fv = u, fu
    return x, fx


def rw3_cr(rd3, kappa, T):                              # common/kappa_koehler.hpp:90-119,153-169 (maximum of the Koehler curve)
    A = kelvin_A(T)
    f = lambda rw3: (A * (rd3 - rw3) * ((kappa - 1) * rd3 + rw3) + 3 * kappa * rd3 * rw3 * math.cbrt(rw3))
    lo, hi = 1e0 * rd3, 1e8 * rd3
    return toms748(f, lo, hi, f(lo), f(hi))


# ---- condensational growth: src/impl/condensation/common/particles_impl_cond_common.ipp:79-338 --------------------------
def drw2_dt(rw2, rhod, rv, T, p, RH_eff, eta, rd3, kpa, vt, lam_D, lam_K):
    rw = math.sqrt(rw2)
    rw3 = rw * rw * rw
    Re = vt * (2.0 * rw) * rhod / eta
    Sc = eta / rhod / D_0
    Pr = c_pd * eta / K_0

    def beta(Kn):                                       # common/transition_regime.hpp:15-19
        return (1 + Kn) / (1 + 1.71 * Kn + 1.33 * Kn * Kn)

    def Nu(P):                                          # common/ventil.hpp:29-45
        pw = math.pow(Re, .077) if Re >= 0 else float("nan")
        return 1.0 + math.cbrt(1.0 + Re * P) * (pw if 1.0 < pw else 1.0)
    D = D_0 * beta(lam_D / rw) * (Nu(Sc) / 2)
    K = K_0 * beta(lam_K / rw) * (Nu(Pr) / 2)
    lv = l_v(T)
    rho_v = rhod * rv
    klv = math.exp(kelvin_A(T) / rw)
    rdrdt = (1.0 - a_w(rw3, rd3, kpa) * klv / RH_eff) / rho_w / (1.0 / D / rho_v + lv / K / RH_eff / T * (lv / R_v / T - 1.0))   # common/maxwell-mason.hpp:33-45
    return 2.0 * rdrdt


def advance_rw2(rw2_old, dt, RH_max, rhod, rv, T, p, RH, eta, rd3, kpa, vt, lam_D, lam_K, cond_mlt=2.0):
    if rw2_old <= 0:
        return rw2_old
    RH_eff = RH_max if RH > RH_max else RH
    g = lambda x: drw2_dt(x, rhod, rv, T, p, RH_eff, eta, rd3, kpa, vt, lam_D, lam_K)
    f = lambda x: (rw2_old + dt * g(x) - x)
    drw2 = dt * g(rw2_old)
    if drw2 == 0:
        return rw2_old
    rd = math.cbrt(rd3)
    rd2 = rd * rd
    a = max(rd2, rw2_old + min(0.0, cond_mlt * drw2))
    b = rw2_old + max(0.0, cond_mlt * drw2)
    if a == b:
        return rw2_old
    if drw2 > 0:
        fa, fb = drw2, f(b)
    else:
        fa, fb = f(a), drw2
    if fa * fb > 0:
        new = rw2_old + drw2
    else:
        new = toms748(f, a, b, fa, fb)
    return rd2 if new < rd2 else new


# ---- terminal velocity (beard77fast): common/vterm.hpp:112-164, hskpng_vterm.ipp:14-36,185-342, init_vterm.ipp:36-59 --------
def vt_beard77_v0(r):
    m_s = [0.105035e2, 0.108750e1, -0.133245, -0.659969e-2]
    m_l = [0.65639e1, -0.10391e1, -0.14001e1, -0.82736e0, -0.34277e0, -0.83072e-1, -0.10583e-1, -0.54208e-3]
    x = math.log(2 * 100 * r)
    y = 0.0
    for i, m in enumerate(m_s if r <= 20e-6 else m_l):
        y += m * math.pow(x, float(i))
    return math.exp(y) / 100.


def vt0_table():
    dlnr = (VT0_LN_R_MAX - VT0_LN_R_MIN) / VT0_N_BIN
    return np.array([vt_beard77_v0(math.exp(VT0_LN_R_MIN + (it + 0.5) * dlnr)) for it in range(VT0_N_BIN)])


def vt_beard77_fact(r, p, rhoa, eta):
    eta_0 = 1.818e-5
    if r <= 20e-6:
        l_0 = 6.62e-8
        l = l_0 * (eta / eta_0) * math.sqrt(p_stp / p * rho_stp / rhoa)
        return (eta_0 / eta) * (1 + 1.255 * (l / r)) / (1 + 1.255 * (l_0 / r))
    eps_s = (eta_0 / eta) - 1
    eps_c = math.sqrt(rho_stp / rhoa) - 1
    return 1.104 * eps_s + ((1.058 * eps_c - 1.104 * eps_s) * (5.52 + math.log(2 * 100 * r)) / 5.01) + 1


def vt_beard76(r, T, p, rhoa, eta):                     # common/vterm.hpp:168-221
    if r <= 9.5e-6:
        l = 6.62e-8 * (eta / 1.818e-5) * (p_stp / p) * math.sqrt(T / 293.15)
        C_ac = 1. + 1.255 * l / r
        return (rho_w - rhoa) * g_acc / (4.5 * eta) * C_ac * r * r
    if r <= 5.035e-4:
        b = (-0.318657e1, 0.992696, -0.153193e-2, -0.987059e-3, -0.578878e-3, 0.855176e-4, -0.327815e-5)
        l = 6.62e-8 * (eta / 1.818e-5) * (p_stp / p) * math.sqrt(T / 293.15)
        C_ac = 1. + 1.255 * l / r
        log_N_Da = math.log((32. / 3.) * r * r * r * rhoa * (rho_w - rhoa) * g_acc / eta / eta)
        Y = 0.
        for i in range(7):
            Y = Y + b[i] * math.pow(log_N_Da, float(i))
        N_Re = C_ac * math.exp(Y)
        return eta * N_Re / rhoa / 2. / r
    b = (-0.500015e1, 0.523778e1, -0.204914e1, 0.475294, -0.542819e-1, 0.238449e-2)
    sg = 0.07275 * (1. - 0.002 * (T - 291.))            # common/kelvin_term.hpp:23-32
    Bo = (16. / 3.) * r * r * (rho_w - rhoa) * g_acc / sg
    N_p = sg * sg * sg * rhoa * rhoa / eta / eta / eta / eta / g_acc / (rho_w - rhoa)
    X = math.log(Bo * math.pow(N_p, 1. / 6.))
    Y = 0.
    for i in range(6):
        Y = Y + b[i] * math.pow(X, float(i))
    N_Re = math.pow(N_p, 1. / 6.) * math.exp(Y)
    return eta * N_Re / rhoa / 2. / r


def vt_khvorostyanov(r, rhoa, eta, spherical):          # common/vterm.hpp:38-105
    X = (32. / 3) * (rho_w - rhoa) / rhoa * g_acc * r * r * r / eta / eta * rhoa * rhoa
    b = (.0902 / 2) * math.sqrt(X) / ((math.sqrt(1. + .0902 * math.sqrt(X)) - 1.) * (math.sqrt(1. + .0902 * math.sqrt(X))))
    pow_hlpr = math.sqrt(1. + .0902 * math.sqrt(X)) - 1.
    a = (9.06 * 9.06 / 4) * pow_hlpr * pow_hlpr / math.pow(X, b)
    if spherical:
        Av = a * math.pow(eta / rhoa * 1e4, 1. - 2. * b) * math.pow((4. / 3) * rho_w / rhoa * g_acc * 1e2, b)
    else:
        lambda_half = 2.35e-3
        ksi = math.exp(-r / lambda_half) + (1. - math.exp(-r / lambda_half)) / (1. + r / lambda_half)
        alfa = PI / 6. * rho_w * ksi
        Av = a * math.pow(eta / rhoa * 1e4, 1. - 2. * b) * math.pow(2.546479 * alfa / rhoa * g_acc * 1e2, b)
    Bv = 3. * b - 1.
    return (Av * math.pow((2 * 1e2) * r, Bv)) / 1e2


def vt_any(kind, rw2, T, p, rhod, eta, table):          # hskpng_vterm.ipp:38-121
    if kind == "beard77fast":
        return vt_beard77fast(rw2, p, rhod, eta, table)
    r = math.sqrt(rw2)
    if kind == "beard76":
        return vt_beard76(r, T, p, rhod, eta)
    if kind == "beard77":
        return vt_beard77_fact(r, p, rhod, eta) * vt_beard77_v0(r)
    if kind == "khvorostyanov_spherical":
        return vt_khvorostyanov(r, rhod, eta, True)
    if kind == "khvorostyanov_nonspherical":
        return vt_khvorostyanov(r, rhod, eta, False)
    return 0.0                                           # vt_t::undefined


def vt_beard77fast(rw2, p, rhod, eta, table):
    dlnr = (VT0_LN_R_MAX - VT0_LN_R_MIN) / VT0_N_BIN
    lnr = .5 * math.log(rw2)
    b = 0 if lnr <= VT0_LN_R_MIN else (VT0_N_BIN - 1 if lnr >= VT0_LN_R_MAX else int((lnr - VT0_LN_R_MIN) / dlnr))
    return vt_beard77_fact(math.sqrt(rw2), p, rhod, eta) * table[b]


# ---- collision kernels: src/detail/kernels.hpp:40-176, kernel_interpolation.hpp:9-64, kernel_utils.hpp:12-29 ---------------
def kernel_index(R):
    return int(R) if R <= 100. else int(100 + (R - 100.) / 10.)


def kernel_vector_index(i, j):
    return int(0.5 * i * (i + 1) + j) if i >= j else int(0.5 * j * (j + 1) + i)


def interpolated_efficiency(eff, r_max, r1, r2):
    r1 *= 1e6
    r2 *= 1e6
    if r1 >= r_max:
        r1 = r_max - 1e-6
    if r2 >= r_max:
        r2 = r_max - 1e-6
    if r1 >= 100.:
        x0, dx = int(math.floor(r1 / 10.) * 10), 10
    else:
        x0, dx = int(math.floor(r1)), 1
    if r2 >= 100.:
        x2, dy = int(math.floor(r2 / 10.) * 10), 10
    else:
        x2, dy = int(math.floor(r2)), 1
    x1, x3 = x0 + dx, x2 + dy
    iv = [kernel_vector_index(kernel_index(a), kernel_index(b)) for a, b in ((x0, x2), (x1, x2), (x0, x3), (x1, x3))]
    w = [r1 - x0, x1 - r1, r2 - x2, x3 - r2]
    return (eff[iv[0]] * w[1] * w[3] + eff[iv[1]] * w[0] * w[3] + eff[iv[2]] * w[1] * w[2] + eff[iv[3]] * w[0] * w[2]) / dx / dy


def coal_kernel(kind, params, n_a, n_b, rw2_a, rw2_b, vt_a, vt_b):
    nmax = float(max(n_a, n_b))
    if kind == "golovin":
        return PI * 4. / 3. * params["b"] * nmax * (rw2_a * math.sqrt(rw2_a) + rw2_b * math.sqrt(rw2_b))
    geo = PI * nmax * abs(vt_a - vt_b) * (rw2_a + rw2_b + 2. * math.sqrt(rw2_a * rw2_b))
    if kind == "geometric":
        return geo * params["mult"] if "mult" in params else geo
    if kind == "efficiencies":
        return interpolated_efficiency(params["eff"], params["r_max"], math.sqrt(rw2_a), math.sqrt(rw2_b)) * geo
    if kind == "long":                                   # kernels.hpp:144-176
        res = geo
        r_L = max(math.sqrt(rw2_a), math.sqrt(rw2_b))
        if r_L < 50.e-6:
            r_s = min(math.sqrt(rw2_a), math.sqrt(rw2_b))
            if r_s <= 3e-6:
                res = 0.
            else:
                res *= 4.5e8 * r_L * r_L * (1. - 3e-6 / r_s)
        return res
    raise ValueError(kind)


# ---- the particle system ----------------------------------------------------------------------------------------------------
class Particles:
    """0-D / 2-D / 3-D box, sd_conc initialisation, per-cell and per-particle (mixing / no mixing / adaptive, activation
    sub-stepping with rc2) condensation sub-stepping, SDM coalescence, implicit / Euler / predictor-corrector advection, sedimentation, subsidence, periodic or open side
    walls, open top / bottom, removal or recycling of used-up SDs.  Call order as the reference (src/particles_step.ipp)."""

    def __init__(self, nx=0, ny=0, nz=0, dx=1., dy=1., dz=1., dt=1., x0=0., y0=0., z0=0., x1=1., y1=1., z1=1., sd_conc=0, n_sd_max=0,
                 sstp_cond=1, sstp_coal=1, kernel=None, kernel_params=None, vt="beard77fast", adve_scheme="implicit",
                 dry_distros=(), RH_max_init=.95, rng_seed=44, sedi_switch=True, coal_switch=True,
                 exact_sstp_cond=False, sstp_cond_mix=True, adaptive_sstp_cond=False, sstp_cond_act=1,
                 sstp_cond_adapt_drw2_eps=1e-4, sstp_cond_adapt_drw2_max=4., rc2_T=10.,
                 sd_const_multi=0, sd_conc_large_tail=False, dry_sizes=(), aerosol_independent_of_rhod=False, aerosol_conc_factor=(),
                 rd_min=-1., rd_max=-1., RH_formula="pv_cc", open_side_walls=False, w_LS=None, th_dry=True, const_p=False):
        self.nx, self.ny, self.nz = nx, ny, nz
        self.dx, self.dy, self.dz, self.dt = dx, dy, dz, dt
        self.x0, self.y0, self.z0, self.x1, self.y1, self.z1 = x0, y0, z0, x1, y1, z1
        self.n_dims = (nx > 0) + (ny > 0) + (nz > 0)
        self.n_cell = max(1, nx) * max(1, ny) * max(1, nz)
        self.sd_conc, self.n_sd_max = sd_conc, n_sd_max
        self.sd_const_multi, self.sd_conc_large_tail = sd_const_multi, sd_conc_large_tail
        self.dry_sizes = list(dry_sizes)               # [(kappa, {radius: (STP concentration, SDs per cell)})] in ascending kappa
        self.aerosol_independent_of_rhod, self.aerosol_conc_factor = aerosol_independent_of_rhod, list(aerosol_conc_factor)
        self.rd_min, self.rd_max = rd_min, rd_max
        self.sstp_cond, self.sstp_coal = sstp_cond, sstp_coal
        self.kernel, self.kernel_params = kernel, kernel_params or {}
        self.vt_kind, self.adve_scheme = vt, adve_scheme
        self.dry_distros = list(dry_distros)          # [(kappa, callable n(ln r))], iterated in ascending kappa like std::map
        self.RH_max_init = RH_max_init
        self.RH_formula = RH_formula
        self.open_side_walls = open_side_walls
        self.distmem = False                             # set by SlabParticles: x faces shared with neighbouring slabs
        self.th_dry, self.const_p = th_dry, const_p     # opts_init.th_dry / const_p: what "th" means and whether p is prescribed
        if const_p and exact_sstp_cond:
            raise NotImplementedError("the restatement covers const_p with per-cell sub-stepping only")
        self.w_LS = None if w_LS is None else np.array(w_LS, dtype=np.float64)     # large-scale subsidence velocity per level (opts_init.w_LS)
        self.rng_seed = rng_seed
        self.rng = HostRNG(rng_seed)
        self.puddle = dict(liquid_volume=0., dry_volume=0., liquid_number=0., particle_number=0.)
        self.table = vt0_table() if vt == "beard77fast" else None
        # per-particle condensation sub-stepping (opts_init.hpp:96-106)
        self.exact_sstp_cond, self.sstp_cond_mix, self.adaptive_sstp_cond = exact_sstp_cond, sstp_cond_mix, adaptive_sstp_cond
        self.sstp_cond_act, self.drw2_eps, self.drw2_max, self.rc2_T = sstp_cond_act, sstp_cond_adapt_drw2_eps, sstp_cond_adapt_drw2_max, rc2_T
        self.allow_sstp_cond = sstp_cond > 1 or sstp_cond_act > 1                       # particles_impl.ipp:376

    # -- grid helpers: src/impl/initialization/particles_impl_init_grid.ipp:13-155 ------------------------------------------
    def unravel(self, c):
        nz1, ny1 = max(1, self.nz), max(1, self.ny)
        return (c // nz1) // ny1, (c // nz1) % ny1, c % nz1

    def cell_volumes(self):
        i, j, k = self.unravel(np.arange(self.n_cell))
        ext = lambda idx, d, a, b: np.minimum((idx + 1) * d, b) - np.maximum(idx * d, a)
        return np.maximum(0., ext(i, self.dx, self.x0, self.x1) * ext(j, self.dy, self.y0, self.y1) * ext(k, self.dz, self.z0, self.z1))

    def hskpng_Tpr(self):                               # hskpng_Tpr.ipp:219-305 (th_dry, variable pressure, four RH formulae)
        for c in range(self.n_cell):
            if self.th_dry:
                T = T_of_th_dry(self.th[c], self.rhod[c])
            else:                                        # "standard" potential temperature: T = th * exner(p), common/theta_std.hpp:36-41
                T = self.th[c] * math.pow(self.p[c] / p_1000, R_d / c_pd)
            p = self.p[c] if self.const_p else self.rhod[c] * (R_d + self.rv[c] * R_v) * T
            self.T[c], self.p[c] = T, p
            self.RH[c] = RH_of(self.RH_formula, p, self.rv[c], T)
            self.eta[c] = visc(T)
        if self.n_dims == 0:
            self.dv = 1.0 / self.rhod

    def hskpng_mfp(self):                               # hskpng_mfp.ipp:42-52, common/mean_free_path.hpp:16-51
        self.lam_D = np.array([2.0 * D_0 / math.sqrt(2.0 * (R_v * T)) for T in self.T])
        self.lam_K = np.array([.8 * (K_0 * T / p) / math.sqrt(2.0 * (R_d * T)) for T, p in zip(self.T, self.p)])

    def hskpng_vterm(self, only_invalid):               # hskpng_vterm.ipp:185-342
        for s in range(self.n_part):
            if self.rw2[s] > 0 and (not only_invalid or self.vt[s] == -1.0):
                c = self.ijk[s]
                self.vt[s] = vt_any(self.vt_kind, self.rw2[s], self.T[c], self.p[c], self.rhod[c], self.eta[c], self.table)

    # -- init: src/particles_init.ipp:16-131 and src/impl/initialization/* ----------------------------------------------------
    def init(self, th, rv, rhod, Cx=None, Cy=None, Cz=None, p=None):
        C = self.n_cell
        self.th, self.rv, self.rhod = (np.array(a, dtype=np.float64).reshape(C).copy() for a in (th, rv, rhod))
        self.Cx, self.Cy, self.Cz = Cx, Cy, Cz
        self.T, self.p, self.RH, self.eta = (np.zeros(C) for _ in range(4))
        if self.const_p:                                 # particles_init.ipp:39-40: the pressure profile comes from the caller and stays
            self.p = np.array(p, dtype=np.float64).reshape(C).copy()
        self.dv = self.cell_volumes() if self.n_dims else np.zeros(C)
        self.hskpng_Tpr()
        parts = {k: [] for k in ("n", "rd3", "rw2", "kpa", "x", "y", "z", "ijk")}

        def finalize(ijk, n, rd3, kappa):                # init_SD_with_distros.ipp:62-107: kappa, init_wet, init_xyz
            n_new = ijk.size
            rw2 = np.empty(n_new)
            for q in range(n_new):
                RH = min(self.RH[ijk[q]], self.RH_max_init)
                rw2[q] = math.pow(rw3_eq(rd3[q], kappa, RH, self.T[ijk[q]]), 2. / 3)     # init_wet.ipp:18-40
            ii, jj, kk = self.unravel(ijk)
            pos = {}
            for name, nn, idx, a, b, d in (("x", self.nx, ii, self.x0, self.x1, self.dx), ("y", self.ny, jj, self.y0, self.y1, self.dy),
                                           ("z", self.nz, kk, self.z0, self.z1, self.dz)):
                if nn == 0:
                    pos[name] = np.zeros(0)
                    continue
                u = self.rng.u01(n_new)                                           # init_xyz.ipp:49-73
                pos[name] = u * np.minimum(b, (idx + 1) * d) + (1. - u) * np.maximum(a, idx * d)
            for k, v in (("n", n), ("rd3", rd3), ("rw2", rw2), ("kpa", np.full(n_new, kappa)), ("x", pos["x"]), ("y", pos["y"]), ("z", pos["z"]), ("ijk", ijk)):
                parts[k].append(v)

        def conc_to_number(conc):                        # init_count_num.ipp:35-64
            arr = np.full(C, conc) * self.dv
            if not self.aerosol_independent_of_rhod:
                arr = self.rhod / rho_stp * arr
            if len(self.aerosol_conc_factor):
                arr = arr * np.asarray(self.aerosol_conc_factor)[np.arange(C) % self.nz]
            return arr

        def const_multi(kappa, fun, multi, forced_min=None):      # init_SD_with_distros_const_multi.ipp / _tail.ipp
            lmin, lmax = self.dist_analysis_const_multi(fun)
            if forced_min is not None:
                lmin = forced_min
            assert lmin < lmax, "Distribution analysis error"
            bin_ = 1e-4                                           # config.hpp: bin_precision
            nb = int((lmax - lmin) / bin_)                        # detail::integrate, init_count_num.ipp:17-27
            integral = (fun(lmin) + fun(lmax)) / 2.
            for i in range(1, nb):
                integral += fun(lmin + i * bin_)
            integral = integral * bin_
            count = (conc_to_number(integral) / multi + 0.5).astype(np.int64)     # init_count_num_hlpr
            ijk = np.repeat(np.arange(C), count)
            n_pt = int((lmax - lmin) / bin_ + 1)                  # calc_CDF, init_dry_const_multi.ipp:25-45
            cdf = np.array([1 * fun(lmin + bin_ * i) for i in range(n_pt)])
            for i in range(1, n_pt):
                cdf[i] = cdf[i - 1] + cdf[i]
            cdf = cdf / cdf[-1]
            u01 = self.rng.u01(ijk.size)
            pos = np.searchsorted(cdf, u01, side="right").astype(np.float64)      # thrust::upper_bound
            rd3 = np.array([math.exp(3 * (lmin + p_ * bin_)) for p_ in pos])
            finalize(ijk, np.full(ijk.size, multi, dtype=np.uint64), rd3, kappa)

        ranges = [self.dist_analysis(fun) for _, fun in self.dry_distros] if self.sd_conc > 0 else []
        tot = sum(r[1] - r[0] for r in ranges)
        for idx_d, (kappa, fun) in enumerate(self.dry_distros):
            if self.sd_conc > 0:
                lmin, lmax, mult = ranges[idx_d]
                fraction = (lmax - lmin) / tot
                mult *= self.sd_conc // int(fraction * self.sd_conc + 0.5)          # integer division: init_SD_with_distros_sd_conc.ipp:28
                per_cell = int(fraction * self.sd_conc)
                n_new = per_cell * C
                ijk = np.repeat(np.arange(C), per_cell)                               # init_ijk.ipp:36-52
                u01 = self.rng.u01(n_new)                                             # init_dry_sd_conc.ipp:48
                s = np.arange(n_new)
                lnrd = lmin + ((s - ijk * per_cell) + u01) * (lmax - lmin) / float(per_cell)
                rd3 = np.array([math.exp(3 * v) for v in lnrd])
                n = np.empty(n_new, dtype=np.uint64)
                for q in range(n_new):
                    v = mult * fun(math.log(rd3[q]) / 3.)                             # init_n.ipp:48-137
                    if not self.aerosol_independent_of_rhod:
                        v = v * self.rhod[ijk[q]] / rho_stp
                    if len(self.aerosol_conc_factor):
                        v = v * self.aerosol_conc_factor[ijk[q] % self.nz]
                    if self.n_dims > 0:
                        v = v * self.dv[ijk[q]] / (self.dx * self.dy * self.dz)
                    n[q] = int(v + 0.5)
                finalize(ijk, n, rd3, kappa)
                if self.sd_conc_large_tail:
                    const_multi(kappa, fun, 1, forced_min=lmax)
            if self.sd_const_multi > 0:
                const_multi(kappa, fun, self.sd_const_multi)
        for (kappa, sizes) in self.dry_sizes:            # init_SD_with_sizes.ipp:16-76 (std::map order: ascending kappa, then radius)
            for radius, (conc, count) in sorted(sizes.items()):
                ijk = np.repeat(np.arange(C), count)
                number = conc_to_number(conc)
                n = (number[ijk] / count + .5).astype(np.uint64)                  # init_n.ipp:147-162
                finalize(ijk, n, np.full(ijk.size, radius * radius * radius), kappa)
        for k, v in parts.items():
            setattr(self, k, np.concatenate(v) if v else np.zeros(0))
        self.ijk = self.ijk.astype(np.int64)
        self.n_part = self.n.size
        self.vt = np.full(self.n_part, -1.0)
        self.hskpng_vterm(True)
        self.rc2 = np.full(self.n_part, -1.0)                                        # particles_impl.ipp:490 (detail::invalid)
        self.hskpng_rc2()
        self.sstp_save()
        self.sort(False)
        self.rng.reseed(self.rng_seed)

    def hskpng_rc2(self):                               # hskpng_rc2.ipp:13-32, particles_diag.ipp:41-62
        if self.sstp_cond_act == 1 or not self.allow_sstp_cond:
            return
        for s in range(self.n_part):
            if self.rc2[s] == -1.0:
                self.rc2[s] = math.pow(rw3_cr(self.rd3[s], self.kpa[s], self.rc2_T + 273.15), 2. / 3)

    def sstp_save(self):                                # condensation/common/sstp_save.ipp:13-29
        if self.exact_sstp_cond and self.allow_sstp_cond:
            self.pp = dict(rv=self.rv[self.ijk].copy(), th=self.th[self.ijk].copy(), rhod=self.rhod[self.ijk].copy())
        self.old = dict(rv=self.rv.copy(), th=self.th.copy(), rhod=self.rhod.copy())

    def dist_analysis_const_multi(self, fun):           # init_dist_analysis.ipp:78-123
        if self.rd_min >= 0 and self.rd_max >= 0:
            return math.log(self.rd_min), math.log(self.rd_max)
        lo, hi = math.log(1e-14), math.log(1e-3)
        x_max, f_max = brent_find_minima(lambda x: -1 * fun(x), lo, hi)
        bound = -f_max / 1e20                            # config.hpp: threshold
        g = lambda x: -bound + fun(x)
        return toms748(g, lo, x_max, g(lo), g(x_max)), toms748(g, x_max, hi, g(x_max), g(hi))

    def dist_analysis(self, fun):                       # init_dist_analysis.ipp:17-75 (automatic range detection)
        vol = self.dv[0] if self.n_dims == 0 else self.dx * self.dy * self.dz
        if self.rd_min >= 0 and self.rd_max >= 0:       # user-defined range
            return math.log(self.rd_min), math.log(self.rd_max), math.log(self.rd_max / self.rd_min) / self.sd_conc * 1.0 * vol
        rd_min, rd_max = 1e-14, 1e-3
        while True:
            mult = math.log(rd_max / rd_min) / self.sd_conc * 1.0 * vol
            lmin, lmax = math.log(rd_min), math.log(rd_max)
            n_min, n_max = int(fun(lmin) * mult), int(fun(lmax) * mult)
            if n_min == 0:
                rd_min *= 1.01
            elif n_max == 0:
                rd_max /= 1.01
            else:
                return lmin, lmax, mult

    # -- sort / shuffle / count: hskpng_sort.ipp:15-70, hskpng_count.ipp:16-48 --------------------------------------------------
    def sort(self, shuffle):
        ids = np.arange(self.n_part)
        if shuffle:
            un = self.rng.un(self.n_part)
            ids = ids[np.argsort(un, kind="stable")]
        self.sorted_id = ids[np.argsort(self.ijk[ids], kind="stable")]
        self.sorted_ijk = self.ijk[self.sorted_id]
        self.count_ijk, self.count_num = np.unique(self.sorted_ijk, return_counts=True)

    # -- moments: particles_impl_moms.ipp:240-387 ---------------------------------------------------------------------------------
    def moment(self, attr, power, n_filtered=None, specific=True):
        w = self.n.astype(np.float64) if n_filtered is None else n_filtered
        out = np.zeros(self.n_cell)
        for pos in range(self.n_part):                   # sequential sum in sorted order, like the serial reduce_by_key
            s = self.sorted_id[pos]
            out[self.ijk[s]] += w[s] * math.pow(attr[s], power)
        if specific and self.n_dims > 0:
            out = out / self.dv / self.rhod
        return out

    # -- diagnostics: src/particles_diag.ipp:148-656 on top of the selectors / moments of particles_impl_moms.ipp:50-387 ---------
    # every diag_* below returns what outbuf() would hold afterwards: a dense per-cell array, zero where no selected SD lives
    def diag_all(self):
        self.n_filtered = self.n.astype(np.float64)

    def _moms_rng(self, lo, hi, vec, cons):              # moms_rng, range_filter: x >= min && x < max ? y : 0
        base = self.n_filtered if cons else self.n.astype(np.float64)
        self.n_filtered = np.where((vec >= lo) & (vec < hi), base, 0.)

    def diag_wet_rng(self, r_min, r_max, cons=False):
        self._moms_rng(math.pow(r_min, 2), math.pow(r_max, 2), self.rw2, cons)

    def diag_dry_rng(self, r_min, r_max, cons=False):
        self._moms_rng(math.pow(r_min, 3), math.pow(r_max, 3), self.rd3, cons)

    def diag_kappa_rng(self, k_min, k_max, cons=False):
        self._moms_rng(k_min, k_max, self.kpa, cons)

    def diag_rw_ge_rc(self):                            # particles_diag.ipp:384-408, moms_cmp
        rc2 = np.array([math.pow(rw3_cr(self.rd3[s], self.kpa[s], self.T[self.ijk[s]]), 2. / 3) for s in range(self.n_part)])
        self.n_filtered = np.where(self.rw2 >= rc2, self.n.astype(np.float64), 0.)

    def diag_RH_ge_Sc(self):                            # particles_diag.ipp:353-381, kappa_koehler.hpp:172-190
        def S_cr(rd3, kpa, T):
            rw3 = rw3_cr(rd3, kpa, T)
            return a_w(rw3, rd3, kpa) * math.exp(kelvin_A(T) / math.cbrt(rw3))
        d = np.array([self.RH[self.ijk[s]] - S_cr(self.rd3[s], self.kpa[s], self.T[self.ijk[s]]) for s in range(self.n_part)])
        self.n_filtered = np.where(d >= 0., self.n.astype(np.float64), 0.)

    def diag_wet_mom(self, k):
        return self.moment(self.rw2, k / 2., self.n_filtered)

    def diag_dry_mom(self, k):
        return self.moment(self.rd3, k / 3., self.n_filtered)

    def diag_kappa_mom(self, k):
        return self.moment(self.kpa, float(k), self.n_filtered)

    def diag_sd_conc(self):                             # particles_diag.ipp:199-220: number of selected SDs, not divided by anything
        out = np.zeros(self.n_cell)
        for pos in range(self.n_part):
            s = self.sorted_id[pos]
            out[self.ijk[s]] += 1. if self.n_filtered[s] > 0. else 0.
        return out

    def diag_precip_rate(self):                         # particles_diag.ipp:561-587: sum of n rw^3 vt, fall speeds refreshed first
        self.hskpng_vterm(False)
        flux = np.array([math.pow(self.rw2[s], 3. / 2) * self.vt[s] for s in range(self.n_part)])
        return self.moment(flux, 1., self.n_filtered, specific=False)

    def diag_max_rw(self):                              # particles_diag.ipp:609-642
        out = np.zeros(self.n_cell)
        for s in range(self.n_part):
            out[self.ijk[s]] = max(out[self.ijk[s]], math.sqrt(self.rw2[s]))
        return out

    # -- condensation: src/particles_step.ipp:161-336, percell/particles_impl_cond.ipp:13-139, update_th_rv.ipp:74-191 ----------
    def step_sync(self, th, rv, rhod=None, RH_max=44., cond=True):
        C = self.n_cell
        self.th, self.rv = (np.array(a, dtype=np.float64).reshape(C).copy() for a in (th, rv))
        var_rho = rhod is not None
        if var_rho:
            self.rhod = np.array(rhod, dtype=np.float64).reshape(C).copy()
        if not cond:
            return self.th, self.rv
        self.hskpng_mfp()
        sstp = self.sstp_cond
        if self.exact_sstp_cond and (sstp > 1 or self.sstp_cond_act > 1):      # particles_step.ipp:199-236
            self.cond_perparticle(RH_max)
            self.sstp_save()
            return self.th, self.rv
        for step in range(sstp):
            if sstp > 1:                                 # sstp_percell_step.ipp:7-47
                for name in (("rv", "th", "rhod") if var_rho else ("rv", "th")):
                    scl = getattr(self, name)
                    if step == 0:
                        self.old[name] = scl - self.old[name]
                        scl = scl - (sstp - 1) * self.old[name] / sstp
                    else:
                        scl = scl + self.old[name] / sstp
                    setattr(self, name, scl)
            self.hskpng_Tpr()
            if step == 0:
                m3_before = self.moment(self.rw2, 1.5)
            dt = self.dt / sstp
            for s in range(self.n_part):
                c = self.ijk[s]
                self.rw2[s] = advance_rw2(self.rw2[s], dt, RH_max, self.rhod[c], self.rv[c], self.T[c], self.p[c], self.RH[c], self.eta[c],
                                          self.rd3[s], self.kpa[s], self.vt[s], self.lam_D[c], self.lam_K[c])
            m3_after = self.moment(self.rw2, 1.5)
            drv = (-m3_before + m3_after) * (rho_w * (4. / 3) * PI)
            self.rv = self.rv - drv
            self.th = self.th - drv * (-self.th / self.T * np.array([l_v(T) for T in self.T]) / c_pd)
            m3_before = m3_after
        self.sstp_save()
        return self.th, self.rv

    # -- per-particle sub-stepping: condensation/perparticle/*.ipp ------------------------------------------------------------------
    def _pp_state(self, th, rv, rhod):                  # cond_perparticle_advance_rw2.ipp:33-78 (th_dry, variable pressure, pv_cc)
        T = T_of_th_dry(th, rhod)
        p = rhod * (R_d + rv * R_v) * T
        return T, p, RH_pv_cc(p, rv, T)

    def _drv(self, drw3, s, rhod_s):                    # rw3diff2drv, cond_common.ipp:24-41
        mlt = -rho_w * (4. / 3) * PI
        nn = float(self.n[s])
        return mlt * drw3 * nn / rhod_s / self.dv[self.ijk[s]] if self.n_dims > 0 else mlt * drw3 * nn

    def cond_perparticle(self, RH_max):
        N, sstp = self.n_part, self.sstp_cond
        pp = self.pp
        dlt = {k: getattr(self, k)[self.ijk] - pp[k] for k in ("rv", "th", "rhod")}     # calculate_noncond_perparticle_sstp_delta.ipp:13-37
        if not self.sstp_cond_mix:
            m3_before = self.moment(self.rw2, 1.5)       # save_liq_ice_content_before_change
        if self.adaptive_sstp_cond:
            for s in range(N):
                self._adaptive_one(s, dlt, RH_max)
        else:
            rw3 = np.zeros(N)
            for step in range(sstp):
                drv, dth, Tp = np.zeros(N), np.zeros(N), np.zeros(N)
                for s in range(N):
                    for k in ("rv", "th", "rhod"):       # apply_noncond_perparticle_sstp_delta.ipp:13-33
                        pp[k][s] = pp[k][s] + dlt[k][s] / sstp
                    drw3 = -(rw3[s] if step > 0 else math.pow(self.rw2[s], 1.5))     # set_perparticle_drwX_to_minus_rwX.ipp:13-38
                    T, p, RH = self._pp_state(pp["th"][s], pp["rv"][s], pp["rhod"][s])
                    c = self.ijk[s]
                    self.rw2[s] = advance_rw2(self.rw2[s], self.dt / sstp, RH_max, pp["rhod"][s], pp["rv"][s], T, p, RH, visc(T),
                                              self.rd3[s], self.kpa[s], self.vt[s], self.lam_D[c], self.lam_K[c])
                    r3 = math.pow(self.rw2[s], 1.5)       # add_perparticle_rwX_to_drwX.ipp:13-44
                    if step < sstp - 1:
                        rw3[s] = r3
                    drw3 = r3 + drw3
                    drv[s] = self._drv(drw3, s, pp["rhod"][s])
                    Tp[s] = T
                # apply_perparticle_drw3_to_perparticle_rv_and_th.ipp:13-62
                if self.sstp_cond_mix:
                    add = self._cell_sums(drv)
                    pp["rv"] = pp["rv"] + add[self.ijk]
                else:
                    pp["rv"] = drv + pp["rv"]
                for s in range(N):
                    dth[s] = drv[s] * (-pp["th"][s] / Tp[s] * l_v(Tp[s]) / c_pd)
                if self.sstp_cond_mix:
                    add = self._cell_sums(dth)
                    pp["th"] = pp["th"] + add[self.ijk]
                else:
                    pp["th"] = dth + pp["th"]
        # apply_perparticle_cond_change_to_percell_rv_and_th.ipp:13-26
        if self.sstp_cond_mix:
            for s in range(N):                           # update_state: every SD writes its cell, the last in storage order stays
                self.rv[self.ijk[s]] = pp["rv"][s]
                self.th[self.ijk[s]] = pp["th"][s]
        else:
            m3_after = self.moment(self.rw2, 1.5)
            drv_c = (-m3_before + m3_after) * (rho_w * (4. / 3) * PI)
            self.rv = self.rv - drv_c
            self.th = self.th - drv_c * (-self.th / self.T * np.array([l_v(T) for T in self.T]) / c_pd)      # T of the last hskpng_Tpr

    def _cell_sums(self, per_sd):                       # update_pstate, update_th_rv.ipp:243-283 (sum in sorted order)
        out = np.zeros(self.n_cell)
        for pos in range(self.n_part):
            s = self.sorted_id[pos]
            out[self.ijk[s]] += per_sd[s]
        return out

    def _adaptive_one(self, s, dlt, RH_max):            # perparticle_nomixing_adaptive_sstp_cond.ipp:49-290
        pp, c = self.pp, self.ijk[s]
        t = {k: pp[k][s] for k in ("rv", "th", "rhod")}
        rw2 = self.rw2[s]
        st = {}

        def shift(m):
            for k in ("rv", "th", "rhod"):
                t[k] += dlt[k][s] * m

        def thermo():
            st["T"], st["p"], st["RH"] = self._pp_state(t["th"], t["rv"], t["rhod"])

        def grow(r2, dt):
            return advance_rw2(r2, dt, RH_max, t["rhod"], t["rv"], st["T"], st["p"], st["RH"], visc(st["T"]),
                               self.rd3[s], self.kpa[s], self.vt[s], self.lam_D[c], self.lam_K[c])
        sstp_max, sstp = self.sstp_cond, self.sstp_cond
        first_done = sstp_max == 1
        drw2 = drw2_new = 0.0
        tr = 1
        while tr <= sstp_max:
            frac = 1.0 if tr == 1 else -1.0 / tr
            shift(frac)
            thermo()
            d = grow(rw2, self.dt / tr) - rw2
            if tr == 1:
                drw2 = d
            else:
                drw2_new = d
                if abs(drw2_new * 2 - drw2) <= self.drw2_eps * rw2 and abs(drw2) < self.drw2_max * rw2:
                    sstp = tr // 2
                    shift(-frac)
                    first_done = True
                    break
                drw2 = drw2_new
            tr *= 2
        if self.sstp_cond_act > 1:
            rc2 = self.rc2[s]
            if (rw2 < rc2 and (rw2 + sstp * drw2) > rc2) or (rw2 > rc2 and (rw2 + sstp * drw2) < rc2):
                sstp = self.sstp_cond_act
                first_done = False
        if not first_done:
            shift(-frac if sstp_max == 1 else frac)
        frac = 1.0 / sstp
        rw3 = 0.0
        for step in range(sstp):
            drw3 = -rw3 if step > 0 else -math.pow(rw2, 1.5)
            if first_done and step == 0:
                rw2 += drw2
            else:
                shift(frac)
                thermo()
                rw2 = grow(rw2, self.dt / sstp)
            if step < sstp - 1:
                rw3 = math.pow(rw2, 1.5)
                drw3 += rw3
            else:
                drw3 += math.pow(rw2, 1.5)
            drw3 = self._drv(drw3, s, t["rhod"])
            t["rv"] += drw3
            drw3 = drw3 * (-t["th"] / st["T"] * l_v(st["T"]) / c_pd)
            t["th"] += drw3
        for k in ("rv", "th", "rhod"):
            pp[k][s] = t[k]
        self.rw2[s] = rw2

    # -- coalescence: coalescence/particles_impl_coal.ipp:99-546 --------------------------------------------------------------------
    def coal(self, dt):
        self.sort(True)
        u01 = self.rng.u01(self.n_part)
        off = np.zeros(self.n_cell + 1, dtype=np.int64)
        off[self.count_ijk + 1] = self.count_num
        off = np.cumsum(off)
        n_coll = 0
        for pos in range(self.n_part - 1):
            c = self.sorted_ijk[pos]
            if (pos - off[c]) % 2 != 0 or self.sorted_ijk[pos + 1] != c:
                continue
            a, b = self.sorted_id[pos], self.sorted_id[pos + 1]
            m = int(off[c + 1] - off[c])
            scl = (float(m * (m - 1)) / 2) / (m // 2) if m > 1 else 0.
            prob = dt / self.dv[c] * scl * coal_kernel(self.kernel, self.kernel_params, int(self.n[a]), int(self.n[b]),
                                                         self.rw2[a], self.rw2[b], self.vt[a], self.vt[b])
            col_no = int(prob)
            if u01[pos] < prob - col_no:
                col_no += 1
            if col_no == 0:
                continue
            hi, lo = (a, b) if self.n[a] >= self.n[b] else (b, a)
            if self.n[lo] > 0:
                col_no = min(col_no, int(self.n[hi]) // int(self.n[lo]))
            n_coll += col_no
            self.n[hi] = int(self.n[hi]) - col_no * int(self.n[lo])
            rw = math.cbrt(col_no * self.rw2[hi] * math.sqrt(self.rw2[hi]) + self.rw2[lo] * math.sqrt(self.rw2[lo]))
            self.rw2[lo] = rw * rw
            rd3_new = col_no * self.rd3[hi] + self.rd3[lo]
            if len(self.dry_distros) > 1:                # kappa mixing: coal.ipp:59-96
                rd3_old = rd3_new - col_no * self.rd3[hi]
                for _ in range(col_no):
                    self.kpa[lo] = (self.kpa[hi] * self.rd3[hi] + self.kpa[lo] * rd3_old) / (self.rd3[hi] + rd3_old)
                    rd3_old += self.rd3[hi]
            self.rd3[lo] = rd3_new
            self.vt[lo] = -1.0
            if self.sstp_cond_act > 1 and self.allow_sstp_cond:
                self.rc2[lo] = -1.0                      # coal.ipp:527-541 (invalidator)
        return n_coll

    # -- transport: advection/particles_impl_adve.ipp:27-165, sedi.ipp:13-24, bcnd.ipp:99-368 -----------------------------------------
    def _halo_courant(self, C, ext_x, h):
        """Courant field with h columns of halo on both x sides, filled the way init_e2l's map does it (init_e2l.ipp:34-114,
        init_sync.ipp:27-44): the linear element index is wrapped once by the size of the caller's array, so column -1 of the
        staggered x-field is the caller's LAST face (nx), not face nx - 1"""
        n = self.nx + ext_x
        idx = np.arange(-h, n + h)
        idx = np.where(idx >= n, idx - n, np.where(idx < 0, idx + n, idx))
        return C[idx]

    def _adve_pred_corr(self):                         # advection/particles_impl_adve.ipp:169-304
        h = 2                                            # halo_size of pred_corr (particles_impl.ipp ctor)
        three = self.n_dims == 3
        Cxh = self._halo_courant(self.Cx, 1, h)
        Cyh = self._halo_courant(self.Cy, 0, h) if three else None
        Czh = self._halo_courant(self.Cz, 0, h) if self.n_dims >= 2 else None

        def cell():                                      # hskpng_ijk in the halo's coordinate system
            i = (self.x / self.dx).astype(np.int64)
            j = (self.y / self.dy).astype(np.int64) if three else np.zeros(self.n_part, dtype=np.int64)
            k = (self.z / self.dz).astype(np.int64) if self.n_dims >= 2 else np.zeros(self.n_part, dtype=np.int64)
            return i, j, k

        def calc(apply):                                 # adve_calc<adve_helper_expl>(apply): all dimensions from the same cell indices
            i, j, k = cell.ijk
            g = (lambda a, di, dj, dk: a[i + di, j + dj, k + dk]) if three else \
                ((lambda a, di, dj, dk: a[i + di, k + dk]) if self.n_dims == 2 else (lambda a, di, dj, dk: a[i + di]))
            f = 1. if apply else 0.
            C_l, C_r = g(Cxh, 0, 0, 0), g(Cxh, 1, 0, 0)
            if self.n_dims == 1:
                C_r = C_l
            self.x = f * self.x + (C_r - C_l) * (self.x - self.dx * i) + self.dx * C_l
            if three:
                C_l, C_r = g(Cyh, 0, 0, 0), g(Cyh, 0, 1, 0)
                self.y = f * self.y + (C_r - C_l) * (self.y - self.dy * j) + self.dy * C_l
            if self.n_dims >= 2:
                C_l, C_r = g(Czh, 0, 0, 0), g(Czh, 0, 0, 1)
                self.z = f * self.z + (C_r - C_l) * (self.z - self.dz * k) + self.dz * C_l

        self.x = self.x + float(h) * self.dx             # coordinates that start at the halo's left edge
        cell.ijk = cell()
        x_old, y_old, z_old = self.x.copy(), self.y.copy(), self.z.copy()
        calc(True)                                       # predictor
        if self.n_dims >= 2:
            self.z = np.where(self.z >= self.z1, self.z1 - 1e-8 * self.dz, self.z)
            self.z = np.where(self.z <= self.z0, self.z0 + 1e-8 * self.dz, self.z)
        if three:
            L_y = self.y1 - self.y0
            y_old = np.where(self.y >= self.y1, y_old + L_y, y_old)
            y_old = np.where(self.y < self.y0, y_old - L_y, y_old)
            self.y = self.y0 + np.fmod((self.y - self.y0) + 10 * L_y, L_y)
        cell.ijk = cell()
        self._k_midpoint = cell.ijk[2]                   # the level index hskpng_ijk leaves behind (read by subs)
        x_old = self.x + x_old
        if three:
            y_old = self.y + y_old
        if self.n_dims >= 2:
            z_old = self.z + z_old
        calc(False)                                      # displacement at the midpoint
        self.x = (self.x + x_old) / 2.
        if three:
            self.y = (self.y + y_old) / 2.
        if self.n_dims >= 2:
            self.z = (self.z + z_old) / 2.
        self.x = self.x - float(h) * self.dx

    def adve(self):
        if self.n_dims == 0:
            return
        if self.adve_scheme == "pred_corr":
            return self._adve_pred_corr()
        i, j, k = self.unravel(self.ijk)
        dims = [("x", i, self.Cx, self.dx, 0)]
        if self.n_dims == 3:
            dims.append(("y", j, self.Cy, self.dy, 1))
        if self.n_dims >= 2:
            dims.append(("z", k, self.Cz, self.dz, self.n_dims - 1))
        for name, idx, Cf, d, axis in dims:
            x = getattr(self, name)
            grid = (i, j, k) if self.n_dims == 3 else ((i, k) if self.n_dims == 2 else (i,))
            lo = list(grid)
            hi = list(grid)
            hi[axis] = hi[axis] + 1
            C_l, C_r = Cf[tuple(lo)], Cf[tuple(hi)]
            if self.n_dims == 1:
                C_r = C_l                                # rgt = lft + nz with nz = 0: init_grid.ipp:110-119
            if self.adve_scheme == "implicit":
                x = (x + d * (C_l - idx * (C_r - C_l))) / (1 - (C_r - C_l))
            else:
                x = 1 * x + (C_r - C_l) * (x - d * idx) + d * C_l
            setattr(self, name, x)

    def bcnd(self):
        if self.n_dims == 0:
            return
        wrap = lambda x, a, b: a + np.fmod((x - a) + 10 * (b - a), b - a)
        if self.distmem:                                 # bcnd.ipp:144-194: ids of the SDs that cross a slab face, in ascending storage order
            self.lft_id = np.nonzero(self.x < self.x0)[0]
            self.rgt_id = np.nonzero(self.x >= self.x1)[0]
        elif not self.open_side_walls:
            self.x = wrap(self.x, self.x0, self.x1)
        else:                                            # bcnd.ipp:132-142: SDs that left through a side wall are flagged for removal
            self.n[(self.x >= self.x1) | (self.x < self.x0)] = 0
        if self.n_dims == 3:
            if not self.open_side_walls:
                self.y = wrap(self.y, self.y0, self.y1)
            else:
                self.n[(self.y >= self.y1) | (self.y < self.y0)] = 0
        if self.n_dims > 1:
            self.n[self.z >= self.z1] = 0
            out = self.z < self.z0
            nf = np.where(out, self.n.astype(np.float64), 0.)
            self.puddle["liquid_volume"] += sum(4. / 3. * PI * nf[s] * math.pow(self.rw2[s], 1.5) for s in range(self.n_part))
            self.puddle["dry_volume"] += sum(4. / 3. * PI * nf[s] * math.pow(self.rd3[s], 1.) for s in range(self.n_part))
            self.puddle["liquid_number"] += float(np.where(self.rw2 == 0, 0., nf).sum())
            self.puddle["particle_number"] += float(nf.sum())
            self.n[out] = 0

    def rcyc(self):                                     # housekeeping/particles_impl_rcyc.ipp:44-139
        n_to_rcyc = n_flagged = int((self.n == 0).sum())
        self.n_recycled = 0
        if n_flagged == 0:
            return False
        order = np.argsort(self.n, kind="stable")       # sort_by_key on (n, storage index): stable in the cpp back-end
        tmp = self.n[order]
        ones = np.nonzero(tmp[::-1] == 1)[0]
        n_splittable = int(ones[0]) if ones.size else self.n_part                      # find() from the large end
        if n_splittable == 0:
            return False
        n_flagged = min(n_flagged, n_splittable)
        src, dst = order[::-1][:n_flagged], order[:n_flagged]
        names = ["rd3", "rw2", "kpa", "vt", "x", "y", "z", "rc2"]
        for q in range(n_flagged):                       # copy_n runs front to back
            for name in names:
                a = getattr(self, name)
                if a.size:
                    a[dst[q]] = a[src[q]]
            if hasattr(self, "pp"):
                for k in self.pp:
                    self.pp[k][dst[q]] = self.pp[k][src[q]]
        for q in range(n_flagged):
            self.n[dst[q]] = self.n[src[q]] - self.n[src[q]] // np.uint64(2)
        for q in range(n_flagged):
            self.n[src[q]] = self.n[src[q]] // np.uint64(2)
        self.n_recycled = n_flagged
        return n_flagged == n_to_rcyc                    # all recycled: nothing left to remove

    def step_async(self, adve=True, sedi=True, coal=True, cond=True, rcyc=False, subs=False):
        self.hskpng_Tpr()
        if sedi or coal or cond:
            self.hskpng_vterm(False)
        n_coll = 0
        if coal:
            for step in range(self.sstp_coal):
                n_coll += self.coal(self.dt / self.sstp_coal)
                if step + 1 != self.sstp_coal:
                    self.hskpng_vterm(True)
            self.hskpng_rc2()                            # particles_step.ipp:402-403
        if adve:
            self.adve()
        if sedi and self.nz:
            self.z = self.z - self.dt * self.vt
        if subs and self.nz:                             # subsidence/particles_impl_subs.ipp:13-25: w_LS at the level index k the SD had
            k = self._k_midpoint if (adve and self.adve_scheme == "pred_corr") else self.unravel(self.ijk)[2]     # when ijk was last computed
            self.z = self.z - self.dt * self.w_LS[k]
        self.bcnd()
        self._n_coll = n_coll
        if self.distmem:                                 # the slab system exchanges the migrants, then calls post_copy on every slab
            return n_coll
        return self.post_copy(rcyc)

    def post_copy(self, rcyc=False):                     # post_copy.ipp:18-35
        n_coll = self._n_coll
        self.n_recycled = 0
        if rcyc:                                         # post_copy.ipp:24-29
            self.rcyc()
        keep = self.n != 0                               # hskpng_remove_n0: hskpng_remove.ipp:20-75
        for name in ("n", "rd3", "rw2", "kpa", "vt", "x", "y", "z", "rc2"):
            a = getattr(self, name)
            if a.size:
                setattr(self, name, a[keep])
        if hasattr(self, "pp"):
            self.pp = {k: v[keep] for k, v in self.pp.items()}
        self.n_part = int(keep.sum())
        ii = (self.x / self.dx).astype(np.int64) if self.nx else 0      # hskpng_ijk.ipp:159-200
        jj = (self.y / self.dy).astype(np.int64) if self.ny else 0
        kk = (self.z / self.dz).astype(np.int64) if self.nz else 0
        self.ijk = ((ii * max(1, self.ny) + jj) * max(1, self.nz) + kk) * np.ones(self.n_part, dtype=np.int64)
        self.sort(False)
        return n_coll


# ---- x-slab decomposition: src/detail/distmem_opts.hpp:10-52, impl_multi_gpu/particles_multi_gpu_impl.ipp:131-170,
#      impl_multi_gpu/particles_multi_gpu_impl_step_async_and_copy.ipp:28-206, distributed_memory/particles_impl_{pack,unpack}.ipp ----
def slab_nx(nx, rank, size):                            # detail::get_dev_nx
    return int(nx / size + .5) if rank < size - 1 else nx - rank * int(nx / size + .5)


def xchng_courants_rule(n_dims, nx, ny, nz, halo=2):
    """Which values of a slab's halo-extended Courant arrays travel to which neighbour, as flat offsets and counts: the index
    arithmetic of src/impl/distributed_memory/particles_impl_xchng_courants.ipp:26-52 (arrays laid out as init_sync.ipp:29-44:
    Cx (nx + 2 h + 1) x-planes, Cy / Cz (nx + 2 h) x-planes, z fastest).  Returns {name: (plane_size, send_to_lft, send_to_rgt,
    recv_from_lft, recv_from_rgt, count)}: `count` values starting at send_to_lft go to the left neighbour, which stores them at
    ITS recv_from_rgt (and the other way round).  Test infrastructure (tests/test_cpu_distributed.py), restated from the reference;
    the product's engine does the same between GPUs (csrc/lcx_transport.cu halo_put / halo_take)."""
    ny1, nz1 = max(ny, 1), max(nz, 1)
    n_cell = nx * ny1 * nz1
    out = {}
    if n_dims >= 1:
        plane = 1 if n_dims == 1 else nz if n_dims == 2 else nz * ny
        halo_x = halo * plane
        size = (nx + 2 * halo + 1) * plane
        out["Cx"] = (plane, (halo + 1) * plane, n_cell, 0, size - halo_x, halo_x)                    # :26-31, :45-46, halo_x values
    if n_dims >= 2:
        plane = (nz + 1) if n_dims == 2 else (nz + 1) * ny
        halo_z = halo * plane
        size = (nx + 2 * halo) * plane
        out["Cz"] = (plane, halo_z, nx * plane, 0, size - halo_z, halo_z)                             # :33-37, :47-48
    if n_dims == 3:
        plane = (ny + 1) * nz
        halo_y = halo * plane
        size = (nx + 2 * halo) * plane
        out["Cy"] = (plane, halo_y, nx * plane, 0, size - halo_y, halo_y)                             # :39-42, :49-50
    return out


class SlabParticles:
    """The periodic domain cut into `size` x-slabs, each an independent Particles object in its own local coordinates (x0 = 0 for every
    slab but the first), super-droplets that cross a slab face handed to the neighbour once per step.

    Restated from the reference's multi_CUDA back-end.  NOT PINNED against a run of the reference: multi_CUDA needs GPUs and the MPI
    variant is not built here.  What tests/test_cpu_oracle.py checks instead are the consequences the reference's own tests demand
    (tests/mpi/mpi_adve_test.cpp: a pattern advected once around the ring returns unchanged) and agreement with the undivided domain.
    Order of arrival (the part that must be bit-exact in a re-implementation): every slab first appends what its RIGHT neighbour sent
    (that neighbour's left-movers, in its ascending storage order), then what its LEFT neighbour sent; the senders' copies get n = 0
    and disappear in post_copy.  Positions: x' = x1_of_receiver + x - x0_of_sender for left-movers, x' = x0_of_receiver + x - x1_of_sender
    for right-movers (pack.ipp:15-26), then pulled inside by config.bcond_tolerance = 5e-4 m if still outside (unpack.ipp:14-31,100-101),
    and - for the batch that came from the right - stepped just below x1 when it landed exactly on it (..._and_copy.ipp:137).
    Every slab seeds its generator with the same opts_init.rng_seed (particles_multi_gpu_impl.ipp:139: the options are copied)."""
    BCOND_TOLERANCE = 5e-4                               # src/detail/config.hpp:31

    def __init__(self, size, **kw):
        assert size > 1 and kw.get("adve_scheme", "implicit") != "pred_corr", "the Courant halo across slabs is not restated"
        self.size, self.nx, self.dx = size, kw["nx"], kw["dx"]
        self.slabs, self.n_x_bfr = [], []
        for rank in range(size):
            k = dict(kw)
            bfr = rank * slab_nx(self.nx, 0, size)
            k["nx"] = slab_nx(self.nx, rank, size)
            k["x0"] = kw.get("x0", 0.) if rank == 0 else 0.
            k["x1"] = k["nx"] * self.dx if rank != size - 1 else kw["x1"] - bfr * self.dx
            k["n_sd_max"] = kw["n_sd_max"] // size + 1
            p = Particles(**k)
            p.distmem = True
            self.slabs.append(p)
            self.n_x_bfr.append(bfr)

    def _cut(self, a, rank, ext=0):
        return None if a is None else a[self.n_x_bfr[rank]: self.n_x_bfr[rank] + self.slabs[rank].nx + ext]

    def init(self, th, rv, rhod, Cx=None, Cy=None, Cz=None):
        for r, p in enumerate(self.slabs):
            p.init(self._cut(th, r), self._cut(rv, r), self._cut(rhod, r), self._cut(Cx, r, 1), self._cut(Cy, r), self._cut(Cz, r))

    def step_sync(self, th, rv, rhod=None, **kw):
        out_th, out_rv = np.array(th, dtype=np.float64), np.array(rv, dtype=np.float64)
        for r, p in enumerate(self.slabs):
            t, q = p.step_sync(self._cut(th, r), self._cut(rv, r), self._cut(rhod, r) if rhod is not None else None, **kw)
            a, b = self.n_x_bfr[r], self.n_x_bfr[r] + p.nx
            out_th[a:b], out_rv[a:b] = t.reshape(out_th[a:b].shape), q.reshape(out_rv[a:b].shape)
        return out_th, out_rv

    def step_async(self, rcyc=False, **kw):
        n_coll = sum(p.step_async(**kw) for p in self.slabs)
        names = ("n", "rd3", "rw2", "kpa", "vt", "x", "y", "z", "rc2")
        size = self.size
        # what every slab sends: copies of the attributes of its leavers with x already in the receiver's coordinates
        out = []
        for r, p in enumerate(self.slabs):
            lft, rgt = self.slabs[(r - 1) % size], self.slabs[(r + 1) % size]
            batch = {}
            for side, ids, shift in (("lft", p.lft_id, lambda x: lft.x1 + x - p.x0), ("rgt", p.rgt_id, lambda x: rgt.x0 + x - p.x1)):
                b = {nm: getattr(p, nm)[ids].copy() for nm in names if getattr(p, nm).size}
                b["x"] = shift(b["x"])
                if hasattr(p, "pp"):
                    b["pp"] = {k: v[ids].copy() for k, v in p.pp.items()}
                batch[side] = b
            out.append(batch)
        for r, p in enumerate(self.slabs):
            tol = self.BCOND_TOLERANCE
            for src, side in (((r + 1) % size, "lft"), ((r - 1) % size, "rgt")):     # from the right neighbour first, then from the left
                b = out[src][side]
                x = b["x"]
                x = np.where(x >= p.x1, x - tol, np.where(x < p.x0, x + tol, x))     # tolerance_away_from_bcond
                if side == "lft":
                    x = np.where(x == p.x1, np.nextafter(x, 0.), x)
                b = dict(b, x=x)
                for nm in names:
                    if getattr(p, nm).size or nm in b:
                        if nm in b:
                            setattr(p, nm, np.concatenate([getattr(p, nm), b[nm]]))
                if hasattr(p, "pp"):
                    p.pp = {k: np.concatenate([v, b["pp"][k]]) for k, v in p.pp.items()}
            p.n[p.lft_id] = 0                            # flag_lft / flag_rgt: the senders' copies are removed by post_copy
            p.n[p.rgt_id] = 0
            p.n_part = p.n.size
        for p in self.slabs:
            p.post_copy(rcyc)
        return n_coll

    def per_cell(self, fun):
        """concatenates a per-cell diagnostic of the slabs into the global (nx, ny, nz) order"""
        return np.concatenate([np.asarray(fun(p)).reshape(p.nx, max(1, p.ny), max(1, p.nz)) for p in self.slabs], axis=0)
