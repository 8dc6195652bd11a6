"""Python view of the `lgrngn` particle-system API (ctypes over bindings/lgrngn_capi.h).

Mirrors what the reference's Boost.Python module offers (reference bindings/python/lib.cpp:217-434):
`opts_init_t`, `opts_t`, the enums, `factory(backend, opts_init)` and a particles object with `init`,
`step_sync`, `step_async`, `diag_*`, `outbuf`.  numpy arrays stand in for `arrinfo_t` (C order, dimension
order x[,y],z).  The class is parametrised by the shared library it drives, so the parity tests can point
the very same calls at the reference build (oracle/_ref) - the product never does.
"""
import ctypes as C
import os

import numpy as np

MAX_MODES, MAX_DISTROS, MAX_KPARAMS, MAX_SIZES = 4, 4, 4, 16


class backend_t:
    undefined, serial, OpenMP, CUDA, multi_CUDA = range(5)


class kernel_t:
    (undefined, geometric, golovin, hall, hall_davis_no_waals, Long, onishi_hall, onishi_hall_davis_no_waals,
     hall_pinsky_1000mb_grav, hall_pinsky_cumulonimbus, hall_pinsky_stratocumulus, vohl_davis_no_waals) = range(12)


class vt_t:
    undefined, beard76, beard77, beard77fast, khvorostyanov_spherical, khvorostyanov_nonspherical = range(6)


class as_t:
    undefined, implicit, euler, pred_corr = range(4)


class RH_formula_t:
    pv_cc, rv_cc, pv_tet, rv_tet = range(4)


class _Distro(C.Structure):
    _fields_ = [("kind", C.c_int), ("kappa", C.c_double), ("rd_insol", C.c_double), ("n_modes", C.c_int),
                ("mean_r", C.c_double * MAX_MODES), ("stdev", C.c_double * MAX_MODES), ("n_tot", C.c_double * MAX_MODES),
                ("r0", C.c_double), ("n0", C.c_double),
                ("fn", C.c_void_p), ("ctx", C.c_void_p)]


DISTRO_FN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)


class _DrySize(C.Structure):
    _fields_ = [("kappa", C.c_double), ("rd_insol", C.c_double), ("radius", C.c_double), ("conc", C.c_double), ("count", C.c_int)]


class _OptsInit(C.Structure):
    _fields_ = [("backend", C.c_int), ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double), ("dt", C.c_double),
                ("sstp_cond", C.c_int), ("sstp_coal", C.c_int),
                ("x0", C.c_double), ("y0", C.c_double), ("z0", C.c_double),
                ("x1", C.c_double), ("y1", C.c_double), ("z1", C.c_double),
                ("sd_conc", C.c_ulonglong), ("sd_const_multi", C.c_ulonglong), ("n_sd_max", C.c_ulonglong),
                ("kernel", C.c_int), ("terminal_velocity", C.c_int), ("adve_scheme", C.c_int), ("RH_formula", C.c_int),
                ("n_kernel_parameters", C.c_int), ("kernel_parameters", C.c_double * MAX_KPARAMS),
                ("coal_switch", C.c_int), ("sedi_switch", C.c_int), ("subs_switch", C.c_int), ("exact_sstp_cond", C.c_int),
                ("turb_adve_switch", C.c_int), ("turb_cond_switch", C.c_int), ("turb_coal_switch", C.c_int),
                ("ice_switch", C.c_int), ("chem_switch", C.c_int),
                ("RH_max", C.c_double),
                ("rng_seed", C.c_int), ("rng_seed_init", C.c_int), ("rng_seed_init_switch", C.c_int),
                ("dev_count", C.c_int), ("dev_id", C.c_int),
                ("rd_min", C.c_double), ("rd_max", C.c_double),
                ("open_side_walls", C.c_int), ("periodic_topbot_walls", C.c_int), ("variable_dt_switch", C.c_int),
                ("th_dry", C.c_int), ("const_p", C.c_int), ("aerosol_independent_of_rhod", C.c_int),
                ("n_distros", C.c_int), ("distros", _Distro * MAX_DISTROS),
                ("n_w_LS", C.c_int), ("w_LS", C.POINTER(C.c_double)),
                ("sd_conc_large_tail", C.c_int), ("no_ccn_at_init", C.c_int),
                ("n_dry_sizes", C.c_int), ("dry_sizes", _DrySize * MAX_SIZES),
                ("n_aerosol_conc_factor", C.c_int), ("aerosol_conc_factor", C.POINTER(C.c_double)),
                ("sstp_cond_mix", C.c_int), ("adaptive_sstp_cond", C.c_int), ("sstp_cond_act", C.c_int),
                ("sstp_cond_adapt_drw2_eps", C.c_double), ("sstp_cond_adapt_drw2_max", C.c_double), ("rc2_T", C.c_double)]


class _Opts(C.Structure):
    _fields_ = [("adve", C.c_int), ("sedi", C.c_int), ("subs", C.c_int), ("cond", C.c_int), ("coal", C.c_int),
                ("rcyc", C.c_int), ("RH_max", C.c_double), ("dt", C.c_double)]


class _Arr(C.Structure):
    _fields_ = [("data", C.c_void_p), ("strides", C.c_long * 3)]      # data: lgc_real * (double, or float for the lgcf_* binding)


class _Prefixed:
    """the flat binding of one precision: `lgc_xyz` resolves to `lgc_xyz` (double) or `lgcf_xyz` (float) of the shared library;
    every other name (back-end specific extras) passes through"""

    def __init__(self, cdll, prefix):
        self._cdll, self._prefix = cdll, prefix

    def __getattr__(self, name):
        if name.startswith("lgc_"):
            return getattr(self._cdll, self._prefix + name[4:])
        return getattr(self._cdll, name)


DIAG = {name: i for i, name in enumerate([
    "all", "rw_ge_rc", "RH_ge_Sc", "dry_rng", "wet_rng", "kappa_rng", "dry_rng_cons", "wet_rng_cons",
    "kappa_rng_cons", "water", "water_cons", "sd_conc", "pressure", "temperature", "RH", "dry_mom", "wet_mom",
    "kappa_mom", "precip_rate", "max_rw", "vel_div", "wet_mass_dens"])}

PUDDLE_KEYS = ["HNO3", "NH3", "CO2", "SO2", "H2O2", "O3", "S_VI", "H", "liquid_volume", "dry_volume",
               "particle_number", "ice_mass", "liquid_number", "ice_number"]


def lognormal(kappa, modes, rd_insol=0.0):
    """dry spectrum: sum of lognormal modes [(mean_r [m], geometric stdev, n_tot [m^-3]), ...]"""
    return {"kind": 0, "kappa": kappa, "rd_insol": rd_insol, "modes": list(modes)}


def callable_distro(kappa, fn, rd_insol=0.0):
    """dry spectrum given as a Python callable n(ln r) [m^-3 per unit ln r at STP], like the reference's Python binding"""
    return {"kind": 2, "kappa": kappa, "rd_insol": rd_insol, "fn": fn}


def expvolume(kappa, r0, n0, rd_insol=0.0):
    """dry spectrum exponential in volume (Shima et al. 2009): n0 * 3 (r/r0)^3 exp(-(r/r0)^3)"""
    return {"kind": 1, "kappa": kappa, "rd_insol": rd_insol, "r0": r0, "n0": n0}


class Library:
    """One loaded implementation of the flat binding (the B200 back-end or, in tests, the reference)."""

    def __init__(self, path, real="f64"):
        """real: "f64" drives factory<double> (lgc_*), "f32" factory<float> (lgcf_*; the reference instantiates both)"""
        if not os.path.exists(path):
            raise OSError("shared library not found: %s (run `python -c 'import __graft_entry__ as g; g.build()'`)" % path)
        assert real in ("f64", "f32")
        self.path = path
        self.real = real
        self.dtype = np.float64 if real == "f64" else np.float32
        self.creal = C.c_double if real == "f64" else C.c_float
        self.cdll = C.CDLL(path, mode=C.RTLD_LOCAL)
        self.lib = lib = _Prefixed(self.cdll, "lgc_" if real == "f64" else "lgcf_")
        P = C.POINTER
        lib.lgc_last_error.restype = C.c_char_p
        lib.lgc_impl_name.restype = C.c_char_p
        lib.lgc_opts_init_defaults.argtypes = [P(_OptsInit)]
        lib.lgc_opts_defaults.argtypes = [P(_Opts)]
        lib.lgc_create.argtypes = [P(_OptsInit), P(C.c_void_p)]
        lib.lgc_destroy.argtypes = [C.c_void_p]
        lib.lgc_init.argtypes = [C.c_void_p] + [P(_Arr)] * 7
        lib.lgc_step_sync.argtypes = [C.c_void_p, P(_Opts)] + [P(_Arr)] * 6
        lib.lgc_sync_in.argtypes = [C.c_void_p] + [P(_Arr)] * 6
        lib.lgc_step_cond.argtypes = [C.c_void_p, P(_Opts)] + [P(_Arr)] * 2
        lib.lgc_step_async.argtypes = [C.c_void_p, P(_Opts)]
        lib.lgc_diag.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
        lib.lgc_n_cell.argtypes = [C.c_void_p]
        lib.lgc_n_cell.restype = C.c_long
        lib.lgc_outbuf.argtypes = [C.c_void_p, P(self.creal), C.c_long]
        lib.lgc_get_attr.argtypes = [C.c_void_p, C.c_char_p, P(self.creal), C.c_long, P(C.c_long)]
        lib.lgc_get_n.argtypes = [C.c_void_p, P(C.c_ulonglong), C.c_long, P(C.c_long)]
        lib.lgc_puddle.argtypes = [C.c_void_p, P(C.c_double)]
        self.name = lib.lgc_impl_name().decode()

    def check(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.lgc_last_error().decode())

    # ---- API objects -------------------------------------------------------------------------------
    def opts_init_t(self):
        return OptsInit(self)

    def opts_t(self):
        return Opts(self)

    def factory(self, backend, opts_init):
        return Particles(self, backend, opts_init)


_NOT_SCALAR = ("distros", "n_distros", "w_LS", "n_w_LS", "kernel_parameters", "n_kernel_parameters", "backend",
               "dry_sizes", "n_dry_sizes", "aerosol_conc_factor", "n_aerosol_conc_factor")


class OptsInit:
    """attribute bag with the defaults of opts_init_t (reference lgrngn/opts_init.hpp:194-249)"""

    def __init__(self, library):
        c = _OptsInit()
        library.lib.lgc_opts_init_defaults(C.byref(c))
        for name, _ in _OptsInit._fields_:
            if name in _NOT_SCALAR:
                continue
            setattr(self, name, getattr(c, name))
        self.kernel_parameters = []
        self.dry_distros = []      # list of lognormal(...) / expvolume(...)
        self.w_LS = []
        self.dry_sizes = {}        # {(kappa, rd_insol) or kappa: {radius: (STP concentration, SDs per cell)}}
        self.aerosol_conc_factor = []

    def _pack(self, backend):
        c = _OptsInit()
        c.backend = int(backend)
        for name, _ in _OptsInit._fields_:
            if name in _NOT_SCALAR:
                continue
            setattr(c, name, getattr(self, name))
        kp = list(self.kernel_parameters)
        c.n_kernel_parameters = len(kp)
        for i, v in enumerate(kp):
            c.kernel_parameters[i] = float(v)
        keep = []
        c.n_distros = len(self.dry_distros)
        for i, d in enumerate(self.dry_distros):
            cd = c.distros[i]
            cd.kind, cd.kappa, cd.rd_insol = d["kind"], d["kappa"], d.get("rd_insol", 0.0)
            if d["kind"] == 0:
                cd.n_modes = len(d["modes"])
                for m, (mean_r, stdev, n_tot) in enumerate(d["modes"]):
                    cd.mean_r[m], cd.stdev[m], cd.n_tot[m] = mean_r, stdev, n_tot
            elif d["kind"] == 1:
                cd.r0, cd.n0 = d["r0"], d["n0"]
            else:
                fn = d["fn"]
                cb = DISTRO_FN(lambda lnr, _ctx, fn=fn: float(fn(lnr)))
                keep.append(cb)                      # must outlive the particles object
                cd.fn = C.cast(cb, C.c_void_p)
        i = 0
        for key, sizes in self.dry_sizes.items():
            kappa, rd_insol = key if isinstance(key, tuple) else (key, 0.0)
            for radius, (conc, count) in sizes.items():
                ds = c.dry_sizes[i]
                ds.kappa, ds.rd_insol, ds.radius, ds.conc, ds.count = kappa, rd_insol, radius, conc, int(count)
                i += 1
        c.n_dry_sizes = i
        if len(self.w_LS):
            keep.append(np.ascontiguousarray(self.w_LS, dtype=np.float64))
            c.n_w_LS = keep[-1].size
            c.w_LS = keep[-1].ctypes.data_as(C.POINTER(C.c_double))
        if len(self.aerosol_conc_factor):
            keep.append(np.ascontiguousarray(self.aerosol_conc_factor, dtype=np.float64))
            c.n_aerosol_conc_factor = keep[-1].size
            c.aerosol_conc_factor = keep[-1].ctypes.data_as(C.POINTER(C.c_double))
        return c, keep


class Opts:
    def __init__(self, library):
        c = _Opts()
        library.lib.lgc_opts_defaults(C.byref(c))
        for name, _ in _Opts._fields_:
            setattr(self, name, getattr(c, name))

    def _pack(self):
        c = _Opts()
        for name, _ in _Opts._fields_:
            setattr(c, name, getattr(self, name))
        return c


def _arr(a, dtype=np.float64):
    """numpy array (of the library's real type, any strides that are multiples of the item size) -> lgc_arr; None -> NULL"""
    if a is None:
        return None
    item = np.dtype(dtype).itemsize
    cai = getattr(a, "__cuda_array_interface__", None)
    if cai is not None and not isinstance(a, np.ndarray):
        # device-resident field (e.g. a torch CUDA tensor): the raw device pointer goes through arrinfo_t, copies stay on the GPU
        assert cai["typestr"][1:] == "f%d" % item, "Eulerian fields must be %s" % np.dtype(dtype).name
        c = _Arr()
        c.data = int(cai["data"][0])
        shape = tuple(cai["shape"])
        if cai.get("strides"):
            st = [s // item for s in cai["strides"]]
        else:
            st = [int(np.prod(shape[i + 1:], dtype=np.int64)) for i in range(len(shape))]
        while len(st) < 3:
            st.append(1)
        for i in range(3):
            c.strides[i] = st[i]
        return C.byref(c)
    assert a.dtype == dtype, "Eulerian fields must be %s" % np.dtype(dtype).name
    c = _Arr()
    c.data = a.ctypes.data
    st = [s // a.itemsize for s in a.strides] if a.ndim else []
    while len(st) < 3:
        st.append(1)
    for i in range(3):
        c.strides[i] = st[i]
    return C.byref(c)


class Particles:
    def __init__(self, library, backend, opts_init):
        self._L = library
        self._lib = library.lib
        c, keep = opts_init._pack(backend)
        self._keep = keep          # arrays and callbacks the C side still points to (dry spectra are evaluated in init())
        h = C.c_void_p()
        library.check(self._lib.lgc_create(C.byref(c), C.byref(h)))
        self._h = h
        self.n_cell = self._lib.lgc_n_cell(h)
        self._cap = int(opts_init.n_sd_max) + 16

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.lgc_destroy(h)
            self._h = None

    def _a(self, *arrays):
        return [_arr(a, self._L.dtype) for a in arrays]

    def init(self, th, rv, rhod, p=None, Cx=None, Cy=None, Cz=None):
        self._L.check(self._lib.lgc_init(self._h, *self._a(th, rv, rhod, p, Cx, Cy, Cz)))

    def step_sync(self, opts, th, rv, rhod=None, Cx=None, Cy=None, Cz=None):
        o = opts._pack()
        self._L.check(self._lib.lgc_step_sync(self._h, C.byref(o), *self._a(th, rv, rhod, Cx, Cy, Cz)))

    def sync_in(self, th, rv, rhod=None, Cx=None, Cy=None, Cz=None):
        self._L.check(self._lib.lgc_sync_in(self._h, *self._a(th, rv, rhod, Cx, Cy, Cz)))

    def step_cond(self, opts, th, rv):
        o = opts._pack()
        self._L.check(self._lib.lgc_step_cond(self._h, C.byref(o), *self._a(th, rv)))

    def step_async(self, opts):
        o = opts._pack()
        self._L.check(self._lib.lgc_step_async(self._h, C.byref(o)))

    def _diag(self, what, a=0.0, b=0.0):
        self._L.check(self._lib.lgc_diag(self._h, DIAG[what], float(a), float(b)))

    def outbuf(self):
        out = np.empty(self.n_cell, dtype=self._L.dtype)
        self._L.check(self._lib.lgc_outbuf(self._h, out.ctypes.data_as(C.POINTER(self._L.creal)), out.size))
        return out

    def get_attr(self, name):
        buf = np.empty(self._cap, dtype=self._L.dtype)
        n = C.c_long()
        self._L.check(self._lib.lgc_get_attr(self._h, name.encode(), buf.ctypes.data_as(C.POINTER(self._L.creal)), buf.size, C.byref(n)))
        return buf[:n.value].copy()

    def get_n(self):
        buf = np.empty(self._cap, dtype=np.uint64)
        n = C.c_long()
        self._L.check(self._lib.lgc_get_n(self._h, buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), buf.size, C.byref(n)))
        return buf[:n.value].copy()

    def diag_puddle(self):
        out = (C.c_double * 14)()
        self._L.check(self._lib.lgc_puddle(self._h, out))
        return dict(zip(PUDDLE_KEYS, list(out)))


def _add_diag(name, nargs):
    if nargs == 0:
        def f(self):
            self._diag(name)
    elif nargs == 1:
        def f(self, k):
            self._diag(name, k)
    else:
        def f(self, a, b):
            self._diag(name, a, b)
    f.__name__ = "diag_" + name
    setattr(Particles, "diag_" + name, f)


for _n in ("all", "rw_ge_rc", "RH_ge_Sc", "water", "water_cons", "sd_conc", "pressure", "temperature", "RH",
           "precip_rate", "max_rw", "vel_div"):
    _add_diag(_n, 0)
for _n in ("dry_mom", "wet_mom", "kappa_mom"):
    _add_diag(_n, 1)
for _n in ("dry_rng", "wet_rng", "kappa_rng", "dry_rng_cons", "wet_rng_cons", "kappa_rng_cons", "wet_mass_dens"):
    _add_diag(_n, 2)


_HERE = os.path.dirname(os.path.abspath(__file__))
B200_LIB_PATH = os.path.join(os.environ.get("LCX_B200_LIBDIR") or os.path.join(_HERE, "lib"), "liblgrngn_b200.so")
_b200 = {}


def b200(real="f64"):
    """the B200-native back-end (hand-written CUDA behind the lgrngn API); fails loudly if it is not built.
    real = "f32": factory<float>, served by the single-precision engine"""
    if real not in _b200:
        _b200[real] = Library(B200_LIB_PATH, real)
    return _b200[real]
