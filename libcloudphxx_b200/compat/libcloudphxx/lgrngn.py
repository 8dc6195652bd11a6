"""`libcloudphxx.lgrngn` as the reference's Python binding exposes it (bindings/python/lib.cpp:217-434)."""
import os

import numpy as np

from libcloudphxx_b200 import lgrngn as _L


def _library():
    """the shared library that serves the calls: the B200 back-end, unless LIBCLOUDPHXX_COMPAT_LIBRARY names another
    library exporting the same flat binding (bindings/lgrngn_capi.h)"""
    path = os.environ.get("LIBCLOUDPHXX_COMPAT_LIBRARY")
    return _L.Library(path) if path else _L.b200()


class _Enum(int):
    """an int that prints like a Boost.Python enum value"""
    _names = {}

    def __new__(cls, value, name):
        obj = int.__new__(cls, value)
        obj._name = name
        return obj

    def __repr__(self):
        return "%s.%s" % (type(self).__name__, self._name)

    __str__ = __repr__


def _enum(name, source, rename=None):
    cls = type(name, (_Enum,), {})
    for k, v in vars(source).items():
        if not k.startswith("_") and isinstance(v, int):
            setattr(cls, (rename or {}).get(k, k), cls(v, (rename or {}).get(k, k)))
    return cls


backend_t = _enum("backend_t", _L.backend_t)
kernel_t = _enum("kernel_t", _L.kernel_t, {"Long": "long"})
vt_t = _enum("vt_t", _L.vt_t)
RH_formula_t = _enum("RH_formula_t", _L.RH_formula_t)
as_t = _enum("as_t", _L.as_t)


class src_t:
    off, simple, matching = 0, 1, 2


class chem_species_t:
    HNO3, NH3, CO2, SO2, H2O2, O3, S_VI, H = range(8)


class opts_t:
    """lgrngn::opts_t (opts.hpp:19-67): process toggles of one time step"""
    _OUT_OF_SCOPE = ("src", "rlx", "chem_dsl", "chem_dsc", "chem_rct", "turb_adve", "turb_cond", "turb_coal", "ice_nucl")

    def __init__(self):
        self.adve = self.sedi = self.cond = self.coal = True
        self.subs = self.src = self.rlx = self.rcyc = False
        self.chem_dsl = self.chem_dsc = self.chem_rct = False
        self.turb_adve = self.turb_cond = self.turb_coal = self.ice_nucl = False
        self.RH_max = 44.0
        self.dt = -1.0
        self.src_dry_distros = {}
        self.src_dry_sizes = {}

    def _flat(self, library):
        o = library.opts_t()
        for k in ("adve", "sedi", "subs", "cond", "coal", "rcyc"):
            setattr(o, k, int(bool(getattr(self, k))))
        o.RH_max, o.dt = float(self.RH_max), float(self.dt)
        return o

    def _check(self):
        texts = {"src": "aerosol source was switched off in opts_init", "rlx": "aerosol relaxation was switched off in opts_init",
                 "chem_dsl": "all chemistry was switched off in opts_init", "chem_dsc": "all chemistry was switched off in opts_init",
                 "chem_rct": "all chemistry was switched off in opts_init", "turb_adve": "turb_adve_switch=False, but turb_adve==True",
                 "turb_cond": "turb_cond_swtich=False, but turb_cond==True", "turb_coal": "turb_coal_swtich=False, but turb_coal==True",
                 "ice_nucl": "ice_switch=False, but ice_nucl==True"}
        for k in self._OUT_OF_SCOPE:
            if getattr(self, k):
                raise RuntimeError("libcloudph++: " + texts[k])


class opts_init_t:
    """lgrngn::opts_init_t (opts_init.hpp:29-253): the defaults come from the library itself"""
    _EXTRA = dict(src_x0=0., src_x1=0., src_y0=0., src_y1=0., src_z0=0., src_z1=0., rlx_switch=False,
                  sstp_chem=1, supstp_rlx=1, rlx_bins=0, rlx_sd_per_bin=0, rlx_timescale=1.,
                  src_type=src_t.off, chem_rho=0., diag_incloud_time=False, time_dep_ice_nucl=False, y0=0., y1=1.)

    def __init__(self):
        object.__setattr__(self, "_flat", _library().opts_init_t())
        for k, v in self._EXTRA.items():
            if not hasattr(self._flat, k):
                object.__setattr__(self, k, v)
        object.__setattr__(self, "SGS_mix_len", [])
        object.__setattr__(self, "rlx_dry_distros", {})
        object.__setattr__(self, "_dry_distros", {})

    # everything the flat structure knows is stored there; the rest lives on this object
    def __getattr__(self, name):
        flat = object.__getattribute__(self, "_flat")
        if name == "dry_distros":
            raise RuntimeError("dry_distros does not feature a getter yet - TODO")     # lgrngn.hpp get_dd
        if hasattr(flat, name):
            v = getattr(flat, name)
            enums = {"kernel": kernel_t, "terminal_velocity": vt_t, "adve_scheme": as_t, "RH_formula": RH_formula_t}
            if name in enums:
                for cand in vars(enums[name]).values():
                    if isinstance(cand, _Enum) and int(cand) == int(v):
                        return cand
            return v
        raise AttributeError(name)

    def __setattr__(self, name, value):
        flat = object.__getattribute__(self, "_flat")
        if name == "dry_distros":
            dd = {}
            for key, fn in dict(value).items():
                kappa, rd_insol = key if isinstance(key, tuple) else (key, 0.0)
                dd[(float(kappa), float(rd_insol))] = fn
            object.__setattr__(self, "_dry_distros", dd)
        elif name == "dry_sizes":
            flat.dry_sizes = {(k if isinstance(k, tuple) else (k, 0.0)): {float(r): (float(cc[0]), int(cc[1])) for r, cc in dict(v).items()}
                              for k, v in dict(value).items()}
        elif hasattr(flat, name):
            setattr(flat, name, type(getattr(flat, name))(value) if isinstance(getattr(flat, name), (int, float)) else value)
        else:
            object.__setattr__(self, name, value)

    def _pack(self):
        flat = object.__getattribute__(self, "_flat")
        # std::map iteration order of the reference: ascending (kappa, rd_insol)
        flat.dry_distros = [_L.callable_distro(k[0], fn, k[1]) for k, fn in sorted(self._dry_distros.items())]
        refused = {"rlx_switch": "aerosol relaxation (rlx_switch)", "diag_incloud_time": "diag_incloud_time"}
        for k, what in refused.items():
            if getattr(self, k, False):
                raise RuntimeError("libcloudph++: %s is not part of the B200 back-end" % what)
        if int(getattr(self, "src_type", 0)) != src_t.off:
            raise RuntimeError("libcloudph++: aerosol sources (src_type) are not part of the B200 back-end")
        return flat


class particles_proto_t:
    """particles_proto_t<double> (particles.hpp:17-134) behind the argument conventions of lgrngn.hpp:76-260"""

    def __init__(self, backend, opts_init):
        lib = _library()
        if os.environ.get("LIBCLOUDPHXX_COMPAT_REDIRECT") == "1" and lib.name == "b200" and int(backend) in (int(backend_t.serial), int(backend_t.OpenMP)):
            backend = backend_t.CUDA
        self._lib = lib
        self._p = lib.factory(int(backend), opts_init._pack())
        self.opts_init = opts_init

    @staticmethod
    def _a(x):
        if x is None:
            return None
        a = np.asarray(x)
        if a.dtype != np.float64:
            raise TypeError("Eulerian fields must be numpy arrays of float64")
        return a

    def init(self, th=None, rv=None, rhod=None, p=None, Cx=None, Cy=None, Cz=None, ambient_chem=None):
        if ambient_chem:
            raise RuntimeError("libcloudph++: chemistry was switched off and ambient_chem is not empty")
        a = self._a
        self._p.init(a(th), a(rv), a(rhod), a(p), a(Cx), a(Cy), a(Cz))

    def step_sync(self, opts, th=None, rv=None, rhod=None, Cx=None, Cy=None, Cz=None, diss_rate=None, ambient_chem=None):
        self._refuse(diss_rate, ambient_chem)
        opts._check()
        a = self._a
        self._p.step_sync(opts._flat(self._lib), a(th), a(rv), a(rhod), a(Cx), a(Cy), a(Cz))

    def sync_in(self, th=None, rv=None, rhod=None, Cx=None, Cy=None, Cz=None, diss_rate=None, ambient_chem=None):
        self._refuse(diss_rate, ambient_chem)
        a = self._a
        self._p.sync_in(a(th), a(rv), a(rhod), a(Cx), a(Cy), a(Cz))

    def step_cond(self, opts, th=None, rv=None, ambient_chem=None):
        self._refuse(None, ambient_chem)
        opts._check()
        self._p.step_cond(opts._flat(self._lib), self._a(th), self._a(rv))

    def step_async(self, opts):
        opts._check()
        self._p.step_async(opts._flat(self._lib))

    @staticmethod
    def _refuse(diss_rate, ambient_chem):
        if ambient_chem:
            raise RuntimeError("libcloudph++: chemistry was switched off and ambient_chem is not empty")
        if diss_rate is not None:
            raise RuntimeError("libcloudph++: turbulent advection, coalescence and condesation are switched off and diss_rate is not empty")

    def outbuf(self):
        """dense per-cell buffer of the last diag_* call; `numpy.frombuffer(prtcls.outbuf())` works as with the reference"""
        return self._p.outbuf()

    def get_attr(self, name):
        return list(self._p.get_attr(name))

    def diag_puddle(self):
        return self._p.diag_puddle()

    def __getattr__(self, name):
        if name.startswith("diag_"):
            flat = getattr(object.__getattribute__(self, "_p"), name, None)
            if flat is not None:
                return flat
            raise RuntimeError("libcloudph++: %s is not part of the B200 back-end" % name)
        raise AttributeError(name)


def factory(backend, opts_init):
    return particles_proto_t(backend, opts_init)
