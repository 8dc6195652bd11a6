"""ctypes view of the parts of the engine C ABI (include/lcx_b200.h) that Python-side plumbing (bench.py, tests) needs:
device timers, the per-kernel profile, launch counters, cell statistics.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBDIR = os.environ.get("LCX_B200_LIBDIR") or os.path.join(_HERE, "lib")
LCX_LIB_PATH = os.path.join(_LIBDIR, "liblcx_b200.so")
LCX_F32_LIB_PATH = os.path.join(_LIBDIR, "liblcx_b200_f32.so")
_lib = None
_lib_f32 = None


class _Suffixed:
    """the single-precision engine exports the same entry points with the suffix _f32 (include/lcx_b200_f32_names.h)"""

    def __init__(self, cdll):
        self._cdll = cdll

    def __getattr__(self, name):
        return getattr(self._cdll, name + "_f32")


def lib(real="f64"):
    global _lib, _lib_f32
    if real == "f32":
        if _lib_f32 is None:
            if not os.path.exists(LCX_F32_LIB_PATH):
                raise OSError("single-precision CUDA engine not built: %s is missing (there is no CPU fallback)" % LCX_F32_LIB_PATH)
            _lib_f32 = _bind(_Suffixed(C.CDLL(LCX_F32_LIB_PATH, mode=C.RTLD_GLOBAL)))
        return _lib_f32
    if _lib is None:
        if not os.path.exists(LCX_LIB_PATH):
            raise OSError("CUDA engine not built: %s is missing (there is no CPU fallback)" % LCX_LIB_PATH)
        _lib = _bind(C.CDLL(LCX_LIB_PATH, mode=C.RTLD_GLOBAL))
    return _lib


def _bind(l):
    if l is not None:
        l.lcx_last_error.restype = C.c_char_p
        l.lcx_version.restype = C.c_char_p
        l.lcx_timer_start.argtypes = [C.c_void_p]
        l.lcx_timer_stop.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        l.lcx_profile_enable.argtypes = [C.c_void_p, C.c_int]
        l.lcx_profile_report.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
        l.lcx_launch_count.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        l.lcx_n_part.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        l.lcx_sync.argtypes = [C.c_void_p]
        l.lcx_cell_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        l.lcx_migr_real_attrs.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        l.lcx_puddle.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        l.lcx_top_loss.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        l.lcx_coal_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    return l


COND_SOLVERS = {"secant": 0, "toms748": 1, "exact": 2}


def set_cond_solver(name):
    """root search of the condensation step, process-wide: "toms748" (default, the reference's trial points), "exact" (TOMS 748 +
    operation-by-operation growth law), "secant" (opt-in fast mode); see include/lcx_b200.h lcx_set_cond_solver"""
    lib().lcx_set_cond_solver(COND_SOLVERS[name])


def get_cond_solver():
    mode = lib().lcx_get_cond_solver()
    return [k for k, v in COND_SOLVERS.items() if v == mode][0]


def set_cond_layout(cells_per_warp, real="f64"):
    """work distribution of the fused per-cell condensation kernel, process-wide (per engine library: real = "f32" addresses the
    single-precision one): 0 automatic (default), -1 eight lanes per cell, k in 1..16 a warp per run of k consecutive cells;
    see include/lcx_b200.h lcx_set_cond_layout"""
    lib(real).lcx_set_cond_layout(int(cells_per_warp))


def get_cond_layout():
    return int(lib().lcx_get_cond_layout())


def set_cond_classed(mode, real="f64"):
    """order in which the run-per-warp condensation kernel walks a run's droplets: -1 automatic (drizzle / rain drops apart from the
    rest once the previous step counted enough of them), 0 storage order always, 1 class by class always; include/lcx_b200.h"""
    lib(real).lcx_set_cond_classed(int(mode))


def get_cond_classed(real="f64"):
    return int(lib(real).lcx_get_cond_classed())


def set_cond_staged(on):
    """phase-grouped form of the run-per-warp condensation kernel (opt-in, default off; bit-identical results either way)"""
    lib().lcx_set_cond_staged(int(bool(on)))


def get_cond_staged():
    return bool(lib().lcx_get_cond_staged())


def check(rc, real="f64"):
    if rc != 0:
        raise RuntimeError(lib(real).lcx_last_error().decode())


class Engine:
    """non-owning handle of the lcx_engine behind a single-slab particle system (real = "f32": an engine of liblcx_b200_f32.so)"""

    def __init__(self, handle, real="f64"):
        if not handle:
            raise RuntimeError("no engine behind this particle system")
        self.h = C.c_void_p(handle)
        self.real = real
        self.l = lib(real)

    def sync(self):
        check(self.l.lcx_sync(self.h), self.real)

    def timer_start(self):
        check(self.l.lcx_timer_start(self.h), self.real)

    def timer_stop(self):
        ms = C.c_float()
        check(self.l.lcx_timer_stop(self.h, C.byref(ms)), self.real)
        return ms.value

    def launches(self):
        v = C.c_uint64()
        check(self.l.lcx_launch_count(self.h, C.byref(v)), self.real)
        return v.value

    def n_part(self):
        v = C.c_int64()
        check(self.l.lcx_n_part(self.h, C.byref(v)), self.real)
        return v.value

    def cell_stats(self):
        a, b = C.c_int64(), C.c_int64()
        check(self.l.lcx_cell_stats(self.h, C.byref(a), C.byref(b)), self.real)
        return a.value, b.value

    def coal_stats(self):
        a, b = C.c_uint64(), C.c_uint64()
        check(self.l.lcx_coal_stats(self.h, C.byref(a), C.byref(b)), self.real)
        return a.value, b.value

    def profile(self, on):
        check(self.l.lcx_profile_enable(self.h, int(on)), self.real)

    def profile_report(self):
        buf = C.create_string_buffer(1 << 16)
        check(self.l.lcx_profile_report(self.h, buf, len(buf)), self.real)
        out = {}
        for line in buf.value.decode().splitlines():
            name, launches, ms = line.rsplit(" ", 2)
            out[name] = (int(launches), float(ms))
        return out

    def migr_real_attrs(self):
        v = C.c_int()
        check(self.l.lcx_migr_real_attrs(self.h, C.byref(v)), self.real)
        return v.value

    def puddle(self):
        out = (C.c_double * 14)()
        check(self.l.lcx_puddle(self.h, out), self.real)
        return list(out)

    def top_loss(self):
        """(dry volume, number of super-droplets) that left through the lid since creation"""
        out = (C.c_double * 2)()
        check(self.l.lcx_top_loss(self.h, out), self.real)
        return out[0], out[1]
