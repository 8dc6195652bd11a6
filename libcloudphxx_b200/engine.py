"""ctypes view of the parts of the engine C ABI (include/lcx_b200.h) that Python-side plumbing (bench.py, tests) needs:
device timers, the per-kernel profile, launch counters, cell statistics.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LCX_LIB_PATH = os.path.join(os.environ.get("LCX_B200_LIBDIR") or os.path.join(_HERE, "lib"), "liblcx_b200.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LCX_LIB_PATH):
            raise OSError("CUDA engine not built: %s is missing (there is no CPU fallback)" % LCX_LIB_PATH)
        l = C.CDLL(LCX_LIB_PATH, mode=C.RTLD_GLOBAL)
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
        _lib = l
    return _lib


COND_SOLVERS = {"secant": 0, "toms748": 1, "exact": 2}


def set_cond_solver(name):
    """root search of the condensation step, process-wide: "toms748" (default, the reference's trial points), "exact" (TOMS 748 +
    operation-by-operation growth law), "secant" (opt-in fast mode); see include/lcx_b200.h lcx_set_cond_solver"""
    lib().lcx_set_cond_solver(COND_SOLVERS[name])


def get_cond_solver():
    mode = lib().lcx_get_cond_solver()
    return [k for k, v in COND_SOLVERS.items() if v == mode][0]


def set_cond_layout(cells_per_warp):
    """work distribution of the fused per-cell condensation kernel, process-wide: 0 automatic (default), -1 eight lanes per
    cell, k in 1..16 a warp per run of k consecutive cells; see include/lcx_b200.h lcx_set_cond_layout"""
    lib().lcx_set_cond_layout(int(cells_per_warp))


def get_cond_layout():
    return int(lib().lcx_get_cond_layout())


def set_cond_staged(on):
    """phase-grouped form of the run-per-warp condensation kernel (opt-in, default off; bit-identical results either way)"""
    lib().lcx_set_cond_staged(int(bool(on)))


def get_cond_staged():
    return bool(lib().lcx_get_cond_staged())


def check(rc):
    if rc != 0:
        raise RuntimeError(lib().lcx_last_error().decode())


class Engine:
    """non-owning handle of the lcx_engine behind a single-slab particle system"""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("no engine behind this particle system")
        self.h = C.c_void_p(handle)
        self.l = lib()

    def sync(self):
        check(self.l.lcx_sync(self.h))

    def timer_start(self):
        check(self.l.lcx_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        check(self.l.lcx_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def launches(self):
        v = C.c_uint64()
        check(self.l.lcx_launch_count(self.h, C.byref(v)))
        return v.value

    def n_part(self):
        v = C.c_int64()
        check(self.l.lcx_n_part(self.h, C.byref(v)))
        return v.value

    def cell_stats(self):
        a, b = C.c_int64(), C.c_int64()
        check(self.l.lcx_cell_stats(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def coal_stats(self):
        a, b = C.c_uint64(), C.c_uint64()
        check(self.l.lcx_coal_stats(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def profile(self, on):
        check(self.l.lcx_profile_enable(self.h, int(on)))

    def profile_report(self):
        buf = C.create_string_buffer(1 << 16)
        check(self.l.lcx_profile_report(self.h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, launches, ms = line.rsplit(" ", 2)
            out[name] = (int(launches), float(ms))
        return out

    def migr_real_attrs(self):
        v = C.c_int()
        check(self.l.lcx_migr_real_attrs(self.h, C.byref(v)))
        return v.value

    def puddle(self):
        out = (C.c_double * 14)()
        check(self.l.lcx_puddle(self.h, out))
        return list(out)

    def top_loss(self):
        """(dry volume, number of super-droplets) that left through the lid since creation"""
        out = (C.c_double * 2)()
        check(self.l.lcx_top_loss(self.h, out))
        return out[0], out[1]
