"""Drop-in stand-in for the reference's Python package `libcloudphxx` (bindings/python/lib.cpp), restricted to the
Lagrangian scheme and the `common` helpers its scripts use:

    import sys; sys.path.insert(0, ".../libcloudphxx_b200/compat")
    from libcloudphxx import lgrngn, common

Same names, argument order, defaults and error texts as the Boost.Python module (lib.cpp:217-434, lgrngn.hpp:41-330), on
top of the flat C binding (`bindings/lgrngn_capi.h`).  Environment:
  LIBCLOUDPHXX_COMPAT_LIBRARY = /path/to/lib.so           another shared library exporting the same flat binding
                                                          (bindings/lgrngn_capi.h) instead of the B200 back-end
  LIBCLOUDPHXX_COMPAT_REDIRECT = 1                        scripts written for backend_t.serial / OpenMP get the CUDA
                                                          back-end instead of "backend was not compiled"
The bulk schemes (blk_1m, blk_2m) are not part of this package.
"""
from . import common, lgrngn          # noqa: F401

__all__ = ["common", "lgrngn"]
