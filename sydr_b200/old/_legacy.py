"""ctypes prototypes of the nine legacy entry points, as the reference's callers declare them
(sydr/old/tracking/tracking_epl_c.py:31-96, sydr/old/acquisition/acquisition_pcps_c.py:32-66), bound to
libsydr_b200.so instead of ./core/c_functions/{tracking,acquisition}.so."""
from __future__ import annotations

import ctypes as C

import numpy as np
from numpy.ctypeslib import ndpointer

from .. import _lib as L

_f64 = lambda nd=None: ndpointer(C.c_double, ndim=nd, flags="C_CONTIGUOUS") if nd else ndpointer(C.c_double)  # noqa: E731
_c128 = ndpointer(np.cdouble, ndim=1, flags="C_CONTIGUOUS")
_i64 = ndpointer(C.c_longlong)
_d, _z, _ll = C.c_double, C.c_size_t, C.c_longlong

PROTOTYPES = {
    # tracking.c
    "getCorrelator": [_f64(1), _f64(1), ndpointer(C.c_int, ndim=1, flags="C_CONTIGUOUS"), _z, _d, _d, _d, _f64(), _f64()],
    "generateReplica": [_f64(1), _z, _d, _d, _f64(), _c128],
    "generateCarrier": [_c128, _c128, _z, _f64(1), _f64(1)],
    "delayLockLoop": [_d] * 10 + [_f64()] * 3,
    "phaseLockLoop": [_d] * 8 + [_f64()] * 3,
    "getLoopCoefficients": [_d] * 3 + [_f64()] * 2,
    # acquisition.c
    "setSatellite": [_f64(1), _z, _c128],
    "PCPS": [_c128, _c128, _ll, _ll, _ll, _d, _d, _f64(1), _z, _f64(2)],
    "twoCorrelationPeakComparison": [_f64(2), _z, _f64(1), _z, _ll, _ll, _d, _f64(), _f64(), _f64(), _i64, _i64, _i64],
}


class LegacyLibrary:
    """The nine functions as attributes; every call is followed by a look at sydr_last_error() (the C functions
    return void, as in the reference)."""

    def __init__(self):
        L.require_device()
        self._lib = L.load()
        for name, argtypes in PROTOTYPES.items():
            fn = getattr(self._lib, name)
            fn.argtypes, fn.restype = argtypes, None
            setattr(self, name, self._checked(name, fn))

    def _checked(self, name, fn):
        lib = self._lib

        def call(*args):
            lib.sydr_clear_error()
            fn(*args)
            msg = lib.sydr_last_error()
            msg = msg.decode() if isinstance(msg, bytes) else (msg or "")
            if msg:
                raise L.SydrError(f"{name}: {msg}")
        return call


_instance = None


def library() -> LegacyLibrary:
    global _instance
    if _instance is None:
        _instance = LegacyLibrary()
    return _instance
