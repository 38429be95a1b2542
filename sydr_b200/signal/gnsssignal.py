"""Drop-in for sydr/signal/gnsssignal.py (same names, arguments and return types); the code
table comes from the device (K-CODE), the index arithmetic of UpsampleCode is host NumPy
exactly as in the reference."""
from __future__ import annotations

import numpy as np

from .. import _lib as L

GPS_L1CA_CODE_SIZE_BITS = 1023
GPS_L1CA_CODE_FREQ = 1.023e6

_code_cache: dict[int, np.ndarray] = {}


def GenerateGPSGoldCode(prn, samplingFrequency=None):
    """sydr/signal/gnsssignal.py:9-31: +-1.0 float64 C/A code, optionally upsampled."""
    prn = int(prn)
    if prn not in _code_cache:
        L.require_device()
        out = np.empty(GPS_L1CA_CODE_SIZE_BITS, dtype=np.float64)
        L.check(L.load().sydr_ca_code(prn, out.ctypes.data), "sydr_ca_code")
        _code_cache[prn] = out
    code = _code_cache[prn].copy()
    if samplingFrequency:
        code = UpsampleCode(code, samplingFrequency)
    return code


def getSamplesPerCode(samplingFrequency: float):
    """sydr/signal/gnsssignal.py:62-70."""
    return round(samplingFrequency / (GPS_L1CA_CODE_FREQ / GPS_L1CA_CODE_SIZE_BITS))


def UpsampleCode(code, samplingFrequency: float):
    """sydr/signal/gnsssignal.py:35-58 (pure index arithmetic; stays on the host)."""
    ts = 1 / samplingFrequency
    tc = 1 / GPS_L1CA_CODE_FREQ
    n = getSamplesPerCode(samplingFrequency)
    idx = np.trunc(ts * np.array(range(n)) / tc).astype(int)
    return code[idx]


def CodeSpectrum(prn, samplingFrequency: float) -> np.ndarray:
    """conj(fft(UpsampleCode(code))) computed on the device in FP64
    (sydr/channel/channel_l1ca_borre.py:281-282)."""
    L.require_device()
    n = getSamplesPerCode(samplingFrequency)
    out = np.empty(n, dtype=np.complex128)
    L.check(L.load().sydr_code_spectrum(int(prn), float(samplingFrequency), out.ctypes.data, n), "sydr_code_spectrum")
    return out
