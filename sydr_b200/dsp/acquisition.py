"""Drop-in for sydr/dsp/acquisition.py: same function names, arguments, return types and
ownership rules (new NumPy arrays / Python scalars out, inputs untouched), computed on the
GPU through libsydr_b200.so."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib as L
from ..engine import AcquisitionEngine, to_device_iq

_engine_cache: dict = {}


def _engine(fs, inter_freq, doppler_range, doppler_step, coh, noncoh) -> AcquisitionEngine:
    key = (float(fs), float(inter_freq), float(doppler_range), float(doppler_step), int(coh), int(noncoh),
           torch.cuda.current_device())
    eng = _engine_cache.get(key)
    if eng is None:
        if len(_engine_cache) > 8:
            _engine_cache.pop(next(iter(_engine_cache))).close()
        eng = AcquisitionEngine(fs, inter_freq, doppler_range, doppler_step, coh, noncoh, [1])
        _engine_cache[key] = eng
    return eng


def PCPS(rfData: np.array, interFrequency: float, samplingFrequency: float, codeFFT: np.array, dopplerRange: tuple,
         dopplerStep: int, samplesPerCode: int, coherentIntegration: int = 1, nonCoherentIntegration: int = 1):
    """sydr/dsp/acquisition.py:9-74.  Returns the float64 (bins, samplesPerCode) correlation map."""
    L.require_device()
    rf = np.squeeze(np.asarray(rfData))
    eng = _engine(samplingFrequency, interFrequency, dopplerRange, dopplerStep, coherentIntegration,
                  nonCoherentIntegration)
    if eng.n_code != int(samplesPerCode):
        raise L.SydrError(f"samplesPerCode {samplesPerCode} does not match round(fs/1000) = {eng.n_code}")
    need = eng.required_samples
    if rf.shape[0] < need:
        raise L.SydrError(f"PCPS needs {need} samples, got {rf.shape[0]}")
    if rf.dtype not in (np.complex64, np.complex128):
        rf = rf.astype(np.complex128)
    eng.set_spectrum(0, np.asarray(codeFFT))
    iq = to_device_iq(rf[:need])
    res = eng.run(iq, want_maps=True)
    cmap = res["maps"][0].astype(np.float64)
    return np.squeeze(np.squeeze(cmap))


def TwoCorrelationPeakComparison(correlationMap: np.array, samplesPerCode: int, samplesPerCodeChip: int):
    """sydr/dsp/acquisition.py:78-115.  Returns ([freq_idx, code_idx], peak1/peak2)."""
    L.require_device()
    m = np.ascontiguousarray(np.atleast_2d(correlationMap), dtype=np.float64)
    import ctypes as C
    fi, ci, ratio = C.c_int(), C.c_int(), C.c_double()
    L.check(L.load().sydr_peak_compare(m.ctypes.data, m.shape[0], m.shape[1], int(samplesPerCodeChip),
                                       C.byref(fi), C.byref(ci), C.byref(ratio)), "sydr_peak_compare")
    return [int(fi.value), int(ci.value)], float(ratio.value)
