"""`Acquisition` of the reference's earlier API (sydr/old/acquisition/acquisition_pcps_c.py:20-164 on top of
acquisition_pcps.py:19-58 and acquisition_abstract.py:15-74): an object per signal that is given a satellite, then
raw samples, and leaves its estimates in attributes.  The three C entry points it binds -- setSatellite, PCPS,
twoCorrelationPeakComparison -- come from libsydr_b200.so (GPU); there is no NumPy path.

Deliberate difference: the reference's acquisition.c works on a
real-input half spectrum and returns a half-width, numerically wrong map; the shims follow the Python twin
(acquisition_pcps.py:90-140), so `codeFFT` has len(code) points and the map is (bins, samplesPerCode).
"""
from __future__ import annotations

import configparser

import numpy as np

from .._legacy import library

GPS_L1CA_CODE_FREQ, GPS_L1CA_CODE_BITS = 1.023e6, 1023


def _section(gnssSignal, name):
    """[ACQUISITION] / [TRACKING] of the signal's configuration: a parsed `config`, a `configFile`, or a plain dict."""
    cfg = getattr(gnssSignal, "config", None)
    if cfg is None and getattr(gnssSignal, "configFile", None):
        cfg = configparser.ConfigParser()
        if not cfg.read(gnssSignal.configFile):
            raise FileNotFoundError(gnssSignal.configFile)
    if cfg is None:
        raise ValueError("gnssSignal carries neither `config` nor `configFile`")
    return cfg[name]


class Acquisition:
    def __init__(self, rfSignal, gnssSignal):
        self._c = library()
        self.rfSignal, self.gnssSignal = rfSignal, gnssSignal
        sec = _section(gnssSignal, "ACQUISITION")
        self.name = "PCPS"
        self.dopplerRange = float(sec["doppler_range"])
        self.dopplerSteps = float(sec["doppler_steps"])
        self.cohIntegration = int(sec["coh_integration"])
        self.nonCohIntegration = int(sec["noncoh_integration"])
        self.metricThreshold = float(sec["metric_threshold"])
        fs = float(rfSignal.samplingFrequency)
        code_freq = float(getattr(gnssSignal, "codeFrequency", GPS_L1CA_CODE_FREQ))
        code_bits = int(getattr(gnssSignal, "codeBits", GPS_L1CA_CODE_BITS))
        self.samplesPerCode = round(fs / (code_freq / code_bits))
        self.samplesPerCodeChip = round(fs / code_freq)
        self.samplingPeriod = 1 / fs
        self.frequencyBins = np.arange(-self.dopplerRange, self.dopplerRange, self.dopplerSteps)
        nan = float("nan")
        self.estimatedCode = self.estimatedFrequency = self.estimatedDoppler = self.acquisitionMetric = nan
        self.idxEstimatedCode = self.idxEstimatedFrequency = nan
        self.correlationMap = np.array([])
        self.isAcquired = False

    # ---- acquisition_abstract.py:48-74
    def setSatellite(self, svid):
        self.svid = svid
        self.code = np.ascontiguousarray(self.gnssSignal.getCode(svid, self.rfSignal.samplingFrequency), dtype=np.float64)
        spectrum = np.empty(len(self.code), dtype=np.complex128)
        self._c.setSatellite(self.code, len(self.code), spectrum)
        self.codeFFT = spectrum

    def getEstimation(self):
        return self.estimatedFrequency, self.estimatedCode

    def getCorrelationMap(self):
        return self.correlationMap

    def getMetric(self):
        return self.acquisitionMetric

    # ---- acquisition_pcps_c.py:85-164
    def run(self, rfData):
        self.correlationMap = self.PCPS(rfData)
        self.twoCorrelationPeakComparison(self.correlationMap)
        if self.acquisitionMetric > self.metricThreshold:
            self.isAcquired = True

    def PCPS(self, rfData):
        cmap = np.empty((len(self.frequencyBins), self.samplesPerCode))
        self._c.PCPS(np.ascontiguousarray(rfData, dtype=np.complex128), np.ascontiguousarray(self.codeFFT),
                     self.cohIntegration, self.nonCohIntegration, self.samplesPerCode, self.samplingPeriod,
                     float(self.rfSignal.interFrequency), np.ascontiguousarray(self.frequencyBins),
                     len(self.frequencyBins), cmap)
        return cmap

    def twoCorrelationPeakComparison(self, correlationMap):
        m = np.ascontiguousarray(np.atleast_2d(correlationMap), dtype=np.float64)
        out_f = [np.empty(1) for _ in range(3)]                      # metric, Doppler, frequency
        out_i = [np.empty(1, dtype=np.int64) for _ in range(3)]      # code, frequency index, code index
        self._c.twoCorrelationPeakComparison(m, m.shape[1], np.ascontiguousarray(self.frequencyBins), m.shape[0],
                                             self.samplesPerCode, self.samplesPerCodeChip,
                                             float(self.rfSignal.interFrequency), *out_f, *out_i)
        self.acquisitionMetric, self.estimatedDoppler, self.estimatedFrequency = (float(a[0]) for a in out_f)
        self.estimatedCode, self.idxEstimatedFrequency, self.idxEstimatedCode = (int(a[0]) for a in out_i)

    def getDatabaseDict(self):
        """acquisition_pcps.py:189-205."""
        return {"type": "acquisition", "frequency": self.estimatedFrequency, "code": self.estimatedCode,
                "frequency_idx": self.idxEstimatedFrequency, "code_idx": self.idxEstimatedCode,
                "correlation_map": self.correlationMap}
