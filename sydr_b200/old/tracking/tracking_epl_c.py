"""`Tracking` of the reference's earlier API (sydr/old/tracking/tracking_epl_c.py:17-240 on top of
tracking_epl.py:17-160 and tracking_abstract.py:39-150): one object per channel, `run(rfData)` consumes one code
period and leaves correlators, loop outputs and the NCO state in attributes.  The six C entry points it binds --
generateReplica, generateCarrier, getCorrelator, delayLockLoop, phaseLockLoop, getLoopCoefficients -- come from
libsydr_b200.so (csrc/legacy.cu, GPU kernels with the operation order of sydr/c_functions/tracking.c)."""
from __future__ import annotations

import numpy as np

from .._legacy import library
from ..acquisition.acquisition_pcps_c import GPS_L1CA_CODE_BITS, GPS_L1CA_CODE_FREQ, _section


class Tracking:
    def __init__(self, rfSignal, gnssSignal):
        self._c = library()
        self.rfSignal, self.gnssSignal = rfSignal, gnssSignal
        sec = _section(gnssSignal, "TRACKING")
        self.pdiCode, self.pdiCarrier = float(sec["pdi_code"]), float(sec["pdi_carrier"])
        self.correlatorSpacing = [float(sec[f"correlator_{i}"]) for i in range(int(sec["correlator_number"]))]
        self.correlatorPrompt = int(sec["correlator_prompt"])
        for loop in ("dll", "pll"):
            for key, attr in (("dumping_ratio", "DumpingRatio"), ("noise_bandwidth", "NoiseBandwidth"), ("loop_gain", "LoopGain")):
                setattr(self, loop + attr, float(sec[f"{loop}_{key}"]))
        self.remCodePhase = self.remCarrierPhase = 0.0
        self.codeNCO = self.codeError = self.carrierNCO = self.carrierError = 0.0
        self.initialFrequency = self.carrierFrequency = 0.0
        self.code, self.iSignal, self.qSignal = [], [], []
        self.dllTau1, self.dllTau2 = self.getLoopCoefficients(self.dllNoiseBandwidth, self.dllDumpingRatio, self.dllLoopGain)
        self.pllTau1, self.pllTau2 = self.getLoopCoefficients(self.pllNoiseBandwidth, self.pllDumpingRatio, self.pllLoopGain)
        self._code_freq0 = float(getattr(gnssSignal, "codeFrequency", GPS_L1CA_CODE_FREQ))
        self._code_bits = int(getattr(gnssSignal, "codeBits", GPS_L1CA_CODE_BITS))
        self.codeFrequency = self._code_freq0
        self._new_epoch_shape()
        self.correlatorResults = [0.0] * 6
        self.pll = self.dll = 0.0
        self.time = np.arange(0, self.samplesRequired + 2) / self.rfSignal.samplingFrequency

    def _new_epoch_shape(self):
        """tracking_epl_c.py:162-163: step and length of the next code period."""
        self.codePhaseStep = self.codeFrequency / self.rfSignal.samplingFrequency
        self.samplesRequired = int(np.ceil((self._code_bits - self.remCodePhase) / self.codePhaseStep))

    # ---- tracking_abstract.py:92-150
    def setInitialValues(self, estimatedFrequency):
        self.initialFrequency = self.carrierFrequency = estimatedFrequency

    def setSatellite(self, svid):
        self.svid = svid
        code = np.asarray(self.gnssSignal.getCode(svid))
        self.code = np.ascontiguousarray(np.r_[code[-1], code, code[0]].astype(np.int32))    # int for the C side (:169-172)

    def getSamplesRequired(self):
        return self.samplesRequired

    def getCorrelatorResults(self):
        return self.correlatorResults

    def getCarrierFrequency(self):
        return self.carrierFrequency

    def getCodeFrequency(self):
        return self.codeFrequency

    def getDLL(self):
        return self.dll

    def getPLL(self):
        return self.pll

    def getPrompt(self):
        return self.correlatorResults[2], self.correlatorResults[3]

    # ---- tracking_epl_c.py:101-240
    def getCorrelator(self, correlatorSpacing):
        i_out, q_out = np.empty(1), np.empty(1)
        self._c.getCorrelator(np.ascontiguousarray(self.iSignal), np.ascontiguousarray(self.qSignal), self.code,
                              self.samplesRequired, self.codePhaseStep, self.remCodePhase, correlatorSpacing, i_out, q_out)
        return i_out[0], q_out[0]

    def generateReplica(self):
        rem, replica = np.empty(1), np.empty(self.samplesRequired, dtype=np.complex128)
        self._c.generateReplica(np.ascontiguousarray(self.time[:self.samplesRequired + 1]), self.samplesRequired,
                                self.carrierFrequency, self.remCarrierPhase, rem, replica)
        self.remCarrierPhase = rem[0]
        return replica

    def run(self, rfData):
        replica = self.generateReplica()
        rf = np.ascontiguousarray(rfData, dtype=np.complex128)
        self.iSignal, self.qSignal = np.empty(len(replica)), np.empty(len(replica))
        self._c.generateCarrier(rf, replica, len(rf), self.iSignal, self.qSignal)
        taps = [self.getCorrelator(sp) for sp in self.correlatorSpacing[:3]]
        self.correlatorResults = [v for tap in taps for v in tap]
        self.delayLockLoop(*taps[0], *taps[2])
        self.phaseLockLoop(*taps[1])
        # code phase left over at the end of the period (np.linspace arithmetic of the reference, :156-158)
        n = self.samplesRequired
        idx = np.linspace(self.remCodePhase, n * self.codePhaseStep + self.remCodePhase, n, endpoint=False)
        self.remCodePhase = idx[n - 1] + self.codePhaseStep - self._code_bits
        self._new_epoch_shape()

    def delayLockLoop(self, iEarly, qEarly, iLate, qLate):
        nco, err, freq = np.empty(1), np.empty(1), np.empty(1)
        self._c.delayLockLoop(iEarly, qEarly, iLate, qLate, self.dllTau1, self.dllTau2, self.pdiCode, self.codeNCO,
                              self.codeError, self._code_freq0, nco, err, freq)
        self.codeNCO, self.codeError, self.codeFrequency = nco[0], err[0], freq[0]
        self.dll = self.codeNCO

    def phaseLockLoop(self, iPrompt, qPrompt):
        nco, err, freq = np.empty(1), np.empty(1), np.empty(1)
        self._c.phaseLockLoop(iPrompt, qPrompt, self.pllTau1, self.pllTau2, self.pdiCarrier, self.carrierNCO,
                              self.carrierError, self.initialFrequency, nco, err, freq)
        self.carrierNCO, self.carrierError, self.carrierFrequency = nco[0], err[0], freq[0]
        self.pll = self.carrierNCO

    def getLoopCoefficients(self, loopNoiseBandwidth, dumpingRatio, loopGain):
        tau1, tau2 = np.empty(1), np.empty(1)
        self._c.getLoopCoefficients(loopNoiseBandwidth, dumpingRatio, loopGain, tau1, tau2)
        return tau1[0], tau2[0]

    def getDatabaseDict(self):
        """tracking_epl.py:143-160."""
        r = self.correlatorResults
        return {"type": "tracking", "i_early": r[0], "q_early": r[1], "i_prompt": r[2], "q_prompt": r[3], "i_late": r[4],
                "q_late": r[5], "dll": self.dll, "pll": self.pll, "carrier_frequency": self.carrierFrequency,
                "code_frequency": self.codeFrequency}
