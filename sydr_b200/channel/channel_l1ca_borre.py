"""GPS L1 C/A channel with the interface of sydr/channel/channel_l1ca_borre.py (ChannelL1CA).

Two ways to run it, same packets either way:
  * stand-alone, like the reference's own channel process: `_processHandler()` once per
    millisecond; `runAcquisition` / `runTracking` call the drop-in `PCPS`,
    `TwoCorrelationPeakComparison` and `EPL` (one GPU call each) and close the loops on the host
    with the reference's scalar expressions;
  * under `ChannelManager` (the fast path): the manager acquires and tracks all channels in
    batched GPU launches and feeds the results back through `_ingestAcquisition` /
    `_ingestEpoch`, which perform the reference's per-epoch bookkeeping (prompt history, bit
    synchronisation, flags, counters, ring-buffer indices, navigation-bit decoding).
"""
from __future__ import annotations

import logging

import numpy as np

from ..dsp.acquisition import PCPS, TwoCorrelationPeakComparison
from ..dsp.decoding import LNAV_CheckPreambule, LNAV_DecodeTOW, Prompt2Bit
from ..dsp.tracking import EPL, BorreLoopFilter, DLL_NNEML, LoopFiltersCoefficients, PLL_costa
from ..signal.gnsssignal import GenerateGPSGoldCode, UpsampleCode
from ..signal.rfsignal import RFSignal
from ..utils.circularbuffer import CircularBuffer
from ..utils.constants import (GPS_L1CA_CODE_FREQ, GPS_L1CA_CODE_MS, GPS_L1CA_CODE_SIZE_BITS, LNAV_MS_PER_BIT,
                               LNAV_SUBFRAME_SIZE, LNAV_WORD_SIZE)
from ..utils.enumerations import ChannelMessage, ChannelState, GNSSSignalType, GNSSSystems, TrackingFlags
from .channel import Channel, ChannelStatus
from .lnav_frame import advance_frame


class ChannelL1CA(Channel):
    MIN_CONVERGENCE_TIME = 100   # ms given to the loops before bit synchronisation is checked (L30)

    def __init__(self, cid: int, sharedBuffer: CircularBuffer, resultQueue, rfSignal: RFSignal, configuration: dict):
        super().__init__(cid, sharedBuffer, resultQueue, rfSignal, configuration)
        # channel_l1ca_borre.py:104-143
        self.codeOffset = 0
        self.codeFrequency = GPS_L1CA_CODE_FREQ
        self.carrierFrequency = 0.0
        self.initialFrequency = 0.0
        self.NCO_code = 0.0
        self.NCO_codeError = 0.0
        self.NCO_remainingCode = 0.0
        self.NCO_carrier = 0.0
        self.NCO_carrierError = 0.0
        self.NCO_remainingCarrier = 0.0
        self.fll = 0.0
        self.fll_vel_memory = 0.0
        self.fll_acc_memory = 0.0
        self.nbPrompt = 0
        self.iPrompt = 0.0
        self.qPrompt = 0.0
        self.iPrompt_sum = 0.0
        self.qPrompt_sum = 0.0
        self.iPrompt_sum2 = 0.0
        self.qPrompt_sum2 = 0.0
        self.codeCounter = 0
        self.navBitBufferSize = LNAV_SUBFRAME_SIZE + 2 * LNAV_WORD_SIZE + 2
        self.navBitsBuffer = np.squeeze(np.empty((1, self.navBitBufferSize), dtype=int))
        self.navBitsCounter = 0
        self.subframeFlags = [False, False, False, False, False]
        self.tow = 0
        self.preambuleFound = False
        self.navPromptSum = 0.0
        self.navPromptSumCounter = 0
        self.setAcquisition(configuration['ACQUISITION'])
        self.setTracking(configuration['TRACKING'])

    # --------------------------------------------------------------------------------------------
    def setSatellite(self, satelliteID):
        """channel_l1ca_borre.py:149-175: code padded with the previous / next chip."""
        super().setSatellite(satelliteID)
        self.systemID = GNSSSystems.GPS
        self.signalID = GNSSSignalType.GPS_L1_CA
        code = GenerateGPSGoldCode(satelliteID)
        self.code = np.r_[code[-1], code, code[0]]

    def setAcquisition(self, configuration: dict):
        """channel_l1ca_borre.py:179-203."""
        self.acq_dopplerRange = float(configuration['doppler_range'])
        self.acq_dopplerSteps = float(configuration['doppler_steps'])
        self.acq_coherentIntegration = int(configuration['coherent_integration'])
        self.acq_nonCoherentIntegration = int(configuration['non_coherent_integration'])
        self.acq_threshold = float(configuration['threshold'])
        self.acq_requiredSamples = int(self.rfSignal.samplingFrequency * 1e-3 *
                                       self.acq_nonCoherentIntegration * self.acq_coherentIntegration)

    def setTracking(self, configuration: dict):
        """channel_l1ca_borre.py:207-259."""
        self.track_correlatorsSpacing = [float(configuration['correlator_early']),
                                         float(configuration['correlator_prompt']),
                                         float(configuration['correlator_late'])]
        self.IDX_I_EARLY, self.IDX_Q_EARLY, self.IDX_I_PROMPT = 0, 1, 2
        self.IDX_Q_PROMPT, self.IDX_I_LATE, self.IDX_Q_LATE = 3, 4, 5
        self.track_dll_tau1, self.track_dll_tau2 = LoopFiltersCoefficients(
            loopNoiseBandwidth=float(configuration['dll_noise_bandwidth']),
            dampingRatio=float(configuration['dll_damping_ratio']),
            loopGain=float(configuration['dll_loop_gain']))
        self.track_pll_tau1, self.track_pll_tau2 = LoopFiltersCoefficients(
            loopNoiseBandwidth=float(configuration['pll_noise_bandwidth']),
            dampingRatio=float(configuration['pll_damping_ratio']),
            loopGain=float(configuration['pll_loop_gain']))
        self.track_fll_tau1, self.track_fll_tau2 = LoopFiltersCoefficients(
            loopNoiseBandwidth=float(configuration['fll_noise_bandwidth']),
            dampingRatio=float(configuration['fll_damping_ratio']),
            loopGain=float(configuration['fll_loop_gain']))
        self.track_dll_pdi = float(configuration['dll_pdi'])
        self.track_pll_pdi = float(configuration['pll_pdi'])
        self.track_fll_pdi = float(configuration['fll_pdi'])
        self.fll_noise_bandwidth = float(configuration['fll_noise_bandwidth'])
        self.pll_noise_bandwidth = float(configuration['pll_noise_bandwidth'])
        self._trackingConfiguration = {k: float(configuration[k]) for k in (
            'correlator_early', 'correlator_prompt', 'correlator_late', 'dll_damping_ratio', 'dll_noise_bandwidth',
            'dll_loop_gain', 'dll_pdi', 'pll_damping_ratio', 'pll_noise_bandwidth', 'pll_loop_gain', 'pll_pdi')}
        self.codeStep = GPS_L1CA_CODE_FREQ / self.rfSignal.samplingFrequency
        self.track_requiredSamples = int(np.ceil((GPS_L1CA_CODE_SIZE_BITS - self.NCO_remainingCode) / self.codeStep))
        self.trackFlags = TrackingFlags.UNKNOWN
        self.maxSizeCorrelatorBuffer = LNAV_MS_PER_BIT
        self.correlatorsBuffer = np.empty((self.maxSizeCorrelatorBuffer, len(self.track_correlatorsSpacing) * 2))
        self.correlatorsBuffer[:, :] = 0.0

    # ---- acquisition -----------------------------------------------------------------------------
    def runAcquisition(self):
        """channel_l1ca_borre.py:263-329 (stand-alone path: one PCPS + one peak search on the GPU)."""
        if self.rfBuffer.getNbUnreadSamples(self.currentSample) < self.acq_requiredSamples:
            return
        code = UpsampleCode(self.code[1:-1], self.rfSignal.samplingFrequency)
        codeFFT = np.conj(np.fft.fft(code))
        samplesPerCode = round(self.rfSignal.samplingFrequency * GPS_L1CA_CODE_SIZE_BITS / GPS_L1CA_CODE_FREQ)
        samplesPerCodeChip = round(self.rfSignal.samplingFrequency / GPS_L1CA_CODE_FREQ)
        correlationMap = PCPS(rfData=self.rfBuffer.getSlice(self.currentSample, self.acq_requiredSamples),
                              interFrequency=self.rfSignal.interFrequency,
                              samplingFrequency=self.rfSignal.samplingFrequency,
                              codeFFT=codeFFT,
                              dopplerRange=self.acq_dopplerRange,
                              dopplerStep=self.acq_dopplerSteps,
                              samplesPerCode=samplesPerCode,
                              coherentIntegration=self.acq_coherentIntegration,
                              nonCoherentIntegration=self.acq_nonCoherentIntegration)
        indices, peakRatio = TwoCorrelationPeakComparison(correlationMap=correlationMap,
                                                          samplesPerCode=samplesPerCode,
                                                          samplesPerCodeChip=samplesPerCodeChip)
        return self._ingestAcquisition(indices, peakRatio, correlationMap)

    def _ingestAcquisition(self, indices, peakRatio, correlationMap):
        """Hand-off scalars and result packet, channel_l1ca_borre.py:301-327."""
        dopplerShift = -((-self.acq_dopplerRange) + self.acq_dopplerSteps * indices[0])
        self.codeOffset = int(np.round(indices[1]))
        self.carrierFrequency = self.rfSignal.interFrequency + dopplerShift
        self.initialFrequency = self.rfSignal.interFrequency + dopplerShift
        self.currentSample = self.currentSample + self.acq_requiredSamples
        self.currentSample -= self.track_requiredSamples
        self.currentSample += self.codeOffset + 1
        self.channelState = ChannelState.TRACKING
        results = self.prepareResultsAcquisition()
        results["carrierFrequency"] = self.carrierFrequency
        results["codeOffset"] = self.codeOffset
        results["frequency_idx"] = indices[0]
        results["code_idx"] = indices[1]
        results["correlation_map"] = correlationMap
        results["peak_ratio"] = peakRatio
        return results

    # ---- tracking --------------------------------------------------------------------------------
    def runTracking(self):
        """channel_l1ca_borre.py:333-451 (stand-alone path: EPL on the GPU, loops on the host)."""
        if self.rfBuffer.getNbUnreadSamples(self.currentSample) < self.track_requiredSamples:
            return
        fs = self.rfSignal.samplingFrequency
        n = self.track_requiredSamples
        corr = EPL(rfData=self.rfBuffer.getSlice(self.currentSample, n), code=self.code, samplingFrequency=fs,
                   carrierFrequency=self.carrierFrequency, remainingCarrier=self.NCO_remainingCarrier,
                   remainingCode=self.NCO_remainingCode, codeStep=self.codeStep,
                   correlatorsSpacing=self.track_correlatorsSpacing)
        remCarrier = self.NCO_remainingCarrier - self.carrierFrequency * 2.0 * np.pi * n / fs      # L364
        remCarrier %= (2 * np.pi)                                                                   # L365
        codeError = DLL_NNEML(iEarly=corr[0], qEarly=corr[1], iLate=corr[4], qLate=corr[5])
        ncoCode = BorreLoopFilter(codeError, self.NCO_codeError, self.track_dll_tau1, self.track_dll_tau2,
                                  self.track_dll_pdi)
        phaseError = PLL_costa(iPrompt=corr[2], qPrompt=corr[3])
        ncoCarrier = BorreLoopFilter(phaseError, self.NCO_carrierError, self.track_pll_tau1, self.track_pll_tau2,
                                     self.track_pll_pdi)
        codeFrequency = self.codeFrequency - ncoCode                                                # L422
        carrierFrequency = self.carrierFrequency + ncoCarrier                                       # L423
        remCode = self.NCO_remainingCode + (n * self.codeStep - GPS_L1CA_CODE_SIZE_BITS)            # L424
        return self._ingestEpoch(corr, ncoCode, ncoCarrier, carrierFrequency, codeFrequency, codeError, phaseError,
                                 remCode, remCarrier)

    def _ingestEpoch(self, corr, ncoCode, ncoCarrier, carrierFrequency, codeFrequency, codeError, phaseError,
                     remCode, remCarrier):
        """Everything runTracking does around the DSP (channel_l1ca_borre.py:367-451): prompt history,
        bit synchronisation, flags and counters, NCO members, ring index, result packet.  The loop
        outputs come from the host (stand-alone path) or from the device records (batched path)."""
        normalisedPower = np.nan
        self.NCO_remainingCarrier = remCarrier
        self.correlatorsBuffer[self.nbPrompt, :] = corr[:]
        self.iPrompt_sum += corr[2]
        self.qPrompt_sum += corr[3]
        self.iPrompt_sum2 += corr[2]
        self.qPrompt_sum2 += corr[3]
        self.nbPrompt += 1
        self.NCO_code = ncoCode
        self.NCO_codeError = codeError
        self.NCO_carrier = ncoCarrier
        self.NCO_carrierError = phaseError
        iPrompt, qPrompt = corr[2], corr[3]
        if not (self.trackFlags & TrackingFlags.BIT_SYNC):
            # bit inversion seen after the convergence time -> bit synchronisation (L401-407)
            if (self.trackFlags & TrackingFlags.CODE_LOCK) and (self.codeCounter > self.MIN_CONVERGENCE_TIME) \
                    and np.sign(self.iPrompt) != np.sign(iPrompt):
                self.trackFlags |= TrackingFlags.BIT_SYNC
                self.resetPrompt()
        else:
            if self.nbPrompt == LNAV_MS_PER_BIT:
                normalisedPower = 0.0                                                               # L413
        self.trackFlags |= TrackingFlags.CODE_LOCK
        self.iPrompt = iPrompt
        self.qPrompt = qPrompt
        self.codeCounter += 1
        self.codeSinceTOW += 1
        self.codeFrequency = codeFrequency
        self.carrierFrequency = carrierFrequency
        self.NCO_remainingCode = remCode
        self.codeStep = self.codeFrequency / self.rfSignal.samplingFrequency                        # L425
        self.currentSample = (self.currentSample + self.track_requiredSamples) % self.rfBuffer.maxSize   # L428
        self.track_requiredSamples = int(np.ceil((GPS_L1CA_CODE_SIZE_BITS - self.NCO_remainingCode) / self.codeStep))
        results = self.prepareResultsTracking()
        results["i_early"] = corr[0]
        results["q_early"] = corr[1]
        results["i_prompt"] = corr[2]
        results["q_prompt"] = corr[3]
        results["i_late"] = corr[4]
        results["q_late"] = corr[5]
        results["dll"] = self.NCO_code
        results["pll"] = self.NCO_carrier
        results["fll"] = self.fll
        results["carrier_frequency"] = self.carrierFrequency
        results["code_frequency"] = self.codeFrequency
        results["cn0"] = normalisedPower
        results["pll_lock"] = 0.0
        results["fll_lock"] = 0.0
        results["lock_state"] = 0
        results["carrier_frequency_error"] = self.NCO_carrierError
        results["code_frequency_error"] = self.NCO_codeError
        return results

    # ---- decoding --------------------------------------------------------------------------------
    def runDecoding(self):
        """channel_l1ca_borre.py:455-573: 20 prompts -> one bit, preamble search, TOW."""
        if not (self.trackFlags & TrackingFlags.BIT_SYNC):
            self.navPromptSum = 0.0
            self.navPromptSumCounter = 0
            return
        self.navPromptSum += self.correlatorsBuffer[self.nbPrompt - 1, self.IDX_I_PROMPT]
        self.navPromptSumCounter += 1
        if not (self.navPromptSumCounter == LNAV_MS_PER_BIT):
            return
        self.navBitsBuffer[self.navBitsCounter] = Prompt2Bit(self.navPromptSum)
        self.navBitsCounter += 1
        self.navPromptSum = 0.0
        self.navPromptSumCounter = 0
        decoded = advance_frame(self)                     # preamble search / subframe sync, L493-533
        if decoded is None:
            return
        tow, subframeID, subframeBits = decoded
        self.subframeFlags[subframeID - 1] = True
        self.codeSinceTOW = 0
        self.trackFlags |= TrackingFlags.TOW_DECODED
        self.trackFlags |= TrackingFlags.TOW_KNOWN
        if not (self.trackFlags & TrackingFlags.EPH_DECODED) and all(self.subframeFlags[0:3]):
            self.trackFlags |= TrackingFlags.EPH_DECODED
            self.trackFlags |= TrackingFlags.EPH_KNOWN
        results = self.prepareResultsDecoding()
        results["type"] = ChannelMessage.DECODING_UPDATE
        results["subframe_id"] = subframeID
        results["tow"] = tow
        results["bits"] = subframeBits
        self.tow = tow
        self.tow += self.navBitsCounter * LNAV_MS_PER_BIT * 1e-3
        logging.getLogger(__name__).debug(f"CID {self.channelID} subframe {subframeID} decoded "
                                          f"(TOW: {tow}, current: {self.tow}).")
        return results

    def resetPrompt(self):
        self.iPrompt_sum = 0.0
        self.qPrompt_sum = 0.0
        self.iPrompt_sum2 = 0.0
        self.qPrompt_sum2 = 0.0
        self.nbPrompt = 0

    # --------------------------------------------------------------------------------------------
    def _processHandler(self):
        """channel_l1ca_borre.py:595-629."""
        _results = []
        if self.channelState == ChannelState.IDLE:
            raise Warning(f"Tracking channel {self.channelID} is in IDLE.")
        elif self.channelState == ChannelState.ACQUIRING:
            _results.append(self.runAcquisition())
        elif self.channelState == ChannelState.TRACKING:
            _results.append(self.runTracking())
            _results.append(self.runDecoding())
        else:
            raise ValueError(f"Channel state {self.channelState} is not valid.")
        results = [i for i in _results if i is not None]
        self._afterTick()
        return results

    def _afterTick(self):
        if self.nbPrompt == self.maxSizeCorrelatorBuffer:          # L626-627
            self.resetPrompt()

    def getTimeSinceTOW(self):
        """channel_l1ca_borre.py:633-653: milliseconds since the last decoded TOW."""
        timeSinceTOW = 0
        timeSinceTOW += self.codeSinceTOW * GPS_L1CA_CODE_MS
        timeSinceTOW += self.rfBuffer.getNbUnreadSamples(self.currentSample) / (self.rfSignal.samplingFrequency / 1e3)
        return timeSinceTOW

    def prepareResultsAcquisition(self):
        mdict = super().prepareResults()
        mdict["type"] = ChannelMessage.ACQUISITION_UPDATE
        return mdict

    def prepareResultsTracking(self):
        mdict = super().prepareResults()
        mdict["type"] = ChannelMessage.TRACKING_UPDATE
        return mdict

    def prepareResultsDecoding(self):
        mdict = super().prepareResults()
        mdict["type"] = ChannelMessage.DECODING_UPDATE
        return mdict


class ChannelStatusL1CA(ChannelStatus):
    """channel_l1ca_borre.py:746-766."""

    def __init__(self, channelID: int, satelliteID: int):
        super().__init__(channelID, satelliteID)
        self.subframeFlags = [False, False, False, False, False]
