"""GPS L1 C/A channel with the interface of sydr/channel/channel_l1ca_kaplan.py
(ChannelL1CA_Kaplan): FLL-assisted PLL, lock indicators, PULL_IN / WIDE_TRACK / NARROW_TRACK
state machine, C/N0.

The correlators are the same K-TRK kernel as the Borre channel's (the drop-in `EPL`, one GPU call
per epoch; acquisition is the drop-in `PCPS` + `TwoCorrelationPeakComparison`); what differs is
the scalar loop closure, which this class performs on the host with the reference's expressions
in the reference's order (one epoch per 1 ms tick, channel_l1ca_kaplan.py:342-368).  Selected in
the reference by swapping one import in sydr/receiver/receiver_gps_l1ca.py:17-19.
"""
from __future__ import annotations

import logging

import numpy as np

from ..dsp.decoding import LNAV_CheckPreambule, LNAV_DecodeTOW, Prompt2Bit
from ..dsp.lockindicator import CN0_Beaulieu, FLL_Lock_Borre, PLL_Lock_Borre
from ..dsp.tracking import (EPL, BorreLoopFilter, DLL_NNEML, FLL_ATAN, FLLassistedPLL_2ndOrder,
                            LoopFiltersCoefficients, PLL_costa)
from ..signal.rfsignal import RFSignal
from ..utils.circularbuffer import CircularBuffer
from ..utils.constants import (GPS_L1CA_CODE_FREQ, GPS_L1CA_CODE_SIZE_BITS, LNAV_MS_PER_BIT, LNAV_SUBFRAME_SIZE,
                               LNAV_WORD_SIZE, TWO_PI, W0_BANDWIDTH_1, W0_BANDWIDTH_2, W0_SCALE_A2)
from ..utils.enumerations import ChannelMessage, ChannelState, LoopLockState, TrackingFlags
from .lnav_frame import advance_frame
from .channel_l1ca_borre import ChannelL1CA


class ChannelL1CA_Kaplan(ChannelL1CA):
    """Acquisition, satellite set-up and packet headers are the Borre channel's (identical code in
    the reference, channel_l1ca_kaplan.py:82-258); tracking and decoding follow the Kaplan file."""

    def __init__(self, cid: int, sharedBuffer: CircularBuffer, resultQueue, rfSignal: RFSignal, configuration: dict):
        # channel_l1ca_kaplan.py:33-45 (Channel.__init__, then the three set-up calls)
        super(ChannelL1CA, self).__init__(cid, sharedBuffer, resultQueue, rfSignal, configuration)
        self.codeOffset = 0
        self.codeFrequency = GPS_L1CA_CODE_FREQ
        self.initialFrequency = 0.0
        self.setAcquisition(configuration['ACQUISITION'])
        self.setTracking(configuration['TRACKING'])
        self.setDecoding()
        self.carrierFrequency = 0.0

    # ---- acquisition hand-off (channel_l1ca_kaplan.py:223-240) -------------------------------------
    def _ingestAcquisition(self, indices, peakRatio, correlationMap):
        results = super()._ingestAcquisition(indices, peakRatio, correlationMap)
        return results

    # ---- tracking ------------------------------------------------------------------------------------
    def setTracking(self, configuration: dict):
        """channel_l1ca_kaplan.py:262-340."""
        wide = float(configuration['correlator_epl_wide'])
        narrow = float(configuration['correlator_epl_narrow'])
        self.dll_epl_wide = [-wide, 0.0, wide]
        self.dll_epl_narrow = [-narrow, 0.0, narrow]
        self.track_correlatorsSpacing = self.dll_epl_wide
        self.correlatorsResults = np.zeros(6)
        self.correlatorsAccum = np.zeros(6)
        self.correlatorsAccumCounter = 0
        self.iPromptPrev = 0.0
        self.qPromptPrev = 0.0
        self.IDX_I_EARLY, self.IDX_Q_EARLY, self.IDX_I_PROMPT = 0, 1, 2
        self.IDX_Q_PROMPT, self.IDX_I_LATE, self.IDX_Q_LATE = 3, 4, 5
        self.track_dll_tau1, self.track_dll_tau2 = LoopFiltersCoefficients(
            loopNoiseBandwidth=float(configuration['dll_noise_bandwidth']),
            dampingRatio=float(configuration['dll_damping_ratio']),
            loopGain=float(configuration['dll_loop_gain']))
        self.track_dll_pdi = float(configuration['dll_pdi'])
        self.dllLockThreshold = float(configuration['dll_threshold'])
        self.cn0_PdPnRatio = 0.0
        self.cn0 = 0.0
        self.iPromptSum = self.qPromptSum = self.iPromptSum2 = self.qPromptSum2 = 0.0
        self.fll_bandwidth_pullin = float(configuration['fll_bandwidth_pullin'])
        self.fll_bandwidth_wide = float(configuration['fll_bandwidth_wide'])
        self.fll_bandwidth_narrow = float(configuration['fll_bandwidth_narrow'])
        self.fll_threshold_wide = float(configuration['fll_threshold_wide'])
        self.fll_threshold_narrow = float(configuration['fll_threshold_narrow'])
        self.pll_bandwidth_wide = float(configuration['pll_bandwidth_wide'])
        self.pll_bandwidth_narrow = float(configuration['pll_bandwidth_narrow'])
        self.pll_threshold_wide = float(configuration['pll_threshold_wide'])
        self.pll_threshold_narrow = float(configuration['pll_threshold_narrow'])
        self.dllDiscrim = self.pllDiscrim = self.fllDiscrim = 0.0
        self.carrierFrequencyError = 0.0
        self.codeFrequencyError = 0.0
        self.fllBandwidth = self.fll_bandwidth_pullin
        self.pllBandwidth = self.pll_bandwidth_wide
        self.dllLockIndicator = self.fllLockIndicator = self.pllLockIndicator = 0.0
        self.fll_vel_memory = 0.0
        self.timeSinceLastState = 0
        self.loopLockState = LoopLockState.PULL_IN
        self.trackFlags = TrackingFlags.UNKNOWN
        # for the batched dispatcher: the [TRACKING] values the device loop closure needs (sydr_kaplan_state)
        self._trackingConfiguration = {k: float(v) for k, v in configuration.items()}
        self._deviceLoop = "kaplan"
        self.remainingCode = 0.0
        self.remainingCarrier = 0.0
        self.codeStep = GPS_L1CA_CODE_FREQ / self.rfSignal.samplingFrequency
        self.track_requiredSamples = int(np.ceil((GPS_L1CA_CODE_SIZE_BITS - self.remainingCode) / self.codeStep))
        self.codeCounter = 0

    def runTracking(self):
        """channel_l1ca_kaplan.py:342-368: one epoch per tick once enough samples are unread."""
        if self.rfBuffer.getNbUnreadSamples(self.currentSample) < self.track_requiredSamples:
            return
        self.runCorrelators()
        dllDiscrim, fllDiscrim, pllDiscrim = self.runDiscriminators()
        carrierFrequencyError = self.runCarrierFrequencyFilter(fllDiscrim=fllDiscrim, pllDiscrim=pllDiscrim)
        codeFrequencyError = self.runCodeFrequencyFilter(dllDiscrim=dllDiscrim)
        self.runLoopIndicators()
        self.postTrackingUpdate(dllDiscrim, fllDiscrim, pllDiscrim, carrierFrequencyError, codeFrequencyError)
        self.trackingStateUpdate()
        return self.prepareResultsTracking()

    def runCorrelators(self):
        """channel_l1ca_kaplan.py:372-403: K-TRK through the drop-in EPL, then the 20 ms accumulator."""
        self.correlatorsResults[:] = EPL(
            rfData=self.rfBuffer.getSlice(self.currentSample, self.track_requiredSamples), code=self.code,
            samplingFrequency=self.rfSignal.samplingFrequency, carrierFrequency=self.carrierFrequency,
            remainingCarrier=self.remainingCarrier, remainingCode=self.remainingCode, codeStep=self.codeStep,
            correlatorsSpacing=self.track_correlatorsSpacing)
        if self.correlatorsAccumCounter == LNAV_MS_PER_BIT:
            self.correlatorsAccumCounter = 0
            self.correlatorsAccum[:] = 0.0
        self.correlatorsAccum += self.correlatorsResults[:]
        self.correlatorsAccumCounter += 1

    def runDiscriminators(self):
        """channel_l1ca_kaplan.py:407-432."""
        fllDiscrim = pllDiscrim = dllDiscrim = 0.0
        if self.loopLockState == LoopLockState.PULL_IN:
            if self.codeCounter > 1:
                fllDiscrim = self.runFrequencyDiscriminator(self.correlatorsResults)
            dllDiscrim = self.runCodeDiscriminator(self.correlatorsResults)
        else:
            fllDiscrim = self.runFrequencyDiscriminator(self.correlatorsResults)
            pllDiscrim = self.runPhaseDiscriminator(self.correlatorsResults)
            dllDiscrim = self.runCodeDiscriminator(self.correlatorsResults)
        return dllDiscrim, fllDiscrim, pllDiscrim

    def runCarrierFrequencyFilter(self, fllDiscrim=0.0, pllDiscrim=0.0, coherentIntegration=1):
        """channel_l1ca_kaplan.py:436-446."""
        carrierFrequencyError, self.fll_vel_memory = FLLassistedPLL_2ndOrder(
            pllDiscrim, fllDiscrim, w0f=self.fllBandwidth / W0_BANDWIDTH_1, w0p=self.pllBandwidth / W0_BANDWIDTH_2,
            a2=W0_SCALE_A2, integrationTime=coherentIntegration * 1e-3, velMemory=self.fll_vel_memory)
        return carrierFrequencyError

    def runCodeFrequencyFilter(self, dllDiscrim: float, coherentIntegration=1):
        """channel_l1ca_kaplan.py:450-456."""
        return BorreLoopFilter(dllDiscrim, self.dllDiscrim, self.track_dll_tau1, self.track_dll_tau2,
                               self.track_dll_pdi * coherentIntegration)

    def runLoopIndicators(self):
        """channel_l1ca_kaplan.py:460-508."""
        if self.codeCounter == 0:
            return
        iprompt = self.correlatorsResults[self.IDX_I_PROMPT]
        qprompt = self.correlatorsResults[self.IDX_Q_PROMPT]
        self.fllLockIndicator = FLL_Lock_Borre(iprompt=iprompt, qprompt=qprompt, iprompt_prev=self.iPromptPrev,
                                               qprompt_prev=self.qPromptPrev, fll_lock_prev=self.fllLockIndicator,
                                               alpha=0.005)
        if self.loopLockState > LoopLockState.PULL_IN:
            self.pllLockIndicator = PLL_Lock_Borre(iprompt=iprompt, qprompt=qprompt,
                                                   pll_lock_prev=self.pllLockIndicator, alpha=0.005)
        self.cn0_PdPnRatio += (iprompt ** 2 + qprompt ** 2) / (abs(iprompt) - abs(qprompt)) ** 2
        self.iPromptSum += abs(iprompt)
        self.qPromptSum += abs(qprompt)
        self.iPromptSum2 += iprompt ** 2
        self.qPromptSum2 += qprompt ** 2
        if self.correlatorsAccumCounter == LNAV_MS_PER_BIT:
            self.cn0 = CN0_Beaulieu(self.cn0_PdPnRatio, self.correlatorsAccumCounter,
                                    self.correlatorsAccumCounter * 1e-3, self.cn0)
            self.cn0_PdPnRatio = 0.0
            self.iPromptSum = self.qPromptSum = self.iPromptSum2 = self.qPromptSum2 = 0.0
        self.dllLockIndicator = self.cn0

    def postTrackingUpdate(self, dllDiscrim, fllDiscrim, pllDiscrim, carrierFrequencyError, codeFrequencyError):
        """channel_l1ca_kaplan.py:512-541: NCO update (GPS value of two pi, as the reference)."""
        fs = self.rfSignal.samplingFrequency
        self.codeCounter += 1
        self.codeSinceTOW += 1
        self.dllDiscrim = dllDiscrim
        self.fllDiscrim = fllDiscrim
        self.pllDiscrim = pllDiscrim
        self.carrierFrequencyError = carrierFrequencyError
        self.codeFrequencyError = codeFrequencyError
        self.remainingCarrier -= self.carrierFrequency * TWO_PI * self.track_requiredSamples / fs
        self.remainingCarrier %= TWO_PI
        self.codeFrequency -= self.codeFrequencyError
        self.carrierFrequency += self.carrierFrequencyError
        self.remainingCode += self.track_requiredSamples * self.codeStep - GPS_L1CA_CODE_SIZE_BITS
        self.codeStep = self.codeFrequency / fs
        self.currentSample = (self.currentSample + self.track_requiredSamples) % self.rfBuffer.maxSize
        self.track_requiredSamples = int(np.ceil((GPS_L1CA_CODE_SIZE_BITS - self.remainingCode) / self.codeStep))

    def trackingStateUpdate(self):
        """channel_l1ca_kaplan.py:545-619: code lock, bit synchronisation, loop bandwidth switching."""
        iprompt = self.correlatorsResults[self.IDX_I_PROMPT]
        if self.loopLockState != LoopLockState.PULL_IN and self.dllLockIndicator > self.dllLockThreshold \
                and not (self.trackFlags & TrackingFlags.CODE_LOCK):
            self.trackFlags |= TrackingFlags.CODE_LOCK
        elif self.dllLockIndicator < self.dllLockThreshold and (self.trackFlags & TrackingFlags.CODE_LOCK):
            self.trackFlags ^= TrackingFlags.CODE_LOCK
        if (self.trackFlags & TrackingFlags.CODE_LOCK) and not (self.trackFlags & TrackingFlags.BIT_SYNC):
            if np.sign(self.iPromptPrev) != np.sign(iprompt):
                self.trackFlags |= TrackingFlags.BIT_SYNC
                self.correlatorsAccum[:] = self.correlatorsResults[:]
                self.correlatorsAccumCounter = 1
                self.cn0_PdPnRatio = 0.0
                self.iPromptSum = self.qPromptSum = self.iPromptSum2 = self.qPromptSum2 = 0.0
                logging.getLogger(__name__).info(f"CID {self.channelID} tracking in {TrackingFlags.BIT_SYNC}.")
        self.iPromptPrev = iprompt
        self.qPromptPrev = self.correlatorsResults[self.IDX_Q_PROMPT]
        if self.loopLockState != LoopLockState.NARROW_TRACK and self.fllLockIndicator >= self.fll_threshold_narrow \
                and self.pllLockIndicator >= self.pll_threshold_narrow:
            self.loopLockState = LoopLockState.NARROW_TRACK
            self.fllBandwidth = self.fll_bandwidth_narrow
            self.pllBandwidth = self.pll_bandwidth_narrow
            self.track_correlatorsSpacing = self.dll_epl_narrow
        elif self.loopLockState != LoopLockState.WIDE_TRACK and self.fllLockIndicator >= self.fll_threshold_wide \
                and self.fllLockIndicator < self.fll_threshold_narrow:
            self.loopLockState = LoopLockState.WIDE_TRACK
            self.fllBandwidth = self.fll_bandwidth_wide
            self.pllBandwidth = self.pll_bandwidth_wide
            self.track_correlatorsSpacing = self.dll_epl_wide
        elif self.loopLockState != LoopLockState.PULL_IN and self.fllLockIndicator <= self.fll_threshold_wide:
            self.loopLockState = LoopLockState.PULL_IN
            self.fllBandwidth = self.fll_bandwidth_pullin
            self.pllBandwidth = 0.0
            self.track_correlatorsSpacing = self.dll_epl_wide
        else:
            self.timeSinceLastState += 1
            return
        self.timeSinceLastState = 0
        logging.getLogger(__name__).debug(f"CID {self.channelID} tracking switched to {self.loopLockState}.")

    def _ingestEpoch(self, rec, kex):
        """Batched path (ChannelManager): take over one epoch closed on the device - the members
        runTracking would have left (channel_l1ca_kaplan.py:342-619) - and build its packet.  `rec` is the
        sydr_trk_epoch record, `kex` its sydr_kaplan_epoch extras."""
        self.correlatorsResults[:] = rec["corr"]
        if self.correlatorsAccumCounter == LNAV_MS_PER_BIT:
            self.correlatorsAccumCounter = 0
            self.correlatorsAccum[:] = 0.0
        self.correlatorsAccum += self.correlatorsResults
        self.correlatorsAccumCounter += 1
        self.dllDiscrim = float(rec["code_err"])
        self.pllDiscrim = float(rec["carrier_err"])
        self.fllDiscrim = float(kex["fll"])
        self.carrierFrequencyError = float(rec["pll"])
        self.codeFrequencyError = float(rec["dll"])
        self.cn0 = self.dllLockIndicator = float(kex["cn0"])
        self.fllLockIndicator = float(kex["fll_lock"])
        self.pllLockIndicator = float(kex["pll_lock"])
        self.codeCounter += 1
        self.codeSinceTOW += 1
        self.carrierFrequency = float(rec["carrier_freq"])
        self.codeFrequency = float(rec["code_freq"])
        self.remainingCode = float(rec["rem_code"])
        self.remainingCarrier = float(rec["rem_carrier"])
        self.codeStep = self.codeFrequency / self.rfSignal.samplingFrequency
        self.currentSample = (self.currentSample + self.track_requiredSamples) % self.rfBuffer.maxSize
        self.track_requiredSamples = int(np.ceil((GPS_L1CA_CODE_SIZE_BITS - self.remainingCode) / self.codeStep))
        # code lock / bit synchronisation as decided on the device; the other flags are host-side (decoding)
        dev = int(kex["flags"]) & int(TrackingFlags.CODE_LOCK | TrackingFlags.BIT_SYNC)
        if (dev & int(TrackingFlags.BIT_SYNC)) and not (self.trackFlags & TrackingFlags.BIT_SYNC):
            self.correlatorsAccum[:] = self.correlatorsResults[:]
            self.correlatorsAccumCounter = 1
        keep = int(self.trackFlags) & ~int(TrackingFlags.CODE_LOCK | TrackingFlags.BIT_SYNC)
        self.trackFlags = keep | dev
        self.iPromptPrev = self.correlatorsResults[self.IDX_I_PROMPT]
        self.qPromptPrev = self.correlatorsResults[self.IDX_Q_PROMPT]
        self.loopLockState = LoopLockState(int(kex["lock_state"]))
        return self.prepareResultsTracking()

    def runFrequencyDiscriminator(self, correlatorResults):
        """channel_l1ca_kaplan.py:623-631."""
        return FLL_ATAN(iPrompt=correlatorResults[self.IDX_I_PROMPT], iPromptPrev=self.iPromptPrev,
                        qPrompt=correlatorResults[self.IDX_Q_PROMPT], qPromptPrev=self.qPromptPrev, deltaT=1e-3)

    def runPhaseDiscriminator(self, correlatorResults):
        """channel_l1ca_kaplan.py:635-641."""
        return PLL_costa(iPrompt=correlatorResults[self.IDX_I_PROMPT], qPrompt=correlatorResults[self.IDX_Q_PROMPT])

    def runCodeDiscriminator(self, correlatorResults):
        """channel_l1ca_kaplan.py:645-653."""
        return DLL_NNEML(iEarly=correlatorResults[self.IDX_I_EARLY], qEarly=correlatorResults[self.IDX_Q_EARLY],
                         iLate=correlatorResults[self.IDX_I_LATE], qLate=correlatorResults[self.IDX_Q_LATE])

    def prepareResultsTracking(self):
        """channel_l1ca_kaplan.py:657-681."""
        results = super().prepareResultsTracking()
        c = self.correlatorsResults
        results["i_early"], results["q_early"] = c[0], c[1]
        results["i_prompt"], results["q_prompt"] = c[2], c[3]
        results["i_late"], results["q_late"] = c[4], c[5]
        results["carrier_frequency"] = self.carrierFrequency
        results["code_frequency"] = self.codeFrequency
        results["carrier_frequency_error"] = self.carrierFrequencyError
        results["code_frequency_error"] = self.codeFrequencyError
        results["cn0"] = self.cn0
        results["pll_lock"] = self.pllLockIndicator
        results["fll_lock"] = self.fllLockIndicator
        results["dll"] = self.dllDiscrim
        results["pll"] = self.pllDiscrim
        results["fll"] = self.fllDiscrim
        results["lock_state"] = self.loopLockState
        return results

    # ---- decoding (channel_l1ca_kaplan.py:685-861) ---------------------------------------------------
    def setDecoding(self):
        self.navPromptSum = 0.0
        self.navPromptSumCounter = 0
        self.navBitBufferSize = LNAV_SUBFRAME_SIZE + 2 * LNAV_WORD_SIZE + 2
        self.navBitsBuffer = np.squeeze(np.empty((1, self.navBitBufferSize), dtype=int))
        self.navBitsCounter = 0
        self.preambuleFound = False
        self.subframeFlags = [False, False, False, False, False]
        self.tow = 0
        self.subframeID = 0
        self.subframeBits = []

    def runDecoding(self):
        if not self.decodeBit() or not self.decodeSubframe() or not self.postDecodingUpdate():
            return
        return self.prepareResultsDecoding()

    def decodeBit(self):
        if not (self.trackFlags & TrackingFlags.BIT_SYNC):
            self.navPromptSum = 0.0
            self.navPromptSumCounter = 0
            return False
        self.navPromptSum += self.correlatorsResults[self.IDX_I_PROMPT]
        self.navPromptSumCounter += 1
        if not (self.navPromptSumCounter == LNAV_MS_PER_BIT):
            return False
        self.navBitsBuffer[self.navBitsCounter] = Prompt2Bit(self.navPromptSum)
        self.navBitsCounter += 1
        self.navPromptSum = 0.0
        self.navPromptSumCounter = 0
        return True

    def decodeSubframe(self):
        """channel_l1ca_kaplan.py:760-823 (the search itself: lnav_frame.advance_frame)."""
        decoded = advance_frame(self)
        if decoded is None:
            return False
        self.tow, self.subframeID, self.subframeBits = decoded
        self.tow += self.navBitsCounter * LNAV_MS_PER_BIT * 1e-3
        return True

    def postDecodingUpdate(self):
        self.codeSinceTOW = 0
        try:
            self.subframeFlags[self.subframeID - 1] = True
            self.trackFlags |= TrackingFlags.TOW_DECODED
            self.trackFlags |= TrackingFlags.TOW_KNOWN
        except IndexError:
            self.trackFlags ^= TrackingFlags.TOW_DECODED
            self.trackFlags ^= TrackingFlags.TOW_KNOWN
            logging.getLogger(__name__).warning(f"CID {self.channelID} Error in subframe ID decoding.")
            return False
        if not (self.trackFlags & TrackingFlags.EPH_DECODED) and all(self.subframeFlags[0:3]):
            self.trackFlags |= TrackingFlags.EPH_DECODED
            self.trackFlags |= TrackingFlags.EPH_KNOWN
        return True

    def prepareResultsDecoding(self):
        results = super().prepareResultsDecoding()
        results["subframe_id"] = self.subframeID
        results["tow"] = int(self.tow)
        results["bits"] = self.subframeBits
        return results

    def _afterTick(self):
        """The Kaplan channel keeps no 20-entry prompt ring (channel_l1ca_kaplan.py:48-78)."""
        return
