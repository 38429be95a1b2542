"""Drop-in for sydr/dsp/tracking.py.  EPL runs on the GPU; the scalar discriminators and loop
filters are per-epoch FP64 scalars (boundary logic) and are restated on
the host with the reference's exact expression order -- in the batched path
(`sydr_b200.engine.TrackingEngine`) the same expressions run inside the CUDA kernel."""
from __future__ import annotations

import numpy as np

from .. import _lib as L
from ..engine import epl_batch, to_device_iq

PI = 3.1415926535898          # sydr/utils/constants.py:4 (GPS ICD pi)
TWO_PI = PI * 2.0
HALF_PI = PI / 2.0

_PRN_BY_CODE: dict[bytes, int] = {}


def _prn_of(code: np.ndarray) -> int:
    """Identify the PRN of a padded (1025) or plain (1023) +-1 C/A code array."""
    from ..signal.gnsssignal import GenerateGPSGoldCode
    c = np.asarray(code)
    core = c[1:-1] if c.shape[0] == 1025 else c
    key = (core > 0).astype(np.uint8).tobytes()
    if not _PRN_BY_CODE:
        for prn in range(1, 38):
            _PRN_BY_CODE[(GenerateGPSGoldCode(prn) > 0).astype(np.uint8).tobytes()] = prn
    try:
        return _PRN_BY_CODE[key]
    except KeyError:
        raise L.SydrError("EPL: `code` is not a GPS L1 C/A code of PRN 1..37") from None


def EPL(rfData: np.array, code: np.array, samplingFrequency: float, carrierFrequency: float,
        remainingCarrier: float, remainingCode: float, codeStep: float, correlatorsSpacing: tuple):
    """sydr/dsp/tracking.py:92-116.  Returns [IE, QE, IP, QP, IL, QL] (Python floats)."""
    L.require_device()
    rf = np.squeeze(np.asarray(rfData))
    if len(correlatorsSpacing) != 3:
        raise L.SydrError("EPL supports exactly three correlators (early, prompt, late)")
    if rf.dtype not in (np.complex64, np.complex128):
        rf = rf.astype(np.complex128)
    args = np.zeros(1, dtype=L.EPL_ARGS_DTYPE)
    args["start"], args["n"], args["prn"] = 0, rf.shape[0], _prn_of(code)
    args["carrier_freq"], args["rem_carrier"] = carrierFrequency, remainingCarrier
    args["rem_code"], args["code_step"] = remainingCode, codeStep
    args["spacing"][0] = [float(s) for s in correlatorsSpacing]
    out = epl_batch(to_device_iq(rf), samplingFrequency, args)
    return [float(v) for v in out[0]]


def LoopFiltersCoefficients(loopNoiseBandwidth: float, dampingRatio: float, loopGain: float):
    """sydr/dsp/tracking.py:39-61."""
    Wn = loopNoiseBandwidth * 8.0 * dampingRatio / (4.0 * dampingRatio ** 2 + 1)
    return loopGain / Wn ** 2, 2.0 * dampingRatio / Wn


def DLL_NNEML(iEarly: float, qEarly: float, iLate: float, qLate: float):
    """sydr/dsp/tracking.py:120-129."""
    e = np.sqrt(iEarly ** 2 + qEarly ** 2)
    l = np.sqrt(iLate ** 2 + qLate ** 2)
    return (e - l) / (e + l)


def PLL_costa(iPrompt: float, qPrompt: float):
    """sydr/dsp/tracking.py:133-142."""
    return np.arctan(qPrompt / iPrompt) / TWO_PI


def FLL_ATAN2(iPrompt, qPrompt, iPromptPrev, qPromptPrev, deltaT):
    """sydr/dsp/tracking.py:146-152."""
    e = np.arctan2(iPromptPrev * iPrompt + qPromptPrev * qPrompt,
                   iPromptPrev * qPrompt - qPromptPrev * iPrompt) / deltaT
    return e / TWO_PI


def phase_unwrap(phase):
    """sydr/dsp/tracking.py:169-176."""
    if phase >= HALF_PI:
        return phase - PI
    if phase <= -HALF_PI:
        return phase + PI
    return phase


def FLL_ATAN(iPrompt, qPrompt, iPromptPrev, qPromptPrev, deltaT):
    """sydr/dsp/tracking.py:156-165."""
    e = np.arctan(qPrompt / iPrompt) - np.arctan(qPromptPrev / iPromptPrev)
    if np.isnan(e):
        e = 0.0
    return phase_unwrap(e) / deltaT / TWO_PI


def BorreLoopFilter(input: float, memory: float, tau1: float, tau2: float, pdi: float):
    """sydr/dsp/tracking.py:180-186."""
    output = tau2 / tau1 * (input - memory)
    output += pdi / tau1 * input
    return output


def FLLassistedPLL_2ndOrder(phaseInput: float, freqInput: float, w0f: float, w0p: float, a2: float,
                            integrationTime: float, velMemory: float):
    """2nd-order PLL assisted by a 1st-order FLL [Kaplan, 2006, p180-182], sydr/dsp/tracking.py:246-279.
    Returns (output, velMemory)."""
    update = (phaseInput * w0p ** 2 + freqInput * w0f) * integrationTime
    output = update + velMemory
    output += phaseInput * a2 * w0p
    return output, update


def FLLassistedPLL_3rdOrder(phaseInput: float, freqInput: float, w0f: float, w0p: float, a2: float, a3: float,
                            b3: float, integrationTime: float, velMemory: float, accMemory: float):
    """3rd-order PLL assisted by a 2nd-order FLL, sydr/dsp/tracking.py:283-325.
    Returns (output, velMemory, accMemory)."""
    acc_update = (phaseInput * w0p ** 3 + freqInput * w0f ** 2) * integrationTime
    output = acc_update + accMemory
    vel_update = (output + (phaseInput * a3 * w0p ** 2 + freqInput * a2 * w0f)) * integrationTime
    output = vel_update + velMemory
    output += phaseInput * b3 * w0p
    return output, vel_update, acc_update
