"""The nine legacy C entry points of libsydr_b200.so (csrc/legacy.cu) and the `Acquisition` / `Tracking` class façade
of the reference's earlier API (sydr/old/acquisition/acquisition_pcps_c.py, sydr/old/tracking/tracking_epl_c.py).

  * getCorrelator, generateReplica, generateCarrier, delayLockLoop, phaseLockLoop, getLoopCoefficients: call by call
    against oracle/_ref/tracking.so -- the reference's own sydr/c_functions/tracking.c compiled by oracle/Makefile
    (prebuilt, travels with the repository) -- and against the NumPy statement of the same lines;
  * setSatellite, PCPS, twoCorrelationPeakComparison: against the NumPy oracle (the reference's acquisition.c is
    numerically wrong against the live Python path, SURVEY.md section 8c, and is not built);
  * the classes: a closed loop of 40 code periods through `Tracking.run` against the same loop driven through
    tracking.so / NumPy, and `Acquisition.run` against oracle.pcps + oracle.two_peak.
"""
import configparser
import ctypes as C
import os

import numpy as np
import pytest

import helpers as H  # noqa: F401

pytestmark = pytest.mark.gpu

REF_SO = os.path.join(H.ROOT, "oracle", "_ref", "tracking.so")
FS = 4e6


class _RF:
    samplingFrequency, interFrequency = FS, 0.0


class _Signal:
    """What the legacy classes read from a GNSSSignal: the configuration and the code generator."""
    codeFrequency, codeBits = 1.023e6, 1023

    def __init__(self):
        self.config = configparser.ConfigParser()
        self.config.read_dict({
            "ACQUISITION": dict(doppler_range="5000", doppler_steps="250", coh_integration="1", noncoh_integration="4",
                                metric_threshold="1.5"),
            "TRACKING": dict(pdi_code="0.001", pdi_carrier="0.001", correlator_number="3", correlator_0="-0.5",
                             correlator_1="0.0", correlator_2="0.5", correlator_prompt="1", dll_dumping_ratio="0.7",
                             dll_noise_bandwidth="1.0", dll_loop_gain="1.0", pll_dumping_ratio="0.7",
                             pll_noise_bandwidth="15.0", pll_loop_gain="0.25")})

    def getCode(self, svid, samplingFrequency=None):
        from sydr_b200.signal.gnsssignal import GenerateGPSGoldCode
        return GenerateGPSGoldCode(svid, samplingFrequency)


def _scenario(ms=60):
    from sydr_b200 import synth
    sc = synth.make_scenario(FS, 8, ms * 1e-3, (3, 7), 4242, 250.0)
    return sc, synth.to_complex(synth.generate_iq(sc)).astype(np.complex128)


def _ref_lib():
    if not os.path.exists(REF_SO):
        return None
    lib = C.CDLL(REF_SO)
    from sydr_b200.old._legacy import PROTOTYPES
    for name in ("getCorrelator", "generateReplica", "generateCarrier", "delayLockLoop", "phaseLockLoop", "getLoopCoefficients"):
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = PROTOTYPES[name], None
    return lib


class _NumpyTracking:
    """tracking.c:40-49, 79-93, 113-119, 144-156, 181-188, 206-209 in NumPy (GPS pi as in tracking.c:21)."""
    PI = 3.1415926535898

    def generateReplica(self, time, size, fc, rem, r_rem, r_rep):
        temp = -(fc * 2.0 * self.PI * time[:size + 1]) + rem
        r_rep[:] = np.cos(temp[:size]) + 1j * np.sin(temp[:size])
        r_rem[0] = np.fmod(temp[size], 2 * self.PI)

    def generateCarrier(self, rf, rep, size, ri, rq):
        ri[:] = rf.real * rep.real - rf.imag * rep.imag
        rq[:] = rf.real * rep.imag + rf.imag * rep.real

    def getCorrelator(self, i_sig, q_sig, code, size, step, rem, spacing, r_i, r_q):
        start, stop = rem + spacing, size * step + rem + spacing
        st = (stop - start) / size
        idx = np.ceil(start + st * np.arange(size)).astype(int)
        r_i[0], r_q[0] = float(np.sum(code[idx] * i_sig)), float(np.sum(code[idx] * q_sig))

    def delayLockLoop(self, ie, qe, il, ql, t1, t2, pdi, nco, err, f0, r_nco, r_err, r_f):
        e, l = np.sqrt(ie * ie + qe * qe), np.sqrt(il * il + ql * ql)
        new = (e - l) / (e + l)
        nco = nco + t2 / t1 * (new - err)
        nco = nco + pdi / t1 * new
        r_nco[0], r_err[0], r_f[0] = nco, new, f0 - nco

    def phaseLockLoop(self, ip, qp, t1, t2, pdi, nco, err, f0, r_nco, r_err, r_f):
        new = np.arctan(qp / ip) / 2.0 / self.PI
        nco = nco + t2 / t1 * (new - err)
        nco = nco + pdi / t1 * new
        r_nco[0], r_err[0], r_f[0] = nco, new, f0 + nco

    def getLoopCoefficients(self, bw, zeta, gain, r1, r2):
        wn = bw * 8.0 * zeta / (4.0 * zeta * zeta + 1)
        r1[0], r2[0] = gain / (wn * wn), 2.0 * zeta / wn


def _checkers():
    out = [("numpy", _NumpyTracking())]
    ref = _ref_lib()
    if ref is not None:
        out.append(("tracking.so", ref))
    return out


def test_tracking_entry_points_call_by_call():
    """Every tracking.c entry point of the GPU library against the reference's compiled tracking.c and NumPy."""
    from sydr_b200.old._legacy import library
    from sydr_b200.signal.gnsssignal import GenerateGPSGoldCode
    lib = library()
    _, x = _scenario(8)
    code = GenerateGPSGoldCode(3)
    code = np.ascontiguousarray(np.r_[code[-1], code, code[0]].astype(np.int32))
    step = 1.023e6 / FS
    time = np.arange(0, 4002) / FS
    rng = np.random.default_rng(5)
    for name, chk in _checkers():
        for fc, rem_c, rem_code, off in ((1250.0, 0.3, 0.21, 100), (-3750.0, 5.9, 0.77, 4001)):
            size = int(np.ceil((1023 - rem_code) / step))                  # a code period, as the callers size it
            # generateReplica
            a_rem, a_rep = np.empty(1), np.empty(size, dtype=np.complex128)
            b_rem, b_rep = np.empty(1), np.empty(size, dtype=np.complex128)
            lib.generateReplica(np.ascontiguousarray(time[:size + 1]), size, fc, rem_c, a_rem, a_rep)
            chk.generateReplica(np.ascontiguousarray(time[:size + 1]), size, fc, rem_c, b_rem, b_rep)
            assert np.abs(a_rep - b_rep).max() <= 1e-12, name
            assert abs(a_rem[0] - b_rem[0]) <= 1e-12, name
            # generateCarrier
            rf = np.ascontiguousarray(x[off:off + size])
            ai, aq, bi, bq = (np.empty(size) for _ in range(4))
            lib.generateCarrier(rf, b_rep, size, ai, aq)
            chk.generateCarrier(rf, b_rep, size, bi, bq)
            assert np.array_equal(ai, bi) and np.array_equal(aq, bq), name
            # getCorrelator, three taps
            for sp in (-0.5, 0.0, 0.5):
                o = [np.empty(1) for _ in range(4)]
                lib.getCorrelator(bi, bq, code, size, step, rem_code, sp, o[0], o[1])
                chk.getCorrelator(bi, bq, code, size, step, rem_code, sp, o[2], o[3])
                scale = np.hypot(o[2][0], o[3][0]) + 1.0
                assert abs(o[0][0] - o[2][0]) <= 1e-9 * scale and abs(o[1][0] - o[3][0]) <= 1e-9 * scale, (name, sp)
        # the scalar loop functions: same FP64 operations in the same order -> equal to the last bit or two
        for _ in range(20):
            v = rng.normal(size=4) * 1e4
            t = [np.empty(1) for _ in range(6)]
            lib.delayLockLoop(*v, 0.9, 1.3, 1e-3, 0.4, -0.02, 1.023e6, t[0], t[1], t[2])
            chk.delayLockLoop(*v, 0.9, 1.3, 1e-3, 0.4, -0.02, 1.023e6, t[3], t[4], t[5])
            assert all(abs(t[k][0] - t[k + 3][0]) <= 1e-15 * max(1.0, abs(t[k + 3][0])) for k in range(3)), name
            lib.phaseLockLoop(v[0], v[1], 0.02, 0.09, 1e-3, 12.0, 0.01, 1250.0, t[0], t[1], t[2])
            chk.phaseLockLoop(v[0], v[1], 0.02, 0.09, 1e-3, 12.0, 0.01, 1250.0, t[3], t[4], t[5])
            assert all(abs(t[k][0] - t[k + 3][0]) <= 1e-13 * max(1.0, abs(t[k + 3][0])) for k in range(3)), name
        for bw, z, g in ((1.0, 0.7, 1.0), (15.0, 0.7, 0.25), (25.0, 0.5, 2.0)):
            t = [np.empty(1) for _ in range(4)]
            lib.getLoopCoefficients(bw, z, g, t[0], t[1])
            chk.getLoopCoefficients(bw, z, g, t[2], t[3])
            assert abs(t[0][0] - t[2][0]) <= 1e-15 * t[2][0] and abs(t[1][0] - t[3][0]) <= 1e-15 * t[3][0], name


def test_acquisition_entry_points_and_class():
    """setSatellite / PCPS / twoCorrelationPeakComparison through the `Acquisition` class against the NumPy oracle
    (oracle.pcps restates sydr/dsp/acquisition.py:9-74, which is acquisition_pcps.py:90-140 with the bins of the new API)."""
    from oracle import sydr_oracle as O
    from sydr_b200.old.acquisition.acquisition_pcps_c import Acquisition
    sc, x = _scenario(12)
    n_code = 4000
    for prn, present in ((3, True), (7, True), (11, False)):
        acq = Acquisition(_RF(), _Signal())
        assert len(acq.frequencyBins) == 40 and acq.samplesPerCode == n_code and acq.samplesPerCodeChip == 4
        acq.setSatellite(prn)
        ref_fft = np.conj(np.fft.fft(acq.code))
        assert np.abs(acq.codeFFT - ref_fft).max() <= 1e-9 * np.abs(ref_fft).max()
        data = x[:n_code * 4]
        acq.run(data)
        # the oracle's bin axis is arange(-range, range + step, step)[:40] for the same 40 rows
        cmap = O.pcps(data[None, :], 0.0, FS, O.code_spectrum(prn, FS), 5000.0, 250.0, n_code, 1, 4)[:40]
        assert acq.correlationMap.shape == cmap.shape
        assert np.abs(acq.correlationMap - cmap).max() <= 1e-4 * cmap.max()
        idx, ratio = O.two_peak(cmap, n_code, 4)
        assert [acq.idxEstimatedFrequency, acq.idxEstimatedCode] == idx
        assert abs(acq.acquisitionMetric - ratio) <= 1e-4 * ratio
        assert acq.estimatedDoppler == -acq.frequencyBins[idx[0]] and acq.estimatedCode == idx[1]
        assert acq.estimatedFrequency == _RF.interFrequency + acq.estimatedDoppler
        assert acq.isAcquired == present
        # the peak search on its own, on a caller-supplied float64 map (ties and edges as the reference resolves them)
        acq.twoCorrelationPeakComparison(cmap)
        assert [acq.idxEstimatedFrequency, acq.idxEstimatedCode] == idx and abs(acq.acquisitionMetric - ratio) <= 1e-12 * ratio
        assert set(acq.getDatabaseDict()) == {"type", "frequency", "code", "frequency_idx", "code_idx", "correlation_map"}


def test_tracking_class_closed_loop():
    """40 code periods through Tracking.run (nine GPU entry points per period) against the same object driven through
    the reference's compiled tracking.c and through NumPy: the NCO trajectory must coincide."""
    from oracle import sydr_oracle as O
    from sydr_b200.old.acquisition.acquisition_pcps_c import Acquisition
    from sydr_b200.old.tracking.tracking_epl_c import Tracking
    sc, x = _scenario(60)
    acq = Acquisition(_RF(), _Signal())
    acq.setSatellite(3)
    acq.run(x[:16000])
    assert acq.isAcquired
    start = int(acq.estimatedCode) + 16000 - 4000 + 1          # hand-off as the receiver does it (channel_l1ca_borre.py:301-311)

    def drive(make_backend):
        trk = Tracking(_RF(), _Signal())
        if make_backend is not None:
            trk._c = make_backend
        trk.setSatellite(3)
        trk.setInitialValues(acq.estimatedFrequency)
        cur, rows = start, []
        for _ in range(40):
            n = trk.getSamplesRequired()
            trk.run(x[cur:cur + n])
            cur += n
            rows.append(trk.getCorrelatorResults() + [trk.getCarrierFrequency(), trk.getCodeFrequency(), trk.remCodePhase,
                                                      trk.remCarrierPhase, float(n)])
        assert set(trk.getDatabaseDict()) >= {"i_prompt", "q_prompt", "dll", "pll", "carrier_frequency", "code_frequency"}
        return np.array(rows)

    ours = drive(None)
    truth = next(s.doppler for s in sc.sats if s.prn == 3)
    assert abs(ours[-10:, 6].mean() - truth) < 30.0              # the loop pulls towards the satellite's Doppler
    for name, chk in _checkers():
        ref = drive(chk)
        assert np.array_equal(ours[:, 10], ref[:, 10]), name     # same epoch lengths
        scale = np.hypot(ref[:, 2], ref[:, 3])[:, None]
        assert (np.abs(ours[:, :6] - ref[:, :6]) <= 1e-7 * scale).all(), name
        assert np.abs(ours[:, 6] - ref[:, 6]).max() <= 1e-6 and np.abs(ours[:, 7] - ref[:, 7]).max() <= 1e-6, name
        assert np.abs(ours[:, 8] - ref[:, 8]).max() <= 1e-9 and np.abs(ours[:, 9] - ref[:, 9]).max() <= 1e-9, name
