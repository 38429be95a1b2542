"""Kaplan channel variant (sydr/channel/channel_l1ca_kaplan.py): the host
mirror ChannelL1CA_Kaplan ticked by `_processHandler()` exactly like the reference's channel
process, compared epoch by epoch with the packets of the live reference channel
(tests/golden/kaplan.npz, made by tests/golden/make_golden.py from /root/reference).

* CPU: the correlator / acquisition calls are replaced by the oracle, so the test pins the host
  logic alone (FLL-assisted PLL, lock indicators, C/N0, state machine, bit sync) - it must follow
  the reference to rounding.
* GPU: the real path (EPL / PCPS through libsydr_b200); FP32 correlator sums perturb the loops
  at the 1e-7 level, so trajectories are compared with the tolerances of the north star.
"""
import numpy as np
import pytest

import helpers as H

ACQ_CFG = {"doppler_range": "5000", "doppler_steps": "250", "coherent_integration": "1",
           "non_coherent_integration": "10", "threshold": "1.5"}
COLS = dict(tick=0, corr=slice(1, 7), dll=7, pll=8, fll=9, carrier=10, code=11, cerr=12, derr=13, cn0=14, pll_lock=15,
            fll_lock=16, state=17, flags=18, rem_code=19, rem_carrier=20, n_req=21, cur=22, navbits=23)


def make_channel(g):
    from sydr_b200 import synth
    from sydr_b200.channel.channel_l1ca_kaplan import ChannelL1CA_Kaplan
    from sydr_b200.signal.rfsignal import RFSignal
    from sydr_b200.utils.circularbuffer import CircularBuffer
    fs, nbits, seed, ms, ds = g["meta"]
    sc = synth.make_scenario(float(fs), int(nbits), int(ms) * 1e-3, tuple(int(p) for p in g["prns"]), int(seed), float(ds))
    iq = synth.generate_iq(sc)
    assert H.sha(iq) == str(g["sha"])
    x = synth.to_complex(iq)
    rf = RFSignal({"filepath": "none", "sampling_frequency": str(float(fs)), "is_complex": "true",
                   "intermediate_frequency": "0.0", "data_size": "8"})

    def build(prn):
        buf = CircularBuffer(int(fs * 1e-3 * 100), np.complex128)
        ch = ChannelL1CA_Kaplan(0, buf, None, rf, {"ACQUISITION": dict(ACQ_CFG), "TRACKING": dict(H.MG.KAPLAN_TRK_CFG)})
        ch.setSatellite(prn)
        return ch, buf
    return x, rf, build


def drive(ch, buf, x, spm, n_ticks):
    from sydr_b200.utils.enumerations import ChannelMessage as M
    rows, acq = [], None
    for t in range(n_ticks):
        buf.shift(x[t * spm:(t + 1) * spm])
        for r in ch._processHandler():
            if r["type"] == M.ACQUISITION_UPDATE:
                acq = (t, r["frequency_idx"], r["code_idx"], r["peak_ratio"], r["carrierFrequency"], ch.currentSample)
            elif r["type"] == M.TRACKING_UPDATE:
                rows.append([t, r["i_early"], r["q_early"], r["i_prompt"], r["q_prompt"], r["i_late"], r["q_late"],
                             r["dll"], r["pll"], r["fll"], r["carrier_frequency"], r["code_frequency"],
                             r["carrier_frequency_error"], r["code_frequency_error"], r["cn0"], r["pll_lock"],
                             r["fll_lock"], int(r["lock_state"]), int(ch.trackFlags), ch.remainingCode,
                             ch.remainingCarrier, ch.track_requiredSamples, ch.currentSample, ch.navBitsCounter])
    return acq, np.array(rows, dtype=np.float64)


def test_host_logic_follows_reference_to_rounding(golden, monkeypatch):
    """Oracle correlators injected: every packet field of 2590 epochs, the PULL_IN -> WIDE -> NARROW
    transitions, code lock and bit synchronisation must be the reference's."""
    from oracle import sydr_oracle as O
    import sydr_b200.channel.channel_l1ca_borre as B
    import sydr_b200.channel.channel_l1ca_kaplan as K
    g = golden("kaplan.npz")
    x, rf, build = make_channel(g)
    monkeypatch.setattr(B, "GenerateGPSGoldCode", lambda prn, samplingFrequency=None: O.ca_code(int(prn)))
    monkeypatch.setattr(K, "EPL", lambda rfData, code, samplingFrequency, carrierFrequency, remainingCarrier,
                        remainingCode, codeStep, correlatorsSpacing:
                        O.epl(rfData, code, samplingFrequency, carrierFrequency, remainingCarrier, remainingCode, codeStep,
                              correlatorsSpacing))
    monkeypatch.setattr(B, "PCPS", lambda rfData, interFrequency, samplingFrequency, codeFFT, dopplerRange, dopplerStep,
                        samplesPerCode, coherentIntegration=1, nonCoherentIntegration=1:
                        O.pcps(rfData, interFrequency, samplingFrequency, codeFFT, dopplerRange, dopplerStep, samplesPerCode,
                               coherentIntegration, nonCoherentIntegration))
    monkeypatch.setattr(B, "TwoCorrelationPeakComparison", lambda correlationMap, samplesPerCode, samplesPerCodeChip:
                        O.two_peak(correlationMap, samplesPerCode, samplesPerCodeChip))
    prn = int(g["prns"][0])
    ch, buf = build(prn)
    n_ticks = 1200                                      # past the switch to NARROW_TRACK (epoch ~532)
    acq, rows = drive(ch, buf, x, rf.samplesPerMs, n_ticks)
    ref = g[f"trk_{prn}"]
    ref = ref[ref[:, 0] < n_ticks]
    ra = g[f"acq_{prn}"]
    assert [acq[0], acq[1], acq[2], acq[5]] == [int(ra[0]), int(ra[1]), int(ra[2]), int(ra[5])] and acq[4] == ra[4]
    assert abs(acq[3] - ra[3]) <= 1e-9 * ra[3]
    assert rows.shape == ref.shape
    for name in ("tick", "state", "flags", "n_req", "cur", "navbits"):
        assert np.array_equal(rows[:, COLS[name]], ref[:, COLS[name]]), name
    assert {1, 2, 3} <= set(ref[:, COLS["state"]].astype(int))
    scale = np.hypot(ref[:, 3], ref[:, 4])[:, None]
    assert (np.abs(rows[:, COLS["corr"]] - ref[:, COLS["corr"]]) / scale).max() <= 1e-9
    for name, tol in (("carrier", 1e-6), ("code", 1e-6), ("dll", 1e-9), ("pll", 1e-9), ("fll", 1e-6), ("cerr", 1e-6),
                      ("derr", 1e-9), ("pll_lock", 1e-9), ("fll_lock", 1e-9), ("rem_code", 1e-9), ("rem_carrier", 1e-6)):
        assert np.abs(rows[:, COLS[name]] - ref[:, COLS[name]]).max() <= tol, name
    assert np.allclose(rows[:, COLS["cn0"]], ref[:, COLS["cn0"]], rtol=1e-6, atol=1e-9)


@pytest.mark.gpu
def test_kaplan_channel_on_the_gpu(golden):
    """The real path: PCPS / peak search / EPL on the B200, Kaplan loops on the host."""
    g = golden("kaplan.npz")
    x, rf, build = make_channel(g)
    for prn in (int(p) for p in g["prns"]):
        ch, buf = build(prn)
        n_ticks = 900
        acq, rows = drive(ch, buf, x, rf.samplesPerMs, n_ticks)
        ref = g[f"trk_{prn}"]
        ref = ref[ref[:, 0] < n_ticks]
        ra = g[f"acq_{prn}"]
        assert [acq[0], acq[1], acq[2]] == [int(ra[0]), int(ra[1]), int(ra[2])] and acq[4] == ra[4]   # bit-exact detection
        assert abs(acq[3] - ra[3]) <= 1e-4 * ra[3]
        assert len(rows) == len(ref) and np.array_equal(rows[:, 0], ref[:, 0])
        # loop outputs within the north star's tolerances
        assert np.abs(rows[:, COLS["carrier"]] - ref[:, COLS["carrier"]]).max() <= 0.5
        assert np.abs(rows[:, COLS["code"]] - ref[:, COLS["code"]]).max() <= 0.5
        # the state machine takes the same path (transitions may move by an epoch or two)
        for st in (2, 3):
            a, b = int(np.argmax(rows[:, COLS["state"]] == st)), int(np.argmax(ref[:, COLS["state"]] == st))
            assert a > 0 and abs(a - b) <= 3, (prn, st, a, b)
        assert int(rows[-1, COLS["flags"]]) == int(ref[-1, COLS["flags"]])
        # correlators where the trajectories have not been separated by an epoch-length flip
        same = rows[:, COLS["n_req"]] == ref[:, COLS["n_req"]]
        e = np.abs(rows[:, COLS["corr"]] - ref[:, COLS["corr"]]).max(axis=1) / np.hypot(ref[:, 3], ref[:, 4])
        assert np.median(e[same]) <= 1e-3
