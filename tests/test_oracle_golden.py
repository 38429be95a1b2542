"""The NumPy oracle against the reference's own outputs (tests/golden, made by make_golden.py),
the IS-GPS-200 first-10-chips table, the replica known-answer vector of
sydr/c_functions/tracking.c:243-247 and, when built, oracle/_ref/tracking.so."""
import ctypes
import os

import numpy as np
import pytest

import helpers as H
from oracle import sydr_oracle as O
from sydr_b200 import synth


def test_ca_codes_match_reference(golden):
    g = golden("codes.npz")
    for prn in range(1, 33):
        assert np.array_equal(O.ca_code(prn).astype(np.int8), g["codes"][prn - 1])
        assert np.array_equal(synth.ca_code_pm1(prn), g["codes"][prn - 1])
        assert O.first10_octal(prn) == O.FIRST10_OCTAL[prn - 1]


@pytest.mark.parametrize("fs", [4e6, 10e6, 25e6, 50e6])
def test_code_spectrum(golden, fs):
    g = golden("codes.npz")
    sel = g[f"sel_{int(fs)}"]
    for prn in (1, 19, 32):
        spec = O.code_spectrum(prn, fs)
        assert np.array_equal(spec[sel], g[f"spec_{int(fs)}_{prn}"])
        up = O.upsample_code(O.ca_code(prn), fs)
        assert np.array_equal(np.array([up.sum(), (up * np.arange(len(up))).sum()]), g[f"upsum_{int(fs)}_{prn}"])


def test_peak_quirks(golden):
    g = golden("peaks.npz")
    maps, n, chip, tie = H.MG.peak_case_maps()
    assert H.sha(maps) == str(g["maps_sha"])
    for m, idx, ratio in zip(maps, g["idx"], g["ratio"]):
        i, r = O.two_peak(m.astype(np.float64), n, chip)
        assert i == list(idx) and r == ratio
    i, r = O.two_peak(tie.astype(np.float64), 64, 2)
    assert i == list(g["tie_idx"]) and r == float(g["tie_ratio"])


@pytest.mark.parametrize("case,prns", [("mini4", None), ("cfg1", (3, 31)), ("cfg2", (1, 3, 19, 22)),
                                       ("cfg3acq", (11, 2)), ("cfg4", (5,))])
def test_acquisition(golden, case, prns):
    g = golden(f"acq_{case}.npz")
    sc, iq, n, p = H.acq_case(case)
    assert H.sha(iq) == str(g["iq_sha"])
    x = synth.to_complex(iq)[None, :]
    table = {int(r[0]): r for r in g["result"]}
    for prn in (prns or p["search"]):
        cmap = O.pcps(x, 0.0, p["fs"], O.code_spectrum(prn, p["fs"]), float(p["doppler_range"]),
                      float(p["doppler_step"]), n, p["coh"], p["noncoh"])
        idx, ratio = O.two_peak(cmap, n, p["chip"])
        ref = table[prn]
        assert idx == [int(ref[1]), int(ref[2])]
        assert ratio == ref[3]
        assert np.array_equal(cmap.max(axis=1), g[f"rowmax_{prn}"])
        if f"map_{prn}" in g:
            assert np.array_equal(cmap.astype(np.float32), g[f"map_{prn}"])


def test_epl(golden):
    g = golden("epl.npz")
    for fs, nbits, seed in H.EPL_SETS:
        iq = H.epl_input(fs, nbits, seed)
        assert H.sha(iq) == str(g[f"sha{int(fs)}"])
        x = synth.to_complex(iq)
        for c in g[f"fs{int(fs)}"]:
            prn, start, n = int(c[0]), int(c[1]), int(c[2])
            out = O.epl(x[start:start + n], O.padded_code(prn), fs, c[3], c[4], c[5], c[6], list(c[7:10]))
            assert np.array_equal(np.array(out), c[10:16])
            for sp in c[7:10]:
                assert np.array_equal(O.code_indices(c[5], sp, c[6], n), O.code_indices_explicit(c[5], sp, c[6], n))
    # the reference's own unit-test fixture (1 ms of a real recording, PRN 2, 3700 Hz)
    u = g["unit_iq"].astype(np.float64)
    rf = u[0::2] + 1j * u[1::2]
    out = O.epl(rf, O.padded_code(2), 10e6, 3700.0, 0.0, 0.0, 1.023e6 / 10e6, [-0.5, 0.0, 0.5])
    assert np.array_equal(np.array(out), g["unit_epl"])


def test_replica_known_answer(golden):
    # sydr/c_functions/tracking.c:243-247, eps 1e-8 (f = -1500 Hz, fs = 10 MHz)
    truth = np.array([1 + 0j, 0.9999995558678348 + 0.000942477656548699j, 0.9999982234717338 + 0.0018849544759281136j,
                      0.9999960028128805 + 0.002827429620969703j, 0.9999928938932473 + 0.0037699022545064132j])
    rep, rem = O.generate_replica(np.arange(6) / 1e7, 5, -1500.0, 0.0)
    assert np.abs(rep - truth).max() < 1e-8
    g = golden("epl.npz")
    assert np.array_equal(rep, g["replica"]) and rem == float(g["replica_rem"])


def test_closed_loop_matches_reference_channel(golden):
    """BorreTrackOracle reproduces the live ChannelL1CA (driven tick by tick through its 100 ms
    ring) bit for bit, so the ring/tick machinery does not influence epoch results."""
    g = golden("loop.npz")
    for name in ("fs4", "fs25"):
        meta, prns = g[f"{name}_meta"], g[f"{name}_prns"]
        sc, iq = H.loop_input(meta, prns)
        assert H.sha(iq) == str(g[f"{name}_sha"])
        x = synth.to_complex(iq)
        fs = float(meta[0])
        n_code = round(fs * 1e-3)
        for prn in prns:
            acq, trk = g[f"{name}_acq_{int(prn)}"], g[f"{name}_trk_{int(prn)}"]
            carrier, code_off, cur = O.acquisition_handoff(int(acq[1]), int(acq[2]), 0.0, 5000.0, float(meta[4]), 0,
                                                           10 * n_code, int(np.ceil(1023 / (1.023e6 / fs))))
            assert carrier == acq[4]
            tr = O.BorreTrackOracle(int(prn), fs, carrier, cur)
            n_ep = min(len(trk), 120 if name == "fs25" else 400)
            # the reference's currentSample wraps in the 100 ms ring; ours is absolute
            ring = int(fs * 0.1)
            for k in range(n_ep):
                r = tr.step(x)
                assert np.array_equal(np.array(r["corr"]), trk[k, 1:7]), (name, prn, k)
                assert r["carrier_frequency"] == trk[k, 9] and r["code_frequency"] == trk[k, 10]
                assert tr.rem_code == trk[k, 14] and tr.rem_carrier == trk[k, 15]
                assert tr.n_req == int(trk[k, 16]) and tr.cur % ring == int(trk[k, 17])


REF_SO = os.path.join(H.ROOT, "oracle", "_ref", "tracking.so")


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/tracking.so not built (needs /root/reference)")
def test_nav_bit_oracle_matches_reference_channel(golden):
    """Bit synchronisation + 20-epoch prompt sums: every tick's navPromptSum, counters and flag of
    the live reference channel, and its final navBitsBuffer."""
    g = golden("nav.npz")
    for prn in g["prns"]:
        ep, bits = g[f"epochs_{prn}"], g[f"bits_{prn}"]
        o = O.NavBitOracle()
        for r in ep:
            o.step(r[0])
            assert o.nav_sum == r[1] and o.nav_count == r[2] and len(o.bits) == r[3] and float(o.bit_sync) == r[4]
        assert np.array_equal(np.array(o.bits, dtype=np.int8), bits) and len(bits) >= 50
    b, s, sync, _ = O.nav_bits(g[f"epochs_{g['prns'][0]}"][:, 0])
    assert sync > 100 and len(b) == len(s)


def test_kaplan_nav_bit_oracle_matches_reference_channel(golden):
    g = golden("nav.npz")
    for prn in g["prns"]:
        ep, bits = g[f"kepochs_{prn}"], g[f"kbits_{prn}"]
        b, s, sync, (pend, cnt) = O.nav_bits_kaplan(ep[:, 0], ep[:, 4] >= 2)
        assert np.array_equal(b, bits)
        assert pend == ep[-1, 1] and cnt == ep[-1, 2] and len(b) == ep[-1, 3]
    assert len(g["kbits_22"]) >= 40 and len(g["kbits_7"]) == 0          # one channel never reaches bit sync in 1.3 s


def test_kaplan_loop_oracle_matches_reference_channel(golden):
    """KaplanTrackOracle (FLL-assisted PLL, lock indicators, C/N0, PULL_IN/WIDE/NARROW, code lock, bit
    sync) teacher-forced with the live reference channel's correlator sums: every packet field of all
    2590 epochs, exactly."""
    g = golden("kaplan.npz")
    cols = dict(dll=7, pll=8, fll=9, carrier_frequency=10, code_frequency=11, carrier_frequency_error=12,
                code_frequency_error=13, cn0=14, pll_lock=15, fll_lock=16, lock_state=17, rem_code=19, rem_carrier=20,
                n_req=21)
    for prn in (int(p) for p in g["prns"]):
        ref, ra = g[f"trk_{prn}"], g[f"acq_{prn}"]
        o = O.KaplanTrackOracle(prn, float(g["meta"][0]), float(ra[4]), 0)
        for k in range(len(ref)):
            r = o.step(None, corr_override=ref[k, 1:7])
            for name, col in cols.items():
                assert float(r[name]) == ref[k, col], (prn, k, name)
            assert (r["flags"] & 3) == (int(ref[k, 18]) & 3), (prn, k)
        assert {1, 2, 3} <= set(ref[:, 17].astype(int))


def test_against_compiled_reference_c(golden):
    """getCorrelator of the reference's own tracking.c (compiled by oracle/Makefile) vs the oracle."""
    lib = ctypes.CDLL(REF_SO)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.getCorrelator.argtypes = [dp, dp, ctypes.POINTER(ctypes.c_int), ctypes.c_size_t, ctypes.c_double,
                                  ctypes.c_double, ctypes.c_double, dp, dp]
    lib.getCorrelator.restype = None
    g = golden("epl.npz")
    u = g["unit_iq"].astype(np.float64)
    rf = u[0::2] + 1j * u[1::2]
    n = len(rf)
    t = np.arange(n) / 10e6
    sig = np.exp(1j * (-(3700.0 * 2.0 * np.pi * t) + 0.0)) * rf
    i_s, q_s = np.ascontiguousarray(sig.real), np.ascontiguousarray(sig.imag)
    code = O.padded_code(2).astype(np.int32)
    ref = g["unit_epl"]
    for k, sp in enumerate((-0.5, 0.0, 0.5)):
        ri, rq = ctypes.c_double(), ctypes.c_double()
        lib.getCorrelator(i_s.ctypes.data_as(dp), q_s.ctypes.data_as(dp), code.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                          n, 1.023e6 / 10e6, 0.0, sp, ctypes.byref(ri), ctypes.byref(rq))
        assert abs(ri.value - ref[2 * k]) <= 1e-11 * abs(ref[2 * k]) + 1e-9
        assert abs(rq.value - ref[2 * k + 1]) <= 1e-11 * abs(ref[2 * k + 1]) + 1e-9
