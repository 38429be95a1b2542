"""Kaplan loop closure on the device (channel_l1ca_kaplan.py:342-619) against the live
reference channel's packets (tests/golden/kaplan.npz) and, teacher-forced, against KaplanTrackOracle."""
import numpy as np
import pytest

import helpers as H
from oracle import sydr_oracle as O

pytestmark = pytest.mark.gpu
COLS = dict(corr=slice(1, 7), dll=7, pll=8, fll=9, carrier=10, code=11, cerr=12, derr=13, cn0=14, pll_lock=15, fll_lock=16,
            state=17, flags=18, rem_code=19, rem_carrier=20, n_req=21)


def _run(g, cluster=0, pieces=1):
    from sydr_b200 import synth
    from sydr_b200.engine import KaplanTrackingEngine, make_kaplan_states, to_device_iq
    fs, nbits, seed, ms, ds = g["meta"]
    prns = [int(p) for p in g["prns"]]
    sc = synth.make_scenario(float(fs), int(nbits), int(ms) * 1e-3, tuple(prns), int(seed), float(ds))
    iq = synth.generate_iq(sc)
    assert H.sha(iq) == str(g["sha"])
    n = len(iq) // 2
    chans = []
    for p in prns:
        ra = g[f"acq_{p}"]                      # tick, freq_idx, code_idx, ratio, carrierFrequency, currentSample (ring index)
        start = 10 * 4000 - 4001 + int(ra[2]) + 1                     # channel_l1ca_kaplan.py:223-240 from sample 0
        chans.append(dict(prn=p, carrier_freq=float(ra[4]), start_sample=start, iq_len=n))
    st, ks = make_kaplan_states(float(fs), chans, H.MG.KAPLAN_TRK_CFG)
    eng = KaplanTrackingEngine(float(fs), st, ks, int(ms) + 8, cluster=cluster)
    d = to_device_iq(iq)
    if pieces == 1:
        eng.launch(d)
    else:
        for k in range(1, pieces + 1):
            eng.launch(d, iq_len=n * k // pieces, append=True)
    return prns, iq, chans, eng.fetch(), eng.fetch_kaplan(), eng


@pytest.mark.parametrize("cluster,pieces", [(0, 1), (2, 3), (1, 1)])
def test_device_kaplan_follows_the_reference_channel(golden, cluster, pieces):
    g = golden("kaplan.npz")
    prns, iq, chans, recs, kex, eng = _run(g, cluster, pieces)
    for c, p in enumerate(prns):
        ref = g[f"trk_{p}"]
        r, k = recs[c], kex[c]
        n = min(len(r), len(ref))
        assert n >= len(ref) - 2 and len(k) == len(r)
        ref = ref[:n]
        # loop outputs within the north star's tolerances
        assert np.abs(r["carrier_freq"][:n] - ref[:, COLS["carrier"]]).max() <= 0.5
        assert np.abs(r["code_freq"][:n] - ref[:, COLS["code"]]).max() <= 0.5
        same = r["n"][1:n] == ref[:-1, COLS["n_req"]]                 # epoch lengths (n_req after epoch k = n of k+1)
        assert same.mean() > 0.99
        e = np.abs(r["corr"][:n] - ref[:, COLS["corr"]]).max(axis=1) / np.hypot(ref[:, 3], ref[:, 4])
        assert np.median(e) <= 1e-3
        # the state machine takes the same path (transitions may move by an epoch or two)
        for st in (2, 3):
            a, b = int(np.argmax(k["lock_state"][:n] == st)), int(np.argmax(ref[:, COLS["state"]] == st))
            assert a > 0 and abs(a - b) <= 3, (p, st, a, b)
        assert (int(k["flags"][n - 1]) & 3) == (int(ref[-1, COLS["flags"]]) & 3) == 3
        assert abs(k["cn0"][n - 1] - ref[-1, COLS["cn0"]]) <= 0.05 * ref[-1, COLS["cn0"]]
        assert np.abs(k["fll_lock"][:n] - ref[:, COLS["fll_lock"]]).max() <= 0.05
        assert np.abs(k["pll_lock"][:n] - ref[:, COLS["pll_lock"]]).max() <= 0.05


def test_device_kaplan_loop_math_teacher_forced(golden):
    """Every epoch's loop update recomputed by the oracle from the device's own correlator sums and
    previous state: the FP64 loop closure on the device equals the Python floats to rounding."""
    g = golden("kaplan.npz")
    prns, iq, chans, recs, kex, eng = _run(g)
    x = iq[0::2] + 1j * iq[1::2]
    for c, p in enumerate(prns):
        o = O.KaplanTrackOracle(p, float(g["meta"][0]), chans[c]["carrier_freq"], chans[c]["start_sample"])
        r, k = recs[c], kex[c]
        for e in range(len(r)):
            assert int(r["start"][e]) == o.cur and int(r["n"][e]) == o.n_req, (p, e)
            if e % 97 == 0:                                            # open-loop correlator check on a subset
                want = np.array(O.epl(x[o.cur:o.cur + o.n_req], o.code, o.fs, o.carrier_freq, o.rem_carrier, o.rem_code,
                                      o.code_step, o.spacings))
                assert np.abs(r["corr"][e] - want).max() <= 1e-4 * np.hypot(want[2], want[3]), (p, e)
            w = o.step(None, corr_override=r["corr"][e])
            tol = lambda a, b, rel=1e-9, ab=1e-12: abs(a - b) <= rel * abs(b) + ab
            assert tol(r["carrier_freq"][e], w["carrier_frequency"]) and tol(r["code_freq"][e], w["code_frequency"]), (p, e)
            assert tol(r["pll"][e], w["carrier_frequency_error"], 1e-7, 1e-10), (p, e)
            assert tol(r["dll"][e], w["code_frequency_error"], 1e-7, 1e-10), (p, e)
            assert tol(r["carrier_err"][e], w["pll"], 1e-9, 1e-13) and tol(k["fll"][e], w["fll"], 1e-7, 1e-9), (p, e)
            assert tol(r["code_err"][e], w["dll"], 1e-9, 1e-13), (p, e)
            assert tol(r["rem_carrier"][e], w["rem_carrier"], 1e-9, 1e-9) and tol(r["rem_code"][e], w["rem_code"], 1e-9, 1e-9)
            assert tol(k["cn0"][e], w["cn0"]) and tol(k["fll_lock"][e], w["fll_lock"]) and tol(k["pll_lock"][e], w["pll_lock"])
            assert int(k["lock_state"][e]) == w["lock_state"] and int(k["flags"][e]) == w["flags"], (p, e)
            # keep the oracle on the device's trajectory (differences at rounding level would otherwise add up)
            o.carrier_freq, o.code_freq = float(r["carrier_freq"][e]), float(r["code_freq"][e])
            o.rem_carrier, o.rem_code = float(r["rem_carrier"][e]), float(r["rem_code"][e])
            o.code_step = o.code_freq / o.fs
            o.n_req = int(np.ceil((O.CODE_CHIPS - o.rem_code) / o.code_step))
            o.vel_memory = o.vel_memory
            o.cn0, o.fll_lock, o.pll_lock = float(k["cn0"][e]), float(k["fll_lock"][e]), float(k["pll_lock"][e])
        ks = eng.kaplan_states()[c]
        assert int(ks["code_counter"]) == len(r) and int(ks["lock_state"]) == 3


def test_pipeline_with_kaplan_loops_at_25_msps():
    """ColdStartPipeline(loop="kaplan") at the headline rate (int16, 25 MS/s, clusters of 8): acquisition, device
    hand-off, Kaplan tracking; every epoch's loop update teacher-forced against the oracle."""
    from sydr_b200 import synth
    from sydr_b200.engine import to_device_iq
    from sydr_b200.pipeline import ColdStartPipeline
    fs, prns = 25e6, (11, 27)
    sc = synth.make_scenario(fs, 16, 0.35, prns, 55, 250.0)
    iq = synth.generate_iq(sc)
    pipe = ColdStartPipeline(fs, 16, list(range(1, 33)), 4, max_seconds=0.35, loop="kaplan")
    out = pipe.finish(pipe.enqueue_device(to_device_iq(iq)), records=True, copy=True)
    assert [c["prn"] for c in out["channels"]] == list(prns)
    for ch, r, k in zip(out["channels"], out["epochs"], out["kaplan"]):
        assert len(r) == len(k) >= 335
        o = O.KaplanTrackOracle(ch["prn"], fs, ch["carrier_freq"], ch["start_sample"])
        for e in range(len(r)):
            assert int(r["start"][e]) == o.cur and int(r["n"][e]) == o.n_req
            w = o.step(None, corr_override=r["corr"][e])
            assert abs(r["carrier_freq"][e] - w["carrier_frequency"]) <= 1e-9 * abs(w["carrier_frequency"]) + 1e-10
            assert abs(k["fll_lock"][e] - w["fll_lock"]) <= 1e-9 and abs(k["cn0"][e] - w["cn0"]) <= 1e-9 * abs(w["cn0"]) + 1e-12
            assert int(k["lock_state"][e]) == w["lock_state"] and int(k["flags"][e]) == w["flags"]
            o.carrier_freq, o.code_freq = float(r["carrier_freq"][e]), float(r["code_freq"][e])
            o.rem_carrier, o.rem_code = float(r["rem_carrier"][e]), float(r["rem_code"][e])
            o.code_step = o.code_freq / o.fs
            o.n_req = int(np.ceil((O.CODE_CHIPS - o.rem_code) / o.code_step))
            o.cn0, o.fll_lock, o.pll_lock = float(k["cn0"][e]), float(k["fll_lock"][e]), float(k["pll_lock"][e])
        sat = [s for s in sc.sats if s.prn == ch["prn"]][0]
        assert abs(float(np.mean(r["carrier_freq"][-100:])) - sat.doppler) < 5.0
    pipe.close()


def test_receiver_with_the_kaplan_ini(tmp_path):
    """ReceiverGPSL1CA picks the Kaplan channel from the channel ini: run() ticks the host class over the GPU
    correlators, run_fast() runs the loop closure on the device; both write Kaplan tracking rows that agree
    within the north star's tolerances, with the same lock-state path."""
    import configparser
    import os
    from sydr_b200 import synth
    from sydr_b200.receiver.receiver_gps_l1ca import ReceiverGPSL1CA
    fs, nbits, ms = 4e6, 8, 700
    sc = synth.make_scenario(fs, nbits, ms * 1e-3 + 0.13, (3, 19), 31, 250.0)
    path = str(tmp_path / "rec.bin")
    synth.write_file(path, synth.generate_iq(sc))

    def config(name):
        cfg = configparser.ConfigParser()
        cfg.read(os.path.join(H.ROOT, "config", "receiver.ini"))
        cfg["DEFAULT"].update({"name": name, "ms_to_process": str(ms), "outfolder": str(tmp_path)})
        cfg["RFSIGNAL"].update({"filepath": path, "sampling_frequency": str(fs), "data_size": str(nbits)})
        cfg["SATELLITES"]["include_prn"] = "3,19"
        cfg["CHANNELS"]["gps_l1ca"] = os.path.join(H.ROOT, "config", "channels", "channel_GPS_L1CA_kaplan.ini")
        return cfg

    # the streaming path with the Kaplan loops, bits included: equal to the oracle's rule on the same records
    from sydr_b200.ingest import StreamingReceiver
    from sydr_b200.signal.rfsignal import RFSignal
    rx = StreamingReceiver(RFSignal(dict(config("x")["RFSIGNAL"])), [3, 19], 2, chunk_seconds=0.2, loop="kaplan",
                           channel_cfg=H.MG.KAPLAN_TRK_CFG)
    so = rx.run_all()
    rx.close()
    for r, k, bits in zip(so["epochs"], so["kaplan"], so["bits"]):
        ob = O.nav_bits_kaplan(r["corr"][:, 2], (k["flags"] & 2) != 0)[0]
        assert np.array_equal(bits, ob)
    assert max(len(b_) for b_ in so["bits"]) >= 10
    b = ReceiverGPSL1CA(config("fast"), overwrite=True)
    assert b.channelClass.__name__ == "ChannelL1CA_Kaplan"
    b.run_fast(chunk_seconds=0.2)
    rows_b = {c: b.database.fetchTracking(c) for c in (0, 1)}
    b.close()
    a = ReceiverGPSL1CA(config("tick"), overwrite=True)
    a.run()
    rows_a = {c: a.database.fetchTracking(c) for c in (0, 1)}
    a.close()
    for c in (0, 1):
        ra, rb = rows_a[c], rows_b[c]
        n = min(len(ra), len(rb))
        assert n >= ms - 15
        cf = lambda rows, k: np.array([r[k] for r in rows[:n]], dtype=np.float64)
        assert np.abs(cf(ra, "carrier_frequency") - cf(rb, "carrier_frequency")).max() <= 0.5
        assert np.abs(cf(ra, "code_frequency") - cf(rb, "code_frequency")).max() <= 0.5
        assert np.abs(cf(ra, "fll_lock") - cf(rb, "fll_lock")).max() <= 0.05
        sa, sb = cf(ra, "lock_state"), cf(rb, "lock_state")
        assert set(sa.astype(int)) == set(sb.astype(int)) and 2 in set(sb.astype(int))
        for st in sorted(set(sb.astype(int)) - {1}):
            assert abs(int(np.argmax(sa == st)) - int(np.argmax(sb == st))) <= 3
        assert set(ra[0]) == set(rb[0])                                   # same columns
        # run() batches the Kaplan channels on the device too (ChannelManager): the very same trajectory
        for key in ("i_prompt", "carrier_frequency", "code_frequency", "cn0", "fll_lock", "pll_lock", "lock_state", "fll"):
            assert [r[key] for r in ra[:n]] == [r[key] for r in rb[:n]], key
