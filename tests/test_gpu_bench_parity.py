"""Parity of exactly what bench.py times ("Next round" item 1).

(a) the headline chunk -- 2 s x 12 channels of 25 MS/s int16 IQ through ColdStartPool (five steps in
    flight, DENSE tracking instantiation), closed loop on the device, against BorreTrackOracle:
      * every epoch re-evaluated by the oracle from the kernel's own pre-epoch state (teacher forcing):
        six correlator sums within 1e-4 of the prompt magnitude, loop outputs equal to 1e-9;
      * the oracle run closed loop on its own: carrier / code frequency within 0.5 Hz and absolute code
        phase within 1e-3 chip over the whole chunk (north_star tolerances);
(b) the throughput instantiation (BASELINE configs[4] shape: several recordings x 12 channels in one
    launch, one CTA per channel / one cluster per recording) with the same two comparisons and
    status == 0 on every channel;
(c) lives in test_gpu_acq.py (cfg4 golden widened to 8 present + 4 absent PRNs);
(d) DENSE added to the closed-loop configurations checked against the reference channel's packets.

Reference: sydr/channel/channel_l1ca_borre.py:333-451, sydr/dsp/tracking.py:92-186.
The oracle legs run one process per channel (fork; the samples are inherited, not pickled).
"""
import multiprocessing as mp
import os

import numpy as np
import pytest

import helpers as H  # noqa: F401  (sys.path)

pytestmark = pytest.mark.gpu

FS = 25e6
TOL_CORR = 1e-4
TOL_HZ = 0.5
TOL_CHIP = 1e-3

_X = None           # complex samples of the recording under test, inherited by the forked workers


def _oracle_channel(a):
    """Teacher-forced and free-running oracle of one channel.  Returns the error figures."""
    from oracle import sydr_oracle as O
    prn, carrier, start, rec = a
    x = _X
    n_ep = len(rec)
    # ---- teacher forced: the oracle computes the sums from the kernel's state, then takes the kernel's sums
    tf = O.BorreTrackOracle(prn, FS, carrier, start)
    e_corr = np.zeros(n_ep)
    e_state = 0.0
    bad_shape = 0
    for k in range(n_ep):
        if tf.cur != int(rec["start"][k]) or tf.n_req != int(rec["n"][k]):
            bad_shape += 1
            break
        o = tf.step(x, corr_override=rec["corr"][k])
        ref = np.asarray(o["corr"])
        e_corr[k] = np.abs(rec["corr"][k] - ref).max() / np.hypot(ref[2], ref[3])
        e_state = max(e_state,
                      abs(rec["carrier_freq"][k] - o["carrier_frequency"]) / max(1.0, abs(o["carrier_frequency"])),
                      abs(rec["code_freq"][k] - o["code_frequency"]) / o["code_frequency"],
                      abs(rec["rem_code"][k] - tf.rem_code), abs(rec["rem_carrier"][k] - tf.rem_carrier))
    # ---- free running: the oracle closes its own loops
    fr = O.BorreTrackOracle(prn, FS, carrier, start)
    d_car = d_code = d_phase = 0.0
    for k in range(n_ep):
        if fr.cur + fr.n_req > len(x):
            break
        o = fr.step(x)
        d_car = max(d_car, abs(o["carrier_frequency"] - rec["carrier_freq"][k]))
        d_code = max(d_code, abs(o["code_frequency"] - rec["code_freq"][k]))
        # absolute code phase of the next epoch: end sample - remCode / codeStep, in chips
        ours = (rec["start"][k] + rec["n"][k]) - rec["rem_code"][k] / (rec["code_freq"][k] / FS)
        theirs = fr.cur - fr.rem_code / fr.code_step
        d_phase = max(d_phase, abs(ours - theirs) * (1.023e6 / FS))
    return dict(prn=prn, bad_shape=bad_shape, e_corr=float(e_corr.max()), e_corr_at=int(e_corr.argmax()),
                e_state=float(e_state), d_car=d_car, d_code=d_code, d_phase=d_phase, epochs=n_ep)


def check_against_oracle(x, chans, recs):
    """x: complex samples per recording (list indexed by chans[i]['rec']) or one array."""
    global _X
    out = []
    by_rec = {}
    for c, r in zip(chans, recs):
        by_rec.setdefault(c.get("rec", 0), []).append((c["prn"], c["carrier_freq"], c["start_sample"], r))
    ctx = mp.get_context("fork")
    for rec_i, tasks in by_rec.items():
        _X = x[rec_i] if isinstance(x, list) else x
        with ctx.Pool(min(len(tasks), len(os.sched_getaffinity(0)))) as pool:
            out += pool.map(_oracle_channel, tasks)
        _X = None
    for o in out:
        assert o["bad_shape"] == 0, f"PRN {o['prn']}: epoch boundaries differ from the reference arithmetic"
        assert o["e_corr"] <= TOL_CORR, f"PRN {o['prn']}: correlators off by {o['e_corr']:.2e} at epoch {o['e_corr_at']}"
        assert o["e_state"] <= 1e-9, f"PRN {o['prn']}: loop closure differs ({o['e_state']:.2e})"
        assert o["d_car"] <= TOL_HZ and o["d_code"] <= TOL_HZ, (o["prn"], o["d_car"], o["d_code"])
        assert o["d_phase"] <= TOL_CHIP, (o["prn"], o["d_phase"])
    return out


def to_c64(iq_int16: np.ndarray) -> np.ndarray:
    """int16 I/Q pairs as complex64: exact (|v| < 2^24), half the memory of the reference's complex128."""
    v = iq_int16.astype(np.float32)
    return v.view(np.complex64)


def test_headline_chunk_dense_pool_vs_oracle():
    """bench.py's timed step: ColdStartPool(lanes=5) -> acquisition, device hand-off, DENSE 12-channel tracking
    of a 2 s chunk; five steps in flight, every lane's records compared."""
    import torch
    from sydr_b200 import synth
    from sydr_b200.pipeline import ColdStartPool
    chunk_s, lanes = 2.0, 5
    sc = synth.make_scenario(FS, 16, chunk_s, synth.PRNS_12, 1003, 250.0)
    d = synth.generate_iq_torch(sc, device="cuda")
    pool = ColdStartPool(lanes=lanes, fs=FS, nbits=16, search_prns=list(range(1, 33)), n_channels=12, max_seconds=chunk_s,
                         doppler_range=5000.0, doppler_step=250.0, coh=1, noncoh=10)
    assert all(p._dense for p in pool.lanes)
    d_iq = pool.lanes[0].device_buffer(d.numel() // 2)
    d_iq.copy_(d)
    torch.cuda.synchronize()
    tickets = [pool.submit_device(d_iq) for _ in range(lanes)]
    outs = [pool.result(t, records=True, copy=True) for t in tickets]
    pool.close()
    first = outs[0]
    assert [c["prn"] for c in first["channels"]] == list(synth.PRNS_12)
    for o in outs[1:]:                                     # steps in flight beside each other: the same bits
        assert o["peaks"].tobytes() == first["peaks"].tobytes()
        for a, b in zip(o["epochs"], first["epochs"]):
            assert a.tobytes() == b.tobytes()
    assert min(len(e) for e in first["epochs"]) >= 1985
    x = to_c64(d.cpu().numpy())
    del d
    res = check_against_oracle(x, first["channels"], first["epochs"])
    print("headline chunk vs oracle:", {k: max(r[k] for r in res) for k in ("e_corr", "e_state", "d_car", "d_code", "d_phase")})


SHAPES = {"cta_per_channel": dict(cluster=1, threads=256, use_tma=False, kernel=2),      # LEAN instantiation of trk.cu
          "pack": dict(cluster=1, threads=0, use_tma=True, dense=2),                     # PACK instantiation (ColdStartBatch, bench.py's timed launch)
          "moments_g3": dict(kernel=1, group=3), "moments_g4": dict(kernel=1, group=4), "moments_g1": dict(kernel=1, group=1)}


def test_batch_step_vs_pipeline_and_oracle():
    """bench.py's timed step: ColdStartBatch -- B recordings (distinct seeds) back to back in one buffer, B acquisitions
    and device hand-offs, ONE tracking launch of B x 12 channels with the PACK instantiation.  Every recording against
    (1) ColdStartPipeline on the same samples (same peak table, same channels, same epoch boundaries, loop outputs within
    the north-star tolerances) and (2) the oracle, teacher-forced and closed loop."""
    import torch
    from sydr_b200 import synth
    from sydr_b200.pipeline import ColdStartBatch, ColdStartPipeline
    B, chunk_s = 3, 1.0
    kw = dict(fs=FS, nbits=16, search_prns=list(range(1, 33)), n_channels=12, max_seconds=chunk_s,
              doppler_range=5000.0, doppler_step=250.0, coh=1, noncoh=10)
    batch = ColdStartBatch(B, **kw)
    xs = []
    for r in range(B):
        sc = synth.make_scenario(FS, 16, chunk_s, synth.PRNS_12, 1103 + r, 250.0)
        batch.slot(r).copy_(synth.generate_iq_torch(sc, device="cuda"))
        xs.append(to_c64(batch.slot(r).cpu().numpy()))
    torch.cuda.synchronize()
    outs = batch.process(records=True, copy=True)
    st = batch._trk.states()
    assert (st["status"] == 0).all()
    again = batch.process(records=True, copy=True)              # a second step over the same buffer: the same bits
    pipe = ColdStartPipeline(**kw)
    chans, recs = [], []
    for r, (o, o2) in enumerate(zip(outs, again)):
        assert [c["prn"] for c in o["channels"]] == list(synth.PRNS_12)
        assert o["peaks"].tobytes() == o2["peaks"].tobytes()
        for a, b in zip(o["epochs"], o2["epochs"]):
            assert a.tobytes() == b.tobytes()
        ref = pipe.finish(pipe.enqueue_device(batch.slot(r)), records=True, copy=True)
        assert ref["peaks"].tobytes() == o["peaks"].tobytes()
        for c, cr, e, er in zip(o["channels"], ref["channels"], o["epochs"], ref["epochs"]):
            assert (c["prn"], c["carrier_freq"], c["start_sample"]) == (cr["prn"], cr["carrier_freq"], cr["start_sample"])
            n = min(len(e), len(er))
            assert n >= 985 and abs(len(e) - len(er)) <= 1
            # (two closed loops whose partial sums are added in different orders: an epoch boundary may fall one sample
            # apart where the code phase sits within 1e-9 of a sample instant; the absolute code phase is what must agree)
            ph = lambda t: ((t["start"][:n] + t["n"][:n]) - t["rem_code"][:n] / (t["code_freq"][:n] / FS)) * (1.023e6 / FS)
            assert np.abs(ph(e) - ph(er)).max() <= TOL_CHIP
            assert np.abs(e["start"][:n] - er["start"][:n]).max() <= 1
            assert np.abs(e["carrier_freq"][:n] - er["carrier_freq"][:n]).max() <= TOL_HZ
            assert np.abs(e["code_freq"][:n] - er["code_freq"][:n]).max() <= TOL_HZ
            # (the correlator sums of two closed loops are not compared: their carrier phases differ by milliradians, which
            # turns I into Q by as much; the 1e-4 tolerance is checked teacher-forced against the oracle below)
        chans += o["channels"]
        recs += o["epochs"]
    pipe.close()
    batch.close()
    res = check_against_oracle(xs, chans, recs)
    print("batch step vs oracle:", {k: max(r[k] for r in res) for k in ("e_corr", "e_state", "d_car", "d_code", "d_phase")})


def test_more_channels_than_a_wave():
    """Automatic shape beyond 296 channels: waves of the PACK shape (296 channels per launch), the remainder in its own
    launch (here 4 channels: clusters of 8).  25 copies of the 12 channels of one recording + the 12 tracked alone: every
    copy inside a wave equals the first bit for bit, all of them follow the lone run within the loop tolerances."""
    import torch
    from sydr_b200 import synth
    from sydr_b200.engine import AcquisitionEngine, TrackingEngine, make_trk_states
    seconds = 0.25
    n = int(round(seconds * FS))
    sc = synth.make_scenario(FS, 16, seconds, synth.PRNS_12, 1201, 250.0)
    buf = torch.zeros(2 * n + 4096, dtype=torch.int16, device="cuda")
    buf[:2 * n] = synth.generate_iq_torch(sc, device="cuda")
    acq = AcquisitionEngine(FS, 0.0, 5000.0, 250.0, 1, 10, list(synth.PRNS_12))
    chans = []
    for p in acq.run(buf[:2 * n])["peaks"]:
        carrier, _, cur = acq.handoff(p)
        chans.append(dict(prn=int(p["prn"]), carrier_freq=carrier, start_sample=cur, iq_len=n))
    acq.close()
    assert len(chans) == 12
    many = chans * 25                                        # 300 channels: one wave of 296 + 4
    eng = TrackingEngine(FS, make_trk_states(FS, many), int(seconds * 1000) + 8)
    eng.launch(buf[:2 * n])
    recs = eng.fetch()
    assert (eng.states()["status"] == 0).all() and min(len(r) for r in recs) >= 235
    lone = TrackingEngine(FS, make_trk_states(FS, chans), int(seconds * 1000) + 8)
    lone.launch(buf[:2 * n])
    ref = lone.fetch()
    for i, r in enumerate(recs):
        if 12 <= i < 296:
            assert r.tobytes() == recs[i % 12].tobytes()
        e = ref[i % 12]
        m = min(len(r), len(e))
        assert abs(len(r) - len(e)) <= 1
        assert np.abs(r["carrier_freq"][:m] - e["carrier_freq"][:m]).max() <= TOL_HZ
        assert np.abs(r["code_freq"][:m] - e["code_freq"][:m]).max() <= TOL_HZ
        assert np.abs(r["start"][:m] - e["start"][:m]).max() <= 1


@pytest.mark.parametrize("shape", list(SHAPES))
def test_throughput_instantiation_vs_oracle(shape):
    """configs[4] shape: 3 recordings x 12 channels x 0.5 s in one launch of the throughput kernels: the
    prefix-moment kernel (trkm.cu, what bench.throughput_stress launches; 3 / 4 / 1 channels per CTA) and the
    per-channel LEAN instantiation of trk.cu."""
    import torch
    from sydr_b200 import synth
    from sydr_b200.engine import AcquisitionEngine, TrackingEngine, make_trk_states
    n_rec, seconds = 3, 0.5
    n = int(round(seconds * FS))
    pad = 2048
    buf = torch.zeros(n_rec * (2 * n + pad) + 4096, dtype=torch.int16, device="cuda")
    acq = AcquisitionEngine(FS, 0.0, 5000.0, 250.0, 1, 10, list(synth.PRNS_12))
    chans, xs = [], []
    for r in range(n_rec):
        sc = synth.make_scenario(FS, 16, seconds, synth.PRNS_12, 1005 + r, 250.0)
        base = r * (2 * n + pad)
        buf[base:base + 2 * n] = synth.generate_iq_torch(sc, device="cuda")
        xs.append(to_c64(buf[base:base + 2 * n].cpu().numpy()))
        for p in acq.run(buf[base:base + 2 * n])["peaks"]:
            carrier, _, cur = acq.handoff(p)
            chans.append(dict(prn=int(p["prn"]), carrier_freq=carrier, start_sample=cur, iq_base=base // 2, iq_len=n, rec=r))
    acq.close()
    st = make_trk_states(FS, chans)
    eng = TrackingEngine(FS, st, int(seconds * 1000) + 8, **SHAPES[shape])
    eng.launch(buf)
    recs = eng.fetch()
    assert (eng.states()["status"] == 0).all()
    assert min(len(r) for r in recs) >= 485
    res = check_against_oracle(xs, chans, recs)
    print(f"throughput launch ({shape}) vs oracle:", {k: max(r[k] for r in res) for k in ("e_corr", "e_state", "d_car", "d_code", "d_phase")})


def test_full_size_recording_properties():
    """BASELINE.json configs[2] at FULL size, exactly as bench.py's step runs it: 60 s of 25 MS/s int16 IQ (6 GB) resident in
    HBM, ColdStartPool -> 32-PRN acquisition, device hand-off, one DENSE tracking launch of 60 000 epochs x 12 channels.  Size-independent properties (the oracle needs ~15 core-minutes for this much): every
    satellite acquired, every channel tracked to the last millisecond with contiguous epochs, Doppler equal to the
    generator's truth, and every navigation bit after the pull-in second equal to the transmitted data (up to the Costas
    half-cycle ambiguity of the whole stream)."""
    import torch
    from sydr_b200 import synth
    from sydr_b200.pipeline import ColdStartPool
    seconds = 60.0
    sc = synth.make_scenario(FS, 16, seconds, synth.PRNS_12, 1003, 250.0)
    d = synth.generate_iq_torch(sc, device="cuda")
    pool = ColdStartPool(lanes=2, fs=FS, nbits=16, search_prns=list(range(1, 33)), n_channels=12, max_seconds=seconds,
                         doppler_range=5000.0, doppler_step=250.0, coh=1, noncoh=10)
    out = pool.result(pool.submit_device(d), records=True, copy=True)
    pool.close()
    del d
    torch.cuda.empty_cache()
    assert [c["prn"] for c in out["channels"]] == list(synth.PRNS_12)
    truth = synth.nav_bits_of(sc)
    for k, (ch, e) in enumerate(zip(out["channels"], out["epochs"])):
        sat = next(s for s in sc.sats if s.prn == ch["prn"])
        assert len(e) >= 59985, (ch["prn"], len(e))
        assert np.array_equal(e["start"][1:], (e["start"] + e["n"])[:-1])                     # contiguous epochs
        assert (e["start"][-1] + e["n"][-1]) / FS > seconds - 0.002                           # to the last millisecond
        assert abs(float(np.mean(e["carrier_freq"][-500:])) - sat.doppler) < 5.0
        assert np.abs(e["n"] - 25000).max() <= 1
        # navigation bits from the prompt sums (the reference's rule: synchronise on the first sign change after 100 epochs,
        # then 20 prompts per bit, channel_l1ca_borre.py:398-413, 455-470)
        ip = e["corr"][:, 2]
        flips = np.nonzero(np.diff(np.sign(ip)) != 0)[0]
        sync = int(flips[flips >= 100][0] + 1)
        n_bits = (len(ip) - sync) // 20
        b = (ip[sync:sync + 20 * n_bits].reshape(n_bits, 20).sum(axis=1) > 0).astype(np.int8)
        t_mid = (e["start"][sync] + 10 * FS * 1e-3 + 20 * FS * 1e-3 * np.arange(n_bits)) / FS
        tx = (truth[ch["prn"]][synth.nav_bit_index(sat, t_mid)] > 0).astype(np.int8)
        late = t_mid > 1.0
        agree = float((tx[late] == b[late]).mean())
        assert agree in (0.0, 1.0) and n_bits >= 2980, (ch["prn"], agree, n_bits)


def test_long_horizon_vs_oracle():
    """10 s x 2 channels (20 000 epochs) of the latency instantiation against BorreTrackOracle, teacher-forced on every epoch
    and free-running over the whole horizon: the closed loop neither drifts nor mis-assigns a sample over 10 000 epochs."""
    import torch
    from sydr_b200 import synth
    from sydr_b200.engine import AcquisitionEngine, TrackingEngine, make_trk_states
    seconds, prns = 10.0, (7, 22)
    sc = synth.make_scenario(FS, 16, seconds, prns, 1711, 250.0)
    d = synth.generate_iq_torch(sc, device="cuda")
    acq = AcquisitionEngine(FS, 0.0, 5000.0, 250.0, 1, 10, list(prns))
    chans = []
    for p in acq.run(d[:2 * 250000])["peaks"]:
        carrier, _, cur = acq.handoff(p)
        chans.append(dict(prn=int(p["prn"]), carrier_freq=carrier, start_sample=cur, iq_len=d.numel() // 2))
    acq.close()
    eng = TrackingEngine(FS, make_trk_states(FS, chans), int(seconds * 1000) + 8)
    recs = eng.run(d)
    assert min(len(r) for r in recs) >= 9985
    x = to_c64(d.cpu().numpy())
    del d
    res = check_against_oracle(x, chans, recs)
    print("10 s horizon vs oracle:", {k: max(r[k] for r in res) for k in ("e_corr", "e_state", "d_car", "d_code", "d_phase")})
