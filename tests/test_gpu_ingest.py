"""Streaming ingest: IQ file -> pinned buffers -> sliding device window.  The
results must not depend on how the file is cut into chunks: bit-identical to the whole recording
resident in HBM, and (through that path's own tests) to the reference channel."""
import numpy as np
import pytest

import helpers  # noqa: F401
from oracle import sydr_oracle as O

pytestmark = pytest.mark.gpu


def _rf(path, fs, nbits):
    from sydr_b200.signal.rfsignal import RFSignal
    return RFSignal({"filepath": path, "sampling_frequency": str(fs), "is_complex": "true",
                     "intermediate_frequency": "0.0", "data_size": str(nbits)})


@pytest.mark.parametrize("fs,nbits,dur,chunk_s,prns", [
    (4e6, 8, 1.25, 0.11, (3, 7, 19)),          # many chunks, ragged last one, bit sync + bits across chunk borders
    (25e6, 16, 0.30, 0.07, (11, 27)),          # int16 at the headline rate (segment path, clusters of 8)
    (4e6, 8, 0.20, 5.0, (7,)),                 # file shorter than one chunk
])
def test_streaming_equals_resident(tmp_path, fs, nbits, dur, chunk_s, prns):
    from sydr_b200 import synth
    from sydr_b200.engine import NavBitEngine, to_device_iq
    from sydr_b200.ingest import StreamingReceiver
    from sydr_b200.pipeline import ColdStartPipeline
    sc = synth.make_scenario(fs, nbits, dur, prns, 91, 250.0)
    iq = synth.generate_iq(sc)
    path = str(tmp_path / "rec.bin")
    synth.write_file(path, iq)

    # whole recording resident in HBM, one launch
    pipe = ColdStartPipeline(fs, nbits, list(range(1, 33)), 8, max_seconds=dur)
    ref = pipe.process_device(to_device_iq(iq))
    nav = NavBitEngine(len(ref["channels"]), max_bits=256)
    nav.launch(pipe._trk)
    ref_bits = [b for b, _ in nav.fetch()]
    ref_ep = pipe.collect()
    pipe.close()

    rx = StreamingReceiver(_rf(path, fs, nbits), list(range(1, 33)), 8, chunk_seconds=chunk_s)
    out = rx.run_all()
    rx.close()
    assert [c["prn"] for c in out["channels"]] == [c["prn"] for c in ref["channels"]] == sorted(prns)
    assert np.array_equal(out["peaks"], ref["peaks"])
    for c in range(len(prns)):
        assert len(out["epochs"][c]) == len(ref_ep[c]) >= int(dur * 1000) - 15
        assert out["epochs"][c].tobytes() == ref_ep[c].tobytes()          # bit-identical trajectory
        assert np.array_equal(out["bits"][c], ref_bits[c])
        o_bits = O.nav_bits(ref_ep[c]["corr"][:, 2])[0]
        assert np.array_equal(out["bits"][c], o_bits)
    if dur > 1.0:
        assert all(len(b) >= 50 for b in out["bits"])


def test_streaming_bits_only_and_skip(tmp_path):
    """want_records=False returns only bits and final states; skip_samples / max_samples select a span."""
    from sydr_b200 import synth
    from sydr_b200.ingest import StreamingReceiver
    fs, nbits = 4e6, 8
    sc = synth.make_scenario(fs, nbits, 0.9, (3, 19), 92, 250.0)
    iq = synth.generate_iq(sc)
    path = str(tmp_path / "rec.bin")
    synth.write_file(path, iq)
    skip, span = 40000, 3_000_000
    full = StreamingReceiver(_rf(path, fs, nbits), [3, 19, 5], 4, chunk_seconds=0.25)
    a = full.run_all(skip_samples=skip, max_samples=span)
    full.close()
    lean = StreamingReceiver(_rf(path, fs, nbits), [3, 19, 5], 4, chunk_seconds=0.4, want_records=False)
    b = lean.run_all(skip_samples=skip, max_samples=span)
    lean.close()
    assert b["epochs"] is None and [c["prn"] for c in b["channels"]] == [3, 19]
    for c in range(2):
        assert np.array_equal(a["bits"][c], b["bits"][c]) and len(a["bits"][c]) > 25
        assert a["states"][c].tobytes() == b["states"][c].tobytes()
        last = a["epochs"][c][-1]
        assert last["start"] + last["n"] <= span and last["start"] + 2 * last["n"] > span - 8


def test_file_to_database(tmp_path):
    """File -> StreamingReceiver -> SQLite in the reference's format: every epoch is a row whose
    values are the device records; cn0 follows the bit-sync rule; time_sample is the emitting tick."""
    from sydr_b200 import synth
    from sydr_b200.ingest import StreamingReceiver
    from sydr_b200.io.database import DatabaseHandler
    fs, nbits = 4e6, 8
    sc = synth.make_scenario(fs, nbits, 0.6, (7, 22), 93, 250.0)
    path = str(tmp_path / "rec.bin")
    synth.write_file(path, synth.generate_iq(sc))
    rx = StreamingReceiver(_rf(path, fs, nbits), [7, 22, 9], 4, chunk_seconds=0.13)
    ref = rx.run_all()
    db = DatabaseHandler(str(tmp_path / "out.db"), overwrite=True)
    tot = rx.run_to_database(db, wall_time=lambda: 123.0)
    sync = rx._nav.states()["sync_epoch"]
    rx.close()
    assert tot["tracking_rows"] == sum(len(e) for e in ref["epochs"])
    assert [r["satellite_id"] for r in db.fetchTable("channel")] == [7, 22]
    acq = db.fetchAcquisition()
    assert [a["code_idx"] for a in acq] == [int(ref["peaks"][ref["peaks"]["prn"] == p]["code_idx"][0]) for p in (7, 22)]
    for cid in range(2):
        rows = db.fetchTracking(cid)
        e = ref["epochs"][cid]
        assert len(rows) == len(e) > 550
        assert [r["i_prompt"] for r in rows] == e["corr"][:, 2].tolist()
        assert [r["carrier_frequency"] for r in rows] == e["carrier_freq"].tolist()
        assert [r["code_frequency_error"] for r in rows] == e["code_err"].tolist()
        ts = np.array([r["time_sample"] for r in rows])
        end = e["start"] + e["n"]
        assert (ts % 4000 == 0).all() and (ts >= end).all() and (ts - end < 4000).all()
        s = int(sync[cid])
        zero = [k for k, r in enumerate(rows) if r["cn0"] == 0.0]
        assert s > 100 and zero == list(range(s + 20, len(rows), 20))
        assert all(r["cn0"] is None for k, r in enumerate(rows) if k not in zero)       # NaN is stored as NULL
    db.close()


def test_pool_results_equal_single_lane():
    """ColdStartPool: steps in flight on several lanes give what one synchronous pipeline gives."""
    import torch
    from sydr_b200 import synth
    from sydr_b200 import _lib as L
    from sydr_b200.pipeline import ColdStartPipeline, ColdStartPool
    fs, nbits, dur = 4e6, 8, 0.25
    recs = []
    for seed in (101, 102, 103):
        sc = synth.make_scenario(fs, nbits, dur, (3, 7, 19, 22)[: 2 + seed % 3], seed, 250.0)
        recs.append(torch.from_numpy(synth.generate_iq(sc)).pin_memory())
    kw = dict(fs=fs, nbits=nbits, search_prns=list(range(1, 33)), n_channels=6, max_seconds=dur)
    single = ColdStartPipeline(**kw)
    want = []
    for h in recs:
        o = single.process_host(h, copy=True)
        want.append(o)
    single.close()
    pool = ColdStartPool(lanes=2, **kw)
    tickets, got = [], []
    order = [0, 1, 2, 1, 0, 2, 2]
    for k in order:
        if len(tickets) == 2:
            got.append(pool.result(tickets.pop(0), copy=True))
        tickets.append(pool.submit_host(recs[k]))
    import pytest as _pt
    with _pt.raises(L.SydrError):
        pool.submit_host(recs[0])                       # both lanes busy
    while tickets:
        got.append(pool.result(tickets.pop(0), copy=True))
    for k, g in zip(order, got):
        w = want[k]
        assert np.array_equal(g["peaks"], w["peaks"]) and g["channels"] == w["channels"]
        assert len(g["epochs"]) == len(w["epochs"]) >= 2
        for a, b in zip(g["epochs"], w["epochs"]):
            assert a.tobytes() == b.tobytes()
    # device-resident submissions
    d = pool.lanes[0].upload(recs[1])
    t1 = pool.submit_device(d)
    t2 = pool.submit_device(d)
    r1, r2 = pool.result(t1, copy=True), pool.result(t2, copy=True)
    for a, b, c in zip(r1["epochs"], r2["epochs"], want[1]["epochs"]):
        assert a.tobytes() == b.tobytes() == c.tobytes()
    pool.close()


def test_receiver_from_ini(tmp_path):
    """main.py's object: ReceiverGPSL1CA built from a receiver.ini, the reference's per-millisecond loop
    over the batched ChannelManager into the SQLite sink; run_fast() gives the same tracking rows."""
    import configparser
    import os
    from sydr_b200 import synth
    from sydr_b200.receiver.receiver_gps_l1ca import ReceiverGPSL1CA
    fs, nbits, ms = 4e6, 8, 300
    sc = synth.make_scenario(fs, nbits, ms * 1e-3 + 0.13, (3, 7), 95, 250.0)      # 120 ms reader chunks: keep a margin
    path = str(tmp_path / "rec.bin")
    synth.write_file(path, synth.generate_iq(sc))

    def config(name):
        cfg = configparser.ConfigParser()
        cfg.read(os.path.join(helpers.ROOT, "config", "receiver.ini"))
        cfg["DEFAULT"].update({"name": name, "ms_to_process": str(ms), "outfolder": str(tmp_path)})
        cfg["RFSIGNAL"].update({"filepath": path, "sampling_frequency": str(fs), "data_size": str(nbits)})
        cfg["SATELLITES"]["include_prn"] = "3,7"
        cfg["CHANNELS"]["gps_l1ca"] = os.path.join(helpers.ROOT, "config", "channels", "channel_GPS_L1CA_borre.ini")
        return cfg

    a = ReceiverGPSL1CA(config("tick"), overwrite=True)
    a.run()
    rows_a = {c: a.database.fetchTracking(c) for c in (0, 1)}
    chan = a.database.fetchTable("channel")
    acq = a.database.fetchAcquisition()
    assert [r["satellite_id"] for r in chan] == [3, 7] and [r["system"] for r in chan] == ["GPS", "GPS"]
    assert len(acq) == 2 and all(r["peak_ratio"] > 1.5 for r in acq)
    assert a.samplesCounter == ms * 4000
    a.close()
    b = ReceiverGPSL1CA(config("fast"), overwrite=True)
    b.run_fast(chunk_seconds=0.1)
    rows_b = {c: b.database.fetchTracking(c) for c in (0, 1)}
    acq_b = b.database.fetchAcquisition()
    b.close()
    assert [r["code_idx"] for r in acq] == [r["code_idx"] for r in acq_b]
    for c in (0, 1):
        n = len(rows_a[c])
        assert n >= ms - 12 and len(rows_b[c]) >= n
        for key in ("i_prompt", "q_late", "carrier_frequency", "code_frequency", "dll", "pll"):
            assert [r[key] for r in rows_a[c]] == [r[key] for r in rows_b[c][:n]], key
        assert [r["time_sample"] for r in rows_a[c]] == [r["time_sample"] for r in rows_b[c][:n]]
    assert os.path.exists(tmp_path / "tick.db") and os.path.exists(tmp_path / "fast.db")
