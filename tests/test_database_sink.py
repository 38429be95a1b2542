"""SQLite result sink against the reference's own DatabaseHandler
(tests/golden/database.json, written by sydr/io/database.py itself): same schema, same rows,
through the packet interface and through the columnar record path."""
import json
import os
import time

import numpy as np

import helpers as H
import make_golden as MG
from sydr_b200 import _lib as L
from sydr_b200.io.database import TRACKING_KEYS, DatabaseHandler, cn0_column
from sydr_b200.utils.enumerations import ChannelMessage


def _golden():
    return json.load(open(os.path.join(H.ROOT, "tests", "golden", "database.json")))


def _packets(golden):
    loop = golden("loop.npz")
    return MG.db_packets(ChannelMessage, loop["fs4_trk_3"][:60]), loop["fs4_trk_3"][:60]


def test_packet_interface_writes_the_reference_database(golden, tmp_path):
    (acq, trk, dec, chan), _ = _packets(golden)
    db = DatabaseHandler(str(tmp_path / "a.db"), overwrite=True)
    db.addData("channel", chan)
    db.addData("acquisition", acq)
    for p in trk[:25]:
        db.addData("tracking", p)
    db.commit()
    db.addData("decoding", dec)
    for p in trk[25:]:
        db.addData("tracking", p)
    db.commit()
    assert MG.db_dump(db) == _golden()
    assert db.fetchTracking(0)[19]["cn0"] == 0.0 and len(db.fetchAcquisition()) == 1
    assert np.array_equal(db.fetchAcquisition(0)[0]["correlation_map"], acq["correlation_map"])
    db.close()


def test_columnar_records_write_the_same_rows(golden, tmp_path):
    """The device's per-epoch records, inserted column-wise, give the rows the packets give."""
    (acq, trk, dec, chan), rows = _packets(golden)
    rec = np.zeros(len(rows), dtype=L.TRK_EPOCH_DTYPE)
    rec["corr"] = rows[:, 1:7]
    rec["dll"], rec["pll"], rec["carrier_freq"], rec["code_freq"] = rows[:, 7], rows[:, 8], rows[:, 9], rows[:, 10]
    rec["carrier_err"], rec["code_err"] = rows[:, 11], rows[:, 12]
    db = DatabaseHandler(str(tmp_path / "b.db"), overwrite=True)
    db.addData("channel", chan)
    db.addData("acquisition", acq)
    k = np.arange(len(rows))
    cn0 = cn0_column(0, len(rows), -1)
    cn0[k % 20 == 19] = 0.0
    for lo, hi in ((0, 25), (25, 60)):
        db.addTrackingRecords(0, rec[lo:hi], time=1000.5 + 0.001 * k[lo:hi], time_sample=44000 + 4000 * k[lo:hi],
                              cn0=cn0[lo:hi])
        if lo == 0:
            db.commit()
            db.addData("decoding", dec)
    db.commit()
    assert MG.db_dump(db) == _golden()
    assert tuple(db.columns["tracking"][4:]) == TRACKING_KEYS[:-3]
    db.close()


def test_cn0_column_follows_the_bit_sync_rule():
    c = cn0_column(100, 100, 131)
    k = np.arange(100, 200)
    assert np.isnan(c[k <= 131]).all()
    assert (np.nonzero(c == 0.0)[0] + 100).tolist() == [151, 171, 191]
    assert np.isnan(cn0_column(0, 50, -1)).all()


def test_batched_commit_is_much_faster_than_one_statement_per_row(tmp_path):
    """720 000 rows per minute of signal: the sink must not be the bottleneck."""
    n = 20000
    rec = np.zeros(n, dtype=L.TRK_EPOCH_DTYPE)
    rec["corr"] = np.random.default_rng(1).normal(size=(n, 6))
    db = DatabaseHandler(str(tmp_path / "c.db"), overwrite=True)
    t0 = time.perf_counter()
    db.addTrackingRecords(3, rec, time=0.0, time_sample=np.arange(n) * 25000)
    db.commit()
    dt = time.perf_counter() - t0
    assert len(db.fetchTable("tracking")) == n
    assert n / dt > 100_000, f"{n / dt:.0f} rows/s"
    db.close()


def test_fast_path_writes_the_decoding_rows(golden, tmp_path):
    """run_to_database (ReceiverGPSL1CA.run_fast / main.py --fast) emits the reference's DECODING_UPDATE rows: the
    navigation bits of every chunk go through the same preamble search / subframe synchronisation as the channel class
    (channel_l1ca_borre.py:455-573).  Driven here without a GPU: a stand-in for the streaming loop delivers synthetic
    parity-clean LNAV bits in uneven chunks; the rows must be the subframes the live reference channel decoded from the
    same bit stream (tests/golden/framing.npz), stamped with the tick of the epoch that completed the subframe."""
    import pickle
    import sqlite3
    from types import SimpleNamespace
    import make_golden as MG
    from sydr_b200 import _lib as L
    from sydr_b200.ingest import StreamingReceiver
    from sydr_b200.io.database import DatabaseHandler
    bits = np.asarray(MG.lnav_stream(seed=5, n_subframes=4, lead=37), dtype=np.int8)
    found = golden("framing.npz")["found_clean"]                  # [bit index, subframe id, tow, 300 bits]
    fs, spm, sync, n_ep0 = 4e6, 4000, 137, 137 + 20 * len(bits)
    start0 = 36001

    def epochs(lo, hi):                                           # epochs lo .. hi-1 of a channel with 4000-sample epochs
        rec = np.zeros(hi - lo, dtype=L.TRK_EPOCH_DTYPE)
        rec["start"] = start0 + spm * np.arange(lo, hi)
        rec["n"] = spm
        return rec

    rx = object.__new__(StreamingReceiver)
    rx.fs, rx.want_records, rx.want_bits = fs, True, True
    rx.acq = SimpleNamespace(required_samples=40000)
    rx._nav = SimpleNamespace(states=lambda: {"sync_epoch": np.array([sync])})
    peaks = np.zeros(1, dtype=L.ACQ_PEAK_DTYPE)
    peaks["prn"], peaks["ratio"] = 5, 3.0

    def fake_run(skip_samples=0, max_samples=None):
        yield dict(peaks=peaks, channels=[dict(prn=5, carrier_freq=1250.0, start_sample=start0)])
        cuts = [0, 1000, 1003, 9000, 15001, n_ep0]               # chunk borders in epochs, on and off bit borders
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            b_lo, b_hi = max(0, (lo - sync) // 20), max(0, (hi - sync) // 20)      # bits completed inside [lo, hi)
            yield dict(epochs=[epochs(lo, hi)], bits=[bits[b_lo:b_hi]])
    rx.run = fake_run
    db = DatabaseHandler(str(tmp_path / "fast.db"), overwrite=True)
    out = rx.run_to_database(db, wall_time=lambda: 0.0)
    db.commit()
    assert out["tracking_rows"] == n_ep0 and out["decoding_rows"] == len(found) >= 2
    con = sqlite3.connect(str(tmp_path / "fast.db"))
    rows = con.execute("select channel_id, time_sample, subframe_id, tow, bits from decoding order by id").fetchall()
    con.close()
    assert len(rows) == len(found)
    for row, ref in zip(rows, found):
        j = int(ref[0])                                           # the bit that completed the subframe
        e = sync + 20 * (j + 1) - 1
        end = start0 + spm * (e + 1)
        assert row[0] == 0 and row[1] == -(-end // spm) * spm
        assert [row[2], row[3]] == [int(ref[1]), int(ref[2])]
        stored = row[4] if isinstance(row[4], str) else pickle.loads(row[4])       # the packet's `bits` as the sink stores its type
        assert [int(c) for c in stored] == [int(c) for c in ref[3:]]
