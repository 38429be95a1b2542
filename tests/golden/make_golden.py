#!/usr/bin/env python
"""Generate the committed golden fixtures by running the REAL reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py            # everything (~3 min)
    python tests/golden/make_golden.py --only epl # one group

Inputs are either tiny and stored in the fixture, or regenerated at test time from
`sydr_b200.synth` with the seeds recorded here (a SHA-256 of the IQ bytes is stored to
detect generator drift).  Outputs are what the reference's own functions returned
(sydr/dsp/acquisition.py, sydr/dsp/tracking.py, sydr/signal/gnsssignal.py,
sydr/channel/channel_l1ca_borre.py), float64, unmodified.
"""
from __future__ import annotations

import argparse
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402
from sydr_b200 import synth  # noqa: E402


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def save(name, **kw):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **kw)
    print(f"  wrote {name} ({os.path.getsize(path) / 1024:.1f} KiB)")


# ----------------------------------------------------------------------------------
def g_codes(R):
    codes = np.stack([R.GenerateGPSGoldCode(p) for p in range(1, 33)]).astype(np.int8)
    spec = {}
    for fs in (4e6, 10e6, 25e6, 50e6):
        n = R.getSamplesPerCode(fs)
        sel = np.unique(np.r_[0:64, np.linspace(0, n - 1, 192).astype(int)])
        for prn in (1, 19, 32):
            up = R.UpsampleCode(R.GenerateGPSGoldCode(prn), fs)
            cf = np.conj(np.fft.fft(up))
            spec[f"sel_{int(fs)}"] = sel
            spec[f"spec_{int(fs)}_{prn}"] = cf[sel]
            spec[f"specabsmax_{int(fs)}_{prn}"] = np.abs(cf).max()
            spec[f"upsum_{int(fs)}_{prn}"] = np.array([up.sum(), (up * np.arange(n)).sum()])
    save("codes.npz", codes=codes, **spec)


# ----------------------------------------------------------------------------------
def peak_case_maps():
    """Deterministic synthetic maps for the second-peak quirks; regenerated at test time
    (float32-exact values so the fp32 GPU rows see identical numbers)."""
    rng = np.random.default_rng(77)
    n, chip, rows = 4000, 4, 5
    maps = []
    pos = [0, 1, chip - 1, chip, chip + 1, n - chip - 2, n - chip - 1, n - chip, n - 2, n - 1, 1234]
    for p in pos:
        for runner in ("last", "near_lo", "near_hi", "other_row", "random"):
            m = rng.uniform(0.0, 1.0, (rows, n)).astype(np.float32)
            r = int(rng.integers(0, rows))
            m[r, p] = 10.0
            if runner == "last":
                m[r, n - 1] = 5.0          # never searched in branches 1 and 3
            elif runner == "near_lo":
                m[r, max(0, p - chip)] = 5.0
            elif runner == "near_hi":
                m[r, min(n - 1, p + chip)] = 5.0
            elif runner == "other_row":
                m[(r + 1) % rows, (p + 100) % n] = 9.0
            maps.append(m)
    tie = np.zeros((3, 64), dtype=np.float32)     # equal maxima -> first in C order
    tie[2, 5] = tie[1, 40] = tie[1, 50] = 7.0
    tie[1, 10] = 3.0
    return np.stack(maps), n, chip, tie


def g_peaks(R):
    """TwoCorrelationPeakComparison quirks (acquisition.py:103-111) on synthetic maps."""
    maps, n, chip, tie = peak_case_maps()
    out_idx, out_ratio = [], []
    for m in maps:
        idx, ratio = R.TwoCorrelationPeakComparison(m.astype(np.float64), n, chip)
        out_idx.append(idx)
        out_ratio.append(ratio)
    tidx, tratio = R.TwoCorrelationPeakComparison(tie.astype(np.float64), 64, 2)
    save("peaks.npz", idx=np.array(out_idx), ratio=np.array(out_ratio), maps_sha=sha(maps),
         n=n, chip=chip, tie_idx=np.array(tidx), tie_ratio=tratio)


# ----------------------------------------------------------------------------------
ACQ_CASES = {
    # name: (fs, nbits, prns_present, seed, doppler_range, doppler_step, coh, noncoh, search_prns, keep_map_prn)
    "mini4": (4e6, 8, (3, 7), 42, 5000, 250, 2, 3, (3, 7, 9), 3),
    "cfg1": (4e6, 8, synth.PRNS_8, 1001, 5000, 100, 5, 10, synth.PRNS_8, None),
    "cfg2": (10e6, 8, synth.PRNS_8, 1002, 5000, 250, 1, 10, tuple(range(1, 33)), 19),
    "cfg3acq": (25e6, 16, synth.PRNS_12, 1003, 5000, 250, 1, 10, tuple(range(1, 33)), None),
    "cfg4": (50e6, 8, synth.PRNS_8, 1004, 5000, 50, 1, 20, tuple(synth.PRNS_8) + (5, 9, 16, 24), None),   # the 8 present + 4 absent
}


def acq_input(case):
    fs, nbits, present, seed, dr, ds, coh, noncoh, search, keep = ACQ_CASES[case]
    n = round(fs * 1e-3)
    dur = coh * noncoh * 1e-3
    sc = synth.make_scenario(fs, nbits, dur, present, seed, float(ds))
    iq = synth.generate_iq(sc)
    return sc, iq, n


def g_acq(R, only=None):
    for case, (fs, nbits, present, seed, dr, ds, coh, noncoh, search, keep) in ACQ_CASES.items():
        if only and case not in only:
            continue
        t0 = time.time()
        sc, iq, n = acq_input(case)
        x = synth.to_complex(iq)[None, :]
        chip = round(fs / 1.023e6)
        res = []
        extra = {}
        for prn in search:
            code = R.GenerateGPSGoldCode(prn)
            cf = np.conj(np.fft.fft(R.UpsampleCode(code, fs)))
            m = R.PCPS(rfData=x, interFrequency=0.0, samplingFrequency=fs, codeFFT=cf,
                       dopplerRange=float(dr), dopplerStep=float(ds), samplesPerCode=n,
                       coherentIntegration=coh, nonCoherentIntegration=noncoh)
            idx, ratio = R.TwoCorrelationPeakComparison(m, n, chip)
            rowmax = m.max(axis=1)
            res.append((prn, idx[0], idx[1], ratio, m[idx[0], idx[1]]))
            extra[f"rowmax_{prn}"] = rowmax
            extra[f"rowarg_{prn}"] = m.argmax(axis=1)
            if keep == prn:
                extra[f"winrow_{prn}"] = m[idx[0]]
                if n <= 4000:
                    extra[f"map_{prn}"] = m.astype(np.float32)
        res = np.array(res, dtype=np.float64)
        save(f"acq_{case}.npz", result=res, iq_sha=sha(iq), present=np.array(present),
             sat_doppler=np.array([s.doppler for s in sc.sats]),
             sat_delay=np.array([s.delay_chips for s in sc.sats]), **extra)
        print(f"  acq {case}: {time.time() - t0:.1f}s")


# ----------------------------------------------------------------------------------
def g_epl(R):
    """Open-loop correlator known answers, incl. awkward NCO states."""
    out = {}
    rng = np.random.default_rng(5)
    for fs, nbits, seed in ((4e6, 8, 11), (10e6, 8, 12), (25e6, 16, 13), (50e6, 16, 14)):
        sc = synth.make_scenario(fs, nbits, 0.0045, (3, 7), seed, 250.0)
        iq = synth.generate_iq(sc)
        x = synth.to_complex(iq)
        step0 = 1.023e6 / fs
        cases = []
        for k in range(6):
            prn = (3, 7)[k % 2]
            code = R.GenerateGPSGoldCode(prn)
            code = np.r_[code[-1], code, code[0]]
            fc = float(rng.uniform(-5000, 5000)) if k != 5 else 9.548e6 * (fs > 2 * 9.548e6)
            remc = float(rng.uniform(0, 2 * np.pi))
            remcode = float(rng.uniform(0, step0)) if k else 0.0
            step = step0 * (1 + float(rng.uniform(-3e-6, 3e-6)))
            n = int(np.ceil((1023 - remcode) / step))
            start = int(rng.integers(0, len(x) - n - 1))
            sp = [-0.5, 0.0, 0.5] if k != 4 else [-0.25, 0.0, 0.25]
            r = R.EPL(x[None, start:start + n], code, fs, fc, remc, remcode, step, sp)
            cases.append([prn, start, n, fc, remc, remcode, step, sp[0], sp[1], sp[2]] + [float(v) for v in r])
        out[f"fs{int(fs)}"] = np.array(cases)
        out[f"sha{int(fs)}"] = sha(iq)
    # the reference's own fixture: 1 ms @ 10 MHz real recording, PRN 2, 3700 Hz
    # (sydr/unitTest/tracking_in_c.py:22-33).  Values are int8-exact.
    p = os.path.join(ref_import.REFERENCE_ROOT, "sydr/unitTest/data/i_rfdata.txt")
    rf = np.loadtxt(p, dtype=np.complex128)
    assert np.all(rf.real == np.round(rf.real)) and np.abs(rf.real).max() < 128
    code = R.GenerateGPSGoldCode(2)
    code = np.r_[code[-1], code, code[0]]
    r = R.EPL(rf[None, :], code, 10e6, 3700.0, 0.0, 0.0, 1.023e6 / 10e6, [-0.5, 0.0, 0.5])
    out["unit_iq"] = np.stack([rf.real, rf.imag], axis=1).astype(np.int8).reshape(-1)
    out["unit_epl"] = np.array(r)
    # replica KAT of tracking.c:243-247 evaluated through the Python twin
    t = np.arange(6) / 1e7
    rep, rem = R.generateReplica(t, 5, -1500.0, 0.0)
    out["replica"] = rep
    out["replica_rem"] = rem
    save("epl.npz", **out)


# ----------------------------------------------------------------------------------
def drive_channel(C, x, fs, prn, ms, acq_cfg, trk_cfg=None):
    """Drive the live reference ChannelL1CA in-process, one 1 ms tick at a time."""
    cfg = {"filepath": "none", "sampling_frequency": str(fs), "is_complex": "true",
           "intermediate_frequency": "0.0", "data_size": "8"}
    rf = C.RFSignal(cfg)
    spm = rf.samplesPerMs
    buf = C.CircularBuffer(int(fs * 1e-3 * 100), np.complex128)
    chcfg = {
        "ACQUISITION": acq_cfg,
        "TRACKING": trk_cfg or {
            "correlator_early": "-0.5", "correlator_prompt": "0", "correlator_late": "0.5",
            "dll_damping_ratio": "0.7", "dll_noise_bandwidth": "1.0", "dll_loop_gain": "1.0", "dll_pdi": "0.001",
            "pll_damping_ratio": "0.7", "pll_noise_bandwidth": "8.0", "pll_loop_gain": "0.25", "pll_pdi": "0.001",
            "fll_damping_ratio": "0.7", "fll_noise_bandwidth": "15.0", "fll_loop_gain": "1.5", "fll_pdi": "0.001"},
    }
    ch = C.ChannelL1CA(0, buf, None, rf, chcfg)
    ch.setSatellite(prn)
    acq = None
    trk = []
    for tick in range(ms):
        buf.shift(x[tick * spm:(tick + 1) * spm])
        res = ch._processHandler()
        upd = ch.prepareChannelUpdate()
        for r in res:
            if r["type"] == C.ChannelMessage.ACQUISITION_UPDATE:
                acq = (tick, r["frequency_idx"], r["code_idx"], r["peak_ratio"], r["carrierFrequency"],
                       ch.currentSample)
            elif r["type"] == C.ChannelMessage.TRACKING_UPDATE:
                trk.append([tick, r["i_early"], r["q_early"], r["i_prompt"], r["q_prompt"], r["i_late"],
                            r["q_late"], r["dll"], r["pll"], r["carrier_frequency"], r["code_frequency"],
                            r["carrier_frequency_error"], r["code_frequency_error"],
                            upd["unprocessed_samples"], ch.NCO_remainingCode, ch.NCO_remainingCarrier,
                            ch.track_requiredSamples, ch.currentSample])
    return acq, np.array(trk)


def g_loop(R):
    C = ref_import.load_channel()
    out = {}
    for name, fs, nbits, seed, ms, prns, ds in (("fs4", 4e6, 8, 21, 700, (3, 7), 250),
                                                ("fs25", 25e6, 16, 23, 260, (11,), 250)):
        t0 = time.time()
        sc = synth.make_scenario(fs, nbits, ms * 1e-3, prns, seed, float(ds))
        iq = synth.generate_iq(sc)
        x = synth.to_complex(iq)
        acq_cfg = {"doppler_range": "5000", "doppler_steps": str(ds), "coherent_integration": "1",
                   "non_coherent_integration": "10", "threshold": "1.5"}
        for prn in prns:
            acq, trk = drive_channel(C, x, fs, prn, ms, acq_cfg)
            out[f"{name}_acq_{prn}"] = np.array(acq, dtype=np.float64)
            out[f"{name}_trk_{prn}"] = trk
        out[f"{name}_sha"] = sha(iq)
        out[f"{name}_meta"] = np.array([fs, nbits, seed, ms, ds], dtype=np.float64)
        out[f"{name}_prns"] = np.array(prns)
        print(f"  loop {name}: {time.time() - t0:.1f}s, epochs={len(trk)}")
    save("loop.npz", **out)


def g_channel(R):
    """Per-tick packet schedule of the live reference channel (types, flags, unread samples) and
    a random-operation trace of the reference CircularBuffer."""
    C = ref_import.load_channel()
    out = {}
    fs, nbits, seed, ms, prns, ds = 4e6, 8, 21, 700, (3, 7), 250          # the "fs4" case of loop.npz
    sc = synth.make_scenario(fs, nbits, ms * 1e-3, prns, seed, float(ds))
    iq = synth.generate_iq(sc)
    x = synth.to_complex(iq)
    acq_cfg = {"doppler_range": "5000", "doppler_steps": str(ds), "coherent_integration": "1",
               "non_coherent_integration": "10", "threshold": "1.5"}
    for prn in prns:
        cfg = {"filepath": "none", "sampling_frequency": str(fs), "is_complex": "true",
               "intermediate_frequency": "0.0", "data_size": "8"}
        rf = C.RFSignal(cfg)
        spm = rf.samplesPerMs
        buf = C.CircularBuffer(int(fs * 1e-3 * 100), np.complex128)
        ch = C.ChannelL1CA(0, buf, None, rf, {"ACQUISITION": acq_cfg, "TRACKING": TRK_CFG})
        ch.setSatellite(prn)
        rows = []
        for tick in range(ms):
            buf.shift(x[tick * spm:(tick + 1) * spm])
            res = ch._processHandler()
            upd = ch.prepareChannelUpdate()
            types = [r["type"] for r in res]
            rows.append([types.count(C.ChannelMessage.ACQUISITION_UPDATE), types.count(C.ChannelMessage.TRACKING_UPDATE),
                         types.count(C.ChannelMessage.DECODING_UPDATE), upd["state"].value, int(upd["tracking_flags"]),
                         upd["time_since_tow"], upd["unprocessed_samples"], upd["code_since_tow"], ch.currentSample,
                         ch.nbPrompt, ch.navBitsCounter])
        out[f"ticks_{prn}"] = np.array(rows, dtype=np.float64)
        out[f"navbits_{prn}"] = np.array(ch.navBitsBuffer[:ch.navBitsCounter], dtype=np.int8)
    out["meta"] = np.array([fs, nbits, seed, ms, ds], dtype=np.float64)
    out["prns"] = np.array(prns)
    out["sha"] = sha(iq)
    # CircularBuffer trace: shifts of 50 into a ring of 400, slices and unread counts at random positions
    rng = np.random.default_rng(9)
    cb = C.CircularBuffer(400, np.float64)
    trace = []
    for k in range(30):
        cb.shift(np.arange(k * 50, (k + 1) * 50, dtype=np.float64))
        cur = int(rng.integers(0, 400))
        nreq = int(rng.integers(1, 120))
        sl = cb.getSlice(cur, nreq)
        trace.append([cb.idxWrite, cb.size, int(cb.full), cur, nreq, cb.getNbUnreadSamples(cur), sl.shape[1],
                      float(np.nansum(sl[:, :min(5, sl.shape[1])]))])
    out["ring_trace"] = np.array(trace, dtype=np.float64)
    save("channel.npz", **out)


def g_nav(R):
    """Bit synchronisation and navigation-bit accumulation of the live reference channel
    (channel_l1ca_borre.py:398-413, 455-491): per tracking epoch the prompt, navPromptSum,
    navPromptSumCounter, navBitsCounter and BIT_SYNC after the tick; the bits at the end.  Kept
    below 62 bits so that the preamble search never rewrites navBitsBuffer."""
    C = ref_import.load_channel()
    out = {}
    fs, nbits, seed, ms, prns, ds = 4e6, 8, 27, 1300, (7, 22), 250
    sc = synth.make_scenario(fs, nbits, ms * 1e-3, prns, seed, float(ds))
    iq = synth.generate_iq(sc)
    x = synth.to_complex(iq)
    acq_cfg = {"doppler_range": "5000", "doppler_steps": str(ds), "coherent_integration": "1",
               "non_coherent_integration": "10", "threshold": "1.5"}
    for prn in prns:
        cfg = {"filepath": "none", "sampling_frequency": str(fs), "is_complex": "true",
               "intermediate_frequency": "0.0", "data_size": "8"}
        rf = C.RFSignal(cfg)
        spm = rf.samplesPerMs
        buf = C.CircularBuffer(int(fs * 1e-3 * 100), np.complex128)
        ch = C.ChannelL1CA(0, buf, None, rf, {"ACQUISITION": acq_cfg, "TRACKING": TRK_CFG})
        ch.setSatellite(prn)
        rows = []
        for tick in range(ms):
            buf.shift(x[tick * spm:(tick + 1) * spm])
            res = ch._processHandler()
            for r in res:
                if r["type"] == C.ChannelMessage.TRACKING_UPDATE:
                    rows.append([r["i_prompt"], ch.navPromptSum, ch.navPromptSumCounter, ch.navBitsCounter,
                                 float(bool(ch.trackFlags & C.TrackingFlags.BIT_SYNC))])
        assert ch.navBitsCounter < 62
        out[f"epochs_{prn}"] = np.array(rows, dtype=np.float64)
        out[f"bits_{prn}"] = np.array(ch.navBitsBuffer[:ch.navBitsCounter], dtype=np.int8)
        print(f"  nav PRN {prn}: {len(rows)} epochs, {ch.navBitsCounter} bits")
    # the Kaplan channel's bit accumulation on the same recording (channel_l1ca_kaplan.py:725-758): BIT_SYNC comes
    # from trackingStateUpdate, the sums start with the synchronisation epoch's own prompt
    K = ref_import.load_channel_kaplan()
    for prn in prns:
        cfg = {"filepath": "none", "sampling_frequency": str(fs), "is_complex": "true",
               "intermediate_frequency": "0.0", "data_size": "8"}
        rf = K.RFSignal(cfg)
        spm = rf.samplesPerMs
        buf = K.CircularBuffer(int(fs * 1e-3 * 100), np.complex128)
        ch = K.ChannelL1CA_Kaplan(0, buf, None, rf, {"ACQUISITION": acq_cfg, "TRACKING": KAPLAN_TRK_CFG})
        ch.setSatellite(prn)
        rows = []
        for tick in range(ms):
            buf.shift(x[tick * spm:(tick + 1) * spm])
            for r in ch._processHandler():
                if r["type"] == K.ChannelMessage.TRACKING_UPDATE:
                    rows.append([r["i_prompt"], ch.navPromptSum, ch.navPromptSumCounter, ch.navBitsCounter,
                                 float(int(ch.trackFlags) & 3)])
        assert ch.navBitsCounter < 62
        out[f"kepochs_{prn}"] = np.array(rows, dtype=np.float64)
        out[f"kbits_{prn}"] = np.array(ch.navBitsBuffer[:ch.navBitsCounter], dtype=np.int8)
        print(f"  nav (Kaplan) PRN {prn}: {len(rows)} epochs, {ch.navBitsCounter} bits")
    out["meta"] = np.array([fs, nbits, seed, ms, ds], dtype=np.float64)
    out["prns"] = np.array(prns)
    out["sha"] = sha(iq)
    save("nav.npz", **out)


def db_packets(CM, trk_rows):
    """Result packets as the channel and the receiver build them (channel_l1ca_borre.py:319-325,
    432-449; receiver.py:322-396), from the tracking rows of loop.npz.  CM = ChannelMessage enum."""
    cmap = (np.arange(12, dtype=np.float64).reshape(3, 4) + 0.5)
    acq = {"cid": 0, "type": CM.ACQUISITION_UPDATE, "carrierFrequency": 1250.0, "codeOffset": 1234,
           "frequency_idx": 15, "code_idx": 1234, "correlation_map": cmap, "peak_ratio": 3.25,
           "channel_id": 0, "time": 1000.5, "time_sample": 40000}
    trk = []
    for k, r in enumerate(trk_rows):
        trk.append({"cid": 0, "type": CM.TRACKING_UPDATE,
                    "i_early": float(r[1]), "q_early": float(r[2]), "i_prompt": float(r[3]), "q_prompt": float(r[4]),
                    "i_late": float(r[5]), "q_late": float(r[6]), "dll": float(r[7]), "pll": float(r[8]), "fll": 0.0,
                    "carrier_frequency": float(r[9]), "code_frequency": float(r[10]),
                    "cn0": 0.0 if (k % 20 == 19) else float("nan"), "pll_lock": 0.0, "fll_lock": 0.0, "lock_state": 0,
                    "carrier_frequency_error": float(r[11]), "code_frequency_error": float(r[12]),
                    "channel_id": 0, "time": 1000.5 + 0.001 * k, "time_sample": 44000 + 4000 * k})
    dec = {"cid": 0, "type": CM.DECODING_UPDATE, "subframe_id": 2, "tow": 345600, "bits": "0110" * 75,
           "channel_id": 0, "time": 1007.0, "time_sample": 26000000}
    chan = {"id": 0, "physical_id": 0, "system": "GPS", "satellite_id": 3, "signal": "GPS_L1_CA",
            "start_time": 1000.0, "start_sample": 0}
    return acq, trk, dec, chan


def db_dump(db):
    """Schema and rows of every table a channel writes to; BLOBs unpickled, NaN/NULL as None."""
    out = {}
    for table in ("channel", "acquisition", "tracking", "decoding"):
        info = db.cursor.execute(f"PRAGMA table_info({table})").fetchall()
        rows = db.sqlRequest(f"SELECT * FROM {table};")
        clean = []
        for r in rows:
            clean.append({k: (np.asarray(v).tolist() if isinstance(v, (list, np.ndarray)) else v) for k, v in r.items()})
        out[table] = {"schema": [[c[1], c[2], c[5]] for c in info], "rows": clean}
    return out


def g_database(R):
    """The reference's own DatabaseHandler (sydr/io/database.py) fed with channel packets: schema and
    rows of the resulting SQLite file."""
    import json
    import tempfile
    D = ref_import.load_database()
    loop = np.load(os.path.join(HERE, "loop.npz"))
    acq, trk, dec, chan = db_packets(D.ChannelMessage, loop["fs4_trk_3"][:60])
    with tempfile.TemporaryDirectory() as tmp:
        db = D.DatabaseHandler(os.path.join(tmp, "ref.db"), overwrite=True)
        db.addData("channel", chan)
        db.addData("acquisition", acq)
        for p in trk[:25]:
            db.addData("tracking", p)
        db.commit()
        db.addData("decoding", dec)
        for p in trk[25:]:
            db.addData("tracking", p)
        db.commit()
        dump = db_dump(db)
        db.close()
    with open(os.path.join(HERE, "database.json"), "w") as f:
        json.dump(dump, f)
    print(f"  wrote database.json ({len(dump['tracking']['rows'])} tracking rows)")


KAPLAN_TRK_CFG = {  # config/channels/channel_GPS_L1CA_kaplan.ini [TRACKING]
    "correlator_epl_wide": "0.5", "correlator_epl_narrow": "0.5", "dll_threshold": "10.0", "dll_damping_ratio": "0.7",
    "dll_noise_bandwidth": "2.0", "dll_loop_gain": "1.0", "dll_pdi": "0.001", "pll_bandwidth_wide": "25.0",
    "pll_bandwidth_narrow": "15.0", "pll_threshold_wide": "0.5", "pll_threshold_narrow": "0.8",
    "fll_bandwidth_pullin": "100.0", "fll_bandwidth_wide": "50.0", "fll_bandwidth_narrow": "15.0",
    "fll_threshold_wide": "0.5", "fll_threshold_narrow": "0.8"}
KAPLAN_CASE = (4e6, 8, 31, 2600, (3, 19), 250)      # fs, bits, seed, ms, PRNs, Doppler step


def g_kaplan(R):
    """The live reference ChannelL1CA_Kaplan driven in-process: every tracking packet plus the NCO
    members after it, and the per-tick flags / state."""
    C = ref_import.load_channel_kaplan()
    fs, nbits, seed, ms, prns, ds = KAPLAN_CASE
    sc = synth.make_scenario(fs, nbits, ms * 1e-3, prns, seed, float(ds))
    iq = synth.generate_iq(sc)
    x = synth.to_complex(iq)
    acq_cfg = {"doppler_range": "5000", "doppler_steps": str(ds), "coherent_integration": "1",
               "non_coherent_integration": "10", "threshold": "1.5"}
    out = {}
    for prn in prns:
        cfg = {"filepath": "none", "sampling_frequency": str(fs), "is_complex": "true",
               "intermediate_frequency": "0.0", "data_size": "8"}
        rf = C.RFSignal(cfg)
        spm = rf.samplesPerMs
        buf = C.CircularBuffer(int(fs * 1e-3 * 100), np.complex128)
        ch = C.ChannelL1CA_Kaplan(0, buf, None, rf, {"ACQUISITION": acq_cfg, "TRACKING": KAPLAN_TRK_CFG})
        ch.setSatellite(prn)
        rows, acq = [], None
        for tick in range(ms):
            buf.shift(x[tick * spm:(tick + 1) * spm])
            for r in ch._processHandler():
                if r["type"] == C.ChannelMessage.ACQUISITION_UPDATE:
                    acq = (tick, r["frequency_idx"], r["code_idx"], r["peak_ratio"], r["carrierFrequency"], ch.currentSample)
                elif r["type"] == C.ChannelMessage.TRACKING_UPDATE:
                    rows.append([tick, r["i_early"], r["q_early"], r["i_prompt"], r["q_prompt"], r["i_late"], r["q_late"],
                                 r["dll"], r["pll"], r["fll"], r["carrier_frequency"], r["code_frequency"],
                                 r["carrier_frequency_error"], r["code_frequency_error"], r["cn0"], r["pll_lock"],
                                 r["fll_lock"], int(r["lock_state"]), int(ch.trackFlags), ch.remainingCode,
                                 ch.remainingCarrier, ch.track_requiredSamples, ch.currentSample, ch.navBitsCounter])
        out[f"acq_{prn}"] = np.array(acq, dtype=np.float64)
        out[f"trk_{prn}"] = np.array(rows, dtype=np.float64)
        st = out[f"trk_{prn}"][:, 17]
        print(f"  kaplan PRN {prn}: {len(rows)} epochs, lock states {sorted(set(st.astype(int)))}, "
              f"first WIDE {np.argmax(st == 2)}, first NARROW {np.argmax(st == 3)}, flags {int(ch.trackFlags)}")
    out["meta"] = np.array([fs, nbits, seed, ms, ds], dtype=np.float64)
    out["prns"] = np.array(prns)
    out["sha"] = sha(iq)
    save("kaplan.npz", **out)


def g_decoding(R):
    """Inputs and the reference's outputs for LNAV_CheckPreambule / LNAV_DecodeTOW / Prompt2Bit."""
    from sydr.dsp import decoding as RD
    from sydr_b200.dsp.decoding import _PARITY_TAPS
    rng = np.random.default_rng(17)

    def word(prev2, data24):
        src = list(prev2) + list(data24)
        par = []
        for taps in _PARITY_TAPS:
            b = 0
            for t in taps:
                b ^= src[t]
            par.append(b)
        return [int(b) ^ int(prev2[1]) for b in data24] + par

    wins, exp = [], []
    pre = [1, 0, 0, 0, 1, 0, 1, 1]
    for trial in range(600):
        if trial % 3 == 0:
            bits = rng.integers(0, 2, 62)
        else:
            prev = [int(v) for v in rng.integers(0, 2, 2)]
            d1 = [b ^ prev[1] for b in pre] + [int(v) for v in rng.integers(0, 2, 16)]
            if trial % 2:
                d1 = [1 - b for b in d1[:8]] + d1[8:]
            w1 = word(prev, d1)
            w2 = word(w1[-2:], [int(v) for v in rng.integers(0, 2, 24)])
            bits = np.array(prev + w1 + w2)
            if trial % 7 == 0:
                bits[int(rng.integers(10, 62))] ^= 1
        wins.append(np.array(bits, dtype=np.int64))
        exp.append(bool(RD.LNAV_CheckPreambule(np.array(bits, dtype=np.int64).copy())))
    sfs, tows = [], []
    for trial in range(60):
        sf = rng.integers(0, 2, 300).astype(np.int64)
        d = int(rng.integers(0, 2))
        tow, sid, txt = RD.LNAV_DecodeTOW(sf.copy(), d)
        sfs.append(np.r_[d, sf])
        tows.append([tow, sid] + [int(c) for c in txt])
    save("decoding.npz", windows=np.array(wins, dtype=np.int8), check=np.array(exp),
         subframes=np.array(sfs, dtype=np.int8), tow=np.array(tows, dtype=np.int32),
         p2b=np.array([RD.Prompt2Bit(v) for v in (-3.0, 0.0, 2.5)]))


def lnav_stream(seed=5, n_subframes=4, lead=37, corrupt_at=None):
    """A navigation bit stream as received: `lead` random bits, then subframes of ten parity-clean words
    (TLM with the preamble, HOW with a TOW count and subframe id, eight data words), optionally one bit
    flipped.  Polarity inverted as a Costas loop may deliver it."""
    from sydr_b200.dsp.decoding import _PARITY_TAPS
    rng = np.random.default_rng(seed)

    def word(prev2, data24):
        src = list(prev2) + list(data24)
        par = []
        for taps in _PARITY_TAPS:
            b = 0
            for t in taps:
                b ^= src[t]
            par.append(b)
        return [int(b) ^ int(prev2[1]) for b in data24] + par

    bits = [int(v) for v in rng.integers(0, 2, lead)]
    prev = bits[-2:]
    tow0 = 34567
    for k in range(n_subframes):
        tlm = [1, 0, 0, 0, 1, 0, 1, 1] + [int(v) for v in rng.integers(0, 2, 16)]
        tow = [(tow0 + k) >> (16 - i) & 1 for i in range(17)]
        how = tow + [0, 0] + [((k % 5) + 1) >> (2 - i) & 1 for i in range(3)] + [0, 0]
        words = [tlm, how] + [[int(v) for v in rng.integers(0, 2, 24)] for _ in range(8)]
        for w in words:
            enc = word(prev, w)
            bits += enc
            prev = enc[-2:]
    bits = np.array(bits, dtype=np.int64)
    if corrupt_at is not None:
        bits[corrupt_at] ^= 1
    return 1 - bits                                              # inverted polarity


def drive_decoding(ch, flags_cls, bits):
    """Feed navigation bits to a Borre-style channel's runDecoding: 20 prompts per bit, bit synchronised."""
    ch.trackFlags |= flags_cls.BIT_SYNC
    rows, found = [], []
    for k, b in enumerate(bits):
        for ms in range(20):
            ch.nbPrompt = 1
            ch.correlatorsBuffer[0, 2] = 1000.0 if b else -1000.0
            r = ch.runDecoding()
            if r is not None:
                found.append([k, int(r["subframe_id"]), int(r["tow"])] + [int(c) for c in r["bits"]])
        rows.append([ch.navBitsCounter, int(ch.trackFlags), int(ch.preambuleFound), float(ch.tow), ch.codeSinceTOW])
    return np.array(rows, dtype=np.float64), np.array(found, dtype=np.int64)


def g_framing(R):
    """Subframe synchronisation of the live reference channel (channel_l1ca_borre.py:455-573) on synthetic
    LNAV bit streams: per-bit navBitsCounter / flags / preambuleFound / tow, and the decoded subframes."""
    C = ref_import.load_channel()
    out = {}
    cases = {"clean": dict(seed=5, n_subframes=4, lead=37), "late": dict(seed=6, n_subframes=4, lead=401),
             "broken": dict(seed=7, n_subframes=5, lead=12, corrupt_at=12 + 300 * 2 + 3)}
    for name, kw in cases.items():
        bits = lnav_stream(**kw)
        cfg = {"filepath": "none", "sampling_frequency": "4e6", "is_complex": "true", "intermediate_frequency": "0.0",
               "data_size": "8"}
        rf = C.RFSignal(cfg)
        buf = C.CircularBuffer(400000, np.complex128)
        acq_cfg = {"doppler_range": "5000", "doppler_steps": "250", "coherent_integration": "1",
                   "non_coherent_integration": "10", "threshold": "1.5"}
        ch = C.ChannelL1CA(0, buf, None, rf, {"ACQUISITION": acq_cfg, "TRACKING": TRK_CFG})
        ch.setSatellite(5)
        rows, found = drive_decoding(ch, C.TrackingFlags, bits)
        out[f"rows_{name}"], out[f"found_{name}"] = rows, found
        print(f"  framing {name}: {len(bits)} bits, {len(found)} subframes decoded, final flags {int(ch.trackFlags)}")
        # the Kaplan channel's decoding on the same stream (channel_l1ca_kaplan.py:725-861)
        K = ref_import.load_channel_kaplan()
        kch = K.ChannelL1CA_Kaplan(0, K.CircularBuffer(400000, np.complex128), None, K.RFSignal(cfg),
                                   {"ACQUISITION": acq_cfg, "TRACKING": KAPLAN_TRK_CFG})
        kch.setSatellite(5)
        kch.trackFlags |= K.TrackingFlags.BIT_SYNC
        kfound, krows = [], []
        for k, b in enumerate(bits):
            for ms in range(20):
                kch.correlatorsResults[kch.IDX_I_PROMPT] = 1000.0 if b else -1000.0
                r = kch.runDecoding()
                if r is not None:
                    kfound.append([k, int(r["subframe_id"]), int(r["tow"])] + [int(c) for c in r["bits"]])
            krows.append([kch.navBitsCounter, int(kch.trackFlags), float(kch.tow), kch.codeSinceTOW])
        out[f"kfound_{name}"], out[f"krows_{name}"] = np.array(kfound, dtype=np.int64), np.array(krows, dtype=np.float64)
    save("framing.npz", **out)


TRK_CFG = {
    "correlator_early": "-0.5", "correlator_prompt": "0", "correlator_late": "0.5",
    "dll_damping_ratio": "0.7", "dll_noise_bandwidth": "1.0", "dll_loop_gain": "1.0", "dll_pdi": "0.001",
    "pll_damping_ratio": "0.7", "pll_noise_bandwidth": "8.0", "pll_loop_gain": "0.25", "pll_pdi": "0.001",
    "fll_damping_ratio": "0.7", "fll_noise_bandwidth": "15.0", "fll_loop_gain": "1.5", "fll_pdi": "0.001"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*", default=None)
    a = ap.parse_args()
    R = ref_import.load()
    groups = {"codes": g_codes, "peaks": g_peaks, "acq": g_acq, "epl": g_epl, "loop": g_loop, "channel": g_channel,
              "kaplan": g_kaplan, "nav": g_nav, "database": g_database, "framing": g_framing,
              "decoding": g_decoding}
    for name, fn in groups.items():
        if a.only and name not in a.only and not (name == "acq" and any(o in ACQ_CASES for o in a.only)):
            continue
        print(f"[{name}]")
        if name == "acq" and a.only and any(o in ACQ_CASES for o in a.only):
            fn(R, only=a.only)
        else:
            fn(R)


if __name__ == "__main__":
    main()
