"""K-TRK parity: open-loop correlators against the reference's EPL outputs (1e-4 of the prompt
magnitude), teacher-forced epochs of the reference channel, and closed-loop trajectories
(carrier / code frequency within 0.5 Hz, absolute code phase within 1e-3 chip)."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

TOL_CORR = 1e-4


@pytest.fixture(autouse=True, params=[0, 1], ids=["auto", "chunks"])
def trk_mode(request):
    """Every test runs with the automatic path selection (half-chip segments where they apply)
    and with the chunk formulations only."""
    from sydr_b200 import _lib as L
    L.check(L.load().sydr_trk_set_mode(request.param))
    yield request.param
    L.check(L.load().sydr_trk_set_mode(0))


def corr_err(got, ref):
    ref = np.asarray(ref, dtype=np.float64)
    scale = np.hypot(ref[..., 2], ref[..., 3])
    return np.abs(np.asarray(got) - ref).max(axis=-1) / scale


def test_epl_known_answers(golden):
    from sydr_b200 import _lib as L
    from sydr_b200.engine import epl_batch, to_device_iq
    g = golden("epl.npz")
    for fs, nbits, seed in H.EPL_SETS:
        iq = H.epl_input(fs, nbits, seed)
        cases = g[f"fs{int(fs)}"]
        args = np.zeros(len(cases), dtype=L.EPL_ARGS_DTYPE)
        args["prn"], args["start"], args["n"] = cases[:, 0], cases[:, 1], cases[:, 2]
        args["carrier_freq"], args["rem_carrier"] = cases[:, 3], cases[:, 4]
        args["rem_code"], args["code_step"], args["spacing"] = cases[:, 5], cases[:, 6], cases[:, 7:10]
        out = epl_batch(to_device_iq(iq), fs, args)
        e = corr_err(out, cases[:, 10:16])
        assert e.max() <= TOL_CORR, (fs, e)
        # same samples as complex64 / complex128 input (the dtype the Python functions receive)
        from sydr_b200 import synth
        x = synth.to_complex(iq)
        for arr in (x.astype(np.complex64), x):
            out2 = epl_batch(to_device_iq(arr), fs, args)
            assert corr_err(out2, cases[:, 10:16]).max() <= TOL_CORR


def test_epl_chip_boundaries_on_sample_instants():
    """Code steps for which chip (and half-chip) boundaries fall exactly on sample instants: the
    chip of such a sample is decided by the rounding of the reference expression
    ceil(fl(fl(i*step')+start)) (tracking.py:111-112), which the kernel must reproduce sample for
    sample (one misplaced sample is ~1e-3 of the prompt magnitude)."""
    from oracle import sydr_oracle as O
    from sydr_b200 import _lib as L, synth
    from sydr_b200.engine import epl_batch, to_device_iq
    fs = 25e6
    sc = synth.make_scenario(fs, 16, 0.0035, (3, 7), 77, 250.0)
    iq = synth.generate_iq(sc)
    x = synth.to_complex(iq)
    cases = []
    for step in (0.04, 0.0390625, 1.0 / 24.5, 0.04 * (1 + 2 ** -40)):
        for rem in (0.0, 0.5, 0.02, 0.25, 0.04, 1e-12, 0.5 - 1e-13):
            for start in (0, 3, 1001):
                cases.append((3, start, 25000, 1234.5, 0.3, rem, step))
    args = np.zeros(len(cases), dtype=L.EPL_ARGS_DTYPE)
    for k, c in enumerate(cases):
        args[k]["prn"], args[k]["start"], args[k]["n"], args[k]["carrier_freq"] = c[0], c[1], c[2], c[3]
        args[k]["rem_carrier"], args[k]["rem_code"], args[k]["code_step"] = c[4], c[5], c[6]
        args[k]["spacing"] = (-0.5, 0.0, 0.5)
    out = epl_batch(to_device_iq(iq), fs, args)
    code = O.padded_code(3)
    for k, c in enumerate(cases):
        ref = O.epl(x[None, c[1]:c[1] + c[2]], code, fs, c[3], c[4], c[5], c[6], [-0.5, 0.0, 0.5])
        e = corr_err(out[k], np.array(ref))
        assert e <= 2e-5, (c, e)          # well below one misplaced sample (~1e-3)


def test_epl_function_on_reference_fixture(golden):
    """EPL() drop-in on the reference's own unit-test recording (PRN 2, 3700 Hz)."""
    from oracle import sydr_oracle as O
    from sydr_b200.dsp.tracking import EPL
    g = golden("epl.npz")
    u = g["unit_iq"].astype(np.float64)
    rf = (u[0::2] + 1j * u[1::2])[None, :]
    out = EPL(rf, O.padded_code(2), 10e6, 3700.0, 0.0, 0.0, 1.023e6 / 10e6, [-0.5, 0.0, 0.5])
    assert isinstance(out, list) and len(out) == 6 and all(isinstance(v, float) for v in out)
    assert corr_err(np.array(out), g["unit_epl"]) <= TOL_CORR


def reference_epoch_inputs(trk, acq, fs, n_code):
    """Pre-epoch NCO state of every reference epoch (teacher forcing)."""
    from oracle import sydr_oracle as O
    carrier, _, cur = O.acquisition_handoff(int(acq[1]), int(acq[2]), 0.0, 5000.0, 250.0, 0, 10 * n_code,
                                            int(np.ceil(1023 / (1.023e6 / fs))))
    n_ep = len(trk)
    start = np.zeros(n_ep, dtype=np.int64)
    n = np.zeros(n_ep, dtype=np.int64)
    fc = np.zeros(n_ep); remc = np.zeros(n_ep); remcode = np.zeros(n_ep); step = np.zeros(n_ep)
    s_cur, s_n, s_fc, s_remc, s_remcode, s_step = cur, int(np.ceil(1023 / (1.023e6 / fs))), carrier, 0.0, 0.0, 1.023e6 / fs
    for k in range(n_ep):
        start[k], n[k], fc[k], remc[k], remcode[k], step[k] = s_cur, s_n, s_fc, s_remc, s_remcode, s_step
        s_cur += s_n
        s_fc, s_remcode, s_remc, s_n = trk[k, 9], trk[k, 14], trk[k, 15], int(trk[k, 16])
        s_step = trk[k, 10] / fs
    return start, n, fc, remc, remcode, step


@pytest.mark.parametrize("name", ["fs4", "fs25"])
def test_teacher_forced_epochs(golden, name):
    from sydr_b200 import _lib as L
    from sydr_b200.engine import epl_batch, to_device_iq
    g = golden("loop.npz")
    meta, prns = g[f"{name}_meta"], g[f"{name}_prns"]
    sc, iq = H.loop_input(meta, prns)
    fs = float(meta[0])
    d_iq = to_device_iq(iq)
    for prn in prns:
        trk, acq = g[f"{name}_trk_{int(prn)}"], g[f"{name}_acq_{int(prn)}"]
        start, n, fc, remc, remcode, step = reference_epoch_inputs(trk, acq, fs, round(fs * 1e-3))
        args = np.zeros(len(trk), dtype=L.EPL_ARGS_DTYPE)
        args["prn"], args["start"], args["n"] = int(prn), start, n
        args["carrier_freq"], args["rem_carrier"], args["rem_code"], args["code_step"] = fc, remc, remcode, step
        args["spacing"] = (-0.5, 0.0, 0.5)
        out = epl_batch(d_iq, fs, args)
        e = corr_err(out, trk[:, 1:7])
        assert e.max() <= TOL_CORR, (name, prn, e.max(), int(e.argmax()))


CONFIGS = [dict(cluster=1, threads=0, use_tma=True), dict(cluster=2, threads=0, use_tma=True),
           dict(cluster=4, threads=128, use_tma=True), dict(cluster=8, threads=0, use_tma=True),
           dict(cluster=1, threads=256, use_tma=False), dict(cluster=8, threads=64, use_tma=False),
           # the DENSE instantiation bench.py's timed region launches (ColdStartPool, several steps in flight)
           dict(cluster=0, threads=0, use_tma=True, dense=True), dict(cluster=8, threads=160, use_tma=True, dense=True),
           # the PACK instantiation (cfg.dense = 2: one CTA per channel, one staged window, two CTAs per SM) of ColdStartBatch
           dict(cluster=1, threads=0, use_tma=True, dense=2), dict(cluster=1, threads=128, use_tma=True, dense=2),
           # the prefix-moment kernel (trkm.cu; int16 IQ only: at fs4 / int8 it must hand every channel to the general kernel)
           dict(cluster=0, threads=0, use_tma=True, kernel=1, group=1), dict(cluster=0, threads=0, use_tma=True, kernel=1, group=2)]


@pytest.mark.parametrize("name", ["fs4", "fs25"])
@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: f"S{c['cluster']}T{c['threads']}{'tma' if c['use_tma'] else 'ldg'}{('pack' if c.get('dense') == 2 else 'dense') if c.get('dense') else ''}{('m%d' % c['group']) if c.get('kernel') == 1 else ''}")
def test_closed_loop_vs_reference_channel(golden, name, cfg, trk_mode):
    from oracle import sydr_oracle as O
    if cfg.get("kernel") == 1 and (name != "fs25" or trk_mode != 0):
        pytest.skip("the prefix-moment kernel serves int16 IQ in the automatic mode")
    from sydr_b200.engine import TrackingEngine, make_trk_states, to_device_iq
    g = golden("loop.npz")
    meta, prns = g[f"{name}_meta"], g[f"{name}_prns"]
    sc, iq = H.loop_input(meta, prns)
    fs = float(meta[0])
    n_code = round(fs * 1e-3)
    chans, refs = [], []
    for prn in prns:
        acq = g[f"{name}_acq_{int(prn)}"]
        carrier, _, cur = O.acquisition_handoff(int(acq[1]), int(acq[2]), 0.0, 5000.0, float(meta[4]), 0, 10 * n_code,
                                                int(np.ceil(1023 / (1.023e6 / fs))))
        chans.append(dict(prn=int(prn), carrier_freq=carrier, start_sample=cur, iq_len=len(iq) // 2))
        refs.append(g[f"{name}_trk_{int(prn)}"])
    st = make_trk_states(fs, chans)
    eng = TrackingEngine(fs, st, max_epochs=max(len(r) for r in refs) + 8, **cfg)
    res = eng.run(to_device_iq(iq))
    for r, ref in zip(res, refs):
        n_ep = len(ref)
        assert len(r) >= n_ep            # the reference stops at its last 1 ms tick, we run to the end of the data
        r = r[:n_ep]
        # loop outputs
        assert np.abs(r["carrier_freq"] - ref[:, 9]).max() <= 0.5
        assert np.abs(r["code_freq"] - ref[:, 10]).max() <= 0.5
        # absolute code phase of the *next* epoch: start - remCode/codeStep, in chips
        step = ref[:, 10] / fs
        n_ref = np.r_[r["n"][0], ref[:-1, 16]]                       # reference epoch lengths
        ref_end = r["start"][0] + np.cumsum(n_ref)                   # reference epoch end samples
        ref_abs = ref_end - ref[:, 14] / step
        ours_abs = (r["start"] + r["n"]) - r["rem_code"] / (r["code_freq"] / fs)
        assert np.abs(ours_abs - ref_abs).max() * (1.023e6 / fs) <= 1e-3, np.abs(ours_abs - ref_abs).max()
        # correlators (closed loop, so compared loosely: 1 % of prompt magnitude outside arctan wraps)
        e = corr_err(r["corr"], ref[:, 1:7])
        assert np.median(e) <= 1e-3


def test_throughput_kernel_hands_over_to_general(golden):
    """Quarter-chip spacing is off the half-chip lattice: the throughput instantiation stops in front
    of the first epoch (status kNeedGeneral) and the general kernel queued behind it tracks the
    channel; the result must equal a launch that used the general kernel from the start."""
    from oracle import sydr_oracle as O
    from sydr_b200.engine import TrackingEngine, make_trk_states, to_device_iq
    g = golden("loop.npz")
    meta, prns = g["fs25_meta"], g["fs25_prns"]
    sc, iq = H.loop_input(meta, prns)
    fs = float(meta[0])
    chans = []
    for prn in prns:
        acq = g[f"fs25_acq_{int(prn)}"]
        carrier, _, cur = O.acquisition_handoff(int(acq[1]), int(acq[2]), 0.0, 5000.0, float(meta[4]), 0, 250000, 25000)
        chans.append(dict(prn=int(prn), carrier_freq=carrier, start_sample=cur, iq_len=len(iq) // 2))
    cfg = dict(correlator_early=-0.25, correlator_late=0.25)
    d_iq = to_device_iq(iq)
    ref = TrackingEngine(fs, make_trk_states(fs, chans, cfg), max_epochs=300, cluster=1, threads=0, use_tma=True)
    a = ref.run(d_iq)
    lean = TrackingEngine(fs, make_trk_states(fs, chans, cfg), max_epochs=300, cluster=1, threads=256, use_tma=False)
    b = lean.run(d_iq)
    assert (lean.states()["status"] == 0).all()
    for ra, rb in zip(a, b):
        assert len(ra) == len(rb) and len(ra) > 100
        assert np.array_equal(ra["start"], rb["start"]) and np.array_equal(ra["n"], rb["n"])
        assert corr_err(rb["corr"], ra["corr"]).max() <= 1e-5
        assert np.abs(ra["carrier_freq"] - rb["carrier_freq"]).max() <= 1e-3


def test_bit_identical_across_launch_splits(golden):
    """Stopping after k epochs and resuming gives the same trajectory as one launch (state round trip)."""
    from oracle import sydr_oracle as O
    from sydr_b200.engine import TrackingEngine, make_trk_states, to_device_iq
    g = golden("loop.npz")
    meta, prns = g["fs4_meta"], g["fs4_prns"]
    sc, iq = H.loop_input(meta, prns)
    fs = float(meta[0])
    chans = []
    for prn in prns:
        acq = g[f"fs4_acq_{int(prn)}"]
        carrier, _, cur = O.acquisition_handoff(int(acq[1]), int(acq[2]), 0.0, 5000.0, 250.0, 0, 40000, 4000)
        chans.append(dict(prn=int(prn), carrier_freq=carrier, start_sample=cur, iq_len=len(iq) // 2))
    d_iq = to_device_iq(iq)
    one = TrackingEngine(fs, make_trk_states(fs, chans), max_epochs=700, cluster=2).run(d_iq)
    eng = TrackingEngine(fs, make_trk_states(fs, chans), max_epochs=100, cluster=2)
    pieces = [[] for _ in chans]
    for _ in range(8):
        for c, r in enumerate(eng.run(d_iq)):
            pieces[c].append(r)
    for c in range(len(chans)):
        joined = np.concatenate(pieces[c])
        assert joined.tobytes() == one[c][:len(joined)].tobytes() and len(joined) == len(one[c])


def test_streaming_append_matches_single_launch(golden):
    """Tracking a recording that arrives in pieces (iq_len raised launch by launch, records appended)
    is bit-identical to one launch over the whole recording."""
    from oracle import sydr_oracle as O
    from sydr_b200.engine import TrackingEngine, make_trk_states, to_device_iq
    g = golden("loop.npz")
    meta, prns = g["fs25_meta"], g["fs25_prns"]
    sc, iq = H.loop_input(meta, prns)
    fs = float(meta[0])
    n = len(iq) // 2
    acq = g[f"fs25_acq_{int(prns[0])}"]
    carrier, _, cur = O.acquisition_handoff(int(acq[1]), int(acq[2]), 0.0, 5000.0, 250.0, 0, 250000, 25000)
    chans = [dict(prn=int(prns[0]), carrier_freq=carrier, start_sample=cur, iq_len=n)]
    d_iq = to_device_iq(iq)
    one = TrackingEngine(fs, make_trk_states(fs, chans), max_epochs=300, cluster=8).run(d_iq)[0]
    eng = TrackingEngine(fs, make_trk_states(fs, chans), max_epochs=300, cluster=8)
    for frac in (0.2, 0.21, 0.5, 0.77, 1.0):
        eng.launch(d_iq, iq_len=int(n * frac), append=True)
    parts = eng.fetch()[0]
    assert len(parts) == len(one) and parts.tobytes() == one.tobytes()


def test_dense_instantiation_matches_latency_instantiation_at_the_headline_rate():
    """The register-lean DENSE instantiation (ColdStartPool) and the latency instantiation walk the same
    trajectory (int16, 25 MS/s, half-chip segment path, clusters of 8): identical epoch boundaries; the
    correlator sums differ only by the FP32 partial-sum order of the two thread counts."""
    from sydr_b200 import synth
    from sydr_b200.engine import AcquisitionEngine, TrackingEngine, make_trk_states, to_device_iq
    fs = 25e6
    sc = synth.make_scenario(fs, 16, 0.12, (8, 17, 30), 61, 250.0)
    d = to_device_iq(synth.generate_iq(sc))
    acq = AcquisitionEngine(fs, 0.0, 5000, 250, 1, 10, [8, 17, 30])
    peaks = acq.run(d)["peaks"]
    chans = [dict(prn=int(p["prn"]), carrier_freq=acq.handoff(p)[0], start_sample=acq.handoff(p)[2], iq_len=d.numel() // 2)
             for p in peaks]
    acq.close()
    out = []
    for dense in (False, True):
        eng = TrackingEngine(fs, make_trk_states(fs, chans), 130, dense=dense)
        out.append(eng.run(d))
    for a, b in zip(*out):
        assert len(a) == len(b) >= 105
        assert np.array_equal(a["start"], b["start"]) and np.array_equal(a["n"], b["n"])
        assert corr_err(b["corr"], a["corr"]).max() <= 1e-5
        assert np.abs(a["carrier_freq"] - b["carrier_freq"]).max() <= 1e-3
        assert np.abs(a["code_freq"] - b["code_freq"]).max() <= 1e-3
