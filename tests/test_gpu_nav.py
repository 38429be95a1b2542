"""K-NAV (bit synchronisation + navigation-bit accumulation on the device) against
the live reference channel's golden (tests/golden/nav.npz) and the oracle, through the C ABI."""
import numpy as np
import pytest

import helpers  # noqa: F401  (path setup)
from oracle import sydr_oracle as O

pytestmark = pytest.mark.gpu


def _records(ip_rows, max_epochs):
    """[n_ch][max_epochs] sydr_trk_epoch records whose i_prompt column is filled from `ip_rows`."""
    from sydr_b200 import _lib as L
    rec = np.zeros((len(ip_rows), max_epochs), dtype=L.TRK_EPOCH_DTYPE)
    for c, ip in enumerate(ip_rows):
        rec["corr"][c, :len(ip), 2] = ip
        rec["corr"][c, :len(ip), 3] = 1.0e9            # q_prompt must not matter
    return rec


def _run(ip_rows, cuts):
    """Feed the prompts in pieces ending at `cuts` (epoch counts); returns per channel (bits, sums), states."""
    import torch
    from sydr_b200.engine import NavBitEngine
    n_ch = len(ip_rows)
    bits = [[] for _ in range(n_ch)]
    sums = [[] for _ in range(n_ch)]
    eng = NavBitEngine(n_ch, max_bits=128)
    lo = 0
    for hi in cuts:
        piece = [ip[lo:hi] for ip in ip_rows]
        mx = max(1, max(len(p) for p in piece))
        rec = _records(piece, mx)
        d_rec = torch.from_numpy(rec.view(np.uint8).reshape(-1)).cuda()
        d_nep = torch.tensor([len(p) for p in piece], dtype=torch.int32).cuda()
        eng.launch_records(d_rec, mx, d_nep)
        for c, (b, s) in enumerate(eng.fetch()):
            bits[c].extend(b.tolist())
            sums[c].extend(s.tolist())
        lo = hi
    return bits, sums, eng.states()


@pytest.mark.parametrize("pieces", ["one", "ragged", "tiny"])
def test_nav_bits_match_reference_channel(golden, pieces):
    g = golden("nav.npz")
    prns = [int(p) for p in g["prns"]]
    ips = [g[f"epochs_{p}"][:, 0] for p in prns]
    n = max(len(ip) for ip in ips)
    if pieces == "one":
        cuts = [n]
    elif pieces == "ragged":
        cuts = sorted(set(np.random.default_rng(5).integers(1, n, 17).tolist() + [n]))
    else:
        cuts = list(range(7, n, 7)) + [n]
    bits, sums, st = _run(ips, cuts)
    for c, p in enumerate(prns):
        ep = g[f"epochs_{p}"]
        ref_bits = g[f"bits_{p}"]
        assert bits[c] == ref_bits.tolist(), p                          # bit-exact with the live channel
        # the 20-epoch sums are the reference's navPromptSum at the tick before it was reset
        done = np.nonzero(np.diff(np.r_[0, ep[:, 3]]) > 0)[0]
        o_bits, o_sums, o_sync, o = O.nav_bits(ep[:, 0])
        assert len(done) == len(o_sums) == len(sums[c])
        assert sums[c] == o_sums.tolist()                               # FP64 sums bit-identical
        assert int(st["sync_epoch"][c]) == o_sync == int(np.nonzero(ep[:, 4])[0][0])
        assert int(st["nav_count"][c]) == int(ep[-1, 2]) and st["nav_sum"][c] == ep[-1, 1]
        assert int(st["n_bits"][c]) == len(ref_bits) and int(st["code_counter"][c]) == len(ep)


def test_nav_bits_edge_cases():
    rng = np.random.default_rng(11)
    # no sign change at all / sync on the first eligible epoch / zeros (np.sign(0) = 0 differs from +-1) /
    # a change exactly at epoch 100 (too early) / empty channel
    a = np.abs(rng.normal(1e5, 1e3, 400))
    b = a.copy(); b[101:] *= -1.0
    c = a.copy(); c[150] = 0.0
    d = a.copy(); d[100] *= -1.0; d[101] *= -1.0; d[102:] *= -1.0
    e = np.zeros(0)
    f = rng.normal(0.0, 1e5, 400)
    rows = [a, b, c, d, e, f]
    bits, sums, st = _run(rows, [123, 124, 300, 400])
    for i, ip in enumerate(rows):
        o_bits, o_sums, o_sync, o = O.nav_bits(ip)
        assert bits[i] == o_bits.tolist(), i
        assert sums[i] == o_sums.tolist(), i
        assert int(st["sync_epoch"][i]) == o_sync, i
        assert int(st["nav_count"][i]) == o.nav_count and st["nav_sum"][i] == o.nav_sum, i
    assert int(st["sync_epoch"][0]) == -1 and int(st["sync_epoch"][1]) == 101 and int(st["sync_epoch"][2]) == 150


def test_nav_bits_from_gpu_tracking_recover_the_data_bits():
    """End to end on the device: acquisition -> closed-loop tracking -> K-NAV; the bits equal the
    oracle's on the same records and reproduce the data bits the generator modulated."""
    from sydr_b200 import synth
    from sydr_b200.engine import NavBitEngine, to_device_iq
    from sydr_b200.pipeline import ColdStartPipeline
    fs, prns = 4e6, (3, 7, 19)
    sc = synth.make_scenario(fs, 8, 1.2, prns, 77, 250.0)
    iq = synth.generate_iq(sc)
    pipe = ColdStartPipeline(fs, 8, list(range(1, 33)), 8, max_seconds=1.2)
    out = pipe.process_device(to_device_iq(iq))
    nav = NavBitEngine(len(out["channels"]), max_bits=128)
    nav.launch(pipe._trk)
    got = nav.fetch()
    recs = pipe.collect()
    assert sorted(c["prn"] for c in out["channels"]) == sorted(prns)
    truth = synth.nav_bits_of(sc)
    for ch, r, (b, s) in zip(out["channels"], recs, got):
        o_bits, o_sums, o_sync, _ = O.nav_bits(r["corr"][:, 2])
        assert b.tolist() == o_bits.tolist() and s.tolist() == o_sums.tolist()
        assert len(b) >= 45
        # compare with the modulated data: bit boundaries of the transmitted stream in receiver epochs
        d = truth[ch["prn"]]
        sat = [s_ for s_ in sc.sats if s_.prn == ch["prn"]][0]
        t_mid = (r["start"][o_sync + 1] + 10 * fs * 1e-3 + 20 * fs * 1e-3 * np.arange(len(b))) / fs
        idx = synth.nav_bit_index(sat, t_mid)
        tx = (d[idx] > 0).astype(np.int8)
        agree = (tx == b).mean()
        assert agree == 1.0 or agree == 0.0, (ch["prn"], agree)          # Costas loop: 180 deg ambiguity
    pipe.close()


def test_device_handoff_equals_host_handoff():
    """K-HAND: selection (threshold, best ratios first, at most n_channels, PRN order) and the hand-off
    scalars of channel_l1ca_borre.py:301-311, bit-identical with the host restatement; idle slots; no peaks."""
    import torch
    from sydr_b200 import _lib as L
    from sydr_b200.engine import make_trk_states
    from sydr_b200.pipeline import ColdStartPipeline
    fs = 4e6
    pipe = ColdStartPipeline(fs, 8, list(range(1, 33)), 5, max_seconds=0.1, doppler_step=250.0)
    rng = np.random.default_rng(4)
    for case in range(4):
        peaks = np.zeros(32, dtype=L.ACQ_PEAK_DTYPE)
        peaks["prn"] = np.arange(1, 33)
        peaks["freq_idx"] = rng.integers(0, 41, 32)
        peaks["code_idx"] = rng.integers(0, 4000, 32)
        peaks["ratio"] = rng.uniform(1.0, 1.45, 32).astype(np.float32)
        if case == 0:                                   # more candidates than channels, with a tie at the cut
            strong = rng.choice(32, 9, replace=False)
            peaks["ratio"][strong] = rng.uniform(2.0, 9.0, 9).astype(np.float32)
            peaks["ratio"][strong[:3]] = np.float32(3.25)
        elif case == 1:                                 # fewer than n_channels
            peaks["ratio"][[4, 20]] = (np.float32(1.5000001), np.float32(7.0))
            peaks["ratio"][9] = np.float32(1.5)         # exactly the threshold: not above it
        elif case == 2:                                 # nothing found
            pass
        else:                                           # freq_idx at both ends (Doppler +5000 / -5000, IF + -0.0)
            peaks["ratio"][[0, 1, 2]] = np.float32(4.0)
            peaks["freq_idx"][[0, 1, 2]] = (0, 40, 20)
        pipe.acq.peaks_device().copy_(torch.from_numpy(peaks.view(np.uint8).reshape(-1)).cuda())
        n = 400000
        pipe._handoff(n)
        torch.cuda.synchronize()
        got = pipe._trk.states()
        chans = pipe._channels_of(peaks, n)
        want = make_trk_states(fs, chans) if chans else np.zeros(0, dtype=L.TRK_STATE_DTYPE)
        assert int(pipe._n_sel.cpu()[0]) == len(chans) <= 5
        assert [int(p) for p in got["prn"][:len(chans)]] == [c["prn"] for c in chans]
        assert got[:len(chans)].tobytes() == want.tobytes(), case          # every member, bit for bit
        assert (got["status"][len(chans):] == 1).all()
        if case == 1:
            assert [c["prn"] for c in chans] == [5, 21]
        if case == 2:
            assert chans == []
    pipe.close()


def test_pipeline_with_nothing_to_track():
    """Noise only: the hand-off leaves every slot idle, the tracking launch does nothing, results are empty."""
    import torch
    from sydr_b200.pipeline import ColdStartPipeline
    fs = 4e6
    rng = np.random.default_rng(8)
    iq = np.clip(np.round(rng.normal(0, 16, 2 * 200000)), -127, 127).astype(np.int8)
    pipe = ColdStartPipeline(fs, 8, [1, 2, 3, 4], 4, max_seconds=0.05, threshold=2.5)
    host = torch.from_numpy(iq).pin_memory()
    out = pipe.process_host(host)
    assert out["channels"] == [] and out["epochs"] == [] and len(out["peaks"]) == 4
    assert (pipe._trk.states()["status"] == 1).all()
    pipe.close()


@pytest.mark.parametrize("cuts", [None, (7, 333, 334, 900)])
def test_nav_bits_kaplan_rule(golden, cuts):
    """K-NAV with the Kaplan channel's rule on the live reference Kaplan channel's prompts and flags
    (tests/golden/nav.npz): bits bit-exact, sums and pending state equal to the oracle's, any chunking."""
    import torch
    from sydr_b200 import _lib as L
    from sydr_b200.engine import NavBitEngine
    g = golden("nav.npz")
    prns = [int(p) for p in g["prns"]]
    eps = [g[f"kepochs_{p}"] for p in prns]
    n = max(len(e) for e in eps)
    bounds = [n] if cuts is None else list(cuts) + [n]
    eng = NavBitEngine(len(prns), max_bits=128)
    bits = [[] for _ in prns]
    sums = [[] for _ in prns]
    lo = 0
    for hi in bounds:
        m = hi - lo
        rec = np.zeros((len(prns), m), dtype=L.TRK_EPOCH_DTYPE)
        krec = np.zeros((len(prns), m), dtype=L.KAPLAN_EPOCH_DTYPE)
        for c, e in enumerate(eps):
            rec["corr"][c, :, 2] = e[lo:hi, 0]
            krec["flags"][c] = e[lo:hi, 4].astype(np.int32)
        d_rec = torch.from_numpy(rec.view(np.uint8).reshape(-1)).cuda()
        d_k = torch.from_numpy(krec.view(np.uint8).reshape(-1)).cuda()
        d_nep = torch.full((len(prns),), m, dtype=torch.int32).cuda()
        L.check(L.load().sydr_nav_bits_kaplan(d_rec.data_ptr(), d_k.data_ptr(), m, d_nep.data_ptr(), 0,
                                              eng._state.data_ptr(), len(prns), eng._bits.data_ptr(), eng._sums.data_ptr(),
                                              eng.max_bits, eng._nbits.data_ptr(), 0))
        for c, (b, s) in enumerate(eng.fetch()):
            bits[c].extend(b.tolist())
            sums[c].extend(s.tolist())
        lo = hi
    st = eng.states()
    for c, p in enumerate(prns):
        e = eps[c]
        ob, osum, osync, (pend, cnt) = O.nav_bits_kaplan(e[:, 0], e[:, 4] >= 2)
        assert bits[c] == g[f"kbits_{p}"].tolist() == ob.tolist()
        assert sums[c] == osum.tolist()
        assert int(st["sync_epoch"][c]) == osync and int(st["nav_count"][c]) == cnt and st["nav_sum"][c] == pend
