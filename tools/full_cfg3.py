"""BASELINE.json configs[2] at full size: 60 s of 25 MS/s int16 IQ (6 GB), 12 channels, from a file
through StreamingReceiver.  Size-independent checks: every satellite acquired and tracked to the last
millisecond, Doppler equal to the generator's truth, navigation bits equal to the transmitted data
(up to the Costas half-cycle ambiguity) for the whole minute.
    python tools/full_cfg3.py [seconds] [chunk_seconds]"""
import os, shutil, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sydr_b200 import synth
from sydr_b200.ingest import StreamingReceiver
from sydr_b200.signal.rfsignal import RFSignal

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
chunk = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
fs, nbits = 25e6, 16
sc = synth.make_scenario(fs, nbits, seconds, synth.PRNS_12, 1003, 250.0)
tmp = tempfile.mkdtemp(prefix="sydr_cfg3_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
path = os.path.join(tmp, "cfg3.bin")
try:
    t0 = time.perf_counter()
    d = synth.generate_iq_torch(sc)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    with open(path, "wb") as f:
        step = 1 << 28
        for lo in range(0, d.numel(), step):
            f.write(d[lo:lo + step].cpu().numpy().tobytes())
    del d
    torch.cuda.empty_cache()
    t2 = time.perf_counter()
    print(f"generated {seconds:g} s ({os.path.getsize(path) / 1e9:.2f} GB) in {t1 - t0:.1f} s, written in {t2 - t1:.1f} s")
    rf = RFSignal({"filepath": path, "sampling_frequency": str(fs), "is_complex": "true",
                   "intermediate_frequency": "0.0", "data_size": str(nbits)})
    rx = StreamingReceiver(rf, list(range(1, 33)), 12, chunk_seconds=chunk)
    best = None
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = rx.run_all()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if it:
            best = dt if best is None else min(best, dt)
    rx.close()
    truth = synth.nav_bits_of(sc)
    ok = True
    print(f"file -> results: {best * 1e3:.1f} ms for {seconds:g} s of signal = RTF {seconds / best:.1f} "
          f"({seconds * fs / best / 1e6:.0f} Msamples/s, {os.path.getsize(path) / best / 1e9:.1f} GB/s from the file)")
    assert sorted(c["prn"] for c in out["channels"]) == sorted(synth.PRNS_12)
    for ch, e, b in zip(out["channels"], out["epochs"], out["bits"]):
        sat = [s for s in sc.sats if s.prn == ch["prn"]][0]
        df = float(np.mean(e["carrier_freq"][-500:])) - sat.doppler
        end = (e["start"][-1] + e["n"][-1]) / fs
        sync = int(np.nonzero(np.abs(np.diff(np.sign(e["corr"][:, 2]))) > 0)[0][np.nonzero(np.abs(np.diff(np.sign(e["corr"][:, 2]))) > 0)[0] >= 100][0] + 1)
        t_mid = (e["start"][sync] + 10 * fs * 1e-3 + 20 * fs * 1e-3 * np.arange(len(b))) / fs
        tx = (truth[ch["prn"]][synth.nav_bit_index(sat, t_mid)] > 0).astype(np.int8)
        # the first second belongs to the loops' pull-in (bit synchronisation is declared 100 ms after the
        # hand-over, reference rule, possibly before the 8 Hz PLL has settled): compared separately
        late = t_mid > 1.0
        agree = float((tx[late] == b[late]).mean())
        early_bad = int(min((tx[~late] == b[~late]).sum(), (tx[~late] != b[~late]).sum()))
        good = abs(df) < 5.0 and end > seconds - 0.002 and (agree == 1.0 or agree == 0.0) and len(b) >= int(seconds * 50) - 10
        ok &= good
        print(f"  PRN {ch['prn']:2d}: {len(e)} epochs to t = {end:.4f} s, Doppler error {df:+.2f} Hz, {len(b)} nav bits, "
              f"agreement with the transmitted data after 1 s {agree:.3f} ({early_bad} pull-in bit errors before) {'ok' if good else 'FAIL'}")
    print("FULL-SIZE CHECK", "PASSED" if ok else "FAILED")
finally:
    shutil.rmtree(tmp, ignore_errors=True)
