"""Acquisition-only timings on the BASELINE.json acquisition configs (one GPU): cfg2 (10 MS/s, 41 bins, 1x10),
the headline dwell (25 MS/s, 41 bins, 1x10) and cfg4 (50 MS/s, 201 bins of 50 Hz, 1x20), 32 PRNs each."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sydr_b200 import synth
from sydr_b200.engine import AcquisitionEngine, to_device_iq

def f_acq(n):
    return 10.0 * n * np.log2(n) + 19.0 * n

for name, fs, nbits, rng, step, noncoh in (("cfg2", 10e6, 8, 5000, 250, 10), ("headline", 25e6, 16, 5000, 250, 10),
                                           ("cfg4", 50e6, 8, 5000, 50, 20)):
    n = int(fs * 1e-3)
    sc = synth.make_scenario(fs, nbits, noncoh * 1e-3 + 0.002, synth.PRNS_8, 1002, float(step))
    d = to_device_iq(synth.generate_iq(sc))
    eng = AcquisitionEngine(fs, 0.0, rng, step, 1, noncoh, list(range(1, 33)))
    for _ in range(3):
        eng.launch(d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.launch(d)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    pk = eng.fetch()["peaks"]
    found = sorted(int(p["prn"]) for p in pk if p["ratio"] > 1.5)
    flop = 32 * eng.n_bins * noncoh * f_acq(n)
    print(f"{name:9s} fs {fs / 1e6:4.0f} MS/s  N {n:6d}  bins {eng.n_bins:4d}  blocks {noncoh:3d}: {ms:8.3f} ms per 32-PRN sweep, "
          f"{flop / ms / 1e9:6.1f} TFLOP/s algorithmic, {32 * eng.n_bins * noncoh * n / ms / 1e6:7.1f} G cell-samples/s, found {found}")
    eng.close()
