"""Tracking kernel timing sweep without instrumentation (run on the GPU box)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from sydr_b200 import _lib as L, synth
from sydr_b200.engine import TrackingEngine, make_trk_states, AcquisitionEngine

fs, dur = 25e6, 1.0
sc = synth.make_scenario(fs, 16, dur, synth.PRNS_12, 1003, 250.0)
d_iq = synth.generate_iq_torch(sc)
d_all = torch.cat([d_iq, torch.zeros(2048, dtype=d_iq.dtype, device="cuda")])[:d_iq.numel()]
acq = AcquisitionEngine(fs, 0.0, 5000, 250, 1, 10, list(synth.PRNS_12))
peaks = acq.run(d_all)["peaks"]
chans = [dict(prn=int(p["prn"]), carrier_freq=acq.handoff(p)[0], start_sample=acq.handoff(p)[2], iq_len=d_all.numel() // 2) for p in peaks]
L.load().sydr_trk_profile_buffer(None)
for cluster, threads in ((8, 0), (8, 256), (8, 224), (8, 192), (8, 160), (8, 128), (8, 96), (4, 0), (4, 288), (2, 0), (1, 0), (1, 352)):
    ts = []
    for rep in range(3):
        eng = TrackingEngine(fs, make_trk_states(fs, chans), 1100, cluster=cluster, threads=threads)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.launch(d_all); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    nep = len(eng.fetch()[0])
    print(f"S={cluster} T={threads:4d}: {min(ts) * 1e3 / nep:6.2f} us/epoch ({nep} epochs)  RTF {nep / min(ts):.0f}")
