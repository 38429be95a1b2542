"""Diagnostics: per-phase cycle breakdown of the tracking kernel's serial chain (run on the GPU box)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from sydr_b200 import _lib as L, synth
from sydr_b200.engine import TrackingEngine, make_trk_states, to_device_iq, AcquisitionEngine

fs, dur = 25e6, 0.5
sc = synth.make_scenario(fs, 16, dur, synth.PRNS_12, 1003, 250.0)
d_iq = synth.generate_iq_torch(sc)
pad = torch.zeros(2048, dtype=d_iq.dtype, device="cuda")
d_all = torch.cat([d_iq, pad])[:d_iq.numel()]
acq = AcquisitionEngine(fs, 0.0, 5000, 250, 1, 10, list(synth.PRNS_12))
peaks = acq.run(d_all)["peaks"]
chans = [dict(prn=int(p["prn"]), carrier_freq=acq.handoff(p)[0], start_sample=acq.handoff(p)[2], iq_len=d_all.numel() // 2) for p in peaks]
names = ["const", "barrier", "win wait", "correlate", "wsum+send", "gather wait", "close", "totals", "pub:math", "pub:store", "w1:->gather", "w1:totals", "w1:close", "w1:const"]
for cluster in (1, 8):
    for tma in (True,):
        prof = torch.zeros(len(chans) * 16, dtype=torch.int64, device="cuda")
        L.load().sydr_trk_profile_buffer(prof.data_ptr())
        eng = TrackingEngine(fs, make_trk_states(fs, chans), 600, cluster=cluster, use_tma=tma)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.launch(d_all); torch.cuda.synchronize()
        eng = TrackingEngine(fs, make_trk_states(fs, chans), 600, cluster=cluster, use_tma=tma)
        e0.record(); eng.launch(d_all); e1.record(); torch.cuda.synchronize()
        p = prof.cpu().numpy().reshape(-1, 16)
        ep = p[:, 15].mean()
        cyc = p[:, :14].mean(axis=0) / ep
        print(f"S={cluster} tma={int(tma)} {e0.elapsed_time(e1) * 1e3 / ep:7.2f} us/epoch | " +
              "  ".join(f"{n}:{c:6.0f}" for n, c in zip(names, cyc)) + f" | sum {cyc[:10].sum():.0f} cyc")
L.load().sydr_trk_profile_buffer(None)
