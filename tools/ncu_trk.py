"""One tracking launch for ncu (12 channels, 25 MS/s int16, 300 epochs)."""
import sys
import torch
sys.path.insert(0, ".")
from sydr_b200 import synth
from sydr_b200.engine import TrackingEngine, make_trk_states, AcquisitionEngine

cluster = int(sys.argv[1]) if len(sys.argv) > 1 else 8
fs, dur = 25e6, 0.32
sc = synth.make_scenario(fs, 16, dur, synth.PRNS_12, 1003, 250.0)
d_iq = synth.generate_iq_torch(sc)
d_all = torch.cat([d_iq, torch.zeros(2048, dtype=d_iq.dtype, device="cuda")])[:d_iq.numel()]
acq = AcquisitionEngine(fs, 0.0, 5000, 250, 1, 10, list(synth.PRNS_12))
peaks = acq.run(d_all)["peaks"]
chans = [dict(prn=int(p["prn"]), carrier_freq=acq.handoff(p)[0], start_sample=acq.handoff(p)[2], iq_len=d_all.numel() // 2) for p in peaks]
eng = TrackingEngine(fs, make_trk_states(fs, chans), 400, cluster=cluster)
eng.launch(d_all)
torch.cuda.synchronize()
print("epochs", [len(r) for r in eng.fetch()])
