"""Device Kaplan loop closure vs Borre: 12 channels, 25 MS/s int16, 1 s, closed loop."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from sydr_b200 import synth
from sydr_b200.engine import (AcquisitionEngine, KaplanTrackingEngine, TrackingEngine, make_kaplan_states, make_trk_states)

fs, dur = 25e6, 1.0
sc = synth.make_scenario(fs, 16, dur, synth.PRNS_12, 1003, 250.0)
d_iq = synth.generate_iq_torch(sc)
d_all = torch.cat([d_iq, torch.zeros(2048, dtype=d_iq.dtype, device="cuda")])[:d_iq.numel()]
acq = AcquisitionEngine(fs, 0.0, 5000, 250, 1, 10, list(synth.PRNS_12))
peaks = acq.run(d_all)["peaks"]
chans = [dict(prn=int(p["prn"]), carrier_freq=acq.handoff(p)[0], start_sample=acq.handoff(p)[2], iq_len=d_all.numel() // 2) for p in peaks]
truth = {s.prn: s.doppler for s in sc.sats}
for name in ("borre", "kaplan"):
    ts = []
    for rep in range(3):
        if name == "borre":
            eng = TrackingEngine(fs, make_trk_states(fs, chans), 1100)
        else:
            st, ks = make_kaplan_states(fs, chans)
            eng = KaplanTrackingEngine(fs, st, ks, 1100)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.launch(d_all); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    res = eng.fetch()
    nep = len(res[0])
    err = max(abs(float(np.mean(r["carrier_freq"][-100:])) - truth[c["prn"]]) for r, c in zip(res, chans))
    extra = ""
    if name == "kaplan":
        k = eng.fetch_kaplan()
        extra = f"  lock states at the end {sorted(set(int(x['lock_state'][-1]) for x in k))}, cn0 {np.mean([x['cn0'][-1] for x in k]):.0f}"
    print(f"{name:7s}: {min(ts) * 1e3 / nep:6.3f} us/epoch ({nep} epochs, {min(ts):.3f} ms)  RTF {dur * 1e3 / min(ts):.0f}  max |df| {err:.2f} Hz{extra}")
