"""Per-channel Kaplan lock state on the headline recording (device), for comparison with the CPU oracle."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from sydr_b200 import synth
from sydr_b200.pipeline import ColdStartPipeline
fs, dur = 25e6, 0.7
sc = synth.make_scenario(fs, 16, dur, synth.PRNS_12, 1003, 250.0)
d = synth.generate_iq_torch(sc)
pipe = ColdStartPipeline(fs, 16, list(range(1, 33)), 12, max_seconds=dur, loop="kaplan")
buf = pipe.device_buffer(d.numel() // 2); buf.copy_(d)
out = pipe.finish(pipe.enqueue_device(buf), records=True, copy=True)
for ch, r, k in zip(out["channels"], out["epochs"], out["kaplan"]):
    first = {s: int(np.argmax(k["lock_state"] == s)) if (k["lock_state"] == s).any() else -1 for s in (2, 3)}
    print(f"PRN {ch['prn']:2d} carrier0 {ch['carrier_freq']:8.1f} start {ch['start_sample']:7d} epochs {len(r)} final state {int(k['lock_state'][-1])} "
          f"first WIDE {first[2]} first NARROW {first[3]} fll_lock {k['fll_lock'][-1]:.3f} pll_lock {k['pll_lock'][-1]:.3f} cn0 {k['cn0'][-1]:.1f} flags {int(k['flags'][-1])}")

# the CPU oracle, closed loop, on the same samples (PRN 7 and PRN 17)
from oracle import sydr_oracle as O          # diagnostics tool: checker only
iq = d.cpu().numpy()
x = iq[0::2].astype(np.float64) + 1j * iq[1::2].astype(np.float64)
for ch, r, k in zip(out["channels"], out["epochs"], out["kaplan"]):
    if ch["prn"] not in (7, 17):
        continue
    o = O.KaplanTrackOracle(ch["prn"], fs, ch["carrier_freq"], ch["start_sample"])
    states, fl = [], []
    while o.cur + o.n_req <= len(x) and len(states) < len(r):
        w = o.step(x)
        states.append(w["lock_state"]); fl.append(w["fll_lock"])
    st = np.array(states)
    first = {s: int(np.argmax(st == s)) if (st == s).any() else -1 for s in (2, 3)}
    print(f"oracle PRN {ch['prn']:2d}: epochs {len(st)} final state {st[-1]} first WIDE {first[2]} first NARROW {first[3]} "
          f"fll_lock {fl[-1]:.3f} | device fll_lock {k['fll_lock'][-1]:.3f}, max |fll_lock diff| {np.abs(np.array(fl) - k['fll_lock'][:len(fl)]).max():.4f}, "
          f"carrier diff at the end {abs(o.carrier_freq - r['carrier_freq'][len(st) - 1]):.3f} Hz")
