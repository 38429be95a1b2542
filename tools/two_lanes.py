"""Experiment: two ColdStartPipeline lanes on two streams, steps alternating (acquisition of step k+1
overlaps the tracking of step k, which occupies 96 of the 148 SMs)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from sydr_b200.pipeline import ColdStartPipeline
dev = torch.device("cuda", 0)
sc, host = B.make_recording(0, 2.0, dev)
lanes = []
for i in range(2):
    p = ColdStartPipeline(B.FS, B.NBITS, B.SEARCH_PRNS, B.N_CHANNELS, max_seconds=2.0, device=dev, **B.ACQ)
    s = torch.cuda.Stream(device=dev, priority=0)
    with torch.cuda.stream(s):
        d = p.upload(host)
    lanes.append((p, s, d))
torch.cuda.synchronize()
def run(n_lanes, steps):
    t0 = time.perf_counter()
    for k in range(steps):
        p, s, d = lanes[k % n_lanes]
        with torch.cuda.stream(s):
            p.process_device(d)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps
for n in (1, 2):
    run(n, 6)
    dt = run(n, 20)
    print(f"{n} lane(s): {dt * 1e3:.3f} ms/step  {2.0 * B.FS / dt / 1e6:.0f} Msamples/s")
# results still right?
p, s, d = lanes[1]
with torch.cuda.stream(s):
    out = p.process_device(d)
    ep = p.collect()
import numpy as np
truth = {x.prn: x.doppler for x in sc.sats}
print("max Doppler error", max(abs(float(np.mean(e["carrier_freq"][-200:])) - truth[c["prn"]]) for c, e in zip(out["channels"], ep)))
