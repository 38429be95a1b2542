"""GPU-side timeline of one end-to-end step (ColdStartPipeline.process_host), from CUDA events."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from sydr_b200 import _lib as L
from sydr_b200.pipeline import ColdStartPipeline
dev = torch.device("cuda", 0)
sc, host = B.make_recording(0, 2.0, dev)
pipe = ColdStartPipeline(B.FS, B.NBITS, B.SEARCH_PRNS, B.N_CHANNELS, max_seconds=2.0, device=dev, **B.ACQ)
pieces = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for _ in range(3):
    pipe.process_host(host, pieces=pieces)
torch.cuda.synchronize()
E = lambda: torch.cuda.Event(enable_timing=True)
def step():
    comp = torch.cuda.current_stream()
    n_el = host.numel(); n = n_el // 2
    first = min(n, pipe.acq.required_samples + 4 * pipe.acq.n_code)
    stp = max(1, -(-(n - first) // pieces))
    bounds = [first]
    while bounds[-1] < n:
        bounds.append(min(n, bounds[-1] + stp))
    d = pipe._d_iq[:n_el]
    t_host0 = time.perf_counter()
    g0 = E(); g0.record(comp)
    events = []
    pipe._copy_stream.wait_stream(comp)
    with torch.cuda.stream(pipe._copy_stream):
        lo = 0
        for hi in bounds:
            d[2 * lo:2 * hi].copy_(host[2 * lo:2 * hi], non_blocking=True)
            ev = E(); ev.record(pipe._copy_stream); events.append(ev); lo = hi
    comp.wait_event(events[0])
    pipe.acq.launch(d[:2 * first])
    a1 = E(); a1.record(comp)
    got = pipe._peaks_to_host_async()
    pipe._handoff(n)
    tl = []
    for hi, ev in zip(bounds, events):
        comp.wait_event(ev)
        s = E(); s.record(comp)
        pipe._trk.launch(d, iq_len=hi, append=True)
        e = E(); e.record(comp)
        tl.append((s, e))
    t_enq = time.perf_counter()
    got.synchronize()
    pipe._n_active = 12
    recs = pipe.collect()
    t_done = time.perf_counter()
    torch.cuda.synchronize()
    print(f"host: enqueued at {1e3 * (t_enq - t_host0):.2f} ms, results on host at {1e3 * (t_done - t_host0):.2f} ms")
    print("gpu : H2D pieces done at", [round(g0.elapsed_time(ev), 2) for ev in events], "ms; acq done", round(g0.elapsed_time(a1), 2))
    print("gpu : trk launches (start, end):", [(round(g0.elapsed_time(s), 2), round(g0.elapsed_time(e), 2)) for s, e in tl])
for _ in range(2):
    step()

# ---- where collect() spends its time
import numpy as np
trk = pipe._trk
for _ in range(3):
    t0 = time.perf_counter()
    trk._nep_host.copy_(trk._nep, non_blocking=True)
    trk._out_host = trk._out_ring[0] if trk._out_ring[0] is not None else torch.empty(trk.n_ch * trk.max_epochs * 128, dtype=torch.uint8, pin_memory=True)
    trk._out_ring[0] = trk._out_host
    trk._out_host.copy_(trk._out, non_blocking=True)
    t1 = time.perf_counter()
    torch.cuda.current_stream().synchronize()
    t2 = time.perf_counter()
    nep = trk._nep_host.numpy()
    out = trk._out_host.numpy().view(L.TRK_EPOCH_DTYPE).reshape(trk.n_ch, trk.max_epochs)
    res = [out[c, :nep[c]].copy() for c in range(trk.n_ch)]
    t3 = time.perf_counter()
    whole = out.copy()
    t4 = time.perf_counter()
    print(f"collect: enqueue {1e3 * (t1 - t0):.3f} ms, D2H wait {1e3 * (t2 - t1):.3f} ms ({trk._out.numel() / 1e6:.2f} MB), "
          f"12 per-channel copies {1e3 * (t3 - t2):.3f} ms, one bulk copy {1e3 * (t4 - t3):.3f} ms")
