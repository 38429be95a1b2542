"""e2e (pinned host -> results) step time vs number of upload pieces."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sydr_b200.pipeline import ColdStartPipeline
dev = torch.device("cuda", 0)
sc, host = bench.make_recording(0, 2.0, dev)
pipe = ColdStartPipeline(bench.FS, bench.NBITS, bench.SEARCH_PRNS, bench.N_CHANNELS, max_seconds=2.0, device=dev, **bench.ACQ)
for pieces in (2, 3, 4, 6, 8, 12, 16):
    for _ in range(3):
        pipe.process_host(host, pieces=pieces)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        pipe.process_host(host, pieces=pieces)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print(f"pieces {pieces:2d}: {dt * 1e3:.3f} ms/step  RTF {2.0 / dt:.0f}")
