"""Diagnostics: H2D bandwidth and the end-to-end step breakdown (run on the GPU box)."""
import sys, time
import torch
sys.path.insert(0, ".")
import bench as B
from sydr_b200.pipeline import ColdStartPipeline

dev = torch.device("cuda", 0)
sc, host = B.make_recording(0, 2.0, dev)
d = torch.empty_like(host, device=dev)
for _ in range(3):
    d.copy_(host, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); d.copy_(host, non_blocking=True); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"H2D {host.numel() * 2 / 1e6:.0f} MB pinned: {ms:.2f} ms = {host.numel() * 2 / ms / 1e6:.1f} GB/s")
pipe = ColdStartPipeline(B.FS, B.NBITS, B.SEARCH_PRNS, B.N_CHANNELS, max_seconds=2.0, device=dev, **B.ACQ)
for pieces in (1, 2, 4, 8, 16, 32):
    for _ in range(2):
        pipe.process_host(host, pieces=pieces)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        pipe.process_host(host, pieces=pieces)
    torch.cuda.synchronize()
    print(f"pieces={pieces:3d}: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms per step")
