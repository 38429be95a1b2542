"""File-ingest sweep: reader threads x chunk length for bench.py's from_file measurement."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
for chunk in (0.25, 0.5, 1.0, 2.0):
    for thr in (2, 4, 8, 16):
        r = bench.file_ingest(dev, 6.0, chunk, reader_threads=thr, reps=3)
        print(f"chunk {chunk:4.2f} s  threads {thr:2d}  wall {r['wall_ms']:7.2f} ms  RTF {r['rtf']:6.1f}  "
              f"{r['file_bytes'] / r['wall_ms'] / 1e6:5.1f} GB/s", flush=True)
