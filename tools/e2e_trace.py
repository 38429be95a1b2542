"""Diagnostics: host-side timeline of ColdStartPipeline.process_host (run on the GPU box)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import bench as B
from sydr_b200.pipeline import ColdStartPipeline
from sydr_b200 import pipeline as PL

dev = torch.device("cuda", 0)
sc, host = B.make_recording(0, 2.0, dev)
pipe = ColdStartPipeline(B.FS, B.NBITS, B.SEARCH_PRNS, B.N_CHANNELS, max_seconds=2.0, device=dev, **B.ACQ)
for _ in range(3):
    pipe.process_host(host)
torch.cuda.synchronize()

marks = []
def mark(name):
    marks.append((name, time.perf_counter()))

orig_launch = pipe.acq.launch
orig_fetch = pipe.acq.fetch
orig_start = pipe._start_tracking
orig_collect = pipe.collect
def launch(*a, **k):
    mark("acq.launch>"); r = orig_launch(*a, **k); mark("acq.launch<"); return r
def fetch(*a, **k):
    mark("acq.fetch>"); r = orig_fetch(*a, **k); mark("acq.fetch<"); return r
def start(*a, **k):
    mark("start_trk>"); r = orig_start(*a, **k); mark("start_trk<"); return r
def collect(*a, **k):
    mark("collect>"); r = orig_collect(*a, **k); mark("collect<"); return r
pipe.acq.launch, pipe.acq.fetch, pipe._start_tracking, pipe.collect = launch, fetch, start, collect
for rep in range(3):
    marks.clear()
    torch.cuda.synchronize()
    mark("begin")
    pipe.process_host(host)
    mark("end")
    t0 = marks[0][1]
    print(" | ".join(f"{n} {1e3 * (t - t0):.2f}" for n, t in marks))
