"""One device-resident step of the headline workload (bench.py's step_device) for ncu captures:
    ncu --set full --clock-control none --import-source on -k regex:"trk_borre|acq_" -o gpurun_out/X python tools/ncu_bench.py"""
import sys
import torch
sys.path.insert(0, ".")
import bench as B  # noqa: E402
from sydr_b200.pipeline import ColdStartPipeline  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
sc, host = B.make_recording(0, 2.0, dev)
pipe = ColdStartPipeline(B.FS, B.NBITS, B.SEARCH_PRNS, B.N_CHANNELS, max_seconds=2.0, device=dev, **B.ACQ)
d_iq = pipe.upload(host)
torch.cuda.synchronize()
out = pipe.process_device(d_iq)
torch.cuda.synchronize()
print("epochs", [len(r) for r in pipe.collect()])
