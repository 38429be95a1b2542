"""ColdStartBatch timing: B recordings x 12 channels in one tracking launch (python tools/batch_time.py B seconds [dense cluster threads] ...)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from sydr_b200 import synth
from sydr_b200.pipeline import ColdStartBatch

FS = 25e6


def run(B, seconds, dense, cluster, threads, alias=0, seeds=2):
    kw = dict(fs=FS, nbits=16, search_prns=list(range(1, 33)), n_channels=12, max_seconds=seconds,
              doppler_range=5000.0, doppler_step=250.0, coh=1, noncoh=10)
    batch = ColdStartBatch(B, dense=dense, cluster=cluster, threads=threads, **kw)
    if alias:            # diagnostics: every recording reads slot 0 (what lanes sharing one buffer do)
        from sydr_b200 import _lib as L
        t = batch._tmpls.cpu().numpy().view(L.TRK_STATE_DTYPE).copy()
        t["iq_base"] = 0
        batch._tmpls.copy_(torch.from_numpy(t.view(np.uint8).reshape(B, -1)))
    t0 = time.time()
    for r in range(B):
        if r < seeds:
            sc = synth.make_scenario(FS, 16, seconds, synth.PRNS_12, 1003 + r, 250.0)
            batch.slot(r).copy_(synth.generate_iq_torch(sc, device="cuda"))
        else:
            batch.slot(r).copy_(batch.slot(r % seeds))
    torch.cuda.synchronize()
    gen_s = time.time() - t0
    res = []
    for it in range(4):
        m = []
        ctx = batch.enqueue(marks=m)
        torch.cuda.synchronize()
        res.append((m[0].elapsed_time(m[1]), m[2].elapsed_time(m[3]), m[0].elapsed_time(m[3])))
    out = batch.finish(ctx, records=False)
    nep = batch._trk._nep.cpu().numpy()
    st = batch._trk.states()
    acq_ms, trk_ms, all_ms = (float(np.mean([x[i] for x in res[1:]])) for i in range(3))
    n = int(seconds * FS)
    print(f"B={B:3d} {seconds:g} s alias={alias} dense={dense} S={cluster} T={threads:3d}: acq+handoff {acq_ms:7.2f} ms, trk {trk_ms:8.2f} ms "
          f"({trk_ms * 1e3 / (seconds * 1e3):6.2f} us per epoch of all {B * 12} channels, {trk_ms * 1e3 / (seconds * 1e3) / B:5.3f} us per recording-epoch), "
          f"step {all_ms:8.2f} ms = {B * n / all_ms / 1e3:8.0f} Msamples/s; trk alone {B * n * 12 * 31 / trk_ms / 1e9:5.1f} TFLOP/s algorithmic; "
          f"channels found {[len(o['channels']) for o in out][:3]}.., epochs min {nep.min()} max {nep.max()}, status!=0: {(st['status'] != 0).sum()}, gen {gen_s:.1f} s", flush=True)
    batch.close()
    del batch
    torch.cuda.empty_cache()


if __name__ == "__main__":
    a = sys.argv[1:]
    B, seconds = int(a[0]), float(a[1])
    cfgs = [tuple(int(v) for v in c.split(",")) for c in a[2:]] or [(2, 1, 0)]
    for c in cfgs:
        run(B, seconds, *c)
