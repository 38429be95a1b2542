"""Throughput stress (BASELINE.json configs[4]): R concurrent recordings x 12 channels of closed-loop
tracking at 25 MS/s int16 on ONE GPU, one launch.  Sweeps the launch shape of K-TRK and reports
per-epoch time over all channels, real-time factor and the algorithmic FP32 rate
(31 flop per sample per channel).  Run on the GPU box:
    python tools/cfg5.py [recordings=32] [seconds=0.5]
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from sydr_b200 import _lib as L, synth  # noqa: E402
from sydr_b200.engine import AcquisitionEngine, TrackingEngine, make_trk_states  # noqa: E402


def build(n_rec, dur, fs=25e6, seed0=1005):
    n = int(round(dur * fs))
    pad = 2048
    buf = torch.zeros(n_rec * (2 * n + pad) + 4096, dtype=torch.int16, device="cuda")
    acq = AcquisitionEngine(fs, 0.0, 5000, 250, 1, 10, list(synth.PRNS_12))
    chans, truth = [], []
    for r in range(n_rec):
        sc = synth.make_scenario(fs, 16, dur, synth.PRNS_12, seed0 + r, 250.0)
        base = r * (2 * n + pad)                       # int16 elements; a multiple of 8 -> 16-byte aligned
        buf[base:base + 2 * n] = synth.generate_iq_torch(sc)
        peaks = acq.run(buf[base:base + 2 * n])["peaks"]
        for p in peaks:
            carrier, _, cur = acq.handoff(p)
            chans.append(dict(prn=int(p["prn"]), carrier_freq=carrier, start_sample=cur, iq_base=base // 2, iq_len=n))
        truth += [s.doppler for s in sc.sats]
    acq.close()
    return buf, chans, np.array(truth), n


def main():
    n_rec = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dur = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
    fs = 25e6
    buf, chans, truth, n = build(n_rec, dur)
    n_ch = len(chans)
    max_ep = int(dur * 1000) + 8
    L.load().sydr_trk_profile_buffer(None)
    import os
    prof = None
    if os.environ.get("TRKM_PROF"):           # cycle counters of the prefix-moment kernel (diagnostics instantiation)
        prof = torch.zeros(n_ch * 16, dtype=torch.int64, device="cuda")
        L.load().sydr_trk_profile_buffer(prof.data_ptr())
    # (cluster, threads, tma[, kernel, group]): kernel 1 = prefix-moment kernel (trkm.cu; group = channels per CTA),
    # 2 = per-channel kernels (trk.cu)
    shapes = [(0, 0, 1, 1, 3), (0, 0, 1, 1, 4), (0, 0, 1, 1, 2), (1, 256, 0, 2, 0)]
    if len(sys.argv) > 3:
        shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[3:]]
    for shp in shapes:
        cluster, threads, tma = shp[:3]
        kernel, group = (shp[3], shp[4]) if len(shp) >= 5 else (2, 0)
        ts = []
        try:
            for rep in range(3):
                eng = TrackingEngine(fs, make_trk_states(fs, chans), max_ep, cluster=cluster, threads=threads, use_tma=bool(tma),
                                     kernel=kernel, group=group)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); eng.launch(buf); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
        except L.SydrError as e:
            print(f"S={cluster} T={threads} tma={tma}: {e}")
            continue
        res = eng.fetch()
        if prof is not None and kernel == 1:
            pc = prof.cpu().numpy().reshape(-1, 16).astype(float)
            pc = pc[pc[:, 0] > 0]
            ep = nep_mean = float(np.mean([len(r) for r in res]))
            print("   one warp per CTA, cycles per epoch (mean / min / max over CTAs): " + "  ".join(
                f"{nm} {(pc[:, i] / ep).mean():.0f}/{(pc[:, i] / ep).min():.0f}/{(pc[:, i] / ep).max():.0f}"
                for i, nm in ((0, "total"), (1, "moments"), (2, "pieces"), (3, "waiting"), (4, "closures"))) +
                f"   per epoch: blocks {pc[:, 5].mean() / ep:.2f} pieces {pc[:, 6].mean() / ep:.2f} closures {pc[:, 7].mean() / ep:.2f}")
        nep = np.array([len(r) for r in res])
        err = np.array([abs(r["carrier_freq"][-1] - t) for r, t in zip(res, truth)])
        ms = min(ts)
        samples = float(sum(r["n"].sum() for r in res))
        tflops = 31.0 * samples / (ms * 1e-3) / 1e12
        print(f"R={n_rec} ch={n_ch} S={cluster} T={threads:3d} tma={tma} kernel={kernel} group={group}: {ms:8.2f} ms  {ms * 1e3 / nep.mean():7.2f} us/epoch(all ch)  "
              f"RTF {dur * 1e3 / ms:6.1f}  {samples / (ms * 1e-3) / 1e9:7.1f} Gsample-ch/s  {tflops:6.2f} TFLOP/s alg  "
              f"epochs {nep.min()}..{nep.max()}  max|df| {err.max():.2f} Hz", flush=True)


if __name__ == "__main__":
    main()
