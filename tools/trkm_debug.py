"""Diagnostics: the prefix-moment kernel against the per-channel general kernel, epoch by epoch, on the
configs[4]-shaped test input (3 recordings x 12 channels x 0.5 s)."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sydr_b200 import synth
from sydr_b200.engine import AcquisitionEngine, TrackingEngine, make_trk_states

FS = 25e6
n_rec, seconds = int(sys.argv[1]) if len(sys.argv) > 1 else 3, float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
group = int(sys.argv[3]) if len(sys.argv) > 3 else 3
n = int(round(seconds * FS)); pad = 2048
buf = torch.zeros(n_rec * (2 * n + pad) + 4096, dtype=torch.int16, device="cuda")
acq = AcquisitionEngine(FS, 0.0, 5000.0, 250.0, 1, 10, list(synth.PRNS_12))
chans = []
for r in range(n_rec):
    sc = synth.make_scenario(FS, 16, seconds, synth.PRNS_12, 1005 + r, 250.0)
    base = r * (2 * n + pad)
    buf[base:base + 2 * n] = synth.generate_iq_torch(sc, device="cuda")
    for p in acq.run(buf[base:base + 2 * n])["peaks"]:
        carrier, _, cur = acq.handoff(p)
        chans.append(dict(prn=int(p["prn"]), carrier_freq=carrier, start_sample=cur, iq_base=base // 2, iq_len=n, rec=r))
acq.close()
st = make_trk_states(FS, chans)
ref = TrackingEngine(FS, st, int(seconds * 1000) + 8, cluster=1, threads=0, use_tma=True, kernel=2)
ref.launch(buf); a = ref.fetch()
m = TrackingEngine(FS, st, int(seconds * 1000) + 8, kernel=1, group=group)
from sydr_b200 import _lib as L
for rep in (0, 4, 5):
    L.load().sydr_trkm_debug(rep)
    m.reset(st)
    m.launch(buf); b = m.fetch()
    print("status", np.unique(m.states()["status"]), "epochs", min(len(r) for r in b), max(len(r) for r in b))
    nbad = 0
    for c, (ra, rb) in enumerate(zip(a, b)):
        k = min(len(ra), len(rb))
        same = np.array_equal(ra["start"][:k], rb["start"][:k]) and np.array_equal(ra["n"][:k], rb["n"][:k])
        sc_ = np.hypot(ra["corr"][:k, 2], ra["corr"][:k, 3])
        e = np.abs(ra["corr"][:k] - rb["corr"][:k]).max(axis=1) / sc_
        bad = np.nonzero(e > 2e-5)[0]
        nbad += int(len(bad) > 0 or not same)
        if (len(bad) or not same or len(ra) != len(rb)) and nbad < 3:
            print(f"ch {c} prn {chans[c]['prn']} rec {chans[c]['rec']}: len {len(ra)}/{len(rb)} same_bounds {same} bad epochs {bad[:8]} e {e[bad[:8]]}"
                  f" start {ra['start'][bad[:4]]} start%2048 {ra['start'][bad[:4]] % 2048}")
            for kk in bad[:2]:
                print("   ref", ra["corr"][kk], "\n   got", rb["corr"][kk], "\n   diff", rb["corr"][kk] - ra["corr"][kk])
    print("rep", rep, "channels with differences:", nbad, "max e overall", max(float((np.abs(ra["corr"][:min(len(ra), len(rb))] - rb["corr"][:min(len(ra), len(rb))]).max(axis=1) / np.hypot(ra["corr"][:min(len(ra), len(rb)), 2], ra["corr"][:min(len(ra), len(rb)), 3])).max()) for ra, rb in zip(a, b)))

# ---- teacher-forced check of the moments kernel against numpy, epoch by epoch, from ITS OWN records
from fractions import Fraction as F
from sydr_b200.synth import ca_code_pm1
L.load().sydr_trkm_debug(int(os.environ.get("TRKM_DEBUG", "0")))
m.reset(st); m.launch(buf); b = m.fetch()
if os.environ.get("TRKM_CHECK_REF"):          # check the per-channel kernel's records instead
    b = a
hbuf = buf.cpu().numpy()
events = 0
for c, rb in enumerate(b):
    if os.environ.get('TRKM_ONLY') and c != int(os.environ['TRKM_ONLY']):
        continue
    ch = chans[c]
    base = ch["iq_base"]
    code = ca_code_pm1(ch["prn"]).astype(np.float64); code = np.r_[code[-1], code, code[0]]
    org = (min(int(r_["start"][0]) for r_, cc in zip(b, chans) if cc["rec"] == ch["rec"]) // 2048) * 2048
    for kk in range(1, len(rb)):
        start, n = int(rb["start"][kk]), int(rb["n"][kk])
        raw = hbuf[2 * (base + start):2 * (base + start + n)].astype(np.float64)
        x = raw[0::2] + 1j * raw[1::2]
        fc, remc, remcode = rb["carrier_freq"][kk - 1], rb["rem_carrier"][kk - 1], rb["rem_code"][kk - 1]
        step = rb["code_freq"][kk - 1] / FS
        t = np.arange(n) / FS
        z = np.exp(1j * (-2 * np.pi * fc * t + remc)) * x
        refc = []
        for sp in (-0.5, 0.0, 0.5):
            idx = np.ceil(np.linspace(remcode + sp, step * n + remcode + sp, n, endpoint=False)).astype(int)
            refc += [float((code[idx] * z.real).sum()), float((code[idx] * z.imag).sum())]
        refc = np.array(refc)
        e = np.abs(rb["corr"][kk] - refc).max() / np.hypot(refc[2], refc[3])
        if e > 2e-5:
            events += 1
            d = (rb["corr"][kk] - refc).reshape(3, 2)
            dz = d[:, 0] + 1j * d[:, 1]
            tap = int(np.argmax(np.abs(dz)))
            cand = np.minimum(np.minimum(np.abs(z * 2 - dz[tap]), np.abs(z * 2 + dz[tap])), np.minimum(np.abs(z - dz[tap]), np.abs(z + dz[tap])))
            j = int(np.argmin(cand))
            ph = np.linspace(remcode, step * n + remcode, n, endpoint=False)
            jabs = start + j
            print(f"ch {c} prn {ch['prn']} epoch {kk}: e {e:.2e} dz {np.round(dz, 1)} tap {tap} sample j={j} of {n} resid {cand[j]:.1f} z {np.round(z[j], 1)} 2*phase {2 * ph[j]:.7f}"
                  f" ring-rel {jabs - org} %16 {(jabs - org) % 16} %512 {(jabs - org) % 512} %2048 {(jabs - org) % 2048} %8192 {(jabs - org) % 8192} start-rel {start - org}")
            for sp in (-0.5, 0.0, 0.5):
                st_, sp_ = remcode + sp, step * n + remcode + sp
                stepp = (sp_ - st_) / n
                lat = np.round(ph[j] + sp)
                Xt = (F(float(lat)) - F(float(st_))) / F(float(stepp))
                vals = [float(np.float64(jj) * stepp + st_) for jj in (j - 1, j, j + 1)]
                print(f"    tap {sp:+.1f}: start {st_!r} step' {stepp!r} lattice {lat} X_true - j = {float(Xt - j):.3e}  fl(phase) at j-1, j, j+1 minus lattice: "
                      f"{vals[0] - lat:.3e} {vals[1] - lat:.3e} {vals[2] - lat:.3e}  ceil idx {[int(np.ceil(v)) for v in vals]}")
            Xp = (F(float(round(2 * ph[j]) / 2)) - F(float(remcode))) * F(1.0 / step)
            print(f"    prompt crossing with 1/code_step: X - j = {float(Xp - j):.3e}; rem_code {remcode!r} code_step {step!r}")
print("teacher-forced events above 2e-5:", events)
