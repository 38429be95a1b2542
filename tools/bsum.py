"""Short summary of a bench.py JSON line."""
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        k = d["roofline"]["kernels"]
        t = d["roofline"].get("throughput_mode", {})
        print(f"value {d['value']:.0f} Msps  step {d['ms_per_step']:.3f} ms  e2e {d['e2e']['value']:.0f}  "
              f"trk {k['trk_borre_kernel']['ms']:.3f} ms (alone {k['trk_borre_kernel'].get('alone_ms', 0):.3f}, {k['trk_borre_kernel'].get('alone_us_per_epoch', 0):.3f} us/epoch)  "
              f"acq {k['acq (fwd+ifft+reduce)']['ms']:.3f} ms (alone {k['acq (fwd+ifft+reduce)'].get('alone_ms', 0):.3f})  lean {t.get('ms', 0):.3f} ms frac {t.get('frac', 0):.3f}  "
              f"file {d['e2e'].get('from_file', {}).get('rtf', 0):.0f}")
