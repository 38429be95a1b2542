"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals.
    python tools/launch_summary.py gpurun_out/launches.csv "header" > profiles/rN_launches.txt
Kernels of libsydr_b200.so are listed first with their share of the library's own time (the
per-launch times are cold-cache and serialised: shares, not absolutes, are comparable with the
CUDA-event timings of bench.py)."""
import csv
import re
import sys
from collections import OrderedDict

OURS = ("trk_borre_kernel", "epl_batch_kernel", "acq_", "ca_code_kernel", "code_spectrum", "peak_rows", "convert_",
        "fp32_peak", "fp64_peak", "kaplan", "bitsync", "nav_bits", "acq_handoff", "sydr")


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name[:110]


def main():
    path, header = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu, ig, ib = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Unit", "Grid Size", "Block Size"))
    agg = OrderedDict()
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1e-3)
        k = short(r[ik])
        a = agg.setdefault(k, [0, 0.0, r[ig], r[ib]])
        a[0] += 1
        a[1] += v
    ours = {k: v for k, v in agg.items() if any(o in k for o in OURS)}
    tot = sum(v[1] for v in ours.values()) or 1.0
    print(header)
    print(f"{len(rows) - 1} launches captured; libsydr_b200 kernels: {sum(v[0] for v in ours.values())} launches, {tot:.1f} us\n")
    print(f"{'kernel':112s} {'n':>4s} {'total us':>10s} {'avg us':>10s} {'share':>7s}  grid / block")
    for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:112s} {v[0]:4d} {v[1]:10.1f} {v[1] / v[0]:10.1f} {100 * v[1] / tot:6.1f}%  {v[2]} / {v[3]}")
    print("\nother kernels (torch: synthetic-input generation, copies, fills):")
    for k, v in sorted(((k, v) for k, v in agg.items() if k not in ours), key=lambda kv: -kv[1][1])[:12]:
        print(f"{k:112s} {v[0]:4d} {v[1]:10.1f} {v[1] / v[0]:10.1f}")


if __name__ == "__main__":
    main()
