"""Shared-memory wavefronts (ideal / excessive = bank conflicts) and stall samples per CUDA source line of one kernel of an
.ncu-rep:   python tools/ncu_smem_lines.py REP.ncu-rep OBJ.o KERNEL_SUBSTRING [MANGLED_NAME_SUBSTRING]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, obj, want = sys.argv[1:4]
mangled = sys.argv[4] if len(sys.argv) > 4 else want
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        sections.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and r:
        cur["data"].append(r)
sec = next(s for s in sections if want in s["name"])
hdr, data = sec["hdr"], sec["data"]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
# the instantiation whose instruction count matches
starts = [i for i, l in enumerate(dis) if l.startswith(".text.") and mangled in l]
best = None
for st in starts:
    lines, stack, fresh = [], [], True
    for l in dis[st + 1:]:
        if l.startswith(".text.") or l.startswith("//-----"):
            break
        m = re.match(r'\s*//## File "(.*?)", line (\d+)', l)
        if m:
            if fresh:
                stack, fresh = [], False
            stack.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            lines.append((list(stack), m.group(2).strip()))
            fresh = True
    if len(lines) == len(data):
        best = lines
        break
assert best is not None, ("no instantiation with", len(data), "instructions among", len(starts))
iw, ie, ii, ismp, iex = (hdr.index(k) for k in ("L1 Wavefronts Shared", "L1 Wavefronts Shared Excessive", "L1 Wavefronts Shared Ideal", "# Samples", "Instructions Executed"))
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
for (st, op), r in zip(best, data):
    # innermost frame inside the .cu / .cuh of this repository
    loc = next((x for x in st if x[0].endswith((".cu", ".cuh", ".inc"))), st[0] if st else ("?", 0))
    a = agg[loc + (op.split()[0] if not op.startswith("@") else op.split()[1],)] if False else agg[loc]
    a[0] += int(r[iw] or 0); a[1] += int(r[ie] or 0); a[2] += int(r[ii] or 0); a[3] += int(r[ismp] or 0); a[4] += int(r[iex] or 0)
tw, te, ts = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values()), sum(a[3] for a in agg.values())
print(f"{sec['name'][:100]}\nshared wavefronts {tw}, excessive {te} ({100 * te / max(tw, 1):.1f} %), stall samples {ts}")
src = {}
def text(loc):
    f, n = loc
    if f not in src:
        p = os.path.join(os.path.dirname(os.path.abspath(obj)), f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    return src[f][n - 1].strip()[:100] if 0 < n <= len(src[f]) else ""
for loc, a in sorted(agg.items(), key=lambda t: -(t[1][0] + 0.0 * t[1][3])):
    if a[0] < 0.01 * tw and a[3] < 0.02 * ts:
        continue
    print(f"  wavefronts {a[0]:10d} ({100 * a[0] / max(tw, 1):5.1f} %)  excessive {a[1]:10d}  samples {a[3]:6d} ({100 * a[3] / max(ts, 1):5.1f} %)  ex {a[4]:10d}  {loc[0]}:{loc[1]:<4d} {text(loc)}")
