"""Executed warp-instructions and stall samples of one .ncu-rep summed over source-line ranges of one file.
    python tools/ncu_ranges.py REP.ncu-rep OBJ.o MANGLED_KERNEL FILE lo-hi[:name] ..."""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, obj, kern, fname = sys.argv[1:5]
ranges = []
for a in sys.argv[5:]:
    r, _, name = a.partition(":")
    lo, hi = r.split("-")
    ranges.append((int(lo), int(hi), name or r))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text." + kern + ":"))
lines, stack, fresh = [], [], True
cur_outer = ("?", 0)
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith("//-----"):
        break
    m = re.match(r'\s*//## File "(.*?)", line (\d+)(.*)', l)
    if m:
        loc = (os.path.basename(m.group(1)), int(m.group(2)))
        if fresh:
            stack = []
            fresh = False
        stack.append(loc)                # innermost frame first, then the frames it was inlined into
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        lines.append((list(stack), m.group(2).strip()))
        fresh = True
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[1], rows[2:]
ismp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
assert len(data) == len(lines)
agg = collections.OrderedDict((n, [0, 0, collections.Counter()]) for _, _, n in ranges)
agg["other"] = [0, 0, collections.Counter()]
for (st, op), r in zip(lines, data):
    s, e = int(r[ismp] or 0), int(r[iex] or 0)
    key = "other"
    for f, n in st:                      # the innermost frame inside the file decides
        if f == fname:
            for lo, hi, name in ranges:
                if lo <= n <= hi:
                    key = name
            break
    a = agg[key]
    a[0] += e; a[1] += s; a[2][op.split()[0].split(".")[0] if not op.startswith("@") else op.split()[1].split(".")[0]] += e
tot_e = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
for k, a in agg.items():
    top = ", ".join(f"{o} {100 * n / max(a[0], 1):.0f}%" for o, n in a[2].most_common(8))
    print(f"{k:24s} executed {a[0]:12d} ({100 * a[0] / tot_e:5.1f}%)  samples {a[1]:7d} ({100 * a[1] / tot_s:5.1f}%)  {top}")
