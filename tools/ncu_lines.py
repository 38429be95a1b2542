"""Join an ncu SASS source page with nvdisasm line info: samples / executed instructions per CUDA line.
    python tools/ncu_lines.py REP.ncu-rep OBJ.o MANGLED_KERNEL [min_samples]
Prints, per source line (innermost inlined location), #SASS rows, executed warp-instructions and stall samples,
split into rows executed by every warp and rows executed by few warps (the serial section)."""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, obj, kern = sys.argv[1:4]
minsmp = int(sys.argv[4]) if len(sys.argv) > 4 else 15
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text." + kern + ":"))
lines, cur = [], ("?", 0)
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith("//-----"):
        break
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        if "inlined at" not in l:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        lines.append((cur, m.group(2).strip()))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[1], rows[2:]
ismp, iex, isrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
assert len(data) == len(lines), (len(data), len(lines))
mx = max(int(r[iex] or 0) for r in data)
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
tot = 0
for (loc, op), r in zip(lines, data):
    s, e = int(r[ismp] or 0), int(r[iex] or 0)
    a = agg[loc]
    a[0] += 1; a[1] += e; a[2] += s
    if e < 0.3 * mx:
        a[3] += s
    tot += s
src = {}
def text(loc):
    f, n = loc
    if f not in src:
        p = os.path.join(os.path.dirname(os.path.abspath(obj)), f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    return src[f][n - 1].strip()[:90] if 0 < n <= len(src[f]) else ""
print(f"total samples {tot}; rows {len(data)}; max executed {mx}")
for loc, a in sorted(agg.items(), key=lambda t: -t[1][2]):
    if a[2] < minsmp:
        break
    print(f"{a[2]:6d} {100 * a[2] / tot:5.1f}%  serial {a[3]:5d}  rows {a[0]:3d}  ex {a[1]:9d}  {loc[0]}:{loc[1]:<5d} {text(loc)}")
