"""Opcode histogram of a kernel from an ncu source page (ncu -i X.ncu-rep --page source --csv):
executed warp instructions and stall samples per opcode, plus the hottest instructions.
    python tools/sass_hist.py X.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = txt.splitlines()
print(lines[0][:160])
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
isrc, iex, ist = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
ops, stalls = defaultdict(int), defaultdict(int)
tot = 0
body = []
for r in rows[1:]:
    if len(r) <= max(isrc, iex, ist):
        continue
    s = r[isrc].strip()
    tok = s.split()
    if not tok:
        continue
    op = tok[1] if tok[0].startswith("@") and len(tok) > 1 else tok[0]
    op = op.rstrip(";")
    base = op.split(".")[0]
    key = base + ("." + op.split(".")[1] if base in ("LDS", "LDG", "STS", "STG", "F2F", "I2F", "F2I", "MUFU") and "." in op else "")
    n, st = int(r[iex] or 0), int(r[ist] or 0)
    ops[key] += n
    stalls[key] += st
    tot += n
    body.append((n, st, s))
print(f"total executed warp instructions: {tot}")
print(f"{'opcode':14s} {'executed':>12s} {'share':>7s} {'stall samples':>14s}")
for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{k:14s} {v:12d} {100 * v / tot:6.1f}% {stalls[k]:14d}")
print("\nhottest instructions by stall samples:")
for n, st, s in sorted(body, key=lambda t: -t[1])[:top]:
    print(f"{st:8d} {n:10d}  {s[:100]}")
