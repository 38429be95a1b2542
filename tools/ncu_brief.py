"""Brief of one .ncu-rep: duration, instruction classes by execution count, issue utilisation, stall mix.
    python tools/ncu_brief.py X.ncu-rep [units_per_launch]   (units = CTA-epochs, to normalise instruction counts)"""
import csv, io, subprocess, sys
from collections import Counter
rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, un, r = rows[0], rows[1], rows[2]
g = lambda k: r[hdr.index(k)] if k in hdr else "n/a"
print("kernel:", g("Kernel Name")[:90], "grid", g("launch__grid_size"), "block", g("launch__block_size"), "regs", g("launch__registers_per_thread"))
for k in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
          "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "lts__t_sectors_op_read.sum", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct"):
    if k in hdr:
        print(f"  {k:78s} {g(k):>16s} {un[hdr.index(k)]}")
st = [(h, float(r[i] or 0)) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
print("  stall cycles per issued instruction:", ", ".join(f"{h.split('stalled_')[1].split('_per_')[0]} {v:.2f}" for h, v in sorted(st, key=lambda t: -t[1])[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(io.StringIO("\n".join(src[1:])))); h2 = rows[0]
iex = h2.index("Instructions Executed")
ex = sorted((int(x[iex] or 0) for x in rows[1:] if len(x) > iex), reverse=True)
if units and ex:
    loop = ex[0]
    cls = Counter()
    for n in ex:
        cls["loop (>= 40% of the hottest)" if n >= 0.4 * loop else "per warp-epoch" if n >= 4 * units else "per CTA-epoch" if n >= 0.5 * units else "rare"] += n
    for k, v in cls.items():
        print(f"  instructions {k:32s} {v / units:10.1f} per CTA-epoch")
