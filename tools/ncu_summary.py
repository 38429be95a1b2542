"""Summarise .ncu-rep captures (read on the CPU box with `ncu -i`) into a text file for profiles/.
    python tools/ncu_summary.py OUT.txt "header line" REP1.ncu-rep [REP2.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__cluster_size',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max']


def main():
    out_path, header, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    with open(out_path, "w") as out:
        out.write(header + "\n")
        for rep in reps:
            txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
            rows = list(csv.reader(io.StringIO(txt)))
            hdr, units = rows[0], rows[1]
            for r in rows[2:]:
                out.write(f"\n== {rep.split('/')[-1]}: {r[hdr.index('Kernel Name')][:120]}\n")
                for w in WANT:
                    if w in hdr:
                        i = hdr.index(w)
                        out.write(f"  {w:68s} {r[i]:>18s} {units[i]}\n")


if __name__ == "__main__":
    main()
