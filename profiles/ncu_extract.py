#!/usr/bin/env python
"""Print the metrics we track from an .ncu-rep (run where ncu is installed; no GPU needed).

    python profiles/ncu_extract.py gpurun_out/prof.ncu-rep
"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
]
STALLS = "smsp__average_warps_issue_stalled_"


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_col = hdr.index("Kernel Name")
    for r in data:
        print("==", r[name_col][:100])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:85s} {r[i]:>18s} {units[i]}")
        st = [(float(r[i]), h[len(STALLS):].replace("_per_issue_active.ratio", "")) for i, h in enumerate(hdr)
              if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")]
        for v, n in sorted(st, reverse=True)[:10]:
            print(f"  stall {n:40s} {v:8.3f} warps per issue")


if __name__ == "__main__":
    main(sys.argv[1])
